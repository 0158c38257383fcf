#!/usr/bin/env python3
"""Headline benchmark: 30 s-clip log-mel clips/sec on B200 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

One "step" is one pass of the hot path over one batch of synthetic clips.  The default workload
(--config 2) is BASELINE.json configs[1]: Whisper-large-v3 features (128 mels) for 4096 synthetic
30 s 16 kHz clips per GPU, resident in HBM (7.86 GB in, 6.29 GB out per step -- far larger than the
126 MB L2, so no L2 flush is needed between iterations).  For N > 1 (torchrun, one rank per GPU)
every rank owns its own shard (weak scaling, no data-path collective); the time is the max over
ranks of the CUDA-event time of the K steps.  The other BASELINE configs are side legs:

    --config 1   Whisper-tiny 80 mels, B = 32, default_rng(0) clips (launch-bound on a GPU)
    --config 3   UrbanSound8K-shaped 8732 x 4 s clips through the n_fft 1024 frontend (--hop 512|128, --mels 128|64)
    --config 4   MIDI-piano clips, B = 1000, 80 mels, mostly zero padding (per-clip lengths)
    --config 5   8192 x 30 s clips, 80 mels, split over the N GPUs (strong scaling)

Printed JSON (one line, rank 0): value = whole-job clips/s with inputs resident in HBM; e2e = the
same metric through the public host-buffer API (pinned host -> device -> pinned host inside the
timed region); roofline = the fused kernel against the measured HBM peak; parity = a 32-clip subset
of the timed batch against the oracle (outside the timed region); cpu_baseline = the reference's own
CPU call on this box's host cores; gpu_library_baseline = the reference's library ops (torch.stft /
cuFFT + cuBLAS, torchaudio) on the same GPU.

`--impl reference` times the reference CPU implementation itself (HF
``WhisperFeatureExtractor.__call__``, the call at /root/reference/AB/fineTune.py:88, or torchaudio's
``MelSpectrogram`` for config 3; the NumPy oracle port if a library is not importable) on a
bounded sample per step, with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "30s-clip log-mel clips/sec"
UNIT = "clips/s"
E2E_CLIPS = 512


class Workload:
    """one BASELINE.json config: shape, algorithmic bytes, how to build the operator"""

    def __init__(self, cfg: int, mels: int | None, hop: int | None, clips: int | None, world: int):
        self.cfg = cfg
        self.kind = "torchaudio" if cfg == 3 else "whisper"
        self.scaling = "strong" if cfg == 5 else "weak"
        if self.kind == "whisper":
            self.n_mels = mels or (128 if cfg == 2 else 80)
            self.n_samples, self.n_fft, self.hop = 480000, 400, 160
            self.n_frames = 3000
        else:
            self.n_mels = mels or 128
            self.n_samples, self.n_fft, self.hop = 64000, 1024, hop or 512
            self.n_frames = 1 + self.n_samples // self.hop
        total = {1: 32, 2: 4096, 3: 8732, 4: 1000, 5: 8192}[cfg]
        if clips:
            total = clips
        self.clips_per_gpu = total // world if cfg == 5 else total
        self.bytes_per_clip = self.n_samples * 4 + self.n_mels * self.n_frames * 4
        self.lengths = cfg in (3, 4)
        names = {
            1: f"whisper-tiny {self.n_mels}-mel log-mel, {total} synthetic 30 s 16 kHz clips, default_rng(0) (BASELINE configs[0])",
            2: f"whisper-large-v3 {self.n_mels}-mel log-mel, {total} synthetic 30 s 16 kHz clips per GPU (BASELINE configs[1])",
            3: f"UrbanSound8K-shaped {total} x 4 s clips, torchaudio MelSpectrogram n_fft 1024 hop {self.hop} {self.n_mels} mels + log(mel+1e-6) (BASELINE configs[2])",
            4: f"MIDI-piano clips, {total} x 30 s containers (1.5-7.7 s of audio, per-clip lengths), {self.n_mels} mels (BASELINE configs[3])",
            5: f"clip-sharded sweep, {total} synthetic 30 s clips split over the GPUs, {self.n_mels} mels (BASELINE configs[4])",
        }
        self.name = names[cfg]
        if cfg == 2 and self.n_mels != 128:
            self.name = f"whisper {self.n_mels}-mel log-mel, {total} synthetic 30 s 16 kHz clips per GPU (side measurement of configs[1] at {self.n_mels} mels)"

    def frontend(self, device: int, variant: int):
        from mlx8_ws_audio_transformer_b200 import LogMelFrontend
        from mlx8_ws_audio_transformer_b200 import _native as N
        from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank, torchaudio_mel_filter_bank
        if self.kind == "whisper":
            return LogMelFrontend(400, 160, slaney_mel_filter_bank(201, self.n_mels), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True,
                                  device=device, variant=variant)
        fb = torchaudio_mel_filter_bank(513, 0.0, 8000.0, self.n_mels, 16000, None, "htk")
        return LogMelFrontend(1024, self.hop, fb, N.LN_PLUS_EPS, 1e-6, False, device=device, variant=variant)

    def make_input(self, dev, rank: int):
        """(wave[B, L] float32 on dev, lengths[B] int32 on dev or None)"""
        import numpy as np
        import torch
        from mlx8_ws_audio_transformer_b200 import synth
        B = self.clips_per_gpu
        if self.cfg == 1:      # the exact seeded batch of SURVEY.md 8d
            return torch.from_numpy(synth.gaussian_clips(B, seed=0)).to(dev), None
        if self.cfg == 4:
            base = min(B, 250)
            w, n = synth.midi_piano_clips(base, seed=rank)
            reps = (B + base - 1) // base
            w = np.tile(w, (reps, 1))[:B]
            n = np.tile(n, reps)[:B]
            return torch.from_numpy(w).to(dev), torch.from_numpy(n).to(dev)
        x = torch.empty((B, self.n_samples), dtype=torch.float32, device=dev)
        x.normal_(0.0, 0.1, generator=torch.Generator(device=dev).manual_seed(rank))
        if self.cfg == 3:      # active length U{8000..64000}, zero padded (spectrogram.py:152-157)
            n = torch.randint(8000, self.n_samples + 1, (B,), generator=torch.Generator(device=dev).manual_seed(100 + rank),
                              device=dev, dtype=torch.int32)
            x *= (torch.arange(self.n_samples, device=dev)[None, :] < n[:, None])
            return x, n
        return x, None

    def oracle(self, x_np, lengths_np):
        import numpy as np
        from oracle import logmel_oracle as O       # checker only (parity field / CPU port)
        if lengths_np is not None:
            x_np = x_np.copy()
            for i, n in enumerate(lengths_np):
                x_np[i, int(n):] = 0.0
        if self.kind == "whisper":
            return O.whisper_logmel(x_np, n_mels=self.n_mels)
        fb = O.htk_mel_filter_bank_f32(513, self.n_mels, 0.0, 8000.0, 16000)
        return O.torchaudio_mel(x_np, fb, 1024, self.hop, 1e-6)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; send all of that to stderr and keep the real
    stdout for the one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict, out_path: str | None = None):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)
    if out_path:
        with open(out_path, "w") as f:
            json.dump(line, f, indent=1)
            f.write("\n")


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(kernel_name: str):
    """per-launch DRAM bytes of the fused kernel from the committed ncu --set full capture, if it is of THIS kernel"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for rec in d.get("kernels", [d]):
            if rec.get("kernel") and rec["kernel"].replace(" ", "") in kernel_name.replace(" ", ""):
                return rec
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        sm, smax, reasons, power = [], [], set(), []
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
                power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference CPU implementation (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is ONE process that owns
    the host, so it takes every core again -- the same CPU path at every N."""
    n = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(n)
        return int(torch.get_num_threads())
    except Exception:
        return 1


def reference_callable(wl: Workload):
    """(fn(batch ndarray) -> features, kind, description).  kind 'reference' = the unmodified library call."""
    import numpy as np
    if wl.kind == "whisper":
        try:
            from transformers import WhisperFeatureExtractor
            fe = WhisperFeatureExtractor(feature_size=wl.n_mels)

            def run(x):
                return fe(list(x), sampling_rate=16000, return_tensors="pt")["input_features"]

            return run, "reference", f"transformers WhisperFeatureExtractor.__call__ (feature_size={wl.n_mels}, torch CPU STFT)"
        except Exception:
            pass
    else:
        try:
            import torch
            import torchaudio
            mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, hop_length=wl.hop, n_mels=wl.n_mels,
                                                       f_min=0, f_max=8000, power=2.0)

            def run(x):
                return torch.log(mel(torch.from_numpy(np.ascontiguousarray(x))) + 1e-6)

            return run, "reference", f"torchaudio MelSpectrogram(n_fft=1024, hop={wl.hop}, n_mels={wl.n_mels}) + torch.log(mel + 1e-6), batched on the CPU"
        except Exception:
            pass

    def run(x):
        return wl.oracle(np.asarray(x), None)

    return run, "port", "oracle/logmel_oracle.py (NumPy float64 port)"


def reference_chunk(wl: Workload) -> int:
    return 32 if wl.kind == "whisper" else 256


def cpu_baseline(wl: Workload, budget_s: float = 12.0):
    import numpy as np
    cores = use_all_host_threads()
    run, kind, desc = reference_callable(wl)
    chunk = reference_chunk(wl)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((chunk, wl.n_samples)) * 0.1).astype(np.float32)
    run(x[:4])                                      # warm-up (thread pools, FFT plans)
    n, t0 = 0, time.perf_counter()
    while True:
        run(x)
        n += chunk
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 4 * wl.clips_per_gpu:
            break
    return {"value": n / el, "unit": UNIT, "cores": cores if kind == "reference" else 1, "kind": kind,
            "sample": f"{n} of the workload's clips in chunks of {chunk} through {desc}; {el:.1f} s wall, "
                      f"os.cpu_count()={os.cpu_count()}"}


def run_reference(args, wl: Workload, rank: int):
    if rank != 0:
        return
    import numpy as np
    cores = use_all_host_threads()
    run, kind, desc = reference_callable(wl)
    chunk = reference_chunk(wl)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((chunk, wl.n_samples)) * 0.1).astype(np.float32)
    for _ in range(max(args.warmup, 1)):
        run(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(x)
    el = time.perf_counter() - t0
    v = chunk * args.steps / el
    cores = cores if kind == "reference" else 1
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": wl.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "clips_per_step": chunk, "n_mels": wl.n_mels, "device": "host CPU",
                   "threads": cores, "os_cpu_count": os.cpu_count()},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"each step = {chunk} clips of the workload through {desc}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line, args.out)


# ---------------------------------------------------------------------------------------------
# the reference's library ops on the SAME GPU (the "existing GPU" comparator)
# ---------------------------------------------------------------------------------------------
def gpu_library_baseline(wl: Workload, x, lengths, steps: int = 3):
    """cuFFT (torch.stft) + cuBLAS + elementwise kernels, exactly the ops the libraries run when the
    reference's calls are placed on CUDA: HF ``_torch_extract_fbank_features`` (feature_extraction_whisper.py:
    135-164) / ``mel_spectrogram.to(device)`` + ``torch.log`` (/root/reference/.charles/spectrogram.py:87,160-162).
    Device-resident in and out, chunked so that the complex64 spectrum fits; CUDA-event timed."""
    import torch
    B = x.shape[0]
    rec = {}
    try:
        if wl.kind == "whisper":
            from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank
            fb = torch.from_numpy(slaney_mel_filter_bank(201, wl.n_mels)).to(x.device, torch.float32)
            window = torch.hann_window(400, device=x.device)
            chunk = 64

            def run_chunk(w):
                stft = torch.stft(w, 400, 160, window=window, return_complex=True)
                mag = stft[..., :-1].abs() ** 2
                mel = fb.T @ mag
                ls = torch.clamp(mel, min=1e-10).log10()
                mx = ls.max(dim=2, keepdim=True)[0].max(dim=1, keepdim=True)[0]
                ls = torch.maximum(ls, mx - 8.0)
                return (ls + 4.0) / 4.0
            what = "torch.stft + abs()**2 + mel_filters.T @ magnitudes + clamp/log10 + max + maximum + affine on cuda, 64-clip chunks"
        else:
            import torchaudio
            mel_t = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, hop_length=wl.hop, n_mels=wl.n_mels,
                                                         f_min=0, f_max=8000, power=2.0).to(x.device)
            chunk = 1024

            def run_chunk(w):
                return torch.log(mel_t(w) + 1e-6)
            what = f"torchaudio MelSpectrogram(hop={wl.hop}).to(cuda) + torch.log(mel + 1e-6), 1024-clip chunks"
        n = min(B, 16 * chunk)
        out = None
        for s in range(0, min(n, 2 * chunk), chunk):
            out = run_chunk(x[s:s + chunk])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            for s in range(0, n, chunk):
                out = run_chunk(x[s:s + chunk])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        rec["device_resident"] = {"value": n / (ms * 1e-3), "unit": UNIT, "clips": n, "ms": ms, "ops": what}
        del out
    except Exception as e:  # noqa: BLE001
        rec["device_resident"] = {"unavailable": repr(e)[:200]}
    if wl.kind == "whisper":
        try:        # the public HF call with device="cuda": host arrays in, host features out
            from transformers import WhisperFeatureExtractor
            fe = WhisperFeatureExtractor(feature_size=wl.n_mels)
            xs = list(x[:32].cpu().numpy())
            fe(xs, sampling_rate=16000, return_tensors="pt", device="cuda")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                fe(xs, sampling_rate=16000, return_tensors="pt", device="cuda")
            torch.cuda.synchronize()
            el = (time.perf_counter() - t0) / steps
            rec["hf_call_device_cuda"] = {"value": 32 / el, "unit": UNIT, "clips": 32, "ms": el * 1e3,
                                          "ops": "WhisperFeatureExtractor(list, sampling_rate=16000, device='cuda'): host in, host out"}
        except Exception as e:  # noqa: BLE001
            rec["hf_call_device_cuda"] = {"unavailable": repr(e)[:200]}
    return rec


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def kernel_geometry(fe, kname):
    """launch geometry of the kernel the timed launch takes (lm_kernel_info describes the handle's CTA-tiled kernel;
    the thread-per-frame kernel's constants are TfGeo in csrc/logmel_tf_kernel.cuh, sizes from cuobjdump -res-usage)"""
    info = fe.kernel_info()
    if "logmel_tf_kernel" in kname:
        return {"n_sm": info["n_sm"], "ctas_per_sm": 1, "threads": 256, "warp_pairs_per_cta": 4, "frames_per_tile": 32,
                "smem_bytes": 174976 + 11568, "tmem_columns": 512}
    return info


def time_steps(fe, x, lengths, out, steps):
    import torch
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        fe.forward(x, lengths=lengths, out=out)
        ev[k + 1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[-1]), [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]


def run_ours(args, wl: Workload, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mlx8_ws_audio_transformer_b200 import launch_count

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fe = wl.frontend(local_rank, args.variant)
    B = wl.clips_per_gpu
    x, lengths = wl.make_input(dev, rank)
    out = torch.empty((B, wl.n_mels, wl.n_frames), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        fe.forward(x, lengths=lengths, out=out)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    n0 = launch_count()
    w0 = time.time()
    total_ms, step_ms = time_steps(fe, x, lengths, out, args.steps)
    w1 = time.time()
    launches = launch_count() - n0
    barrier()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    max_ms = max_over_ranks(total_ms)
    checksum = float(out[:: max(1, B // 8)].double().mean().item())
    kname = fe.kernel_name(B, wl.n_samples)

    # ---- parity of a subset of the timed batch against the oracle (rank 0, outside the timed region)
    parity = None
    if rank == 0 and not args.no_parity:
        idx = np.unique(np.linspace(0, B - 1, min(B, 32)).astype(np.int64))
        ref = wl.oracle(x[idx].cpu().numpy(), None if lengths is None else lengths[idx].cpu().numpy())
        got = out[idx].cpu().numpy()
        d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        parity = {"clips": int(len(idx)), "max_abs": float(d.max()), "mean_abs": float(d.mean()),
                  "tolerance": {"max_abs": 1e-3, "mean_abs": 1e-5}, "ok": bool(d.max() < 1e-3 and d.mean() < 1e-5),
                  "checker": "oracle/logmel_oracle.py (float64) on clips of the timed batch"}

    # ---- end to end: pinned host buffers through the public host API ----------------------
    eb = min(E2E_CLIPS, B)
    hx = torch.empty((eb, wl.n_samples), dtype=torch.float32).pin_memory()
    hx.copy_(x[:eb])
    hl = None if lengths is None else lengths[:eb].cpu()
    hy = torch.empty((eb, wl.n_mels, wl.n_frames), dtype=torch.float32).pin_memory()
    for _ in range(2):
        fe.forward_host(hx, lengths=hl, out=hy)
    e_steps = max(3, min(args.steps, 10))
    barrier()
    e0 = time.perf_counter()
    for _ in range(e_steps):
        fe.forward_host(hx, lengths=hl, out=hy)      # returns only when hy is complete on the host
    torch.cuda.synchronize()
    e_el = max_over_ranks(time.perf_counter() - e0)
    e2e_value = world * eb * e_steps / e_el
    e2e_diff = float((hy[:4] - out[:4].cpu()).abs().max())

    # ---- fused 16-bit PCM ingest, a SEPARATE metric: the same clips as int16 mono frames, converted
    #      in the kernel's tile loader; device-resident and end to end (H2D bytes halve)
    pcm = None
    if wl.kind == "whisper" and wl.cfg == 2 and not args.no_extras:
        xi = (x * 32768.0).clamp_(-32768, 32767).round_().to(torch.int16)          # [B, T] mono PCM
        fe.forward(xi, out=out)
        barrier()
        p_total, _ = time_steps(fe, xi, None, out, max(3, args.steps // 2))
        p_ms = max_over_ranks(p_total) / max(3, args.steps // 2)
        nb = min(B, 640)                 # enough clips for the float32 launch to take the same kernel as the timed one
        same = float((out[:4] - fe.forward(xi[:nb].float() / 32768.0)[:4]).abs().max())
        hxi = torch.empty((eb, wl.n_samples), dtype=torch.int16).pin_memory()
        hxi.copy_(xi[:eb])
        for _ in range(2):
            fe.forward_host(hxi, out=hy)
        barrier()
        p0 = time.perf_counter()
        for _ in range(e_steps):
            fe.forward_host(hxi, out=hy)
        torch.cuda.synchronize()
        p_el = max_over_ranks(time.perf_counter() - p0)
        bpc = wl.n_samples * 2 + wl.n_mels * wl.n_frames * 4
        pcm = {"value": world * B / (p_ms * 1e-3), "unit": UNIT, "ms_per_step": p_ms, "algorithmic_bytes_per_clip": bpc,
               "roofline_frac": bpc * B / (p_ms * 1e-3) / 1e9 / load_peak()[0],
               "max_abs_vs_float32_path_on_the_converted_clips": same,
               "e2e": {"value": world * eb * e_steps / p_el, "unit": UNIT, "h2d_bytes_per_step": eb * wl.n_samples * 2,
                       "d2h_bytes_per_step": eb * wl.n_mels * wl.n_frames * 4, "api": "lm_forward_host_pcm16"},
               "input": "the timed batch quantised to int16 mono (what AB/memoToWav.py:19 writes), x / 32768 inside the kernel"}
        del xi, hxi

    # ---- the same kernel on an input where EVERY tile needs the max-8 clamp (noise 80 dB under one burst)
    clamp = None
    if wl.kind == "whisper" and wl.cfg in (2, 5) and not args.no_extras:
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        x.normal_(0.0, 1e-5, generator=g)
        tt = torch.arange(400, device=dev) / 16000.0
        pos = torch.randint(0, wl.n_samples - 400, (B,), generator=g, device=dev)
        burst = 0.9 * torch.sin(2 * np.pi * 440.0 * tt)
        x[torch.arange(B, device=dev)[:, None], pos[:, None] + torch.arange(400, device=dev)[None, :]] += burst[None, :]
        fe.forward(x, out=out)
        barrier()
        c_total, _ = time_steps(fe, x, None, out, max(3, args.steps // 4))
        c_ms = max_over_ranks(c_total) / max(3, args.steps // 4)
        rng_ok = bool(((out[:8].reshape(8, -1).max(dim=1).values - out[:8].reshape(8, -1).min(dim=1).values) - 2.0).abs().max() < 1e-4)
        clamp = {"ms_per_step": c_ms, "value": world * B / (c_ms * 1e-3), "unit": UNIT, "clamp_active_in_every_clip": rng_ok,
                 "input": "1e-5 Gaussian noise + one 25 ms 0.9-amplitude burst per clip: every tile is revisited by the max-8 pass"}

    # ---- strong-scaling side record at N > 1 on the default workload: 8192 clips split N ways
    strong = None
    if wl.cfg == 2 and world > 1 and not args.no_extras:
        strong = {}
        for nm in (80, 128):
            w5 = Workload(5, nm, None, None, world)
            f5 = w5.frontend(local_rank, args.variant)
            b5 = w5.clips_per_gpu
            x5 = x[:b5]
            x5.normal_(0.0, 0.1, generator=torch.Generator(device=dev).manual_seed(rank))
            o5 = torch.empty((b5, nm, 3000), dtype=torch.float32, device=dev)
            for _ in range(3):
                f5.forward(x5, out=o5)
            barrier()
            t5, _ = time_steps(f5, x5, None, o5, args.steps)
            m5 = max_over_ranks(t5) / args.steps
            strong[f"mels{nm}"] = {"clips_total": b5 * world, "clips_per_gpu": b5, "ms_per_step": m5,
                                   "value": b5 * world / (m5 * 1e-3), "unit": UNIT, "kernel": f5.kernel_name(b5, 480000)}
            del o5, f5

    if rank == 0:
        peak, peak_src = load_peak()
        kern_ms = statistics.mean(step_ms)           # one fused launch per step
        # algorithmic bytes of this launch: samples read once + features written once; with per-clip lengths the
        # padding of a container is never read (the features of its frames are still written)
        algo_launch = wl.bytes_per_clip * B
        if lengths is not None:
            algo_launch = int(lengths.clamp(0, wl.n_samples).sum().item()) * 4 + wl.n_mels * wl.n_frames * 4 * B
        achieved = algo_launch / (kern_ms * 1e-3) / 1e9
        traffic = load_traffic(kname)
        step_gb = algo_launch / 1e9
        line = {
            "metric": METRIC, "value": world * B * args.steps / (max_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps,
            "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "baseline_config": wl.cfg, "clips_per_gpu": B, "n_mels": wl.n_mels,
                       "n_samples": wl.n_samples, "frames": wl.n_frames, "per_clip_lengths": wl.lengths,
                       "parallelism": f"clip-sharded x{world}, no collective",
                       "l2": (f"inputs+outputs {step_gb:.2f} GB per step >> 126 MB L2, no flush needed" if step_gb > 1.0 else
                              f"inputs+outputs {step_gb * 1e3:.0f} MB per step: fits the 126 MB L2, launch-bound shape, "
                              "reported but not graded against the HBM roofline"),
                       "kernel": kernel_geometry(fe, kname), "kernel_name": kname, "variant": args.variant, "checksum": checksum},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if not traffic else traffic.get("dram_bytes_per_launch"),
                         "peak_source": peak_src, "kernel": kname,
                         "algorithmic_bytes_per_clip": wl.bytes_per_clip,
                         "algorithmic_bytes_per_launch": algo_launch, "kernel_ms": kern_ms,
                         "kernel_ms_min": min(step_ms), "traffic_source": None if not traffic else traffic.get("source")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eb * wl.n_samples * 4 + (eb * 4 if hl is not None else 0),
                    "d2h_bytes_per_step": eb * wl.n_mels * wl.n_frames * 4, "clips_per_step": eb, "steps": e_steps,
                    "api": "LogMelFrontend.forward_host -> lm_forward_host (pinned host in/out)",
                    "max_abs_vs_device_path": e2e_diff},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if parity is not None:
            line["parity"] = parity
        if pcm is not None:
            line["pcm16_ingest"] = pcm
        if clamp is not None:
            line["clamp_everywhere"] = clamp
        if strong is not None:
            line["strong_8192"] = strong
        if world == 1 and not args.no_extras:
            x2, l2 = wl.make_input(dev, rank)
            line["gpu_library_baseline"] = gpu_library_baseline(wl, x2, l2)
            del x2
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(wl)
        emit(line, args.out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json config (1-based); 2 = configs[1], the headline")
    ap.add_argument("--clips", type=int, default=None, help="override the workload's clip count (per GPU; config 5: total)")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (lm_config.variant)")
    ap.add_argument("--mels", type=int, default=None, choices=[64, 80, 128])
    ap.add_argument("--hop", type=int, default=None, choices=[128, 512], help="config 3 only")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed batch")
    ap.add_argument("--no-extras", action="store_true", help="skip clamp_everywhere / strong_8192 / gpu_library_baseline")
    ap.add_argument("--out", default=None, help="also write the JSON line (indented) to this file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when asked for N > 1 from a plain python call
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    wl = Workload(args.config, args.mels, args.hop, args.clips, world if args.impl == "ours" else max(args.gpus, 1))
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    run_ours(args, wl, rank, local_rank, world)


if __name__ == "__main__":
    main()
