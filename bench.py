#!/usr/bin/env python3
"""Headline benchmark: 30 s-clip log-mel clips/sec on B200 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one pass of the hot path over one batch of synthetic clips.  At N = 1 the
workload is BASELINE.json configs[1]: Whisper-large-v3 features (128 mels) for 4096 synthetic
30 s 16 kHz clips resident in HBM (7.86 GB in, 6.29 GB out -- far larger than the 126 MB L2,
so no L2 flush is needed between iterations).  For N > 1 (torchrun, one rank per GPU) every
rank owns its own 4096-clip shard (weak scaling, no data-path collective); the time is the
max over ranks of the CUDA-event time of the K steps.

Printed JSON (one line, rank 0): value = whole-job clips/s with inputs resident in HBM;
e2e = the same metric through the public host-buffer API (pinned host -> device -> pinned
host inside the timed region); roofline = the fused kernel against the measured HBM peak;
cpu_baseline = the reference's own CPU call timed on this box's host cores.

`--impl reference` times the reference CPU implementation itself (HF
``WhisperFeatureExtractor.__call__``, the call at /root/reference/AB/fineTune.py:88; the NumPy
oracle port if transformers is not importable) on a bounded sample per step.
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

N_MELS = 128
N_SAMPLES = 480000
N_FRAMES = 3000
CLIPS_PER_GPU = 4096
E2E_CLIPS = 512
BYTES_PER_CLIP = N_SAMPLES * 4 + N_MELS * N_FRAMES * 4          # 3,456,000 B
METRIC = "30s-clip log-mel clips/sec"
UNIT = "clips/s"
WORKLOAD = "whisper-large-v3 128-mel log-mel, 4096 synthetic 30 s 16 kHz clips per GPU (BASELINE configs[1])"


def set_mels(n: int):
    """side measurements on the 80-mel shape; the default run is untouched"""
    global N_MELS, BYTES_PER_CLIP, WORKLOAD
    N_MELS = n
    BYTES_PER_CLIP = N_SAMPLES * 4 + N_MELS * N_FRAMES * 4
    WORKLOAD = f"whisper {n}-mel log-mel, synthetic 30 s 16 kHz clips (side measurement, not BASELINE configs[1])"


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; send all of that to stderr and keep the real
    stdout for the one JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """per-launch DRAM bytes of the fused kernel from the committed ncu --set full capture, if any"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return d
    except Exception:
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
def reference_callable():
    """(fn(list_of_clips) -> features, kind, description).  kind 'reference' = the unmodified HF call."""
    import numpy as np
    try:
        from transformers import WhisperFeatureExtractor
        fe = WhisperFeatureExtractor(feature_size=N_MELS)

        def run(x):
            return fe(list(x), sampling_rate=16000, return_tensors="pt")["input_features"]

        return run, "reference", "transformers WhisperFeatureExtractor.__call__ (feature_size=128, torch CPU STFT)"
    except Exception:
        from oracle import logmel_oracle as O

        def run(x):
            return O.whisper_logmel(np.asarray(x), n_mels=N_MELS)

        return run, "port", "oracle/logmel_oracle.py whisper_logmel (NumPy float64 port)"


def cpu_threads():
    try:
        import torch
        return int(torch.get_num_threads())
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(budget_s: float = 12.0, chunk: int = 32):
    import numpy as np
    run, kind, desc = reference_callable()
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((chunk, N_SAMPLES)) * 0.1).astype(np.float32)
    run(x[:4])                                      # warm-up (thread pools, FFT plans)
    n, t0 = 0, time.perf_counter()
    while True:
        run(x)
        n += chunk
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 4096:
            break
    return {"value": n / el, "unit": UNIT, "cores": cpu_threads() if kind == "reference" else 1, "kind": kind,
            "sample": f"{n} of the workload's clips in chunks of {chunk} through {desc}; {el:.1f} s wall, "
                      f"os.cpu_count()={os.cpu_count()}"}


def run_reference(args, rank: int):
    if rank != 0:
        return
    import numpy as np
    run, kind, desc = reference_callable()
    chunk = 32
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((chunk, N_SAMPLES)) * 0.1).astype(np.float32)
    for _ in range(max(args.warmup, 1)):
        run(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(x)
    el = time.perf_counter() - t0
    v = chunk * args.steps / el
    cores = cpu_threads() if kind == "reference" else 1
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": chunk, "n_mels": N_MELS, "device": "host CPU"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"each step = {chunk} clips of the workload through {desc}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mlx8_ws_audio_transformer_b200 import LogMelFrontend, launch_count
    from mlx8_ws_audio_transformer_b200 import _native as N
    from mlx8_ws_audio_transformer_b200.filters import slaney_mel_filter_bank

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fe = LogMelFrontend(400, 160, slaney_mel_filter_bank(201, N_MELS), N.LOG10_CLAMP_WHISPER_NORM, 1e-10, True,
                        device=local_rank, variant=args.variant)
    B = args.clips
    x = torch.empty((B, N_SAMPLES), dtype=torch.float32, device=dev)
    x.normal_(0.0, 0.1, generator=torch.Generator(device=dev).manual_seed(rank))
    out = torch.empty((B, N_MELS, N_FRAMES), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        fe.forward(x, out=out)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    n0 = launch_count()
    w0 = time.time()
    ev[0].record()
    for k in range(args.steps):
        fe.forward(x, out=out)
        ev[k + 1].record()
    torch.cuda.synchronize()
    w1 = time.time()
    launches = launch_count() - n0
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    checksum = float(out[:: max(1, B // 8)].double().mean().item())

    # ---- end to end: pinned host buffers through the public host API ----------------------
    eb = min(E2E_CLIPS, B)
    hx = torch.empty((eb, N_SAMPLES), dtype=torch.float32).pin_memory()
    hx.copy_(x[:eb])
    hy = torch.empty((eb, N_MELS, N_FRAMES), dtype=torch.float32).pin_memory()
    for _ in range(2):
        fe.forward_host(hx, out=hy)
    e_steps = max(3, min(args.steps, 10))
    barrier()
    e0 = time.perf_counter()
    for _ in range(e_steps):
        fe.forward_host(hx, out=hy)                 # returns only when hy is complete on the host
    torch.cuda.synchronize()
    e_el = time.perf_counter() - e0
    te = torch.tensor([e_el], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * eb * e_steps / float(te.item())
    e2e_ok = bool(torch.equal(hy[:4], out[:4].cpu()))

    if rank == 0:
        peak, peak_src = load_peak()
        kern_ms = statistics.mean(step_ms)           # one fused launch per step (plus a 16 KB memset)
        achieved = BYTES_PER_CLIP * B / (kern_ms * 1e-3) / 1e9
        traffic = load_traffic()
        info = fe.kernel_info()
        line = {
            "metric": METRIC, "value": world * B * args.steps / (max_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_gpu": B, "n_mels": N_MELS, "n_samples": N_SAMPLES,
                       "frames": N_FRAMES, "parallelism": f"clip-sharded x{world}, no collective",
                       "l2": "inputs+outputs 14.2 GB per step >> 126 MB L2, no flush needed",
                       "kernel": info, "variant": args.variant, "checksum": checksum},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if not traffic else traffic.get("dram_bytes_per_launch"),
                         "peak_source": peak_src, "kernel": "lm::logmel_ws_kernel<Geo<400,160,2,1>,3>",
                         "algorithmic_bytes_per_launch": BYTES_PER_CLIP * B, "kernel_ms": kern_ms,
                         "kernel_ms_min": min(step_ms), "traffic_source": None if not traffic else traffic.get("source")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eb * N_SAMPLES * 4,
                    "d2h_bytes_per_step": eb * N_MELS * N_FRAMES * 4, "clips_per_step": eb, "steps": e_steps,
                    "api": "LogMelFrontend.forward_host -> lm_forward_host (pinned host in/out)",
                    "matches_device_path": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU, help="clips per GPU per step")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (lm_config.variant)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--mels", type=int, default=N_MELS, choices=[80, 128],
                    help="128 = BASELINE configs[1] (the headline); 80 = the Whisper-tiny shape of configs[0]/[4], side measurement")
    args = ap.parse_args()
    if args.mels != N_MELS:
        set_mels(args.mels)
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when asked for N > 1 from a plain python call
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
