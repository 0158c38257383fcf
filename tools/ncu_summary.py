#!/usr/bin/env python3
"""Summarise an ncu report (.ncu-rep) of the fused log-mel kernel into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep <clips in the profiled launch> [bytes per clip] [frames per clip] \
        > profiles/rNN_ncu_full_summary.txt
"""
import csv
import io
import subprocess
import sys

rep, clips = sys.argv[1], int(sys.argv[2])
bytes_per_clip = int(sys.argv[3]) if len(sys.argv) > 3 else 3456000
frames_per_clip = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
frames = clips * frames_per_clip
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
print(f"# ncu --set full --clock-control none --import-source on   ({rep}, {clips} clips x {frames_per_clip} frames)")
print(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
keys = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
]
for k in keys:
    if k in d:
        print(f"{k:78s} {u[k]:14s} {d[k]}")
try:
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
    alg = clips * bytes_per_clip
    print(f"DRAM traffic {(rd + wr) / 1e9:.3f} GB  vs algorithmic {alg / 1e9:.3f} GB  -> x{(rd + wr) / alg:.3f}")
    ms = float(d["gpu__time_duration.sum"])
    print(f"cycles per frame per SM: {float(d['sm__cycles_elapsed.avg']) * 148 / frames:.1f}   instructions per frame: "
          f"{float(d['smsp__inst_executed.sum']) / frames:.1f}   (kernel {ms:.3f} ms under ncu)")
except Exception as e:  # noqa: BLE001
    print("derived figures unavailable:", e)
print("warp stall samples (all):")
st = {k.split("stalled_")[1]: int(float(v)) for k, v in d.items() if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k}
tot = sum(st.values()) or 1
print("  " + "  ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v))
