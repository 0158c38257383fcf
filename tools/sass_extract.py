#!/usr/bin/env python3
"""SASS evidence for the headline kernel: opcode histogram and the mnemonics that prove the design.

    python tools/sass_extract.py [path/to/liblogmel_b200.so] > profiles/rNN_sass_extract.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mlx8_ws_audio_transformer_b200", "csrc", "liblogmel_b200.so")
KERNEL = "logmel_tf_kernelILi128ELi3000ELb0E"

sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
ops, on = [], False
for line in sass.splitlines():
    if "Function :" in line:
        on = KERNEL in line
        continue
    if not on:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops.append(m.group(1))
usage = ""
lines = res.splitlines()
for i, line in enumerate(lines):
    if KERNEL in line and i + 1 < len(lines):
        usage = lines[i + 1].strip()
        break
full = collections.Counter(ops)
base = collections.Counter(o.split(".")[0] for o in ops)


def n(prefix):
    return sum(v for k, v in full.items() if k.startswith(prefix))


print(f"# SASS extract of the headline kernel (cuobjdump -sass {os.path.basename(so)}, function lm::logmel_tf_kernel<128, 3000, false>)")
print(f"# {len(ops)} instructions ({len(ops) * 16 / 1024:.1f} KB); nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo")
print(f"# cuobjdump -res-usage: {usage}")
print("#")
print("# what proves the design (B200_PROFILING.md, 'What proves a Blackwell-native kernel'):")
print(f"#   tensor memory as lane-private scratch : LDTM {n('LDTM')}  STTM {n('STTM')}  (tcgen05.ld / tcgen05.st; UTCATOMSWS {n('UTCATOMSWS')} = tcgen05.alloc / dealloc)")
print(f"#   no tensor-core math                   : UTC*MMA {sum(v for k, v in full.items() if k.startswith('UTC') and 'MMA' in k)}  HMMA {n('HMMA')}   (FP32 SIMT by design, DESIGN.md 4.4)")
print(f"#   packed FP32 pipe                      : FFMA2 {n('FFMA2')}  FADD2 {n('FADD2')}  FMUL2 {n('FMUL2')}   scalar FFMA {base['FFMA']} FADD {base['FADD']} FMUL {base['FMUL']}")
print(f"#   3-input min/max, warp reduce          : FMNMX3 {n('FMNMX3')}  CREDUX {n('CREDUX')}  MUFU.LG2 {n('MUFU.LG2')}")
print(f"#   async copies                          : LDGSTS {n('LDGSTS')} (cp.async 16 B)   LDS.128 {n('LDS.128')}  LDS.64 {n('LDS.64')}")
print(f"#   16-bit PCM ingest                     : PRMT {n('PRMT')} (splice into the mantissa)   I2F {n('I2F.')} / I2FP {n('I2FP')} only in the clip-edge sample loader")
print(f"#   named barriers (warp pairs)           : BAR {n('BAR')}   local memory: LDL {n('LDL')} STL {n('STL')}")
print("#")
print("# opcode histogram:")
for k, v in base.most_common():
    print(f"#   {k:14s} {v}")
