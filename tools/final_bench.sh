#!/bin/bash
# One B200: the bench line of every BASELINE config (JSON into gpurun_out/b/), the reference arm, then the ncu
# launch list and one `--set full` capture of the same bench command (summaries only travel back).
#   gpurun --timeout 2400 -- 'bash tools/final_bench.sh'
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/b
mkdir -p $O
run() { name=$1; shift; python bench.py "$@" --out $O/$name.json > $O/$name.log 2> $O/$name.err; echo "$name rc=$? $(python -c "
import json; d=json.load(open('$O/$name.json')); print(d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))" 2>/dev/null)"; }
run cfg2 --steps 20 --warmup 3
run cfg2_reference --impl reference --steps 3 --warmup 1
run cfg2_80mel --mels 80 --steps 20 --warmup 3 --no-cpu
run cfg1 --config 1 --steps 50 --warmup 5
run cfg3_hop512_128 --config 3 --steps 20 --warmup 3
run cfg3_hop128_128 --config 3 --hop 128 --steps 20 --warmup 3 --no-cpu
run cfg3_hop512_64 --config 3 --mels 64 --steps 20 --warmup 3 --no-cpu
run cfg4 --config 4 --steps 50 --warmup 5
run cfg5 --config 5 --steps 10 --warmup 3
# ncu: launch list of the plain bench command, then the full set on the headline kernel
python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-parity > $O/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-parity > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:logmel_tf --launch-skip 2 --launch-count 1 -o /tmp/full \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-parity > $O/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/full.ncu-rep 4096 > $O/ncu_full_summary.txt
ncu -i /tmp/full.ncu-rep --page raw --csv > $O/ncu_full_raw.csv
ls -la $O | head -50
