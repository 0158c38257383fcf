#!/bin/bash
# 8-GPU box: NCCL sharding test, then the bench line at N = 2, 4, 8 (config 2, weak) and config 5 (8192 clips split) at N = 8.
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/final_bench_multi.sh'
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/b
mkdir -p $O
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2
run() { name=$1; n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $n "$@" --out $O/$name.json > $O/$name.log 2> $O/$name.err
  echo "$name rc=$? $(python -c "
import json; d=json.load(open('$O/$name.json')); print(d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), json.dumps(d.get('strong_8192'))[:300])" 2>/dev/null)"; }
run cfg2_n8 8 --steps 10 --warmup 3 --no-cpu
run cfg5_n8 8 --config 5 --steps 10 --warmup 3 --no-cpu
run cfg2_n4 4 --steps 10 --warmup 3 --no-cpu
[ -n "${LM_MULTI_SHORT:-}" ] || run cfg2_n2 2 --steps 10 --warmup 3 --no-cpu
