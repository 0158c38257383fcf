#!/bin/bash
# On a B200 box: GPU tests, then one `ncu --set full` capture each of the thread-per-frame kernel on 16-bit PCM and on
# float32 input (same process, after the plain command exited 0); reports stay in /tmp, summaries and the source /
# raw pages go to gpurun_out/n/ (small enough to travel back).
#   gpurun --timeout 1500 -- 'bash tools/profile_tf.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/n
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python tools/pcm_check.py | tee gpurun_out/n/pcm_check.txt
python tools/pcm_check.py --stereo --clips 2048 | tee -a gpurun_out/n/pcm_check.txt
python tools/tf_check.py --no-parity --sweep | tee gpurun_out/n/sweep.txt
python tools/pcm_check.py --clips 1184 --iters 2 || exit 1
NCU="ncu --set full --clock-control none --import-source on -k regex:logmel_tf --launch-count 1"
$NCU --launch-skip 4 -o /tmp/pcm python tools/pcm_check.py --clips 1184 --iters 2 > gpurun_out/n/ncu_pcm.log 2>&1
$NCU --launch-skip 9 -o /tmp/flt python tools/pcm_check.py --clips 1184 --iters 2 > gpurun_out/n/ncu_flt.log 2>&1
for r in pcm flt; do
  ncu -i /tmp/$r.ncu-rep --page source --csv > gpurun_out/n/${r}_source.csv 2>/dev/null
  ncu -i /tmp/$r.ncu-rep --page raw --csv > gpurun_out/n/${r}_raw.csv
done
python tools/ncu_summary.py /tmp/pcm.ncu-rep 1184 2496000 > gpurun_out/n/pcm_summary.txt
python tools/ncu_summary.py /tmp/flt.ncu-rep 1184 > gpurun_out/n/flt_summary.txt
ls -la gpurun_out/n /tmp/*.ncu-rep
