#!/bin/bash
# dense inverse tests + Newton setup A/B, row-split mid-shape A/B on the large grids
set -x
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_ks.py -x -q 2>&1 | tail -8 > gpurun_out/r02d_tests.log
cat gpurun_out/r02d_tests.log
HANK_NEWTON_TRACE=1 HANK_NO_JBAR_CACHE=1 timeout 300 python tools/newton_time.py > gpurun_out/r02d_newton_gj.log 2>&1
HANK_CUSOLVER=1 HANK_NEWTON_TRACE=1 HANK_NO_JBAR_CACHE=1 timeout 300 python tools/newton_time.py > gpurun_out/r02d_newton_cusolver.log 2>&1
tail -4 gpurun_out/r02d_newton_gj.log gpurun_out/r02d_newton_cusolver.log
: > gpurun_out/r02d_mid.jsonl
for mid in 0 1 2 3; do
  HANK_RS_MID=$mid timeout 200 python tools/sweep_times.py --shape 2000 11 500 --lanes 64 --tag mid$mid >> gpurun_out/r02d_mid.jsonl 2>&1
  HANK_RS_MID=$mid timeout 200 python tools/sweep_times.py --shape 1000 7 300 --lanes 64 --tag mid$mid >> gpurun_out/r02d_mid.jsonl 2>&1
done
HANK_NO_WIDE=1 timeout 200 python tools/sweep_times.py --shape 1000 7 300 --lanes 296 444 --tag nowide >> gpurun_out/r02d_mid.jsonl 2>&1
timeout 200 python tools/sweep_times.py --shape 1000 7 300 --lanes 296 --tag default >> gpurun_out/r02d_mid.jsonl 2>&1
cut -c1-330 gpurun_out/r02d_mid.jsonl
