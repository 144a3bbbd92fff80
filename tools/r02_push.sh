#!/bin/bash
# the one-lane PUSH row-split kernels: GPU suite, K = 1..6 sweep times with and without them, race check, Newton times
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02p_gputests.log
tail -3 gpurun_out/r02p_gputests.log
timeout 200 python tools/kernel_times.py --lanes 1 4 6 > gpurun_out/r02p_kt_push.jsonl 2>&1
HANK_NO_RS_PUSH=1 timeout 200 python tools/kernel_times.py --lanes 1 4 6 > gpurun_out/r02p_kt_nopush.jsonl 2>&1
cut -c1-400 gpurun_out/r02p_kt_push.jsonl gpurun_out/r02p_kt_nopush.jsonl
timeout 400 python tools/jacobian_repeat_check.py 150 > gpurun_out/r02p_repeat.log 2>&1
tail -3 gpurun_out/r02p_repeat.log
timeout 300 python tools/newton_time.py > gpurun_out/r02p_newton.log 2>&1
cat gpurun_out/r02p_newton.log | cut -c1-300
timeout 200 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 > gpurun_out/r02p_c5_k1.jsonl 2>&1
cut -c1-300 gpurun_out/r02p_c5_k1.jsonl
