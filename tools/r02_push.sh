#!/bin/bash
# the one-lane st.async / bulk-push row-split kernels: GPU suite, K = 1..6 sweep times for each variant, race check, Newton times
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02p_gputests.log
tail -3 gpurun_out/r02p_gputests.log
timeout 200 python tools/kernel_times.py --lanes 1 4 6 18 > gpurun_out/r02p_kt_st.jsonl 2>&1
HANK_NO_RS_ST=1 timeout 200 python tools/kernel_times.py --lanes 1 4 6 18 > gpurun_out/r02p_kt_push.jsonl 2>&1
cut -c1-400 gpurun_out/r02p_kt_st.jsonl gpurun_out/r02p_kt_push.jsonl
timeout 400 python tools/jacobian_repeat_check.py 150 > gpurun_out/r02p_repeat.log 2>&1
tail -3 gpurun_out/r02p_repeat.log
timeout 300 python tools/newton_time.py > gpurun_out/r02p_newton.log 2>&1
cat gpurun_out/r02p_newton.log | cut -c1-300
