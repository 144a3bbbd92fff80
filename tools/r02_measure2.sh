#!/bin/bash
# tests after the primal de-triplication + ncu captures of the row-split kernels on the large grids
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02b_gputests.log
tail -3 gpurun_out/r02b_gputests.log
timeout 300 python tools/kernel_times.py --lanes 1 592 > gpurun_out/r02b_kernel_times.jsonl 2>&1
cat gpurun_out/r02b_kernel_times.jsonl | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(backward|forward)_tangent' --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02b_prof_c4_k64 python tools/sweep_times.py --shape 2000 11 500 --lanes 64 --reps 1 > gpurun_out/r02b_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(backward|forward)_tangent' --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02b_prof_c5_k64 python tools/sweep_times.py --shape 1000 7 300 --lanes 64 --reps 1 > gpurun_out/r02b_ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(backward|forward)_tangent' --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02b_prof_c5_k444 python tools/sweep_times.py --shape 1000 7 300 --lanes 444 --reps 1 > gpurun_out/r02b_ncu_c5b.log 2>&1
ls -la gpurun_out/
