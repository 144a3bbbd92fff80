#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -x -q -k "not between_the_compiled and not kronecker" 2>&1 | tail -4 > gpurun_out/r02h_gputests.log
cat gpurun_out/r02h_gputests.log
HANK_NEWTON_TRACE=1 HANK_NO_JBAR_CACHE=1 timeout 300 python tools/newton_time.py > gpurun_out/r02h_newton_gj.log 2>&1
HANK_CUSOLVER=1 HANK_NEWTON_TRACE=1 HANK_NO_JBAR_CACHE=1 timeout 300 python tools/newton_time.py > gpurun_out/r02h_newton_cusolver.log 2>&1
grep -h "setup" gpurun_out/r02h_newton_gj.log gpurun_out/r02h_newton_cusolver.log
HANK_DENSE_TIME=1 python tools/dense_time.py 1196 596 1996 > gpurun_out/r02h_dense.log 2>&1
grep -h "hank_dense" gpurun_out/r02h_dense.log | sort | uniq | head
timeout 900 python bench.py > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
tail -c 600 gpurun_out/r02h_bench_n1.json
