#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): GPU tests, bench, per-config sweep times, ncu launch list + full capture.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_gputests.log
tail -3 gpurun_out/r02_gputests.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 1500 gpurun_out/r02_bench_n1.json
timeout 300 python tools/kernel_times.py --lanes 1 4 64 148 592 1156 > gpurun_out/r02_kernel_times.jsonl 2>&1
timeout 300 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 64 444 > gpurun_out/r02_sweeps_c5.jsonl 2>&1
timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 1 2 64 > gpurun_out/r02_sweeps_c4.jsonl 2>&1
timeout 300 python bench.py --workload ks_1000x7_T300 --lanes 64 --steps 10 --no-cpu > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
timeout 300 python bench.py --workload ks_2000x11_T500 --lanes 64 --steps 10 --no-cpu > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
cat gpurun_out/r02_kernel_times.jsonl gpurun_out/r02_sweeps_c5.jsonl gpurun_out/r02_sweeps_c4.jsonl | cut -c1-420
timeout 300 python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02_ncu_lc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_(backward|forward)_tangent' --launch-skip 6 --launch-count 2 -f -o gpurun_out/r02_prof_sweeps python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02_ncu_fc.log 2>&1
ls -la gpurun_out/
