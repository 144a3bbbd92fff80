#!/bin/bash
# Final round-2 pass on one B200: GPU tests, bench lines (default, C5, C4), large-grid sweeps, Newton trace, ncu launch list
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02z_gputests.log
tail -2 gpurun_out/r02z_gputests.log
timeout 900 python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err
timeout 300 python bench.py --workload ks_1000x7_T300 --lanes 64 --steps 10 --no-cpu > gpurun_out/r02z_bench_c5.json 2> gpurun_out/r02z_bench_c5.err
timeout 300 python bench.py --workload ks_2000x11_T500 --lanes 64 --steps 10 --no-cpu > gpurun_out/r02z_bench_c4.json 2> gpurun_out/r02z_bench_c4.err
timeout 300 python tools/kernel_times.py --lanes 1 4 64 148 592 1156 > gpurun_out/r02z_kernel_times.jsonl 2>&1
: > gpurun_out/r02z_sweeps_large.jsonl
timeout 300 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 64 444 >> gpurun_out/r02z_sweeps_large.jsonl 2>&1
timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 1 64 148 592 >> gpurun_out/r02z_sweeps_large.jsonl 2>&1
HANK_RS_NO_MULTI=1 timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 148 --tag one_cta_fallback --reps 1 >> gpurun_out/r02z_sweeps_large.jsonl 2>&1
cut -c1-330 gpurun_out/r02z_sweeps_large.jsonl
HANK_NEWTON_TRACE=1 HANK_NO_JBAR_CACHE=1 timeout 300 python tools/newton_time.py > gpurun_out/r02z_newton.log 2>&1
grep setup gpurun_out/r02z_newton.log
python tools/jacobian_shard_times.py 1 2 4 8 > gpurun_out/r02z_shards.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02z_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02z_ncu_lc.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/r02z_bench_n1.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['frac_by_kernel'], d['jacobian_build']['ms'], d['newton']['ms_per_solve'], d['newton']['batched_jacobian_mode']['ms_per_solve'])"
