set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/kernel_times.py --lanes 592 888 1156 > gpurun_out/kt_r01c.jsonl 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1_r01c.json 2> gpurun_out/bench_n1_r01c.err
tail -c 3000 gpurun_out/bench_n1_r01c.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/plain_r01c.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/ncu_lc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_(backward|forward)_(tangent|primal)' --launch-skip 12 --launch-count 4 -f -o gpurun_out/prof_sweeps_r01c python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/ncu_fc.log 2>&1
ls -la gpurun_out/prof_sweeps_r01c.ncu-rep
