#!/bin/bash
# Final round-2 pass after the st.async one-lane kernels and the pipelined linearisation: GPU tests, bench lines
# (default, C5, C4), kernel times per lane count, Newton times, race check, ncu launch list + full capture
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02f_gputests.log
tail -2 gpurun_out/r02f_gputests.log
timeout 900 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
HANK_NO_PIPE=1 timeout 300 python bench.py --no-newton --no-cpu > gpurun_out/r02f_bench_n1_nopipe.json 2> gpurun_out/r02f_bench_n1_nopipe.err
timeout 300 python bench.py --workload ks_1000x7_T300 --lanes 64 --steps 10 --no-cpu > gpurun_out/r02f_bench_c5.json 2> gpurun_out/r02f_bench_c5.err
timeout 300 python bench.py --workload ks_2000x11_T500 --lanes 64 --steps 10 --no-cpu > gpurun_out/r02f_bench_c4.json 2> gpurun_out/r02f_bench_c4.err
timeout 300 python tools/kernel_times.py --lanes 1 4 64 148 592 1156 > gpurun_out/r02f_kernel_times.jsonl 2>&1
timeout 300 python tools/newton_time.py > gpurun_out/r02f_newton.log 2>&1
cut -c1-250 gpurun_out/r02f_newton.log
timeout 300 python tools/jacobian_repeat_check.py 100 > gpurun_out/r02f_repeat.log 2>&1
tail -2 gpurun_out/r02f_repeat.log
python tools/jacobian_shard_times.py 1 2 4 8 > gpurun_out/r02f_shards.log 2>&1
cat gpurun_out/r02f_shards.log | cut -c1-250
timeout 300 python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02f_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02f_ncu_lc.log 2>&1
tail -2 gpurun_out/r02f_ncu_lc.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(backward|forward)_tangent_rs_st' --launch-skip 4 --launch-count 2 -f -o gpurun_out/r02f_prof_st python tools/kernel_times.py --lanes 1 > gpurun_out/r02f_ncu_st.log 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json
for f in ("n1", "n1_nopipe", "c5", "c4"):
    try:
        d = json.loads(open(f"gpurun_out/r02f_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"]["frac_by_kernel"], d.get("jacobian_build", {}).get("ms"),
              (d.get("newton") or {}).get("ms_per_solve"), ((d.get("newton") or {}).get("batched_jacobian_mode") or {}).get("ms_per_solve"), d.get("jvp_regimes"))
    except Exception as e:
        print(f, "failed", e)
PY
