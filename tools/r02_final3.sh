#!/bin/bash
# after restricting the fused primal launch to multi-wave passes: tests, bench, Newton, shard times, race check, launch list
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02g_gputests.log
tail -2 gpurun_out/r02g_gputests.log
timeout 900 python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err
timeout 300 python tools/newton_time.py > gpurun_out/r02g_newton.log 2>&1
cut -c1-250 gpurun_out/r02g_newton.log
timeout 300 python tools/jacobian_repeat_check.py 100 > gpurun_out/r02g_repeat.log 2>&1
tail -2 gpurun_out/r02g_repeat.log
python tools/jacobian_shard_times.py 1 2 4 8 > gpurun_out/r02g_shards.log 2>&1
cat gpurun_out/r02g_shards.log | cut -c1-250
timeout 300 python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02g_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-newton --no-cpu > gpurun_out/r02g_ncu_lc.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02g_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"]["frac_by_kernel"], d.get("jacobian_build", {}).get("ms"),
      (d.get("newton") or {}).get("ms_per_solve"), ((d.get("newton") or {}).get("batched_jacobian_mode") or {}).get("ms_per_solve"), d.get("jvp_regimes"), d.get("cpu_baseline"))
PY
