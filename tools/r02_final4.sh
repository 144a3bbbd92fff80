#!/bin/bash
# three-mode pipe hint: tests, smoke, bench default + K=592 A/B, Newton, shard times, race check
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02i_gputests.log
tail -2 gpurun_out/r02i_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err
timeout 300 python bench.py --lanes 592 --no-newton --no-cpu > gpurun_out/r02i_bench_k592.json 2> gpurun_out/r02i_bench_k592.err
HANK_NO_PIPE=1 timeout 300 python bench.py --lanes 592 --no-newton --no-cpu > gpurun_out/r02i_bench_k592_nopipe.json 2> gpurun_out/r02i_bench_k592_nopipe.err
timeout 300 python tools/newton_time.py > gpurun_out/r02i_newton.log 2>&1
cut -c1-250 gpurun_out/r02i_newton.log
timeout 300 python tools/jacobian_repeat_check.py 100 > gpurun_out/r02i_repeat.log 2>&1
tail -2 gpurun_out/r02i_repeat.log
python tools/jacobian_shard_times.py 1 2 4 8 > gpurun_out/r02i_shards.log 2>&1
cat gpurun_out/r02i_shards.log | cut -c1-250
python - <<'PY'
import json
for f in ("n1", "k592", "k592_nopipe"):
    try:
        d = json.loads(open(f"gpurun_out/r02i_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"]["frac_by_kernel"], d.get("jacobian_build", {}).get("ms"),
              (d.get("newton") or {}).get("ms_per_solve"), ((d.get("newton") or {}).get("batched_jacobian_mode") or {}).get("ms_per_solve"))
    except Exception as e:
        print(f, "failed", e)
PY
