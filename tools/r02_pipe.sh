#!/bin/bash
# pipelined linearisation (backward primal next to the backward tangent): GPU suite, race check, bench / Newton A/B
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02q_gputests.log
tail -3 gpurun_out/r02q_gputests.log
timeout 400 python tools/jacobian_repeat_check.py 100 > gpurun_out/r02q_repeat.log 2>&1
tail -3 gpurun_out/r02q_repeat.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-newton --no-cpu > gpurun_out/r02q_bench_pipe.json 2> gpurun_out/r02q_bench_pipe.err
HANK_NO_PIPE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-newton --no-cpu > gpurun_out/r02q_bench_nopipe.json 2> gpurun_out/r02q_bench_nopipe.err
python - <<'PY'
import json
for f in ("pipe", "nopipe"):
    try:
        d = json.loads(open(f"gpurun_out/r02q_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"].get("frac_by_kernel"), d.get("jacobian_build", {}).get("ms"), d.get("parity_max_err_over_tol"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/r02q_bench_{f}.err").read()[-800:])
PY
timeout 300 python tools/newton_time.py > gpurun_out/r02q_newton.log 2>&1
cut -c1-300 gpurun_out/r02q_newton.log
