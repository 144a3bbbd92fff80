set -x
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_t7.log
tail -6 gpurun_out/r02_t7.log
timeout 300 python tools/sweep_times.py --shape 500 7 300 --lanes 1 14 74 --tag v6 > gpurun_out/r02_rs7.jsonl 2> gpurun_out/r02_rs7.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_rs7.jsonl'):
    d=json.loads(l); print(d['shape'],d['K'],d['tag'],d['us_per_period'],d['frac_of_measured_hbm'])
PY
