set -x
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_baseline_shapes.py -x -q -k "row_split or c4 or c5 or c2 or horizons" 2>&1 | tail -4 > gpurun_out/r02_t11.log
tail -3 gpurun_out/r02_t11.log
rm -f gpurun_out/r02_rs11.jsonl
timeout 300 python tools/sweep_times.py --shape 500 7 300 --lanes 1 14 74 --tag v8 >> gpurun_out/r02_rs11.jsonl 2>> gpurun_out/r02_rs11.err
timeout 300 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 64 --tag v8 >> gpurun_out/r02_rs11.jsonl 2>> gpurun_out/r02_rs11.err
timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 1 64 --tag v8 >> gpurun_out/r02_rs11.jsonl 2>> gpurun_out/r02_rs11.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_rs11.jsonl'):
    d=json.loads(l); print(d['shape'],d['K'],d['tag'],d['us_per_period'],d['frac_of_measured_hbm'])
PY
