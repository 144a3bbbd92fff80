set -x
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_sweeps.py tests/test_gpu_ks.py -x -q 2>&1 | tail -15 > gpurun_out/r02_t5.log
tail -5 gpurun_out/r02_t5.log
rm -f gpurun_out/r02_rs5.jsonl
timeout 300 python tools/sweep_times.py --shape 500 7 300 --lanes 1 8 14 32 64 74 --tag v4 >> gpurun_out/r02_rs5.jsonl 2>> gpurun_out/r02_rs5.err
timeout 300 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 14 64 --tag v4 >> gpurun_out/r02_rs5.jsonl 2>> gpurun_out/r02_rs5.err
timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 1 14 64 --tag v4 >> gpurun_out/r02_rs5.jsonl 2>> gpurun_out/r02_rs5.err
HANK_RS_NO_MULTI=1 timeout 300 python tools/sweep_times.py --shape 500 7 300 --lanes 32 64 75 --tag v3nomulti >> gpurun_out/r02_rs5.jsonl 2>> gpurun_out/r02_rs5.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_rs5.jsonl'):
    d=json.loads(l); print(d['shape'],d['K'],d['tag'],d['us_per_period'],d['frac_of_measured_hbm'])
PY
