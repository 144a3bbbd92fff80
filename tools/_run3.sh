set -x
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_sweeps.py -x -q -k "row_split or fallback or padding or minimal or block_matches" 2>&1 | tail -15 > gpurun_out/r02_t3.log
tail -5 gpurun_out/r02_t3.log
rm -f gpurun_out/r02_rs3.jsonl
timeout 300 python tools/sweep_times.py --shape 500 7 300 --lanes 1 8 16 18 --tag v2 >> gpurun_out/r02_rs3.jsonl 2>> gpurun_out/r02_rs3.err
timeout 300 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 16 64 --tag v2 >> gpurun_out/r02_rs3.jsonl 2>> gpurun_out/r02_rs3.err
timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 1 16 64 --tag v2 >> gpurun_out/r02_rs3.jsonl 2>> gpurun_out/r02_rs3.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_rs3.jsonl'):
    d=json.loads(l); print(d['shape'],d['K'],d['us_per_period'],d['frac_of_measured_hbm'])
PY
