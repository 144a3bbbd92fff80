set -x
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_sweeps.py -x -q -k "row_split or block_matches or horizons" 2>&1 | tail -15 > gpurun_out/r02_t6.log
tail -3 gpurun_out/r02_t6.log
rm -f gpurun_out/r02_rs6.jsonl
timeout 300 python tools/sweep_times.py --shape 500 7 300 --lanes 1 14 64 74 --tag v5 >> gpurun_out/r02_rs6.jsonl 2>> gpurun_out/r02_rs6.err
timeout 300 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 64 --tag v5 >> gpurun_out/r02_rs6.jsonl 2>> gpurun_out/r02_rs6.err
timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 1 64 --tag v5 >> gpurun_out/r02_rs6.jsonl 2>> gpurun_out/r02_rs6.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_rs6.jsonl'):
    d=json.loads(l); print(d['shape'],d['K'],d['tag'],d['us_per_period'],d['frac_of_measured_hbm'])
PY
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench6.json')); print(d['value'], d['jacobian_build'], d['newton'])"
