set -x
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_ks.py -x -q -k "full_size or horizons or golden_ks" 2>&1 | tail -4 > gpurun_out/r02_t12.log
tail -3 gpurun_out/r02_t12.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err
HANK_NEWTON_TRACE=1 python tools/newton_time.py > gpurun_out/r02_newton12.log 2>&1
tail -5 gpurun_out/r02_newton12.log
python -c "
import json; d=json.load(open('gpurun_out/r02_bench12.json')); print(d['value'], d['jacobian_build']['ms'], d['newton']['ms_per_solve'], d['newton']['batched_jacobian_mode']['ms_per_solve'])"
