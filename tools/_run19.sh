set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_t19.log
tail -3 gpurun_out/r02_t19.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-newton > gpurun_out/r02_bench19.json 2> gpurun_out/r02_bench19.err
HANK_NO_RING_NE=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-newton > gpurun_out/r02_bench19b.json 2> gpurun_out/r02_bench19b.err
python tools/kernel_times.py --lanes 148 296 592 > gpurun_out/r02_kt19.jsonl 2>&1
python -c "
import json
for f in ('gpurun_out/r02_bench19.json','gpurun_out/r02_bench19b.json'):
    d=json.load(open(f)); print(f, round(d['value']), d['roofline']['frac_by_kernel'], d['roofline']['kernel_ms_per_launch'])
print(open('gpurun_out/r02_kt19.jsonl').read())"
