set -x
timeout 1200 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_edges.py tests/test_gpu_baseline_shapes.py -x -q 2>&1 | tail -4 > gpurun_out/r02_t20.log
tail -3 gpurun_out/r02_t20.log
python tools/kernel_times.py --lanes 148 296 592 1156 > gpurun_out/r02_kt20.jsonl 2>&1
cat gpurun_out/r02_kt20.jsonl | cut -c1-330
