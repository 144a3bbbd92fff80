set -x
timeout 600 python -m pytest tests/test_gpu_edges.py -x -q -k "row_split or fallback or padding or minimal" 2>&1 | tail -15 > gpurun_out/r02_t1.log
tail -5 gpurun_out/r02_t1.log
for env in "" "HANK_NO_ROWSPLIT=1"; do
  env $env timeout 300 python tools/sweep_times.py --shape 500 7 300 --lanes 1 8 18 --tag "$env" >> gpurun_out/r02_rs1.jsonl 2>> gpurun_out/r02_rs1.err
  env $env timeout 300 python tools/sweep_times.py --shape 1000 7 300 --lanes 1 18 64 --tag "$env" >> gpurun_out/r02_rs1.jsonl 2>> gpurun_out/r02_rs1.err
  env $env timeout 300 python tools/sweep_times.py --shape 2000 11 500 --lanes 1 18 64 --tag "$env" >> gpurun_out/r02_rs1.jsonl 2>> gpurun_out/r02_rs1.err
done
cat gpurun_out/r02_rs1.jsonl | cut -c1-400
