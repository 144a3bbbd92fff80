: > gpurun_out/jac_scaling.jsonl
python tools/jacobian_scaling.py 2>gpurun_out/js1.err | tail -1 >> gpurun_out/jac_scaling.jsonl
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) tools/jacobian_scaling.py 2>gpurun_out/js$n.err | tail -1 >> gpurun_out/jac_scaling.jsonl
done
cat gpurun_out/jac_scaling.jsonl
