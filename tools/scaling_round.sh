#!/bin/bash
# 1/2/4/8-GPU weak scaling of bench.py and strong scaling of the full Jacobian (one box, run under gpurun --gpus 8)
set -x
: > gpurun_out/weak_scaling.jsonl
python bench.py --gpus 1 --no-newton --no-cpu 2>gpurun_out/ws1.err | tail -1 >> gpurun_out/weak_scaling.jsonl
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --no-newton --no-cpu 2>gpurun_out/ws$n.err | tail -1 >> gpurun_out/weak_scaling.jsonl
done
cat gpurun_out/weak_scaling.jsonl | cut -c1-200
: > gpurun_out/jac_scaling.jsonl
python tools/jacobian_scaling.py 2>gpurun_out/js1.err | tail -1 >> gpurun_out/jac_scaling.jsonl
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) tools/jacobian_scaling.py 2>gpurun_out/js$n.err | tail -1 >> gpurun_out/jac_scaling.jsonl
done
cat gpurun_out/jac_scaling.jsonl
