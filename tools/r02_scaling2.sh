#!/bin/bash
# 8/4/2/1-GPU bench lines at the final commit + the multi-rank NCCL Jacobian test
set -x
for N in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r02s_scale_n$N.json 2> gpurun_out/r02s_scale_n$N.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02s_scale_n1.json 2> gpurun_out/r02s_scale_n1.err
timeout 300 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -x -q -k multi_gpu 2>&1 | tail -2
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r02s_scale_n{N}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(N,"ERR",e); continue
    print(N, "JVP/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "strong jac ms", d.get("jacobian_build_strong_scaling",{}).get("ms"), "1gpu jac ms", d.get("jacobian_build",{}).get("ms"))
PY
