#!/bin/bash
# 1/2/4/8-GPU bench lines (weak-scaled JVP batch + strong-scaled full Jacobian build with NCCL all-gather)
set -x
for N in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r02_scale_n{N}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(N,"ERR",e); continue
    print(N, "JVP/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "strong jac ms", d.get("jacobian_build_strong_scaling",{}).get("ms"), "1gpu jac ms", d.get("jacobian_build",{}).get("ms"))
PY
