set -x
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench18_n$N.json 2> gpurun_out/r02_bench18_n$N.err
done
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench18_n1.json 2> gpurun_out/r02_bench18_n1.err
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench18_n{N}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(N,"ERR",e); continue
    print(N, round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("jacobian_build_strong_scaling",{}).get("ms"), d.get("jacobian_build",{}).get("ms"))
PY
