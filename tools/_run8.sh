set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench8_n1.json 2> gpurun_out/r02_bench8_n1.err
python bench.py --steps 10 --warmup 3 --workload ks_1000x7_T300 --lanes 64 > gpurun_out/r02_bench8_c5.json 2> gpurun_out/r02_bench8_c5.err
python bench.py --steps 5 --warmup 3 --workload ks_2000x11_T500 --lanes 64 > gpurun_out/r02_bench8_c4.json 2> gpurun_out/r02_bench8_c4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench8_n2.json 2> gpurun_out/r02_bench8_n2.err
tail -3 gpurun_out/r02_bench8_*.err
python - <<'PY'
import json
for f in ("n1","c5","c4","n2"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench8_{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"ERR",e); continue
    print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "frac", {k:round(v,3) for k,v in d["roofline"]["frac_by_kernel"].items()}, d["roofline"]["us_per_period"])
    for k in ("newton","jacobian_build","jacobian_build_strong_scaling","cpu_baseline"):
        if k in d: print("   ",k,json.dumps(d[k])[:700])
PY
