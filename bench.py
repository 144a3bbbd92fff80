#!/usr/bin/env python
"""bench.py — headline benchmark of the hankb200 household block.

Workload (BASELINE.json configs[1]/[2]): Krusell-Smith, T=300 (P=299), 500 assets x 7 income
states, model-consistent inputs (steady-state record from tests/golden/ss_500x7_T300.npz, shock
Z_t = 1 + 0.8^t, RunMain.jl:50-51).  One STEP = one pass of the hot path over one batch:
linearise F at x (primal EGM backward sweep + primal lottery forward sweep + residuals), then K
dense JVP directions as batched tangent lanes (backward tangent sweep, forward tangent sweep,
aggregation, residual tangents).  metric = JVPs per second (direction x full sweep pair).
With --gpus N (one process per GPU, torchrun) every rank carries K lanes (weak scaling) and the
n x K column blocks are all-gathered with NCCL (hank_allgather_columns_dev).

Prints ONE JSON line (rank 0).  `value` has inputs resident in HBM; `e2e` goes through the
host-pointer C ABI (hank_ks_fjvp = linearise + K-lane JVP) with pinned host buffers, copies inside the
timed region.  `--impl reference` times the CPU oracle (single-thread C++ restatement of the
reference's Julia path; Julia is not installed in this image) on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "julia-newtonraphsonhank_b200"))

WORKLOADS = {
    "ks_500x7_T300": dict(fixture="ss_500x7_T300.npz", desc="Krusell-Smith T=300, 500 assets x 7 income states (BASELINE configs 2/3)"),
    "ks_1000x7_T300": dict(fixture="ss_1000x7_T300.npz", desc="Krusell-Smith T=300, 1000 assets x 7 income states (BASELINE config 5: --lanes 64)"),
    # no steady-state record at this size: synthetic-throughput inputs of SURVEY.md 8d-ii through hank_block (r, w paths + K lanes)
    "ks_2000x11_T500": dict(shape=(2000, 11, 500), desc="2000 assets x 11 income states, T=500, synthetic-throughput inputs (BASELINE config 4: --lanes 64)"),
}
METRIC, UNIT = "jvps_per_sec", "JVP/s"


def synthetic_inputs(n_a, n_e, T, rbar=0.015, wbar=1.35, c0=0.1):
    """Synthetic-throughput regime: closed-form monotone terminal value, uniform D0, smooth paths."""
    from hankb200 import model as M
    grid = M.double_exponential(n_a, 0.0, 200.0)
    z, Pi = M.rouwenhorst_discretization(n_e, 0.966, 0.283)
    P = T - 1
    vT = (1 + rbar) * ((rbar * grid[None, :] + wbar * z[:, None]) + c0) ** (-2.0)
    t = np.arange(1, P + 1)
    return dict(grid=grid, z=z, Pi=Pi, beta=0.98, gamma=2.0, borrow_cons=0.0, T=T, P=P, n_a=n_a, n_e=n_e, vT=vT,
                D0=np.full((n_e, n_a), 1.0 / (n_a * n_e)), r=rbar * (1 + 0.1 * 0.9 ** t), w=wbar * (1 + 0.05 * 0.9 ** t))


def load_fixture(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", WORKLOADS[name]["fixture"]))
    T = int(g["T"]); P = T - 1
    vars_ = g["ss_vars"]
    return dict(g=g, T=T, P=P, n=4 * P, n_a=int(g["n_a"]), n_e=int(g["n_e"]),
                x0=np.tile(vars_[:4], P), Z=1.0 + 0.8 ** np.arange(1, P + 1),
                ks=(float(g["alpha"]), float(g["delta"]), float(vars_[1])))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        # a sample is stamped when its line is read, i.e. up to one query (~20-50 ms) after the clocks were
        # taken; a timed region shorter than that falls back to the samples that bracket it
        inside = [r for r in self.rows if t0 <= r[0] <= t1 + 0.05]
        window = "timed region"
        if not inside:
            inside = [r for r in self.rows if t0 - 0.3 <= r[0] <= t1 + 0.3]
            window = "timed region +-0.3 s (region shorter than one nvidia-smi query)"
        for ts, line in inside:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "window": window}


_W = {}


def is_synth(workload):
    return "shape" in WORKLOADS[workload]


def workload_inputs(workload):
    """Fixture-based (Krusell-Smith full function) or synthetic (household block only) inputs of a workload."""
    if is_synth(workload):
        sy = synthetic_inputs(*WORKLOADS[workload]["shape"])
        return dict(sy=sy, T=sy["T"], P=sy["P"], n=4 * sy["P"], n_a=sy["n_a"], n_e=sy["n_e"])
    return load_fixture(workload)


def _ref_worker_init(workload):
    from oracle import oracle as O
    fx = workload_inputs(workload)
    _W["fx"] = fx
    if is_synth(workload):
        sy = fx["sy"]
        _W["orc"] = O.Oracle(sy["grid"], sy["z"], sy["Pi"], sy["beta"], sy["gamma"], sy["borrow_cons"], sy["T"])
    else:
        g = fx["g"]
        _W["orc"] = O.Oracle(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])


def _ref_worker(V):
    fx = _W["fx"]
    if "sy" in fx:   # household block only: V = (dr, dw)
        sy = fx["sy"]
        pol, dpol, _, _ = _W["orc"].backward(sy["vT"], sy["r"], sy["w"], V[0], V[1])
        KD, dKD = _W["orc"].forward(sy["D0"], pol, dpol)
        return float(dKD.sum())
    g = fx["g"]
    F, JV = _W["orc"].ks_fjvp(fx["ks"], g["ss_value"], g["ss_D"], fx["Z"], fx["x0"], V)
    return float(JV.sum())


def make_config(args, world, fx):
    """The workload description both arms print (the reference arm times a bounded sample of it)."""
    alg = 8.0 * fx["n_a"] * fx["n_e"] * fx["P"] * args.lanes
    step = ("backward + forward sweep of the household block with K tangent lanes (hank_block)" if is_synth(args.workload) else
            "linearise F(x) (primal sweeps) + K-lane JVP (tangent sweeps) + residual tangents")
    return {"workload": args.workload, "desc": WORKLOADS[args.workload]["desc"], "lanes_per_gpu_per_step": args.lanes,
            "step": step + (" + NCCL all-gather of n x K columns" if world > 1 else ""),
            "l2": "inputs larger than L2 (policy-tangent stream %.2f GB per step vs 126 MB L2)" % (alg / 1e9)
                  if alg > 2.5e8 else "L2 flushed between timed steps (256 MB device memset)"}


def run_reference(args):
    """CPU arm: the oracle (kind "port") on the same workload, bounded sample per step.  The reference has no
    threading of its own; independent JVP directions are farmed out to one process per host core (what a user
    would do with Distributed.pmap), each process recomputing the primal like each GPU rank does."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    fx = workload_inputs(args.workload)
    procs = max(1, args.cpu_procs or (os.cpu_count() or 1))
    Kc = args.cpu_lanes                      # lanes per process per step
    if is_synth(args.workload):
        Kc = max(1, Kc // 8)                 # ~10x the points per lane of the 500 x 7 case
    rng = np.random.default_rng(1234)
    if is_synth(args.workload):
        Vs = [(rng.standard_normal((Kc, fx["P"])), rng.standard_normal((Kc, fx["P"]))) for _ in range(procs)]
    else:
        Vs = [rng.standard_normal((Kc, fx["n"])) for _ in range(procs)]
    with mp.get_context("fork").Pool(procs, initializer=_ref_worker_init, initargs=(args.workload,)) as pool:
        def step():
            pool.map(_ref_worker, Vs, chunksize=1)
        for _ in range(min(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    val = procs * Kc * args.steps / dt
    sample = (f"{procs} processes x {Kc} lanes per step (each: primal + {Kc} tangent lanes) of the GPU arm's "
              f"{args.lanes} lanes, {args.steps} steps")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, args.gpus, fx),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
                         "sample_lanes_per_step": procs * Kc,
                         "note": "C++ restatement of the reference CPU path (oracle/); Julia is not installed. The "
                                 "reference is single-threaded: lanes are spread over one process per host core"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hankb200", choices=["hankb200", "reference"])
    ap.add_argument("--workload", default="ks_500x7_T300", choices=sorted(WORKLOADS))
    ap.add_argument("--lanes", type=int, default=1156, help="tangent lanes per GPU per step: 4 lanes x (2 x 148 - 7) CTAs = two waves with the 7-CTA forward-primal cluster overlapped in the first; the full T=300 Jacobian has 1196 columns")
    ap.add_argument("--cpu-lanes", type=int, default=32, help="lanes per step (per process in the reference arm) of the CPU sample")
    ap.add_argument("--cpu-procs", type=int, default=0, help="processes of the reference arm (0 = one per host core)")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from hankb200 import HouseholdBlock

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hankb200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    synth = is_synth(args.workload)
    fx = workload_inputs(args.workload)
    P, n, K = fx["P"], fx["n"], args.lanes
    if synth:
        sy = fx["sy"]
        blk = HouseholdBlock(sy["grid"], sy["z"], sy["Pi"], sy["beta"], sy["gamma"], sy["borrow_cons"], sy["T"], device=local)
        blk.set_terminal(sy["vT"]); blk.set_initial_dist(sy["D0"])
    else:
        g = fx["g"]
        blk = HouseholdBlock(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]),
                             fx["T"], device=local)
        blk.set_terminal(g["ss_value"]); blk.set_initial_dist(g["ss_D"])
        blk.ks_configure(*fx["ks"])
    L = blk._L; h = blk.handle
    blk.reserve_lanes(max(K, n // 2) if (world == 1 and not synth) else K)   # n/2 household lanes for the Jacobian build
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt = torch.tensor(list(HouseholdBlock.comm_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(idt, 0)
        blk.comm_init(world, rank, bytes(idt.cpu().numpy().tolist()))

    # ---- device-resident inputs (value) and pinned host inputs (e2e) ----------------------------
    rng = np.random.default_rng(1234 + rank)
    vp = lambda t: C.c_void_p(t.data_ptr())
    dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    if synth:
        rh, wh = pin(sy["r"]), pin(sy["w"])
        drh, dwh = pin(rng.standard_normal((K, P))), pin(rng.standard_normal((K, P)))
        KDh = torch.empty(P, dtype=torch.float64).pin_memory(); JVh = torch.empty((K, P), dtype=torch.float64).pin_memory()
        rd, wd, drd, dwd = rh.to(dev), wh.to(dev), drh.to(dev), dwh.to(dev)
        KDd = torch.empty(P, dtype=torch.float64, device=dev); JVd = torch.empty((K, P), dtype=torch.float64, device=dev)
        out_cols = P   # doubles per lane in the gathered block
        h2d, d2h = 8 * (2 * P + 2 * K * P), 8 * (P + K * P)
    else:
        Vh = pin(rng.standard_normal((K, n)))
        xh = pin(fx["x0"].copy()); Zh = pin(fx["Z"].copy())
        Fh = torch.empty(n, dtype=torch.float64).pin_memory(); JVh = torch.empty((K, n), dtype=torch.float64).pin_memory()
        Vd = Vh.to(dev); xd = xh.to(dev); Zd = Zh.to(dev)
        Fd = torch.empty(n, dtype=torch.float64, device=dev); JVd = torch.empty((K, n), dtype=torch.float64, device=dev)
        out_cols = n
        h2d, d2h = 8 * (n * K + n + P), 8 * (n * K + n)
    ALLd = torch.empty((world * K, out_cols), dtype=torch.float64, device=dev) if world > 1 else None
    alg = 8.0 * fx["n_a"] * fx["n_e"] * P * K
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if alg <= 2.5e8 else None   # working set below 2 x L2: flush
    torch.cuda.synchronize()

    def step_dev():
        if synth:
            blk._ck(L.hank_block_dev(h, vp(rd), vp(wd), K, vp(drd), vp(dwd), vp(KDd), vp(JVd)))
        else:
            blk._ck(L.hank_ks_linearize_dev(h, vp(xd), vp(Zd), vp(Fd)))
            blk._ck(L.hank_ks_jvp_dev(h, K, vp(Vd), vp(JVd)))
        if world > 1:
            blk._ck(L.hank_allgather_columns_dev(h, vp(JVd), K * out_cols, vp(ALLd)))

    def step_e2e():
        if synth:
            blk._ck(L.hank_block(h, dp(rh), dp(wh), K, dp(drh), dp(dwh), dp(KDh), dp(JVh)))
        else:
            blk._ck(L.hank_ks_fjvp(h, dp(xh), dp(Zh), K, dp(Vh), dp(Fh), dp(JVh)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        blk.sync()

    def timed(fn, steps):
        """CUDA-event time of `steps` calls on the context's stream, barrier + synchronize on both sides, max over ranks.
        With the L2 flush each step is timed on its own (the flush is outside the events)."""
        barrier()
        t0 = time.time()
        if flush is None:
            blk.timer_start()
            for _ in range(steps):
                fn()
            ms = blk.timer_stop()
        else:
            ms = 0.0
            for _ in range(steps):
                flush.zero_(); torch.cuda.synchronize()
                blk.timer_start(); fn(); ms += blk.timer_stop()
        blk.sync()
        t1 = time.time()
        barrier()
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, t0, t1

    sampler = ClockSampler(local) if rank == 0 else None   # started before the warm-up so it is already streaming
    for _ in range(max(args.warmup, 3)):
        step_dev()
    blk.sync()
    blk.profile(True); blk.kernel_times(reset=True)
    lc0 = blk.launch_count()
    ms, t0, t1 = timed(step_dev, args.steps)
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = blk.launch_count() - lc0
    kt = blk.kernel_times(reset=True)
    blk.profile(False)
    value = world * K * args.steps / (ms * 1e-3)

    # ---- e2e through the host-pointer C ABI ------------------------------------------------------
    for _ in range(2):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    e2e_val = world * K * args.steps / (ms_e2e * 1e-3)

    # ---- N > 1: the north star's split — ONE Jacobian's columns over the ranks (strong scaling) ---
    strong = None
    if world > 1 and not synth:
        from hankb200.sharding import period_round_robin
        cols = period_round_robin(n, world, rank)
        kmax = max(len(period_round_robin(n, world, r)) for r in range(world))
        blk.reserve_lanes(kmax)
        ones_d = torch.ones(P, dtype=torch.float64, device=dev)
        loc = torch.zeros((kmax, n), dtype=torch.float64, device=dev)
        allb = torch.empty((world * kmax, n), dtype=torch.float64, device=dev)
        cptr = cols.ctypes.data_as(C.POINTER(C.c_int))

        def build():
            blk._ck(L.hank_ks_linearize_dev(h, vp(xd), vp(ones_d), vp(Fd)))
            blk._ck(L.hank_ks_jacobian_column_list_dev(h, len(cols), cptr, vp(loc)))
            blk._ck(L.hank_allgather_columns_dev(h, vp(loc), kmax * n, vp(allb)))
        for _ in range(3):
            build()
        nb = max(5, min(args.steps, 20))
        ms_b, _, _ = timed(build, nb)
        strong = {"metric": "ms per full Jacobian build (linearise + %d columns + NCCL all-gather)" % n, "ms": ms_b / nb,
                  "scaling": "strong", "n_gpus": world, "columns": n, "household_lanes_per_rank": int(np.sum((cols - 1) % 4 >= 2)),
                  "partition": "periods dealt round-robin over the ranks (equal mix of seed horizons)",
                  "allgather_bytes_per_rank": int(kmax * n * 8), "steps": nb}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes: 8*G*P*K per tangent sweep) ----------
    G = fx["n_a"] * fx["n_e"]
    peak, peak_src = peaks()
    per = {k: (v[0] / v[1] if v[1] else None) for k, v in kt.items()}
    dom = max(("backward_tangent", "forward_tangent"), key=lambda k: per[k] or 0.0)
    ach = alg / (per[dom] * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            per_lane = json.load(open(prof)).get(args.workload, {}).get(dom + "_bytes_per_lane")
            traffic = per_lane * K if per_lane else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                "kernel_ms_per_launch": per,
                "frac_by_kernel": {k: alg / (per[k] * 1e-3) / 1e9 / peak for k in ("backward_tangent", "forward_tangent")},
                "sweep_pair_GBps": 2 * alg / ((per["backward_tangent"] + per["forward_tangent"]) * 1e-3) / 1e9,
                "us_per_period": {k: 1e3 * v / P for k, v in per.items() if v},
                "note": "event-timed on the launching stream inside the timed steps; with the pipelined linearisation "
                        "(default; HANK_NO_PIPE=1 serialises) backward_primal is the fused launch of both primal sweeps and "
                        "the backward tangent's time contains the late start of the CTAs displaced by the primal cluster"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, world, fx),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if strong:
        out["jacobian_build_strong_scaling"] = strong

    # ---- N = 1: the few-lane regimes of SURVEY 8d (one JVP = direction x full sweep pair, at a fixed linearisation) ---
    if world == 1 and not synth:
        reg = {}
        blk._ck(L.hank_ks_linearize_dev(h, vp(xd), vp(Zd), vp(Fd)))
        for kk in (1, 64):
            if kk > K:
                continue
            for _ in range(3):
                blk._ck(L.hank_ks_jvp_dev(h, kk, vp(Vd), vp(JVd)))
            blk.sync(); blk.timer_start()
            for _ in range(20):
                blk._ck(L.hank_ks_jvp_dev(h, kk, vp(Vd), vp(JVd)))
            t_ms = blk.timer_stop() / 20
            reg["K%d" % kk] = {"ms_per_pass": t_ms, "jvps_per_sec": kk / (t_ms * 1e-3)}
        out["jvp_regimes"] = dict(reg, note="K-lane JVP passes at a fixed linearisation (tape resident), CUDA events, mean of 20: "
                                            "K=1 is Newton's inner JVP, K=64 the batched stress of BASELINE config 5")

    # ---- N = 1: the metric's other halves — Jacobian build and ms per Newton solve, CPU beside them ---
    if world == 1 and not synth and not args.no_newton:
        ones = np.ones(P)
        blk.linearize(fx["x0"], ones)
        Jd = torch.empty((n, n), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        jac_ms = 1e9
        for _ in range(3):
            blk.sync(); blk.timer_start()
            blk._ck(L.hank_ks_jacobian_columns_dev(h, 1, n + 1, vp(Jd)))
            jac_ms = min(jac_ms, blk.timer_stop())
        Jbar = Jd.cpu().numpy().T.copy()   # device buffer is column-major

        def newton_timed(solver, reps):
            """CUDA-event time on the context's stream around hank_newton_solve (host buffers in and out)."""
            best, res = 1e9, None
            for _ in range(reps):
                blk.sync(); blk.timer_start()
                res = blk.newton_solve(Jbar, fx["x0"], fx["Z"], solver=solver)
                best = min(best, blk.timer_stop())
            return best, res
        newton_ms, (x, st) = newton_timed("lu", 2)
        nb_ms, (xb, stb) = newton_timed("lu_batched", 3)
        Fx = blk.linearize(x, fx["Z"])
        out["jacobian_build"] = {"ms": jac_ms, "columns": n, "household_lanes": n // 2,
                                 "note": "full n x n sequence-space Jacobian by unit-seed lanes at a given linearisation (min of 3); "
                                         "Y/KS columns skip the sweeps"}
        out["newton"] = {"metric": "ms per Newton solve (KS T=%d, %d x %d)" % (fx["T"], fx["n_a"], fx["n_e"]),
                         "ms_per_solve": newton_ms, "solver": "lu (exact preconditioner solve), one single-lane JVP sweep pair per inner iteration "
                                                              "as in NewtonRaphson.jl:94-105",
                         "outer": st["outer"], "jvps": st["jvps"], "inner": st["inner"], "jvps_per_sec_k1": st["jvps"] / (newton_ms * 1e-3),
                         "ms_per_inner_iteration": newton_ms / max(st["jvps"], 1),
                         "residual_norm": float(np.linalg.norm(Fx)), "timing": "CUDA events on the context's stream around hank_newton_solve (min of 2)",
                         "batched_jacobian_mode": {"ms_per_solve": nb_ms, "outer": stb["outer"], "inner": stb["inner"],
                                                   "max_abs_diff_vs_sequential": float(np.max(np.abs(xb - x))),
                                                   "note": "J(x) assembled once per outer iteration from unit-seed lanes; inner J(x)y by FP64 GEMV"}}
        if not args.no_cpu:
            from oracle import oracle as O
            O.build()
            orc = O.Oracle(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])
            tw = time.perf_counter()
            xo, so = orc.newton(fx["ks"], g["ss_value"], g["ss_D"], fx["Z"], Jbar, fx["x0"], solver="lu")
            dtn = time.perf_counter() - tw
            out["newton"]["cpu_baseline"] = {"ms_per_solve": 1e3 * dtn, "cores": 1, "kind": "port", "outer": so["outer"], "inner": so["inner"],
                                             "sample": "the whole solve, same Jbar, same x0 and shock, LU preconditioner solve",
                                             "max_abs_diff_gpu_vs_cpu_path": float(np.max(np.abs(x - xo))),
                                             "speedup_sequential": 1e3 * dtn / newton_ms, "speedup_batched": 1e3 * dtn / nb_ms}
    if world == 1 and not args.no_cpu:
        from oracle import oracle as O
        O.build()
        Kc = args.cpu_lanes if not synth else max(1, args.cpu_lanes // 8)
        if synth:
            orc = O.Oracle(sy["grid"], sy["z"], sy["Pi"], sy["beta"], sy["gamma"], sy["borrow_cons"], sy["T"])
            orc.backward(sy["vT"], sy["r"], sy["w"])   # warm-up (primal only)
            tw = time.perf_counter()
            pol, dpol, _, _ = orc.backward(sy["vT"], sy["r"], sy["w"], drh.numpy()[:Kc], dwh.numpy()[:Kc])
            _, JVc = orc.forward(sy["D0"], pol, dpol)
            dtc = time.perf_counter() - tw
        else:
            orc = O.Oracle(g["grid"], g["z"], g["Pi"], float(g["beta"]), float(g["gamma"]), float(g["borrow_cons"]), fx["T"])
            Vc = Vh.numpy()[:Kc]
            orc.ks_fjvp(fx["ks"], g["ss_value"], g["ss_D"], fx["Z"], fx["x0"])  # warm-up (primal only)
            tw = time.perf_counter()
            Fc, JVc = orc.ks_fjvp(fx["ks"], g["ss_value"], g["ss_D"], fx["Z"], fx["x0"], Vc)
            dtc = time.perf_counter() - tw
        # the CPU sample doubles as an in-run parity check of the GPU result
        err = float(np.max(np.abs(JVh.numpy()[:Kc] - JVc) / (1e-12 * max(1.0, np.abs(JVc).max()) + 1e-10 * np.abs(JVc))))
        out["cpu_baseline"] = {"value": Kc / dtc, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"1 step of primal + {Kc} of the {K} lanes ({dtc:.1f} s)",
                               "host_nproc": os.cpu_count(), "parity_max_err_over_tol": err,
                               "note": "single-thread C++ restatement (oracle/); Julia is not installed in this image"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
