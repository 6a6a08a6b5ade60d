#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 batched MCMC engine.

Metric (BASELINE.json): chain-iterations/s of AMWG on the `seeds` random-effects logistic model
(configs[1]): scheme [AMWG(alpha0..alpha12, 0.1), AMWG(b, 0.01), AMWG(s2, 0.1)],
125,000 chains per GPU (10^6 chains on 8 GPUs, weak scaling), one step = one mcmc() call of 2,000
iterations (burn-in 1,000, thin 10) for every chain, followed by the on-device Gelman-Rubin /
summary reductions (all-reduced over NCCL when N > 1 — the only collective on the path).

  value : chain-iterations/s over all GPUs with the inputs resident in HBM
  e2e   : the same through the C ABI with HOST buffers: per-chain initial values copied from pinned host
          memory every step, final chain states + diagnostics read back every step
  --impl reference : the CPU restatement of the reference's algorithm (oracle/) on all host threads,
          on a bounded sample of the same workload (the Julia reference cannot run here: no julia).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

CHAINS_PER_GPU = 125_000
ITERS, BURNIN, THIN = 2000, 1000, 10
SEED = 123
# Algorithmic FP64 work of one chain-iteration in the minimal form the fused kernel implements, DESIGN.md §4:
# 25 exp (one per alpha proposal + one per b_i) x 28 flop + 68 plate terms x (log 40 + 8) flop
# + 13 Box-Muller pairs x 110 flop (log, sqrt, sincos) + 26 MH tests x 20 flop + 26 x 2 + 60.
ALGO_FLOP_PER_CHAIN_ITER = 25 * 28 + 68 * 48 + 13 * 110 + 26 * 20 + 26 * 2 + 60
# Algorithmic HBM bytes of one chain-iteration: thinned output only (5 monitored doubles every THIN iterations)
ALGO_BYTES_PER_CHAIN_ITER = 5 * 8 / THIN

SCHEME = [dict(kind="amwg", nodes=[0, 1, 2, 3], scale=0.1), dict(kind="amwg", nodes=[5], scale=0.01),
          dict(kind="amwg", nodes=[4], scale=0.1)]


# dram__bytes_read.sum + dram__bytes_write.sum of one seeds_fast_kernel launch (ncu --set full capture of this command,
# profiles/r1_seeds_fast_summary.md), keyed by (chains per GPU, iterations per launch)
SEEDS_KERNEL_DRAM_BYTES = {(125000, 2000): 86.63e6 + 117.40e6}


def seeds_inits():
    x = np.zeros((2, 26)); x[0, 4] = 0.01; x[1, 4] = 1.0   # doc/examples/seeds.jl:60-65
    return x


def per_chain_inits(n, offset):
    """Host-side initial values for every chain: the two reference init records cycled, plus a deterministic jitter."""
    base = seeds_inits()
    idx = (np.arange(n) + offset) % 2
    x = base[idx].copy()
    rng = np.random.default_rng(SEED + offset)
    x[:, :4] += 0.1 * rng.standard_normal((n, 4))
    x[:, 4] *= np.exp(0.1 * rng.standard_normal(n))
    x[:, 5:] += 0.1 * rng.standard_normal((n, 21))
    return x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def run_oracle_sample(n_chains, iters, burnin, thin, nthreads):
    import helpers
    import pyoracle
    orc = pyoracle.Oracle("seeds")
    orc.set_scheme([helpers.oracle_block(b) for b in SCHEME])
    t0 = time.perf_counter()
    orc.run(n_chains, seeds_inits(), iters, burnin=burnin, thin=thin, seed=SEED, nthreads=nthreads, store=False)
    dt = time.perf_counter() - t0
    return n_chains * iters / dt, dt


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path: the oracle port, all host threads, bounded sample."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_chains = 64 * cores
    for _ in range(args.warmup):
        run_oracle_sample(cores, 200, 100, THIN, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_oracle_sample(n_chains, ITERS, BURNIN, THIN, cores)
    dt = time.perf_counter() - t0
    value = args.steps * n_chains * ITERS / dt
    sample = f"{n_chains} chains (64 per host thread) x {ITERS} iterations per step, same model/scheme/burn-in/thinning"
    line = {
        "impl": "reference", "metric": "chain_iters_per_sec", "value": value, "unit": "chain-iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "seeds random-effects logistic regression (21 plates, doc/examples/seeds.jl), AMWG(alpha0..alpha12)+AMWG(b)+AMWG(s2)",
                   "chains_per_gpu": 125000, "chains_total": 125000 * args.gpus, "iters": ITERS, "burnin": BURNIN, "thin": THIN,
                   "sample_chains": n_chains, "note": "Julia reference not runnable (no julia in image); CPU restatement of its algorithm (oracle/) timed instead"},
        "cpu_baseline": {"value": value, "unit": "chain-iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "chain-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def glm_synthetic(N, d, family="logit"):
    """X[:,0] = 1, X[:,1:] ~ N(0,1), beta* ~ N(0, I/d), y ~ Bernoulli(invlogit(X beta*));
    the other members of the GLM family: y ~ Poisson(exp(X beta*)), y ~ Normal(X beta*, 1)."""
    rng = np.random.default_rng(1)
    X = rng.standard_normal((N, d)); X[:, 0] = 1.0
    beta = np.random.default_rng(2).standard_normal(d) / np.sqrt(d)
    eta = X @ beta
    r3 = np.random.default_rng(3)
    if family == "poisson":
        y = r3.poisson(np.exp(np.clip(eta, -20, 5))).astype(np.float64)
    elif family == "normal":
        y = eta + r3.standard_normal(N)
    else:
        y = (r3.uniform(size=N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return X, y, beta


def glm_bench(args, rank, local_rank, world):
    """configs[3]: NUTS on Bayesian logistic regression; reports the gradient-pass rate against the tensor roofline
    and the chain-iteration / leapfrog rate of a short NUTS run (same JSON contract as the headline line)."""
    import torch
    from mambacuda.engine import Engine
    torch.cuda.set_device(local_rank)
    N, d, C = args.glm_n, args.glm_d, args.glm_chains
    X, y, beta_true = glm_synthetic(N, d, args.glm_family)
    eng = Engine("glm", C, seed=SEED, chain_offset=rank * C, device=local_rank)
    eng.set_data("X", X); eng.set_data("y", y)
    eng.set_data("family", np.array([{"logit": 0.0, "poisson": 1.0, "normal": 2.0}[args.glm_family]]))
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    beta = 0.1 * np.random.default_rng(5).standard_normal((C, d))
    # gradient pass alone (the dominant kernel): tensor-core kernel vs FP64 reference kernel
    times = {}
    for impl, reps in ((1, max(args.steps, 3) + 3), (0, 2)):
        ms = []
        for r in range(reps):
            lp, g = eng.glm_gradient(beta, impl=impl)
            ms.append(eng.last_kernel_ms())
        times[impl] = float(np.mean(ms[3:])) if impl == 1 else float(ms[-1])
        if impl == 1:
            lp1, g1 = lp, g
        else:
            err_lp = float(np.max(np.abs(lp1 - lp) / np.abs(lp)))
            err_g = float(np.max(np.abs(g1 - g) / np.abs(g).max(axis=1, keepdims=True)))
    flops = 4.0 * C * N * d
    peaks, peak_src = measured_peaks()
    # short NUTS run through the tick engine (adaptation on): leapfrogs = gradient passes ("ticks")
    eng.set_inits(np.zeros((1, d)), jitter_sd=0.1)
    sampler = ClockSampler(local_rank); sampler.start()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    launches0 = eng.launch_count()
    eng.run(args.glm_iters, burnin=args.glm_iters // 2, thin=1, store=False, out=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    ticks = launches // 3      # advance + tensor-core pass + fold per tick
    summ = eng.summary_streaming()
    if rank == 0:
        ach = flops / (times[1] * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        line = {
            "metric": "chain_iters_per_sec", "value": world * C * args.glm_iters / dt, "unit": "chain-iterations/s", "n_gpus": world,
            "steps": 1, "warmup": 3, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16x2-split tensor (f32 accumulate) + f64 NUTS state", "data": "synthetic",
            "config": {"workload": f"Bayesian {args.glm_family} regression N={N}, d={d}, NUTS(beta), {C} chains/GPU (configs[3])",
                       "iters": args.glm_iters, "burnin": args.glm_iters // 2, "gradient_passes": int(ticks),
                       "l2": "X (449 MB packed) exceeds L2; every pass streams it from HBM",
                       "posterior_mean_abs_err_vs_truth": float(np.mean(np.abs(summ[:, 0] - beta_true)))},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "effective_peak_note": "operands are split fp16 pairs: 3 tensor products per algorithmic product, so the reachable fraction is 1/3",
                         "traffic": None, "kernel": "glm_tc_kernel", "kernel_ms": times[1], "algo_flops_per_pass": flops,
                         "peak_source": peak_src + " (sustained bf16; fp16 runs at the same rate)",
                         "fp64_reference_kernel_ms": times[0], "max_rel_err_logf_vs_fp64": err_lp, "max_rel_err_grad_vs_fp64": err_g,
                         "hbm": {"achieved": (np.ceil(C / 128) * (N * 112 * 4)) / (times[1] * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s"}},
            "e2e": {"value": world * C * args.glm_iters / dt, "unit": "chain-iterations/s", "h2d_bytes_per_step": int(d * 8), "d2h_bytes_per_step": int(d * 5 * 8)},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)


def small_model_bench(args, rank, local_rank, world):
    """The other configurations of BASELINE.json on the generic engine kernel (one chain per thread):
    configs[0] line (3 chains x 10,000, CPU-scale sanity), configs[2] rats (NUTS + Slice, 65,536 chains),
    configs[4] pumps (chain sweep with on-device Gelman-Rubin).  One JSON line per run."""
    import helpers
    import torch
    from mambacuda.engine import Engine
    torch.cuda.set_device(local_rank)
    runs = []
    if args.workload == "line":
        runs = [("line_amwg_slice", 3, 10000, 1000, 1)]
    elif args.workload == "rats":
        runs = [("rats_nuts_slice", 65536, 2000, 1000, 5), ("rats_slice_amwg", 65536, 2000, 1000, 5)]
    else:
        # BASELINE.json configs[4]: Gibbs + AMWG, chain sweep 10^3 .. 10^7 with on-device Gelman-Rubin; the reference's own Slice scheme beside it
        runs = [("pumps_gibbs_amwg", n, 2000, 1000, 10) for n in (10**3, 10**4, 10**5, 10**6, 10**7)] + [("pumps_slice", 10**6, 2000, 1000, 10)]
    for name, C, iters, burnin, thin in runs:
        tpl, blocks, inits = helpers.scheme(name)
        eng = Engine(tpl, C, seed=SEED, chain_offset=rank * C, device=local_rank)
        eng.set_scheme(blocks)
        eng.set_inits(inits, jitter_sd=0.05 if C > 16 else 0.0)
        eng.run(min(iters, 20), burnin=min(burnin, 10), thin=1, store=False, out=False)     # warm-up launch
        eng.set_inits(inits, jitter_sd=0.05 if C > 16 else 0.0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.run(iters, burnin=burnin, thin=thin, store=False, out=False)
        kms = eng.last_kernel_ms()
        psrf = eng.gelman(0.05, True) if C >= 2 else None
        summ = eng.summary_streaming() if (iters - burnin) // thin >= 200 else None
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if rank == 0:
            line = {"metric": "chain_iters_per_sec", "value": world * C * iters / dt, "unit": "chain-iterations/s", "n_gpus": world, "steps": 1,
                    "warmup": 1, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                    "data": "reference data set", "config": {"workload": name, "chains_per_gpu": C, "iters": iters, "burnin": burnin, "thin": thin,
                                                             "kernel": {"rats_nuts_slice": "rats_warp_kernel (one warp per chain)", "rats_slice_amwg": "rats_fast_kernel (fused)", "pumps_slice": "pumps_fast_kernel (fused)", "pumps_gibbs_amwg": "pumps_gibbs_kernel (fused)"}.get(name, "run_generic_kernel"), "kernel_ms": kms,
                                                             "psrf_max": None if psrf is None else float(np.nanmax(psrf[:, 0])),
                                                             "names": eng.names(1)[:12],
                                                             "posterior_mean": None if summ is None else [float(v) for v in summ[:12, 0]],
                                                             "ess_per_sec_min": None if summ is None else float(np.nanmin(summ[:, 4]) * world * C / dt)},
                    "gpu_launches": int(eng.launch_count())}
            print(json.dumps(line), flush=True)
        eng.close()


def _protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1), so fd 1 is
    pointed at stderr for the whole run and the JSON lines go to a private duplicate of the original stdout."""
    import builtins
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    orig_print = builtins.print

    def print_json(*a, **k):
        if k.get("file") is None and a and isinstance(a[0], str) and a[0].startswith("{"):
            k["file"] = real
            k["flush"] = True
        return orig_print(*a, **k)
    builtins.print = print_json


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains-per-gpu", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--iters", type=int, default=ITERS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-generic", action="store_true", help="time the generic engine kernel instead of the fused one")
    ap.add_argument("--workload", default="seeds", choices=["seeds", "glm", "rats", "pumps", "line"],
                    help="seeds = headline (configs[1]); glm = configs[3]: NUTS logistic regression N=1e6, d=100 (tensor-core likelihood)")
    ap.add_argument("--glm-n", type=int, default=1_000_000)
    ap.add_argument("--glm-d", type=int, default=100)
    ap.add_argument("--glm-chains", type=int, default=512, help="chains per GPU (4096 chains on 8 GPUs)")
    ap.add_argument("--glm-iters", type=int, default=40)
    ap.add_argument("--glm-family", default="logit", choices=["logit", "poisson", "normal"], help="member of the GLM family (tensor-core epilogue)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.workload == "glm":
        glm_bench(args, rank, local_rank, world)
        return
    if args.workload in ("rats", "pumps", "line"):
        small_model_bench(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    from mambacuda import distributed as mdist
    from mambacuda.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    C = args.chains_per_gpu
    iters = args.iters
    burnin = min(BURNIN, iters // 2)
    eng = Engine("seeds", C, seed=SEED, chain_offset=rank * C, device=local_rank)
    eng.set_scheme(SCHEME)
    inits2 = seeds_inits()
    peak_fp64 = eng.fp64_peak_tflops()
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        l2_flush.fill_(1)                                  # evict L2 between steps
        eng.set_inits(inits2, jitter_sd=0.1)               # 2 init records, cycled + Philox jitter on the device
        eng.run(iters, burnin=burnin, thin=THIN, store=False, out=False, force_generic=args.force_generic)
        ms = eng.last_kernel_ms()
        psrf = mdist.global_gelman(eng, 0.05, True, device)
        return ms, psrf

    # pinned host buffers for the end-to-end arm
    host_inits = torch.empty((C, 26), dtype=torch.float64).pin_memory()
    host_inits.numpy()[:] = per_chain_inits(C, rank * C)
    host_state = torch.empty((C, 26), dtype=torch.float64).pin_memory()

    def step_e2e():
        l2_flush.fill_(1)
        eng.set_inits(host_inits.numpy())                  # H2D: C x 26 doubles from pinned memory
        eng.run(iters, burnin=burnin, thin=THIN, store=False, out=False, force_generic=args.force_generic)
        psrf = mdist.global_gelman(eng, 0.05, True, device)
        summ = mdist.global_summary(eng, device)
        import ctypes as Ct
        dp = host_state.numpy().ctypes.data_as(Ct.POINTER(Ct.c_double))
        it = Ct.c_int64()
        eng._chk(eng.L.mcu_get_state(eng.h, dp, None, Ct.byref(it)))   # D2H: final states into pinned memory
        return psrf, summ

    # ---- resident arm ---------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    launches0 = eng.launch_count()
    t0 = time.perf_counter()
    kernel_ms = []
    psrf = None
    for _ in range(args.steps):
        ms, psrf = step_resident()
        kernel_ms.append(ms)
    sync_all()
    dt = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    clocks = sampler.stop()
    tmax = torch.tensor([dt], dtype=torch.float64, device=device)
    kmax = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item()); kms = float(kmax.item())
    value = world * C * iters * args.steps / dt

    # ---- end-to-end arm -------------------------------------------------------------------------
    for _ in range(2):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        psrf_e, summ_e = step_e2e()
    sync_all()
    dte = time.perf_counter() - t0
    tmax = torch.tensor([dte], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dte = float(tmax.item())
    e2e_value = world * C * iters * args.steps / dte
    h2d = C * 26 * 8
    d2h = C * 26 * 8 + 5 * 2 * 8 + 5 * 5 * 8 + 4 * 5 * 7 * 8

    if rank == 0:
        peaks, peak_src = measured_peaks()
        kernel_rate = C * iters / (kms * 1e-3)             # chain-iterations/s of the dominant kernel on one GPU
        achieved_tflops = kernel_rate * ALGO_FLOP_PER_CHAIN_ITER / 1e12
        kept = (iters - burnin) // THIN
        hbm_bytes = C * (kept * 5 * 8)                     # algorithmic: the thinned monitored values
        line = {
            "metric": "chain_iters_per_sec", "value": value, "unit": "chain-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "seeds random-effects logistic regression (21 plates, doc/examples/seeds.jl), AMWG(alpha0..alpha12)+AMWG(b)+AMWG(s2)",
                "chains_per_gpu": C, "chains_total": world * C, "iters": iters, "burnin": burnin, "thin": THIN,
                "kernel": "generic" if args.force_generic else "seeds_fast (fused)",
                "parallelism": f"chains sharded over {world} GPU(s), no data-path collective; all-reduce of 7p moment sums for PSRF",
                "l2": "L2 flushed between steps (256 MiB fill); chain state is register/shared-memory resident inside a step",
                "psrf_max": float(np.max(psrf[:, 0])) if psrf is not None else None,
                "psrf_note": "SURVEY.md §8d config 2 runs 2,000 iterations from the reference's two dispersed initial records (s2 = 0.01 / 1): "
                             "not yet mixed in s2 (the reference runs 12,500); the converged check against doc/examples/seeds.rst is "
                             "tests/test_gpu_parity.py::test_seeds_fast_posterior_within_3_mcse_of_reference",
            },
            "roofline": {
                "bound": "fp64", "achieved": achieved_tflops, "peak": peak_fp64, "unit": "TFLOP/s",
                "frac": achieved_tflops / peak_fp64 if peak_fp64 > 0 else None, "traffic": None,
                "peak_source": "DFMA microbenchmark run by this process (mcu_fp64_peak_tflops); MEASURED_PEAKS.json has no FP64 figure",
                "kernel": "seeds_fast_kernel<96>" if not args.force_generic else "run_generic_kernel<SeedsModel>",
                "kernel_ms": kms, "algo_flop_per_chain_iter": ALGO_FLOP_PER_CHAIN_ITER,
                "note": "CUDA-core FP64 kernel: neither HBM- nor tensor-bound (SURVEY.md §8d); the hbm object shows the memory side is idle",
                "hbm": {"bound": "hbm", "achieved": hbm_bytes / (kms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": hbm_bytes / (kms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 1.0), "peak_source": peak_src},
            },
            "e2e": {"value": e2e_value, "unit": "chain-iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * dte / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        # ESS/s (BASELINE.json metric, second half): summarystats' ESS is (SD / MCSE)^2 with batch-means MCSE over the draws of
        # ALL chains (src/output/stats.jl:85-94, mcse.jl:10-19); the reference then caps it at the per-chain draw count, which
        # is meaningless at 10^6 chains, so the uncapped value over all chains is reported, per second of the end-to-end step
        ess_all = (summ_e[:, 1] / summ_e[:, 3]) ** 2
        line["ess"] = {"ess_per_sec_min": float(np.nanmin(ess_all) / (dte / args.steps)), "ess_min": float(np.nanmin(ess_all)),
                       "kept_draws": int(world * C * kept), "names": ["alpha0", "alpha1", "alpha2", "alpha12", "s2"],
                       "ess_per_param": [float(v) for v in ess_all],
                       "definition": "(SD/MCSE_bm)^2 over the kept draws of all chains, batch size 100, uncapped; per second of one end-to-end step"}
        line["roofline"]["traffic"] = SEEDS_KERNEL_DRAM_BYTES.get((C, iters))
        line["roofline"]["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one seeds_fast_kernel launch, ncu --set full "
                                              "(profiles/r1_seeds_fast_summary.md); null for other chain / iteration counts")
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_s = 128 * cores    # ~10-20 s of CPU work
            v, secs = run_oracle_sample(n_s, ITERS, BURNIN, THIN, cores)
            line["cpu_baseline"] = {"value": v, "unit": "chain-iterations/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_s} chains x {ITERS} iterations of the same model/scheme on {cores} host threads ({secs:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
