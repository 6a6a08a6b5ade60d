#!/usr/bin/env python
"""bench.py — benchmark of the B200 batched MCMC engine: ONE JSON line per run.

Headline (BASELINE.json metric, configs[1]): chain-iterations/s of AMWG on the `seeds` random-effects logistic
model, scheme [AMWG(alpha0..alpha12, 0.1), AMWG(b, 0.01), AMWG(s2, 0.1)], 125,000 chains per GPU (10^6 chains on 8 GPUs, weak scaling);
one step = one mcmc() call of 2,000 iterations (burn-in 1,000, thin 10) for every chain + the on-device Gelman-Rubin / summary
reductions over the chains of ALL GPUs (NCCL inside libmambacuda when N > 1 — the only collective on the path).

  value      chain-iterations/s over all GPUs with the inputs resident in HBM
  e2e        the same through the C ABI with HOST buffers: per-chain initial values copied from pinned host memory every step, final
             chain states + PSRF + summary statistics read back every step
  e2e_full   an mcmc() that returns ModelChains.value itself (4,096 chains x 1,000 kept draws into pinned host memory)
  ess        ESS/s from a CONVERGED run: the reference's own run length for this model (12,500 iterations, burn-in 2,500, thin 2:
             doc/examples/seeds.jl:75) on all 125,000 chains per GPU, PSRF reported beside it
  configs    the other BASELINE.json configurations measured in the same process after the headline: configs[0] line, configs[2] rats
             (NUTS + Slice and the reference's Slice + AMWG), configs[3] GLM / NUTS (N = 10^6, d = 100, tensor-core likelihood),
             configs[4] pumps (10^7 chains, on-device PSRF) — each with roofline, cpu_baseline, e2e (host buffers) and clocks
  --impl reference   the CPU restatement of the reference's algorithm (oracle/) on all host threads, on a bounded sample of the
             headline workload (the Julia reference cannot run here: no julia in the image)
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
# ---- algorithmic work per unit (DESIGN.md §4 states each model; redundant work cannot inflate them) --------------------------------
# seeds, minimal form the fused kernel implements: 25 exp (one per alpha proposal + one per b_i) x 28 flop + 68 plate terms x (log 40
# + 8) flop + 13 Box-Muller pairs x 110 flop (log, sqrt, sincos) + 26 MH tests x 20 flop + 26 x 2 + 60.
ALGO_FLOP_PER_CHAIN_ITER = 25 * 28 + 68 * 48 + 13 * 110 + 26 * 20 + 26 * 2 + 60
# rats, one leapfrog of the 62-dimensional NUTS block: 150 residuals x 7 (value + three gradient sums) + 60 hierarchical terms x 6
# + 62 momentum / position updates x 6 + kinetic energy and logp assembly 144
RATS_FLOP_PER_LEAPFROG = 150 * 7 + 60 * 6 + 62 * 6 + 144
# rats, reference scheme (Slice + AMWG): 60 AMWG component updates x (25 flop MH test on sufficient statistics + half a Box-Muller pair
# 55 + log-uniform 40) + ~15 slice evaluations x 30
RATS_FAST_FLOP_PER_ITER = 60 * (25 + 55 + 40) + 15 * 30
# pumps Gibbs + AMWG: 11 Gamma variates x (normal 55 + cube / squeeze test 30 + 10 % full test with two logs 8) + AMWG(alpha):
# lgamma 60 + 10 log theta x 40 + exp / log-uniform / normal 140
PUMPS_FLOP_PER_ITER = 11 * (55 + 30 + 8) + 60 + 400 + 140
# line AMWG(beta) + Slice(s2): 3 + ~4 block evaluations x (5 residuals x 4 + 20)
LINE_FLOP_PER_ITER = 7 * 40 + 2 * 55 + 6 * 40

SCHEME = [dict(kind="amwg", nodes=[0, 1, 2, 3], scale=0.1), dict(kind="amwg", nodes=[5], scale=0.01),
          dict(kind="amwg", nodes=[4], scale=0.1)]
SEEDS_WORKLOAD = "seeds random-effects logistic regression (21 plates, doc/examples/seeds.jl), AMWG(alpha0..alpha12)+AMWG(b)+AMWG(s2)"

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels (ncu --set full captures, profiles/), keyed by shape
SEEDS_KERNEL_DRAM_BYTES = {(125000, 2000): 85.71e6 + 117.95e6}   # profiles/r2_seeds_dram_2000.csv (round-2 kernel; round 1: 86.63e6 + 117.40e6)
GLM_KERNEL_DRAM_BYTES = {512: 453e6 + 5e6, 4096: 546.2e6 + 38.0e6}   # profiles/r1_glm_tc_summary.md, profiles/r2_glm_tc_c4096_ncu.csv


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
        return self

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


# ---- CPU baseline: the oracle port on the box's host threads, bounded samples --------------------------------------------------------
def oracle_rate(template, blocks, inits, n_chains, iters, burnin, thin, nthreads, jitter_sd=0.0, glm=None, max_depth=10):
    """chain-iterations/s of the CPU restatement of the reference's algorithm (oracle/) on `nthreads` host threads."""
    import helpers
    import pyoracle
    orc = pyoracle.Oracle(template, glm_d=glm[0].shape[1] if glm else 0)
    if glm:
        orc.set_data("X", glm[0]); orc.set_data("y", glm[1])
    ob = [helpers.oracle_block(b) for b in blocks]
    for b in ob:
        if b["kind"] in ("nuts", 4):
            b["max_depth"] = max_depth
    orc.set_scheme(ob)
    t0 = time.perf_counter()
    orc.run(n_chains, inits, iters, burnin=burnin, thin=thin, seed=SEED, nthreads=nthreads, store=False, jitter_sd=jitter_sd)
    dt = time.perf_counter() - t0
    return n_chains * iters / dt, dt


def run_oracle_sample(n_chains, iters, burnin, thin, nthreads):
    return oracle_rate("seeds", SCHEME, seeds_inits(), n_chains, iters, burnin, thin, nthreads)


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
        "config": {"workload": SEEDS_WORKLOAD,
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


class Ctx:
    """What every leg of the bench needs: ranks, device, barrier + max-over-ranks timing."""

    def __init__(self, rank, local_rank, world):
        import torch
        self.torch = torch
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.device = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
        self.peaks, self.peak_src = measured_peaks()
        self.l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)
        self.peak_fp64 = None

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.device)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.device)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def engine(self, template, n_chains, chain_offset, seed=SEED):
        """A handle on this rank's GPU; with N > 1 it joins an NCCL communicator of its own (mcu_comm_init) for mcu_diag_global."""
        from mambacuda import distributed as mdist
        from mambacuda.engine import Engine
        eng = Engine(template, n_chains, seed=seed, chain_offset=chain_offset, device=self.local_rank)
        if self.dist is not None:
            mdist.init_comm(eng)
        return eng

    def pinned(self, shape, order="C"):
        t = self.torch.empty(int(np.prod(shape)), dtype=self.torch.float64).pin_memory()
        return t, t.numpy().reshape(shape, order=order)


def fp64_roofline(ctx, flop_per_unit, units, kernel_ms, kernel, note, traffic=None, traffic_source=None, hbm_bytes=None):
    ach = flop_per_unit * units / (kernel_ms * 1e-3) / 1e12
    r = {"bound": "fp64", "achieved": ach, "peak": ctx.peak_fp64, "unit": "TFLOP/s", "frac": ach / ctx.peak_fp64 if ctx.peak_fp64 else None,
         "traffic": traffic, "kernel": kernel, "kernel_ms": kernel_ms, "algo_flop_per_unit": flop_per_unit, "units_per_launch": units,
         "peak_source": "DFMA microbenchmark run by this process (mcu_fp64_peak_tflops; profiles/r2_fp64_peak.md sets it beside the datasheet figure); "
                        "MEASURED_PEAKS.json has no FP64 figure",
         "note": note}
    if traffic_source:
        r["traffic_source"] = traffic_source
    if hbm_bytes is not None:
        g = hbm_bytes / (kernel_ms * 1e-3) / 1e9
        r["hbm"] = {"bound": "hbm", "achieved": g, "peak": ctx.peaks.get("hbm_gbs"), "unit": "GB/s", "frac": g / ctx.peaks.get("hbm_gbs", 1.0),
                    "peak_source": ctx.peak_src}
    return r


# ---- the other BASELINE.json configurations ------------------------------------------------------------------------------------------
def small_config(ctx, name, scheme_name, C_total, iters, burnin, thin, flop_per_unit, unit_is_leapfrog, kernel, cpu_chains, cpu_iters, strong=False,
                 jitter=0.05, note=""):
    """One configuration on a fused / generic small-model kernel.  Timed once, end to end through host buffers:
    H2D of per-chain initial values (pinned), mcmc(), diagnostics over all GPUs, D2H of the final states."""
    import helpers
    tpl, blocks, inits = helpers.scheme(scheme_name)
    C = max(1, C_total // ctx.world) if strong else C_total
    eng = ctx.engine(tpl, C, ctx.rank * C)
    eng.set_scheme(blocks)
    D = eng.dims()[0]
    # warm-up launch (module load, attribute calls), then per-chain host inits: the script's records cycled + jitter, materialised on the device once
    eng.set_inits(inits, jitter_sd=jitter if C > 16 else 0.0)
    eng.run(min(iters, 20), burnin=min(burnin, 10), thin=1, store=False, out=False)
    if ctx.world * C >= 2:
        eng.diag_global(0.05, True)               # first collective of this handle's communicator: NCCL sets its connections up here, not in the timed call
    eng.set_inits(inits, jitter_sd=jitter if C > 16 else 0.0)
    st0, _, _ = eng.get_state()
    keep_in, host_in = ctx.pinned((C, D)); host_in[:] = st0
    keep_out, host_out = ctx.pinned((C, D))
    import ctypes as Ct
    ctx.l2_flush.fill_(1)
    sampler = ClockSampler(ctx.local_rank).start()
    w0, _ = eng.work_count(); cap0 = eng.nuts_cap_hits
    ctx.sync_all(); t0 = time.perf_counter()
    eng.set_inits(host_in)                                                         # H2D
    eng.run(iters, burnin=burnin, thin=thin, store=False, out=False)
    kms = eng.last_kernel_ms()
    t1 = time.perf_counter()
    psrf, summ, _ = eng.diag_global(0.05, True) if ctx.world * C >= 2 else (None, eng.summary_streaming(), None)
    t2 = time.perf_counter()
    it = Ct.c_int64()
    eng._chk(eng.L.mcu_get_state(eng.h, host_out.ctypes.data_as(Ct.POINTER(Ct.c_double)), None, Ct.byref(it)))   # D2H
    ctx.sync_all(); dt = time.perf_counter() - t0
    clocks = sampler.stop()
    w1, _ = eng.work_count(); cap1 = eng.nuts_cap_hits
    dt, kms, diag_s = ctx.max_over_ranks(dt, kms, t2 - t1)
    (leap,) = ctx.sum_over_ranks(w1 - w0)
    chains_total = C * ctx.world
    kept = max(0, (iters - burnin) // thin)
    out = None
    if ctx.rank == 0:
        units = (w1 - w0) if unit_is_leapfrog else C * iters
        res_time = kms * 1e-3 + diag_s
        ess_all = None
        if summ is not None and kept >= 200:
            ess_all = (summ[:, 1] / summ[:, 3]) ** 2
        out = {
            "metric": "chain_iters_per_sec", "value": chains_total * iters / res_time, "unit": "chain-iterations/s", "n_gpus": ctx.world,
            "ms_per_step": 1e3 * res_time, "scaling": "strong" if strong else "weak", "dtype": "f64", "data": "reference data set",
            "config": {"workload": name, "scheme": scheme_name, "chains_per_gpu": C, "chains_total": chains_total, "iters": iters, "burnin": burnin,
                       "thin": thin, "kernel": kernel, "l2": "L2 flushed before the timed call; chain state stays on chip inside it",
                       "psrf_max": None if psrf is None else float(np.nanmax(psrf[:, 0])), "names": eng.names(1)[:12],
                       "posterior_mean": None if summ is None else [float(v) for v in summ[:12, 0]]},
            "roofline": fp64_roofline(ctx, flop_per_unit, units, kms, kernel, note, hbm_bytes=C * (kept * eng.dims()[1] * 8 + 2 * D * 8)),
            "e2e": {"value": chains_total * iters / dt, "unit": "chain-iterations/s", "h2d_bytes_per_step": int(C * D * 8),
                    "d2h_bytes_per_step": int(C * D * 8 + eng.dims()[1] * 7 * 8), "ms_per_step": 1e3 * dt},
            "gpu_launches": int(eng.launch_count()), "clocks": clocks,
        }
        if unit_is_leapfrog:
            out["config"]["leapfrogs_all_gpus"] = int(leap)
            out["config"]["leapfrogs_per_chain_iter"] = leap / (chains_total * iters)
            out["config"]["nuts_depth_cap_hits_per_chain_iter"] = (cap1 - cap0) / (C * iters)     # reference: no cap (nuts.jl:106-124); device: 10 doublings
        if ess_all is not None:
            out["ess"] = {"ess_per_sec_min": float(np.nanmin(ess_all) / dt), "ess_min": float(np.nanmin(ess_all)),
                          "definition": "(SD/MCSE_bm)^2 over the kept draws of all chains, batch size 100 (batches never straddle chains), uncapped; per second of the end-to-end call"}
        if cpu_chains:
            cores = os.cpu_count() or 1
            v, secs = oracle_rate(tpl, blocks, inits, cpu_chains, cpu_iters, min(burnin, cpu_iters // 2), thin, cores, jitter_sd=jitter if cpu_chains > 16 else 0.0)
            out["cpu_baseline"] = {"value": v, "unit": "chain-iterations/s", "cores": min(cores, cpu_chains), "kind": "port",
                                   "sample": f"{cpu_chains} chains x {cpu_iters} iterations of the same model / scheme on {min(cores, cpu_chains)} host threads ({secs:.1f} s)"}
    eng.close()
    return out


def glm_config(ctx, args):
    """configs[3]: NUTS on Bayesian logistic regression, N = 10^6, d = 100: gradient-pass rate against the tensor roofline, and a NUTS run
    through the tick engine, end to end through host buffers."""
    N, d = args.glm_n, args.glm_d
    C = args.glm_chains if args.glm_chains > 0 else (4096 if ctx.world == 1 else 512)
    iters = args.glm_iters
    X, y, beta_true = glm_synthetic(N, d, args.glm_family)
    eng = ctx.engine("glm", C, ctx.rank * C)
    eng.set_data("X", X); eng.set_data("y", y)
    eng.set_data("family", np.array([{"logit": 0.0, "poisson": 1.0, "normal": 2.0}[args.glm_family]]))
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    beta = 0.1 * np.random.default_rng(5 + ctx.rank).standard_normal((C, d))
    ms = []
    for r in range(8):
        lp1, g1 = eng.glm_gradient(beta, impl=1)
        ms.append(eng.last_kernel_ms())
    pass_ms = float(np.mean(ms[3:]))
    sub = slice(0, min(C, 512))
    ref = eng if C <= 512 else None
    if ref is None:      # the FP64 CUDA-core kernel is ~370x slower: compare on 512 chains of the same positions
        from mambacuda.engine import Engine
        ref = Engine("glm", 512, seed=SEED, device=ctx.local_rank)
        ref.set_data("X", X); ref.set_data("y", y); ref.set_scheme([dict(kind="nuts", nodes=[0])])
    lp0, g0 = ref.glm_gradient(beta[sub], impl=0)
    ref_ms = ref.last_kernel_ms()
    err_lp = float(np.max(np.abs(lp1[sub] - lp0) / np.abs(lp0)))
    err_g = float(np.max(np.abs(g1[sub] - g0) / np.abs(g0).max(axis=1, keepdims=True)))
    if ref is not eng:
        ref.close()
    # NUTS run, end to end: per-chain initial positions from pinned host memory, diagnostics over all GPUs, final states back
    eng.set_inits(np.zeros((1, d)), jitter_sd=0.1)
    eng.run(4, burnin=1, thin=1, store=False, out=False)      # warm-up: tick-engine buffers, and the first collective of the communicator
    eng.diag_global(0.05, False)
    keep_in, host_in = ctx.pinned((C, d)); host_in[:] = 0.1 * np.random.default_rng(7 + ctx.rank).standard_normal((C, d))
    keep_out, host_out = ctx.pinned((C, d))
    import ctypes as Ct
    sampler = ClockSampler(ctx.local_rank).start()
    w0, k0 = eng.work_count(); s0 = eng.glm_pass_slots; cap0 = eng.nuts_cap_hits; l0 = eng.launch_count()
    ctx.sync_all(); t0 = time.perf_counter()
    eng.set_inits(host_in)
    eng.run(iters, burnin=iters // 2, thin=1, store=False, out=False)
    kms = eng.last_kernel_ms()
    psrf, summ, _ = eng.diag_global(0.05, False)
    it = Ct.c_int64()
    eng._chk(eng.L.mcu_get_state(eng.h, host_out.ctypes.data_as(Ct.POINTER(Ct.c_double)), None, Ct.byref(it)))
    ctx.sync_all(); dt = time.perf_counter() - t0
    clocks = sampler.stop()
    w1, k1 = eng.work_count(); s1 = eng.glm_pass_slots; cap1 = eng.nuts_cap_hits; launches = eng.launch_count() - l0
    dt, kms = ctx.max_over_ranks(dt, kms)
    out = None
    if ctx.rank == 0:
        flops = 4.0 * C * N * d
        ach = flops / (pass_ms * 1e-3) / 1e12
        peak = ctx.peaks.get("bf16_tflops_sustained", ctx.peaks.get("bf16_tflops"))
        ticks = k1 - k0
        out = {
            "metric": "chain_iters_per_sec", "value": ctx.world * C * iters / (kms * 1e-3), "unit": "chain-iterations/s", "n_gpus": ctx.world,
            "ms_per_step": kms, "scaling": "weak", "dtype": "f16x2-split tensor (f32 accumulate) + f64 NUTS state", "data": "synthetic",
            "config": {"workload": f"Bayesian {args.glm_family} regression N={N}, d={d}, NUTS(beta), {C} chains/GPU (configs[3])", "chains_per_gpu": C,
                       "iters": iters, "burnin": iters // 2, "gradient_passes": int(ticks), "tick_ms": kms / max(ticks, 1), "pass_ms": pass_ms,
                       "useful_chain_gradients": int(w1 - w0), "pass_chain_slots": int(s1 - s0), "useful_fraction": (w1 - w0) / max(1, s1 - s0),
                       "useful_fraction_without_compaction": (w1 - w0) / max(1, ticks * C),
                       "nuts_depth_cap_hits_per_chain_iter": (cap1 - cap0) / (C * iters),
                       "l2": "X (449 MB packed) exceeds L2; every pass streams it from HBM",
                       "psrf_max": float(np.nanmax(psrf[:, 0])), "posterior_mean_abs_err_vs_truth": float(np.mean(np.abs(summ[:, 0] - beta_true)))},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "effective_peak_note": "operands are split fp16 pairs: 3 tensor products per algorithmic product, so the reachable fraction is 1/3",
                         "traffic": GLM_KERNEL_DRAM_BYTES.get(C), "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one glm_tc_kernel launch (profiles/); null for chain counts without a capture",
                         "kernel": "glm_tc_kernel", "kernel_ms": pass_ms, "algo_flops_per_pass": flops,
                         "peak_source": ctx.peak_src + " (sustained bf16; fp16 runs at the same rate)",
                         "fp64_reference_kernel_ms_512_chains": ref_ms, "max_rel_err_logf_vs_fp64": err_lp, "max_rel_err_grad_vs_fp64": err_g},
            "e2e": {"value": ctx.world * C * iters / dt, "unit": "chain-iterations/s", "h2d_bytes_per_step": int(C * d * 8), "d2h_bytes_per_step": int(C * d * 8 + d * 7 * 8),
                    "ms_per_step": 1e3 * dt},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        # CPU baseline at reduced size, extrapolated linearly in N: analytic gradient (engine mode) and the reference's own
        # forward differences (d + 2 density evaluations per gradient, src/model/simulation.jl:47-51)
        cores = os.cpu_count() or 1
        base = {}
        for mode, Ns, its in [] if args.no_cpu_baseline else (("analytic", 100000, 10), ("forward", 10000, 6)):
            Xs, ys = X[:Ns], y[:Ns]
            blocks = [dict(kind="nuts", nodes=[0], grad=mode)]
            v, secs = oracle_rate("glm", blocks, 0.1 * np.random.default_rng(9).standard_normal((cores, d)), cores, its, its // 2, 1, cores, glm=(Xs, ys))
            base[mode] = {"value_at_sample_N": v, "sample_N": Ns, "seconds": secs, "value_extrapolated_to_N": v * Ns / N}
        if base:
            out["cpu_baseline"] = {"value": base["analytic"]["value_extrapolated_to_N"], "unit": "chain-iterations/s", "cores": cores, "kind": "port",
                                   "sample": f"{cores} chains, NUTS, N = {base['analytic']['sample_N']} rows (analytic gradient) and N = {base['forward']['sample_N']} (forward differences, "
                                             f"as the reference computes gradients), extrapolated linearly in N to {N}: an extrapolation, not a measurement at full size",
                                   "reference_fd_gradient_value": base["forward"]["value_extrapolated_to_N"], "detail": base}
    eng.close()
    return out


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
    ap.add_argument("--configs", default="all", help="comma list of line,rats,glm,pumps measured after the headline ('all', 'none')")
    ap.add_argument("--glm-n", type=int, default=1_000_000)
    ap.add_argument("--glm-d", type=int, default=100)
    ap.add_argument("--glm-chains", type=int, default=0, help="chains per GPU (default: 4096 on one GPU, 512 per GPU otherwise = 4096 chains on 8 GPUs)")
    ap.add_argument("--glm-iters", type=int, default=100)
    ap.add_argument("--glm-family", default="logit", choices=["logit", "poisson", "normal"], help="member of the GLM family (tensor-core epilogue)")
    ap.add_argument("--pumps-chains", type=int, default=10_000_000, help="total chains of the pumps configuration (split over the GPUs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    ctx = Ctx(rank, local_rank, world)

    C = args.chains_per_gpu
    iters = args.iters
    burnin = min(BURNIN, iters // 2)
    eng = ctx.engine("seeds", C, rank * C)
    eng.set_scheme(SCHEME)
    inits2 = seeds_inits()
    ctx.peak_fp64 = eng.fp64_peak_tflops()
    l2_flush = ctx.l2_flush

    def step_resident():
        l2_flush.fill_(1)                                  # evict L2 between steps
        eng.set_inits(inits2, jitter_sd=0.1)               # 2 init records, cycled + Philox jitter on the device
        eng.run(iters, burnin=burnin, thin=THIN, store=False, out=False, force_generic=args.force_generic)
        ms = eng.last_kernel_ms()
        psrf, _, _ = eng.diag_global(0.05, True)           # both protocol rounds on the device, NCCL all-reduce when N > 1
        return ms, psrf

    # pinned host buffers for the end-to-end arm
    keep_i, host_inits = ctx.pinned((C, 26)); host_inits[:] = per_chain_inits(C, rank * C)
    keep_s, host_state = ctx.pinned((C, 26))
    import ctypes as Ct

    def step_e2e():
        l2_flush.fill_(1)
        eng.set_inits(host_inits)                          # H2D: C x 26 doubles from pinned memory
        eng.run(iters, burnin=burnin, thin=THIN, store=False, out=False, force_generic=args.force_generic)
        psrf, summ, _ = eng.diag_global(0.05, True)
        it = Ct.c_int64()
        eng._chk(eng.L.mcu_get_state(eng.h, host_state.ctypes.data_as(Ct.POINTER(Ct.c_double)), None, Ct.byref(it)))   # D2H: final states into pinned memory
        return psrf, summ

    # ---- resident arm ---------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank).start()
    ctx.sync_all()
    launches0 = eng.launch_count()
    t0 = time.perf_counter()
    kernel_ms = []
    psrf = None
    for _ in range(args.steps):
        ms, psrf = step_resident()
        kernel_ms.append(ms)
    ctx.sync_all()
    dt = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    clocks = sampler.stop()
    dt, kms = ctx.max_over_ranks(dt, float(np.mean(kernel_ms)))
    value = world * C * iters * args.steps / dt

    # ---- end-to-end arm -------------------------------------------------------------------------
    for _ in range(2):
        step_e2e()
    ctx.sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        psrf_e, summ_e = step_e2e()
    ctx.sync_all()
    (dte,) = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * C * iters * args.steps / dte
    h2d = C * 26 * 8
    d2h = C * 26 * 8 + 5 * (11 + 15) * 8

    # ---- ESS/s from a converged run: the reference's run length for this model (doc/examples/seeds.jl:75) -------------------------------
    ess = None
    if iters == ITERS and not args.force_generic:
        E_IT, E_BURN, E_THIN = 12500, 2500, 2
        l2_flush.fill_(1)
        ctx.sync_all(); t0 = time.perf_counter()
        eng.set_inits(host_inits)
        eng.run(E_IT, burnin=E_BURN, thin=E_THIN, store=False, out=False)
        ess_kms = eng.last_kernel_ms()
        psrf_c, summ_c, _ = eng.diag_global(0.05, True)
        ctx.sync_all()
        (dtc,) = ctx.max_over_ranks(time.perf_counter() - t0)
        kept_c = (E_IT - E_BURN) // E_THIN
        ess_all = (summ_c[:, 1] / summ_c[:, 3]) ** 2
        ess = {"ess_per_sec_min": float(np.nanmin(ess_all) / dtc), "ess_min": float(np.nanmin(ess_all)), "seconds": dtc, "kernel_ms": ess_kms,
               "psrf_max": float(np.max(psrf_c[:, 0])), "psrf": [float(v) for v in psrf_c[:, 0]],
               "iters": E_IT, "burnin": E_BURN, "thin": E_THIN, "chains_total": world * C, "kept_draws": int(world * C * kept_c),
               "names": ["alpha0", "alpha1", "alpha2", "alpha12", "s2"], "ess_per_param": [float(v) for v in ess_all],
               "posterior_mean": [float(v) for v in summ_c[:, 0]], "chain_iters_per_sec": world * C * E_IT / dtc,
               "ess_reference_cap": [float(v) for v in summ_c[:, 4]],
               "definition": "timed leg of its own: the reference's run length for this model (12,500 iterations, burn-in 2,500, thin 2) on every chain, end to end with host "
                             "buffers; ESS = (SD/MCSE_bm)^2 over the kept draws of ALL chains, batch size 100 = the reference's chain-major batches exactly (100 divides the 5,000 kept draws "
                             "of a chain, so no batch straddles chains: mcse.jl:10-19); stats.jl:92 then caps ESS at the draws of ONE chain (5,000 — ess_reference_cap), which is "
                             "meaningless for 10^5 chains, so the uncapped value is what ess_per_sec_min divides by the leg's seconds"}

    # ---- mcmc() that returns ModelChains.value: 4,096 chains x 1,000 kept draws into pinned host memory --------------------------------
    e2e_full = None
    if not args.force_generic:
        from mambacuda.engine import Engine
        CF, KF = 4096, 1000
        ef = Engine("seeds", CF, seed=SEED, chain_offset=rank * CF, device=local_rank)
        ef.set_scheme(SCHEME)
        keep_v, value_host = ctx.pinned((KF, 5, CF), order="F")
        fin = per_chain_inits(CF, rank * CF)
        for rep in range(3):                               # the first call sizes the handle's device buffers; the third is the one reported
            ctx.sync_all(); t0 = time.perf_counter()
            ef.set_inits(fin)
            ef.run(2000, burnin=1000, thin=1, wait=False)  # samples stay on the device ...
            ef.wait()
            ef.samples(into=value_host)                    # ... one transpose kernel + one D2H into pinned memory (no allocation on the call)
            ps_f, su_f, _ = ef.diag_global(0.05, True)
            ctx.sync_all()
            (dtf,) = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e_full = {"value": world * CF * 2000 / dtf, "unit": "chain-iterations/s", "ms_per_step": 1e3 * dtf, "chains_per_gpu": CF, "iters": 2000, "burnin": 1000, "thin": 1,
                    "h2d_bytes_per_step": CF * 26 * 8, "d2h_bytes_per_step": KF * 5 * CF * 8 + 5 * 26 * 8,
                    "returns": "ModelChains.value [1000 x 5 x 4096] (column-major, src/Mamba.jl:172-185) + PSRF + summary",
                    "value_mean_check": float(value_host[:, 0, :].mean())}
        ef.close()

    line = None
    if rank == 0:
        kept = (iters - burnin) // THIN
        line = {
            "metric": "chain_iters_per_sec", "value": value, "unit": "chain-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": SEEDS_WORKLOAD,
                "chains_per_gpu": C, "chains_total": world * C, "iters": iters, "burnin": burnin, "thin": THIN,
                "kernel": "generic" if args.force_generic else "seeds_fast (fused)",
                "parallelism": f"chains sharded over {world} GPU(s), no data-path collective; PSRF / summary over all chains: two NCCL all-reduces of 11p / 15p doubles inside libmambacuda (mcu_diag_global)",
                "l2": "L2 flushed between steps (256 MiB fill); chain state is register/shared-memory resident inside a step",
                "psrf_max": float(np.max(psrf[:, 0])) if psrf is not None else None,
                "psrf_note": "2,000 iterations from the reference's two dispersed initial records (s2 = 0.01 / 1) is the throughput shape of SURVEY.md §8d config 2, not a converged run: "
                             "ESS/s comes from the converged `ess` leg (the reference's 12,500 iterations)",
            },
            "roofline": fp64_roofline(ctx, ALGO_FLOP_PER_CHAIN_ITER, C * iters, kms,
                                      "seeds_fast_kernel" if not args.force_generic else "run_generic_kernel<SeedsModel>",
                                      "CUDA-core FP64 kernel: neither HBM- nor tensor-bound (SURVEY.md §8d); the hbm object shows the memory side is idle",
                                      traffic=SEEDS_KERNEL_DRAM_BYTES.get((C, iters)),
                                      traffic_source="dram__bytes_read.sum + dram__bytes_write.sum of one seeds_fast_kernel launch, ncu --set full (profiles/); null for other chain / iteration counts",
                                      hbm_bytes=C * (kept * 5 * 8)),
            "e2e": {"value": e2e_value, "unit": "chain-iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * dte / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line["roofline"]["algo_flop_per_chain_iter"] = ALGO_FLOP_PER_CHAIN_ITER
        if ess is not None:
            line["ess"] = ess
        if e2e_full is not None:
            line["e2e_full"] = e2e_full
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_s = 128 * cores    # ~10-20 s of CPU work
            v, secs = run_oracle_sample(n_s, ITERS, BURNIN, THIN, cores)
            line["cpu_baseline"] = {"value": v, "unit": "chain-iterations/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_s} chains x {ITERS} iterations of the same model/scheme on {cores} host threads ({secs:.1f} s)"}
    eng.close()

    # ---- the other configurations, measured after the headline regions ------------------------------------------------------------------
    want = [] if args.configs == "none" else (["line", "rats", "glm", "pumps"] if args.configs == "all" else args.configs.split(","))
    cfgs = {}
    cpu = not args.no_cpu_baseline
    if "line" in want:      # configs[0]: 3 chains x 10,000 on every GPU (CPU-scale sanity; the reference's own CPU-runnable case)
        cfgs["line"] = small_config(ctx, "configs[0] tutorial line regression (5 obs), AMWG(beta) + Slice(s2), 3 chains x 10,000", "line_amwg_slice", 3, 10000, 1000, 1,
                                    LINE_FLOP_PER_ITER, False, "run_generic_kernel<LineModel>", 3 if cpu else 0, 10000, jitter=0.0,
                                    note="3 chains cannot fill a GPU: launch latency of one warp; reported for completeness")
    if "rats" in want:      # configs[2]: 65,536 chains per GPU
        cfgs["rats_nuts_slice"] = small_config(ctx, "configs[2] rats hierarchical normal growth model (30 x 5), NUTS(alpha, beta, mu_alpha, mu_beta) + Slice(s2_c, s2_alpha, s2_beta), 65,536 chains x 2,000",
                                               "rats_nuts_slice", 65536, 2000, 1000, 5, RATS_FLOP_PER_LEAPFROG, True, "rats_warp_kernel (one warp per chain)",
                                               4 * (os.cpu_count() or 1) if cpu else 0, 1500,
                                               note="one chain per warp, 30 of 32 lanes own a rat: bound by shuffle / dependent-issue latency of the leapfrog butterflies, not by the FP64 pipe")
        cfgs["rats_slice_amwg"] = small_config(ctx, "configs[2] rats, the reference's own scheme (doc/examples/rats.jl:112-116): Slice + AMWG, 65,536 chains x 2,000",
                                               "rats_slice_amwg", 65536, 2000, 1000, 5, RATS_FAST_FLOP_PER_ITER, False, "rats_fast_kernel (fused)",
                                               16 * (os.cpu_count() or 1) if cpu else 0, 2000, note="RNG / dependent-issue latency bound (as seeds_fast_kernel)")
    if "glm" in want:
        cfgs["glm_nuts"] = glm_config(ctx, args)
    if "pumps" in want:     # configs[4]: 10^7 chains in total, split over the GPUs (strong scaling), on-device PSRF
        cfgs["pumps_gibbs_amwg"] = small_config(ctx, f"configs[4] pumps gamma-Poisson hierarchy, Gibbs(theta) + Gibbs(beta) + AMWG(alpha), {args.pumps_chains:.0e} chains x 2,000, on-device Gelman-Rubin",
                                                "pumps_gibbs_amwg", args.pumps_chains, 2000, 1000, 10, PUMPS_FLOP_PER_ITER, False, "pumps_gibbs_kernel (fused)",
                                                512 * (os.cpu_count() or 1) if cpu else 0, 2000, strong=True, note="issue-slot bound: 11 Gamma variates per iteration")
    if rank == 0:
        if cfgs:
            line["configs"] = cfgs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
