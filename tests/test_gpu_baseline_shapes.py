"""GPU parity at the BASELINE.json shapes (configs[1..4] at their own chain counts / data sizes), through the C ABI.

Philox is keyed by the GLOBAL chain id, so the oracle can re-run any subset of a 10^5 .. 10^6-chain launch: each test runs the
full-size launch on the device and compares scattered chains (final state, tune records with exact integer counters, every kept
sample) with the oracle.  No fraction of diverging chains is tolerated: a chain may part from the oracle only where the oracle
itself took a decision within 1e-9 of its threshold (helpers.audit_divergence).  NUTS schemes amplify rounding chaotically
(leapfrog trajectories, dual averaging), so they are compared one iteration at a time from a common state ("resync"), which is
exactly north_star's per-step parity: same state + same stream => same decisions and the same next state."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

NTHREADS = os.cpu_count() or 4


def scattered_ids(C, n, rng_seed=0, block=96):
    """n global chain ids spread over [0, C): both ends, CTA / wave boundaries and random picks."""
    rng = np.random.default_rng(rng_seed)
    fixed = [0, 1, 31, 32, block - 1, block, 2 * block + 5, C // 2, C - block, C - 2, C - 1]
    ids = set(i for i in fixed if 0 <= i < C)
    while len(ids) < n:
        ids.add(int(rng.integers(0, C)))
    return np.array(sorted(ids), dtype=np.int64)


def kept_iterations(iter0, iters, burnin, thin):
    return np.array([i for i in range(iter0 + 1, iter0 + iters + 1) if i > burnin and (i - burnin) % thin == 0], dtype=np.int64)


def run_full_and_sample(oracle, name, C, iters, burnin, thin, n_sample, seed, jitter_sd, blocks=None, block=96):
    from mambacuda.engine import Engine
    tpl, blocks0, inits = helpers.scheme(name)
    blocks = blocks or blocks0
    eng = Engine(tpl, C, seed=seed)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=jitter_sd)
    out = eng.run(iters, burnin=burnin, thin=thin)
    st, tune, it = eng.get_state()
    assert it == iters
    ids = scattered_ids(C, n_sample, block=block)
    orc = oracle.Oracle(tpl)
    orc.set_scheme([helpers.oracle_block(b) for b in blocks])
    out_o, st_o, tune_o, marg = orc.run(0, inits, iters, burnin=burnin, thin=thin, seed=seed, jitter_sd=jitter_sd, chain_ids=ids,
                                        nthreads=NTHREADS, margins=True)
    g = (out[:, :, ids], st[ids], tune[ids])
    return g, (out_o, st_o, tune_o), marg, eng


def test_seeds_fused_kernel_at_the_baseline_shape(oracle):
    # BASELINE.json configs[1] per GPU: 125,000 chains x 2,000 iterations, burn-in 1,000, thin 10, the bench's scheme and jitter
    C, iters, burnin, thin = 125_000, 2000, 1000, 10
    g, o, marg, eng = run_full_and_sample(oracle, "seeds_amwg", C, iters, burnin, thin, 256, seed=123, jitter_sd=0.1)
    assert eng.launch_count() > 0
    # tune layout: block 0 [m, adapt, sigma[4], accept[4]], block 1 [m, adapt, sigma[21], accept[21]], block 2 [m, adapt, sigma, accept]
    int_cols = [0, 1] + list(range(6, 10)) + [10, 11] + list(range(33, 54)) + [54, 55, 57]
    n_same, ties = helpers.audit_divergence(g, o, marg, kept_iterations(0, iters, burnin, thin), 0, int_tune_cols=int_cols)
    print(f"seeds 125,000 x 2,000: {n_same}/256 sampled chains reproduce the oracle; threshold ties: {ties}")
    assert n_same >= 250          # a tie at 1e-9 is a ~1e-4 event per chain over 1.2e5 decisions


def test_seeds_reference_scheme_amm_at_the_baseline_shape(oracle):
    # the reference's own scheme (AMM + AMWG + AMWG, doc/examples/seeds.jl:69-71) through the fused kernel, 125,000 chains.
    # AMM's proposal at iteration t uses the factor cholfact(Sigma, Val{true}) computed at the end of an EARLIER iteration, and while a
    # chain has hardly moved its running covariance is singular, so the rank / pivot decisions (amm.jl:88-91) are taken on rounding
    # noise in the reference itself.  Hence per-step parity: state AND tune record (Mv, Mvv, SigmaLm) after every single step from the
    # oracle's state; a differing step needs one of the oracle's decisions of THAT step — MH test, pivot choice or rank test, all
    # noted as margins (samplers.hpp) — at rounding distance from its threshold.
    from mambacuda.engine import Engine
    C, seed = 125_000, 7
    tpl, blocks, inits = helpers.scheme("seeds_amm")
    eng = Engine(tpl, C, seed=seed); eng.set_scheme(blocks)
    orc = oracle.Oracle(tpl); orc.set_scheme([helpers.oracle_block(b) for b in blocks])
    ids = scattered_ids(C, 96)
    compared, ties = helpers.resync_audit(eng, orc, ids, inits, 60, 0, seed, 0.1, rtol=1e-7, tie=1e-9, tune_rtol=1e-4)   # SigmaLm: Mvv - Mv Mv' cancels
    print(f"seeds AMM 125,000 chains: {compared} single steps compared, {len(ties)} threshold ties: {ties[:6]}")
    assert len(ties) <= 0.02 * compared


def test_rats_fused_slice_amwg_kernel_at_the_baseline_shape(oracle):
    # BASELINE.json configs[2] chain count with the reference's own scheme (doc/examples/rats.jl:112-116): 65,536 chains x 2,000
    C, iters, burnin, thin = 65_536, 2000, 1000, 5
    g, o, marg, _ = run_full_and_sample(oracle, "rats_slice_amwg", C, iters, burnin, thin, 96, seed=31, jitter_sd=0.05, block=128)
    n_same, ties = helpers.audit_divergence(g, o, marg, kept_iterations(0, iters, burnin, thin), 0)
    print(f"rats Slice+AMWG 65,536 x 2,000: {n_same}/96 reproduce the oracle; ties: {ties}")
    assert n_same >= 90


@pytest.mark.parametrize("name", ["pumps_gibbs_amwg", "pumps_slice"])
def test_pumps_fused_kernels_at_the_baseline_shape(oracle, name):
    # BASELINE.json configs[4]: 10^6 chains x 2,000 iterations (Gibbs + AMWG), and the reference's Slice scheme (pumps.jl:52-53)
    C, iters, burnin, thin = 1_000_000, 2000, 1000, 50
    g, o, marg, _ = run_full_and_sample(oracle, name, C, iters, burnin, thin, 256, seed=17, jitter_sd=0.05, block=128)
    n_same, ties = helpers.audit_divergence(g, o, marg, kept_iterations(0, iters, burnin, thin), 0)
    print(f"{name} 1e6 x 2,000: {n_same}/256 reproduce the oracle; ties: {ties}")
    assert n_same >= 250


def test_rats_warp_nuts_kernel_per_step_parity_at_the_baseline_shape(oracle):
    # BASELINE.json configs[2]: 65,536 chains, NUTS(alpha, beta, mu_alpha, mu_beta) + Slice(s2_c, s2_alpha, s2_beta); 200 adaptive
    # iterations (dual averaging + nutsepsilon at iteration 1), then 40 at the adapted step size; 64 sampled chains re-synchronised
    # with the oracle's recursive buildtree (nuts.jl:139-180) before every step
    from mambacuda.engine import Engine
    C, seed = 65_536, 5
    tpl, blocks, inits = helpers.scheme("rats_nuts_slice")
    eng = Engine(tpl, C, seed=seed); eng.set_scheme(blocks)
    ob = [helpers.oracle_block(b) for b in blocks]; ob[0]["max_depth"] = 10
    orc = oracle.Oracle(tpl); orc.set_scheme(ob)
    ids = scattered_ids(C, 64, block=4)
    compared, ties = helpers.resync_audit(eng, orc, ids, inits, 240, 200, seed, 0.05, rtol=1e-7, tie=1e-7)
    print(f"rats NUTS+Slice 65,536 chains: {compared} single steps compared, {len(ties)} threshold ties: {ties[:6]}")
    assert len(ties) <= 0.002 * compared


def test_glm_tensor_core_gradient_at_the_baseline_shape(oracle):
    # BASELINE.json configs[3]: N = 10^6, d = 100; 512 chains (one GPU's share of 4,096) and 4,096 chains.  The tcgen05 kernel against
    # the FP64 CUDA-core kernel for ALL chains and against the oracle's logpdf! / analytic gradient for 8 of them (north_star: 1e-5)
    from mambacuda.engine import Engine
    N, d = 1_000_000, 100
    rng = np.random.default_rng(1)
    X = rng.standard_normal((N, d)); X[:, 0] = 1.0
    beta_true = np.random.default_rng(2).standard_normal(d) / np.sqrt(d)
    y = (np.random.default_rng(3).uniform(size=N) < 1.0 / (1.0 + np.exp(-(X @ beta_true)))).astype(np.float64)
    orc = oracle.Oracle("glm", glm_d=d)
    orc.set_data("X", X); orc.set_data("y", y)
    orc.set_scheme([dict(kind=4, nodes=[0])])
    for C in (512, 4096):
        eng = Engine("glm", C, seed=1)
        eng.set_data("X", X); eng.set_data("y", y)
        eng.set_scheme([dict(kind="nuts", nodes=[0])])
        # positions a chain visits: around the posterior mode (+- a few posterior sds ~ 2e-3) and the bench's dispersed inits
        beta = beta_true + np.random.default_rng(4).normal(scale=0.01, size=(C, d))
        beta[C // 2:] = 0.1 * np.random.default_rng(5).standard_normal((C - C // 2, d))
        lp0, g0 = eng.glm_gradient(beta, impl=0)
        lp1, g1 = eng.glm_gradient(beta, impl=1)
        gscale = np.abs(g0).max(axis=1, keepdims=True)
        err_lp = np.max(np.abs(lp1 - lp0) / np.abs(lp0)); err_g = np.max(np.abs(g1 - g0) / gscale)
        print(f"GLM N=1e6 d=100 C={C}: tensor-core vs FP64 kernel max rel err logf {err_lp:.2e}, gradient {err_g:.2e}")
        assert err_lp < 1e-5 and err_g < 1e-5
        sel = np.array([0, 1, C // 2 - 1, C // 2, C // 2 + 1, C - 3, C - 2, C - 1])
        lp_o, g_o = orc.gradlogpdf(0, beta[sel], mode=0)
        prior_lp = -0.5 * (d * np.log(2 * np.pi) + d * np.log(1000.0)) - (beta[sel] ** 2).sum(axis=1) / 2000.0   # beta ~ MvNormal(d, sqrt(1000))
        prior_g = -beta[sel] / 1000.0
        np.testing.assert_allclose(lp0[sel] + prior_lp, lp_o, rtol=1e-10)          # FP64 kernel = the oracle's block density (10^6-term sums)
        assert np.max(np.abs((g0[sel] + prior_g) - g_o) / np.abs(g_o).max(axis=1, keepdims=True)) < 1e-9
        np.testing.assert_allclose(lp1[sel] + prior_lp, lp_o, rtol=1e-5)           # tensor-core kernel within north_star's tolerance
        assert np.max(np.abs((g1[sel] + prior_g) - g_o) / np.abs(g_o).max(axis=1, keepdims=True)) < 1e-5
        eng.close()
