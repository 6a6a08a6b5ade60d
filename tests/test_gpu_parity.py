"""GPU parity: the CUDA engine, called through the C ABI, against the CPU oracle on the same
Philox stream.  Tolerances (FP64 paths, north_star): logpdf / gradient
<= 1e-12 relative; trajectories on a shared stream <= 1e-8 relative with identical integer
tune counters (accept counts, m)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

RTOL_LP = 1e-12


def make_pair(oracle, name, n_chains, seed=123, glm=None):
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme(name)
    eng = Engine(tpl, n_chains, seed=seed)
    orc = oracle.Oracle(tpl)
    eng.set_scheme(blocks)
    orc.set_scheme([helpers.oracle_block(b) for b in blocks])
    return eng, orc, inits


def random_states(inits, B, rng, D_pos):
    """Perturbed copies of the init records; positive-support entries stay positive."""
    st = np.repeat(inits, (B + len(inits) - 1) // len(inits), axis=0)[:B].copy()
    noise = rng.normal(scale=0.3, size=st.shape)
    for e in range(st.shape[1]):
        if e in D_pos:
            st[:, e] *= np.exp(noise[:, e])
        else:
            st[:, e] += noise[:, e] * (1.0 + np.abs(st[:, e]) * 0.05)
    return st


POS = {"line": {2}, "seeds": {4}, "rats": {2, 3, 4}, "pumps": set(range(12)), "surgical": {1}, "dyes": {0, 2}, "salm": {0}, "equiv": {0, 1}, "blocker": {0}, "stacks": {4}}


@pytest.mark.parametrize("name", ["line_amwg_slice", "line_nuts_all", "seeds_amwg", "seeds_amm", "rats_slice_amwg",
                                  "rats_nuts_slice", "pumps_slice", "pumps_amwg_nuts", "surgical_nuts_slice", "surgical_amwg", "dyes_nuts_slice", "dyes_mala_slice", "dyes_hmc_slice",
                                  "salm_slice_amwg", "equiv_nuts_slice", "equiv_amwg", "blocker_amwg_slice", "blocker_nuts_slice", "stacks_nuts_slice", "stacks_amwg"])
def test_logpdf_matches_oracle(oracle, name):
    eng, orc, inits = make_pair(oracle, name, 4)
    tpl = helpers.SCHEMES[name][0]
    rng = np.random.default_rng(7)
    st = random_states(inits, 64, rng, POS[tpl])
    for b in range(len(helpers.SCHEMES[name][1])):
        lp_o = orc.logpdf(b, st)
        lp_g = eng.logpdf(b, st)
        np.testing.assert_allclose(lp_g, lp_o, rtol=RTOL_LP, atol=1e-12)
        # explicit block vector on the sampler's scale
        x = np.stack([orc.unlist(b, s) for s in st]) + rng.normal(scale=0.1, size=(64, orc.unlist(b, st[0]).size))
        np.testing.assert_allclose(eng.logpdf(b, st, x), orc.logpdf(b, st, x), rtol=RTOL_LP, atol=1e-12)


def test_logpdf_out_of_support_is_minus_inf(oracle):
    # Slice samples s2 on the constrained scale: negative proposals must give -Inf (distributionstruct.jl:138-140)
    eng, orc, inits = make_pair(oracle, "rats_slice_amwg", 2)
    st = inits.copy()
    x = np.array([[-1.0], [-0.5]])
    assert np.all(np.isneginf(orc.logpdf(0, st, x)))
    assert np.all(np.isneginf(eng.logpdf(0, st, x)))
    eng2, orc2, in2 = make_pair(oracle, "pumps_slice", 2)
    x = np.array([[-0.1, 1.0], [1.0, -2.0]])
    assert np.all(np.isneginf(eng2.logpdf(0, in2, x))) and np.all(np.isneginf(orc2.logpdf(0, in2, x)))


@pytest.mark.parametrize("name", ["line_nuts_all", "line_nuts_slice", "rats_nuts_slice", "pumps_amwg_nuts", "seeds_amwg", "surgical_nuts_slice", "dyes_nuts_slice", "equiv_nuts_slice", "blocker_nuts_slice", "stacks_nuts_slice"])
def test_gradient_matches_oracle(oracle, name):
    eng, orc, inits = make_pair(oracle, name, 4)
    tpl = helpers.SCHEMES[name][0]
    rng = np.random.default_rng(11)
    st = random_states(inits, 32, rng, POS[tpl])
    for b in range(len(helpers.SCHEMES[name][1])):
        k = orc.unlist(b, st[0]).size
        lp_o, g_o = orc.gradlogpdf(b, st, mode=0)
        lp_g, g_g = eng.gradlogpdf(b, st, k, mode="analytic")
        np.testing.assert_allclose(lp_g, lp_o, rtol=RTOL_LP)
        np.testing.assert_allclose(g_g, g_o, rtol=1e-11, atol=1e-11)
        # the reference's finite differences (simulation.jl:47-51): same formula on both sides, but the
        # quotient amplifies libm rounding by 1/h, so the tolerance is the FD noise floor
        _, gf_o = orc.gradlogpdf(b, st, mode=1)
        _, gf_g = eng.gradlogpdf(b, st, k, mode="forward")
        scale = 1.0 + np.abs(g_o)
        noise = 3e-7 * np.maximum(1.0, np.abs(lp_o))[:, None]    # ~ 20 eps |f| / h with h = sqrt(eps)
        assert np.all(np.abs(gf_g - gf_o) / scale < noise + 1e-6)
        assert np.all(np.abs(gf_g - g_g) / scale < noise + 1e-4)  # + O(h f'') truncation
        _, gc_g = eng.gradlogpdf(b, st, k, mode="central")
        assert np.all(np.abs(gc_g - g_g) / scale < 1e-9 * np.maximum(1.0, np.abs(lp_o))[:, None] + 1e-5)


def test_glm_density_and_gradient(oracle):
    from mambacuda.engine import Engine
    X, y, _ = helpers.glm_data(N=300, d=10)
    eng = Engine("glm", 4)
    eng.set_data("X", X); eng.set_data("y", y)
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    orc = oracle.Oracle("glm", glm_d=10)
    orc.set_data("X", X); orc.set_data("y", y)
    orc.set_scheme([dict(kind=4, nodes=[0])])
    st = np.random.default_rng(3).normal(scale=0.5, size=(16, 10))
    lp_o, g_o = orc.gradlogpdf(0, st, mode=0)
    lp_g, g_g = eng.gradlogpdf(0, st, 10)
    np.testing.assert_allclose(lp_g, lp_o, rtol=RTOL_LP)
    np.testing.assert_allclose(g_g, g_o, rtol=1e-11, atol=1e-11)


def run_pair(oracle, name, n_chains, iters, burnin, thin, seed=99, force_generic=True):
    """The same mcmc() on the device and on the oracle; the oracle also reports its smallest decision margin per iteration."""
    eng, orc, inits = make_pair(oracle, name, n_chains, seed=seed)
    eng.set_inits(inits)
    out_g = eng.run(iters, burnin=burnin, thin=thin, force_generic=force_generic)
    st_g, tune_g, it = eng.get_state()
    out_o, st_o, tune_o, marg = orc.run(n_chains, inits, iters, burnin=burnin, thin=thin, seed=seed, nthreads=4, margins=True)
    assert it == iters
    return (out_g, st_g, tune_g), (out_o, st_o, tune_o, marg), eng, orc


def kept_iterations(iters, burnin, thin, iter0=0):
    return [i for i in range(iter0 + 1, iter0 + iters + 1) if i > burnin and (i - burnin) % thin == 0]


def assert_same_run(g, o, iters, burnin, thin, rtol=1e-8, tune_rtol=1e-6):
    """Every chain reproduces the oracle to rounding, or parts from it in an iteration where the oracle's own decision sat within
    1e-9 of its threshold (helpers.audit_divergence) — no fraction of unexplained chains is tolerated.  Returns the tied chains."""
    n_same, ties = helpers.audit_divergence(g, o[:3], o[3], kept_iterations(iters, burnin, thin), 0, rtol=rtol, tune_rtol=tune_rtol)
    assert len(ties) <= max(1, g[1].shape[0] // 16), f"implausibly many threshold ties: {ties}"
    return [t[0] for t in ties]


def assert_same_chains(a, b, skip=(), rtol=1e-9, atol=1e-12):
    """Two device runs of the same scheme (fused kernel / generic kernel / restarted) chain by chain, except chains in `skip`."""
    for c in range(a.shape[2]):
        if c not in skip:
            np.testing.assert_allclose(a[:, :, c], b[:, :, c], rtol=rtol, atol=atol, err_msg=f"chain {c}")


def resync(oracle, name, n_chains, iters, burnin, seed, rtol, tie=helpers.TIE, force_generic=True, run_kw=None):
    """Gradient-based samplers (NUTS, HMC, MALA) amplify rounding along a trajectory, and dual averaging amplifies it from one
    iteration to the next, so whole runs cannot be compared to a fixed tolerance; every single step from a common state can."""
    eng, orc, inits = make_pair(oracle, name, n_chains, seed=seed)
    tpl, blocks, _ = helpers.scheme(name)
    ob = [helpers.oracle_block(b) for b in blocks]
    for b in ob:
        if b["kind"] == "nuts":
            b["max_depth"] = 10           # the device stops doubling after 10 doublings (documented deviation)
    orc.set_scheme(ob)
    kw = dict(force_generic=force_generic); kw.update(run_kw or {})
    compared, ties = helpers.resync_audit(eng, orc, np.arange(n_chains), inits, iters, burnin, seed, 0.0, rtol, tie, run_kw=kw, nthreads=4)
    assert len(ties) <= max(1, compared // 200), f"implausibly many threshold ties: {ties}"
    _, tune, _ = eng.get_state()
    return eng, tune


@pytest.mark.parametrize("name,iters,burnin,thin", [
    ("line_amwg_slice", 600, 200, 2),
    ("seeds_amwg", 300, 150, 3),
    ("seeds_amm", 300, 150, 3),
    ("rats_slice_amwg", 200, 100, 2),
    ("pumps_slice", 300, 100, 2),
    ("pumps_gibbs_amwg", 300, 100, 2),
    ("line_gibbs", 500, 100, 1),
    ("surgical_amwg", 300, 150, 2),
    ("salm_slice_amwg", 300, 150, 2),
    ("equiv_amwg", 300, 150, 2),
    ("blocker_amwg_slice", 300, 150, 2),
    ("stacks_amwg", 300, 150, 2),
    ("dyes_rwm_slice", 300, 0, 1),
    ("line_rwm", 500, 0, 1),
    ("line_rwm_unif", 500, 0, 1),
    ("line_rwm_tri", 500, 0, 1),
    ("line_rwm_cos", 500, 0, 1),
    ("line_rwm_epa", 500, 0, 1),
    ("line_rwm_biw", 500, 0, 1),
    ("line_rwm_trw", 500, 0, 1),
    ("line_slice_uni", 400, 100, 1),
    ("line_amm", 500, 250, 1),
])
def test_trajectories_match_oracle(oracle, name, iters, burnin, thin):
    g, o, _, _ = run_pair(oracle, name, 16, iters, burnin, thin)
    # AMM's SigmaLm comes from Mvv - Mv Mv' (cancellation): its entries agree to fewer digits than the chain does
    assert_same_run(g, o, iters, burnin, thin, tune_rtol=1e-4 if "amm" in name else 1e-6)


@pytest.mark.parametrize("name,iters", [("line_hmc", 120), ("line_hmc_sigma", 120), ("dyes_hmc_slice", 60), ("line_mala", 120), ("line_mala_sigma", 120)])
def test_hmc_and_mala_steps_match_oracle(oracle, name, iters):
    resync(oracle, name, 16, iters, 0, seed=99, rtol=1e-8)


def test_mala_is_statistically_equivalent_to_the_published_posterior(oracle):
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("line_mala")
    eng = Engine(tpl, 256, seed=2)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(6000, burnin=1000, thin=1, store=False, out=False, force_generic=True)
    summ = eng.summary_streaming()
    assert abs(summ[0, 0] - 0.5971) < 0.05 and abs(summ[1, 0] - 0.8017) < 0.02   # doc/tutorial.rst:432-436


@pytest.mark.parametrize("name", ["line_nuts_slice", "line_nuts_all", "rats_nuts_slice", "pumps_amwg_nuts", "surgical_nuts_slice", "equiv_nuts_slice"])
def test_nuts_adaptive_steps_match_oracle(oracle, name):
    # The oracle runs the reference's recursive buildtree (nuts.jl:139-180), the device the unrolled leaf-by-leaf form.  Every step of
    # 30 adaptive (nutsepsilon at iteration 1, dual averaging) + 10 non-adaptive iterations of 32 chains, each from the oracle's state
    eng, tune = resync(oracle, name, 32, 40, 30, seed=5, rtol=1e-7, tie=1e-7)
    assert tune[:, -1].max() >= 4          # nalpha of the last doubling: trees deeper than one doubling were built


@pytest.mark.parametrize("name,iters", [("line_nuts_slice", 100), ("line_nuts_all", 50), ("rats_nuts_slice", 40), ("pumps_amwg_nuts", 100)])
def test_nuts_fixed_stepsize_steps_match_oracle(oracle, name, iters):
    # burnin = 0: model-based NUTS adapts only while iter <= burnin (nuts.jl:52), epsilon stays at nutsepsilon()
    resync(oracle, name, 16, iters, 0, seed=6, rtol=1e-7, tie=1e-7)


def test_seeds_fast_kernel_with_amm_block_matches_oracle_and_generic(oracle):
    # the reference's own seeds scheme (AMM + AMWG + AMWG, doc/examples/seeds.jl:69-71) through the fused kernel
    g, o, eng, _ = run_pair(oracle, "seeds_amm", 32, 300, 150, 3, force_generic=False)
    tied = assert_same_run(g, o, 300, 150, 3, tune_rtol=1e-4)
    tpl, blocks, inits = helpers.scheme("seeds_amm")
    e2 = Engine_(tpl, 32, blocks, inits, seed=99)
    out_gen = e2.run(300, burnin=150, thin=3, force_generic=True)
    assert_same_chains(g[0], out_gen, skip=tied, rtol=1e-8, atol=1e-10)
    # restart through the fused kernel
    e3 = Engine_(tpl, 8, blocks, inits, seed=5); full = e3.run(90, burnin=30, thin=3)
    e4 = Engine_(tpl, 8, blocks, inits, seed=5); a = e4.run(45, burnin=30, thin=3); b = e4.run(45, burnin=30, thin=3)
    np.testing.assert_allclose(np.concatenate([a, b], axis=0), full, rtol=1e-12)


# ---- fused pumps Slice kernel (mamba.jl_b200/csrc/pumps_fast.cu) ---------------------------------------------
def test_pumps_fast_kernel_matches_oracle_generic_and_published_table(oracle):
    g, o, _, _ = run_pair(oracle, "pumps_slice", 32, 300, 100, 2, force_generic=False)
    tied = assert_same_run(g, o, 300, 100, 2)
    tpl, blocks, inits = helpers.scheme("pumps_slice")
    e2 = Engine_(tpl, 32, blocks, inits, seed=99)
    out_gen = e2.run(300, burnin=100, thin=2, force_generic=True)
    assert_same_chains(g[0], out_gen, skip=tied, rtol=1e-8, atol=1e-10)
    from mambacuda.engine import Engine
    eng = Engine(tpl, 2048, seed=8); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(3000, burnin=1500, thin=1, store=False, out=False)
    summ = eng.summary_streaming()
    ref = np.array([0.6968, 0.9304, 0.0599, 0.1013, 0.0891, 0.1153, 0.5997, 0.6097, 0.8677, 0.8545, 1.5572, 1.9848])   # doc/examples/pumps.rst:43-56
    np.testing.assert_allclose(summ[:, 0], ref, rtol=0.05)


# ---- fused rats Slice + AMWG kernel (mamba.jl_b200/csrc/rats_fast.cu) -------------------------------------
def test_rats_fast_kernel_trajectories_match_oracle_and_generic(oracle):
    g, o, eng, _ = run_pair(oracle, "rats_slice_amwg", 32, 200, 100, 2, force_generic=False)
    tied = assert_same_run(g, o, 200, 100, 2)
    tpl, blocks, inits = helpers.scheme("rats_slice_amwg")
    eng2 = Engine_(tpl, 32, blocks, inits, seed=99)
    out_gen = eng2.run(200, burnin=100, thin=2, force_generic=True)
    assert_same_chains(g[0], out_gen, skip=tied, rtol=1e-8, atol=1e-10)


def test_rats_fast_kernel_restart_and_posterior(oracle):
    # mcmc(mc, iters) restart through the fused kernel, and the published table doc/examples/rats.rst:42-46
    tpl, blocks, inits = helpers.scheme("rats_slice_amwg")
    e1 = Engine_(tpl, 16, blocks, inits, seed=3); full = e1.run(60, burnin=20, thin=2)
    e2 = Engine_(tpl, 16, blocks, inits, seed=3); a = e2.run(24, burnin=20, thin=2); b = e2.run(36, burnin=20, thin=2)
    np.testing.assert_allclose(np.concatenate([a, b], axis=0), full, rtol=1e-12)
    from mambacuda.engine import Engine
    eng = Engine(tpl, 512, seed=4); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(6000, burnin=3000, thin=1, store=False, out=False)
    summ = eng.summary_streaming()
    ref = np.array([6.1831, 106.626, 37.254]); ref_mcse = np.array([0.0018, 0.0527, 0.234]); ref_sd = np.array([0.108, 3.459, 6.027])
    assert np.all(np.abs(summ[:, 0] - ref) < 3 * np.hypot(ref_mcse, summ[:, 3]) + 0.03 * ref_sd)
    np.testing.assert_allclose(summ[:, 1], ref_sd, rtol=0.1)


def Engine_(tpl, n, blocks, inits, seed):
    from mambacuda.engine import Engine
    e = Engine(tpl, n, seed=seed); e.set_scheme(blocks); e.set_inits(inits)
    return e


# ---- warp-per-chain rats kernel (mamba.jl_b200/csrc/rats_warp.cu) ------------------------------------------
@pytest.mark.parametrize("iters,burnin", [(40, 30), (40, 0)])
def test_rats_warp_kernel_steps_match_oracle(oracle, iters, burnin):
    # the same draws in the same order as the reference's recursion: adaptive (dual averaging + nutsepsilon) and fixed step size
    eng, tune = resync(oracle, "rats_nuts_slice", 64, iters, burnin, seed=5, rtol=1e-7, tie=1e-7, force_generic=False)
    assert tune[:, -1].max() >= 4


def test_rats_warp_kernel_matches_generic_kernel_and_restarts(oracle):
    # warp-per-chain kernel against the one-chain-per-thread kernel, step by step from common states (the two evaluate the same
    # leapfrogs in a different floating-point order), and mcmc(mc, iters) restart (mcmc.jl:3-16) through the warp kernel
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("rats_nuts_slice")
    a = Engine(tpl, 96, seed=21); a.set_scheme(blocks); a.set_inits(inits, jitter_sd=0.05)
    b = Engine(tpl, 96, seed=21); b.set_scheme(blocks); b.set_inits(inits, jitter_sd=0.05)
    n_diff = 0
    for i in range(1, 31):
        a.run(1, burnin=10, thin=1, store=False, out=False, partial=True, force_generic=True)
        b.run(1, burnin=10, thin=1, store=False, out=False, partial=True)
        sa, ta, _ = a.get_state(); sb, tb, _ = b.get_state()
        same = np.array([np.allclose(sa[c], sb[c], rtol=1e-7, atol=1e-9) for c in range(96)])
        n_diff += int((~same).sum())
        b.set_state(sa, ta, i)           # both continue from the generic kernel's state
    assert n_diff <= 3, f"{n_diff} of 2,880 single steps differ between the warp kernel and the generic kernel"
    e1 = Engine(tpl, 64, seed=22); e1.set_scheme(blocks); e1.set_inits(inits, jitter_sd=0.05)
    full = e1.run(30, burnin=10, thin=2)
    e2 = Engine(tpl, 64, seed=22); e2.set_scheme(blocks); e2.set_inits(inits, jitter_sd=0.05)
    p1 = e2.run(12, burnin=10, thin=2); p2 = e2.run(18, burnin=10, thin=2)
    np.testing.assert_array_equal(np.concatenate([p1, p2], axis=0), full)


def test_rats_warp_kernel_posterior_matches_published_table(oracle):
    # doc/examples/rats.rst:42-46 (Slice+AMWG run of the reference): mu_beta 6.1831 (0.108), alpha0 106.626 (3.459), s2_c 37.254 (6.027)
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("rats_nuts_slice")
    eng = Engine(tpl, 512, seed=4)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(1500, burnin=750, thin=1, store=False, out=False)
    summ = eng.summary_streaming()   # rows: mu_beta, alpha0, s2_c ; columns: mean, sd, naive se, mcse, ess
    ref = np.array([6.1831, 106.626, 37.254]); ref_mcse = np.array([0.0018, 0.0527, 0.234]); ref_sd = np.array([0.108, 3.459, 6.027])
    assert np.all(np.abs(summ[:, 0] - ref) < 3 * np.hypot(ref_mcse, summ[:, 3]) + 0.02 * ref_sd)
    np.testing.assert_allclose(summ[:, 1], ref_sd, rtol=0.08)
    assert (eng.gelman(0.05, True)[:, 0] < 1.05).all()


@pytest.mark.parametrize("name", ["dyes_nuts_slice", "dyes_hmc_slice", "dyes_mala_slice"])
def test_dyes_schemes_match_published_table(oracle, name):
    # doc/examples/dyes.rst (scheme 1): theta 1526.72 (24.5), s2_within 2887.6 (1075), mu[1] 1511.48, mu[5] 1578.66, mu[6] 1487.19;
    # the other schemes of the script (MALA, HMC on theta / mu with Sigma = I) target the same posterior
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme(name)
    eng = Engine(tpl, 1024, seed=6)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.01)
    eng.run(5000, burnin=2500, thin=1, store=False, out=False)
    summ = eng.summary_streaming(); names = eng.names(1)
    ref = {"theta": (1526.7186, 0.377, 24.55), "s2_within": (2887.5853, 76.9, 1075.2), "mu[1]": (1511.4798, 0.52, 20.8),
           "mu[5]": (1578.6636, 1.29, 25.5), "mu[6]": (1487.1934, 1.24, 24.7)}
    for nm, (mean, mcse_ref, sd) in ref.items():
        j = names.index(nm)
        assert abs(summ[j, 0] - mean) < 3 * np.hypot(mcse_ref, summ[j, 3]) + 0.05 * sd, (name, nm, summ[j, 0], mean)


def test_surgical_posterior_matches_published_table(oracle):
    # doc/examples/surgical.rst: mu -2.5503 (0.152), pop_mean 0.07306 (0.0101), s2 0.1831 (0.161), p[4] 0.05986, p[8] 0.12230
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("surgical_nuts_slice")
    eng = Engine(tpl, 1024, seed=6)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(3000, burnin=1500, thin=1, store=False, out=False)
    summ = eng.summary_streaming()
    names = eng.names(1)
    ref = {"mu": (-2.550263247, 0.0035, 0.1518), "pop_mean": (0.073062651, 0.00023, 0.0101), "s2": (0.183080212, 0.0063, 0.1612),
           "p[4]": (0.059863573, 0.00033, 0.0082), "p[8]": (0.122296440, 0.00086, 0.0233)}
    for nm, (mean, mcse_ref, sd) in ref.items():
        j = names.index(nm)
        assert abs(summ[j, 0] - mean) < 3 * np.hypot(mcse_ref, summ[j, 3]) + 0.02 * sd, (nm, summ[j, 0], mean)
    # pop_mean and p[i] are Logical columns in (0, 1): link(c) takes their logit (chains.jl:241-243)
    assert (eng.gelman(0.05, False)[:, 0] < 1.05).all()
    assert (eng.gelman(0.05, True)[:, 0] < 1.05).all()


def test_gelman_logit_link_for_logical_columns_in_the_unit_interval(oracle):
    # link(c::ModelChains) (modelchains.jl:57-76): mu identity, s2 log, pop_mean and p[1..12] Logical with every value in (0, 1) -> logit
    g, o, eng, orc = run_pair(oracle, "surgical_amwg", 8, 700, 100, 2)
    out_g = g[0]
    codes = eng.link_codes(True)
    names = eng.names(1)
    assert codes[names.index("s2")] == 1 and codes[names.index("mu")] == 0 and codes[names.index("pop_mean")] == 2 and codes[names.index("p[3]")] == 2
    psrf_g = eng.gelman(0.05, True)
    psrf_o = oracle.gelmandiag(out_g, 0.05, [{0: 0, 1: 1, 2: -1}[int(c)] for c in codes])
    np.testing.assert_allclose(psrf_g, psrf_o, rtol=1e-7)
    # the same through the host-array entry point on the materialised draws
    from mambacuda import api
    psrf_h = api._chains_gelman(out_g, 0.05, codes, False)
    np.testing.assert_allclose(psrf_h, psrf_o, rtol=1e-7)


@pytest.mark.parametrize("tpl_scheme", ["line_amwg_slice", "seeds_amwg", "rats_slice_amwg", "pumps_slice", "surgical_amwg", "dyes_nuts_slice", "salm_slice_amwg", "equiv_amwg", "blocker_amwg_slice", "stacks_amwg"])
def test_node_logpdf_matches_oracle(oracle, tpl_scheme):
    # logpdf(mc, nodekeys) (modelstats.jl:16-58): observed nodes only (the deviance of dic), every stochastic node, one parameter node
    eng, orc, inits = make_pair(oracle, tpl_scheme, 4)
    tpl = helpers.scheme(tpl_scheme)[0]
    st = random_states(inits, 24, np.random.default_rng(8), POS[tpl])
    nn, nf = eng.factor_counts()
    assert nf == nn + (2 if tpl == "blocker" else 1)   # blocker has two observed nodes (rc, rt)
    for mask in (1 << nn, (1 << nf) - 1, 1, 1 << (nn - 1), 1 << (nf - 1)):
        np.testing.assert_allclose(eng.logpdf_nodes(mask, st), orc.logpdf_nodes(mask, st), rtol=RTOL_LP, atol=1e-12)
    # the block density is the sum of the block's own node densities and its targets': for a block holding every parameter node
    # this is the joint (simulation.jl:77-90)
    joint = eng.logpdf_nodes((1 << nf) - 1, st)
    assert np.isfinite(joint).all()


@pytest.mark.parametrize("tpl_scheme", ["line_amwg_slice", "seeds_amwg", "rats_slice_amwg", "pumps_slice", "surgical_amwg", "dyes_nuts_slice", "salm_slice_amwg", "equiv_amwg", "blocker_amwg_slice", "stacks_amwg"])
def test_predict_matches_oracle(oracle, tpl_scheme):
    # predict(mc) (modelstats.jl:63-96): rand of the observed node at each state, same Philox stream on both sides
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme(tpl_scheme)
    eng = Engine(tpl, 4, seed=31); eng.set_scheme(blocks)
    orc = oracle.Oracle(tpl); orc.set_scheme([helpers.oracle_block(b) for b in blocks])
    st = random_states(inits, 40, np.random.default_rng(9), POS[tpl])
    g = eng.predict(st, stream_id=5); o = orc.predict(st, 31, stream_id=5)
    assert g.shape == o.shape and g.shape[0] == 40
    if tpl in ("seeds", "pumps", "surgical", "salm", "blocker"):    # counts: identical integers
        assert (g == np.round(g)).all() and (g >= 0).all()
        assert (g == o).mean() > 0.999           # a uniform within rounding of a CDF step may fall on either side
    else:
        np.testing.assert_allclose(g, o, rtol=1e-10, atol=1e-10)
    assert not np.array_equal(g, eng.predict(st, stream_id=6))


def test_predict_glm_families():
    # the three GLM families draw Bernoulli / Poisson / Normal observations with the right means
    from mambacuda.engine import Engine
    rng = np.random.default_rng(3)
    X = np.column_stack([np.ones(400), rng.normal(size=(400, 2))]); beta = np.array([0.3, -0.5, 0.8])
    eta = X @ beta
    for family, mean in ((0, 1 / (1 + np.exp(-eta))), (1, np.exp(eta)), (2, eta)):
        eng = Engine("glm", 2, seed=1); eng.set_data("X", X); eng.set_data("y", np.zeros(400))
        eng.set_data("family", np.array([float(family)])); eng.set_data("sigma", np.array([0.5]))
        d = eng.predict(np.tile(beta, (2000, 1)))
        assert d.shape == (2000, 400)
        sd = np.sqrt(mean * (1 - mean)) if family == 0 else np.sqrt(mean) if family == 1 else np.full(400, 0.5)
        assert (np.abs(d.mean(axis=0) - mean) < 5 * sd / np.sqrt(2000)).all()
        if family == 2:
            np.testing.assert_allclose(d.std(axis=0), 0.5, rtol=0.1)


def test_equiv_posterior_matches_published_table(oracle):
    # doc/examples/equiv.rst:43-50 (NUTS(delta) + Slice([mu, phi, pi]) + Slice([s2_1, s2_2], Univariate)): mean (MCSE, SD)
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("equiv_nuts_slice")
    eng = Engine(tpl, 1024, seed=12)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(5000, burnin=2500, thin=1, store=False, out=False)
    summ = eng.summary_streaming(); names = eng.names(1)
    ref = {"s2_2": (0.0173121833, 0.0007329722, 0.014549568), "s2_1": (0.0184397014, 0.0005689492, 0.013837972),
           "pi": (-0.1874240524, 0.0032257037, 0.086420302), "phi": (-0.0035569545, 0.0035141650, 0.087590520),
           "theta": (1.0002921934, 0.0036227671, 0.088250458), "equiv": (0.9751, 0.0036666529, 0.155828169), "mu": (1.4387396416, 0.0013735876, 0.042269208)}
    for nm, (mean, mcse_ref, sd) in ref.items():
        j = names.index(nm)
        assert abs(summ[j, 0] - mean) < 3 * np.hypot(mcse_ref, summ[j, 3]) + 0.02 * sd, (nm, summ[j, 0], mean)
    psrf = eng.gelman(0.05, True)      # theta is a Logical column > 0 (log link); equiv is a 0/1 indicator (identity: min > 0 fails)
    codes = eng.link_codes(True)
    assert codes[names.index("theta")] == 1 and codes[names.index("equiv")] == 0
    assert (psrf[2:, 0] < 1.03).all() and (psrf[:2, 0] < 1.15).all()   # the variances mix slowly under the constrained-scale slice (published ESS 394 / 592 of 10,000)


def test_blocker_posterior_dic_and_predict(oracle):
    # doc/examples/blocker.rst:46-49 through the API mirror; dic / predict with two observed nodes
    from mambacuda import api
    m = api.Model("blocker")
    api.setsamplers(m, [api.AMWG("mu", 0.1), api.AMWG(["delta", "delta_new"], 0.1), api.Slice(["d", "s2"], 1.0)])        # blocker.jl:84-86
    inits = [dict(d=0.0, delta_new=0.0, s2=1.0, mu=np.zeros(22), delta=np.zeros(22)), dict(d=2.0, delta_new=2.0, s2=10.0, mu=np.full(22, 2.0), delta=np.full(22, 2.0))]
    sim = api.mcmc(m, {}, inits * 16, 10000, burnin=2500, thin=2, chains=32)
    assert sim.names == ["s2", "d", "delta_new"]
    ss, _, _ = api.summarystats(sim)
    ref = {"s2": (0.01822186, 0.0014150714, 0.021121265), "d": (-0.25563567, 0.0040205781, 0.061841945), "delta_new": (-0.25005767, 0.0050219145, 0.150325282)}
    for j, nm in enumerate(sim.names):
        assert abs(ss[j, 0] - ref[nm][0]) < 3 * np.hypot(ref[nm][1], ss[j, 3]) + 0.02 * ref[nm][2], (nm, ss[j], ref[nm])
    with pytest.raises(api.ArgumentError, match="chain values are missing for nodes : mu, delta"):     # the arms depend on unmonitored nodes
        api.dic(sim)
    with pytest.raises(api.ArgumentError, match="chain values are missing for nodes"):
        api.predict(sim, "rt")
    # with every node available (a ModelChains over the full state) both run: deviance = -2 (logpdf(rc) + logpdf(rt))
    eng = sim.engine
    vals, _, _ = eng.get_state()
    full = api.ModelChains(vals.T[None, :, :].copy(), sim.model, engine=None, names=eng.names(0), start=10000, thin=2)
    full.engine = eng; full._streaming = False
    d, rows, cols = api.dic(full)
    lp = api.logpdf(full, ["rc", "rt"]).value[0, 0, :]
    assert d.shape == (2, 2) and np.isfinite(d).all()
    np.testing.assert_allclose(-2.0 * lp.mean() , d[0, 0] - d[0, 1], rtol=1e-12)                        # DIC - pD = mean deviance
    pp = api.predict(full)
    assert pp.names[0] == "rc[1]" and pp.names[22] == "rt[1]" and pp.value.shape == (1, 44, 32)
    assert api.predict(full, "rt").names == [f"rt[{i + 1}]" for i in range(22)]
    nt = np.array([38, 114, 69, 1533, 355, 59, 945, 632, 278, 1916, 873, 263, 291, 858, 154, 207, 251, 151, 174, 209, 391, 680])
    assert (pp.value[0, 22:, :] <= nt[:, None]).all() and (pp.value >= 0).all()


def test_salm_posterior(oracle):
    # doc/examples/salm.rst:43-47.  The reference's own run mixes slowly in the multivariate slice block (ESS 93-185), so its table is
    # matched within 0.75 of the published SD; the long-run values this scheme converges to (CPU oracle, 8 x 100,000 iterations:
    # alpha 2.176, beta 0.311, gamma -9.7e-4, s2 0.074) are matched within Monte Carlo error.
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("salm_slice_amwg")
    eng = Engine(tpl, 2048, seed=13)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.02)
    eng.run(12000, burnin=4000, thin=1, store=False, out=False)
    summ = eng.summary_streaming(); names = eng.names(1)
    pub = {"s2": (0.0690769709, 0.04304237136), "gamma": (-0.0011250515, 0.00034536546), "beta": (0.3543443166, 0.07160779229), "alpha": (2.0100584321, 0.26156942610)}
    longrun = {"s2": (0.0740, 0.047), "gamma": (-0.000973, 0.00045), "beta": (0.3106, 0.1031), "alpha": (2.176, 0.379)}
    for nm in pub:
        j = names.index(nm)
        assert abs(summ[j, 0] - pub[nm][0]) < 0.75 * pub[nm][1], (nm, summ[j, 0])
        assert abs(summ[j, 0] - longrun[nm][0]) < 0.06 * longrun[nm][1] + 3 * summ[j, 3], (nm, summ[j, 0], summ[j, 3])
    assert (eng.gelman(0.05, True)[:, 0] < 1.1).all()


def test_gibbs_is_rejected_where_no_conjugate_form_is_registered(oracle):
    from mambacuda.engine import Engine
    from mambacuda._lib import MambaCudaError
    eng = Engine("pumps", 2)
    with pytest.raises(MambaCudaError):
        eng.set_scheme([dict(kind="gibbs", nodes=[0])])           # alpha has no conjugate full conditional
    eng2 = Engine("seeds", 2)
    with pytest.raises(MambaCudaError):
        eng2.set_scheme([dict(kind="gibbs", nodes=[5])])


def test_pumps_gibbs_kernel_matches_oracle_and_generic_and_restarts(oracle):
    # fused [Gibbs(theta), Gibbs(beta), AMWG(alpha)] kernel (pumps_fast.cu): the Gibbs draws are the generic kernel's to the last bit or two (same accept decisions; reciprocals by Newton steps), the AMWG
    # target is evaluated on sufficient statistics
    g, o, eng, _ = run_pair(oracle, "pumps_gibbs_amwg", 64, 300, 100, 2, force_generic=False)
    tied = assert_same_run(g, o, 300, 100, 2)
    tpl, blocks, inits = helpers.scheme("pumps_gibbs_amwg")
    gen = Engine_(tpl, 64, blocks, inits, seed=99)
    out_gen = gen.run(300, burnin=100, thin=2, force_generic=True)
    assert_same_chains(g[0], out_gen, skip=tied)
    two = Engine_(tpl, 64, blocks, inits, seed=99)        # mcmc(mc, iters): 130 + 170 iterations in two calls
    a = two.run(130, burnin=100, thin=2); b = two.run(170, burnin=100, thin=2)
    np.testing.assert_array_equal(np.concatenate([a, b], axis=0), g[0])
    st2, tune2, _ = two.get_state()
    np.testing.assert_array_equal(st2, g[1]); np.testing.assert_array_equal(tune2, g[2])


def test_pumps_gibbs_amwg_posterior_and_psrf(oracle):
    # doc/examples/pumps.rst:43-56: beta 0.9304, alpha 0.6968, theta[1] 0.0599, theta[10] 1.9848
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("pumps_gibbs_amwg")
    eng = Engine(tpl, 2048, seed=8)
    eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(2000, burnin=1000, thin=1, store=False, out=False)
    summ = eng.summary_streaming()
    ref = np.array([0.6968, 0.9304, 0.0599, 0.1013, 0.0891, 0.1153, 0.5997, 0.6097, 0.8677, 0.8545, 1.5572, 1.9848])
    # theta[7], theta[8] share y = 1, t = 1.05; the published single run has MCSE 0.029 on them (0.8677 vs 0.8545): 5 % covers 3 MCSE
    np.testing.assert_allclose(summ[:, 0], ref, rtol=0.05)
    assert abs(summ[8, 0] - summ[9, 0]) < 0.01
    assert (eng.gelman(0.05, True)[:, 0] < 1.02).all()


def test_nuts_fd_gradient_statistically_equivalent(oracle):
    # reference mode (forward differences) and engine mode (analytic) target the same posterior
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("line_nuts_fd")
    eng = Engine(tpl, 64, seed=3)
    eng.set_scheme(blocks); eng.set_inits(inits)
    out = eng.run(1500, burnin=500, thin=1, force_generic=True)
    mean_b1 = out[:, 1, :].mean()
    assert abs(mean_b1 - 0.80) < 0.05   # doc/tutorial.rst:432-436: beta[2] = 0.8017


def test_external_stream_matches_oracle(oracle):
    # the north_star "shim": both sides consume the same host-supplied uniform stream
    name = "line_amwg_slice"
    eng, orc, inits = make_pair(oracle, name, 8)
    u = np.random.default_rng(42).uniform(size=(8, 4000))
    eng.set_external_stream(u)
    eng.set_inits(inits)
    out_g = eng.run(200, burnin=50, thin=1, force_generic=True)
    st_g, tune_g, _ = eng.get_state()
    out_o, st_o, tune_o, marg = orc.run(8, inits, 200, burnin=50, thin=1, ext_u=u, margins=True)
    assert_same_run((out_g, st_g, tune_g), (out_o, st_o, tune_o, marg), 200, 50, 1)


def test_restart_continues_the_chain(oracle):
    # mcmc(mc, iters): src/model/mcmc.jl:3-16 — two calls == one call; also via get_state/set_state
    from mambacuda.engine import Engine
    name = "seeds_amwg"
    tpl, blocks, inits = helpers.scheme(name)
    a = Engine(tpl, 8, seed=4); a.set_scheme(blocks); a.set_inits(inits)
    full = a.run(240, burnin=100, thin=2, force_generic=True)
    b = Engine(tpl, 8, seed=4); b.set_scheme(blocks); b.set_inits(inits)
    p1 = b.run(130, burnin=100, thin=2, force_generic=True)
    vals, tune, it = b.get_state()
    c = Engine(tpl, 8, seed=4); c.set_scheme(blocks); c.set_state(vals, tune, it)
    p2 = c.run(110, burnin=100, thin=2, force_generic=True)
    assert p1.shape[0] + p2.shape[0] == full.shape[0]
    np.testing.assert_array_equal(np.concatenate([p1, p2], axis=0), full)


def test_sharding_does_not_change_chains(oracle):
    # Philox key = global chain id: chains 8..15 of a 16-chain handle == an 8-chain handle at offset 8
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("seeds_amwg")
    a = Engine(tpl, 16, seed=8); a.set_scheme(blocks); a.set_inits(inits, jitter_sd=0.1)
    b = Engine(tpl, 8, seed=8, chain_offset=8); b.set_scheme(blocks); b.set_inits(inits, jitter_sd=0.1)
    oa = a.run(100, burnin=50, thin=5, force_generic=True); ob = b.run(100, burnin=50, thin=5, force_generic=True)
    np.testing.assert_array_equal(oa[:, :, 8:], ob)


def test_jitter_matches_oracle(oracle):
    eng, orc, inits = make_pair(oracle, "seeds_amwg", 8, seed=77)
    eng.set_inits(inits, jitter_sd=0.2)
    st_g, _, _ = eng.get_state()
    _, st_o, _ = orc.run(8, inits, 1, burnin=0, thin=1, seed=77, jitter_sd=0.2)
    # oracle state is after one iteration; compare the initial draw through b (block 2 updates s2 only after)
    assert np.isfinite(st_g).all() and (st_g[:, 4] > 0).all()
    out_g = eng.run(50, burnin=0, thin=1, force_generic=True)
    out_o, _, _ = orc.run(8, inits, 50, burnin=0, thin=1, seed=77, jitter_sd=0.2)
    np.testing.assert_allclose(out_g, out_o, rtol=1e-8, atol=1e-10)


def test_gelman_and_summary_match_oracle(oracle):
    g, o, eng, orc = run_pair(oracle, "rats_slice_amwg", 8, 1200, 200, 2)
    out_g = g[0]
    # exact reference semantics on the materialised samples
    ss_g = eng.summarystats("bm", 100)
    ss_o = oracle.summarystats(out_g, 0, 100)
    np.testing.assert_allclose(ss_g, ss_o, rtol=1e-9)
    ss_i = eng.summarystats("imse")
    np.testing.assert_allclose(ss_i, oracle.summarystats(out_g, 1), rtol=1e-8)
    # streaming device reductions: kept = 500 per chain is a multiple of the batch size, so batches coincide
    st = eng.summary_streaming()
    np.testing.assert_allclose(st, ss_o, rtol=1e-8)
    for transform in (False, True):
        psrf_g = eng.gelman(0.05, transform)
        linkcode = [0, -1, 1] if transform else None   # mu_beta (identity), alpha0 (Logical: heuristic), s2_c (log)
        psrf_o = oracle.gelmandiag(out_g, 0.05, linkcode)
        np.testing.assert_allclose(psrf_g, psrf_o, rtol=1e-7)


def test_error_codes(oracle):
    from mambacuda.engine import Engine, MambaCudaError
    tpl, blocks, inits = helpers.scheme("line_amwg_slice")
    e = Engine(tpl, 2); e.set_scheme(blocks)
    with pytest.raises(MambaCudaError, match="initial values must be set"):
        e.run(10)
    e.set_inits(inits)
    with pytest.raises(MambaCudaError, match="burnin is greater than or equal to iters"):   # mcmc.jl:22-23
        e.run(10, burnin=10)
    with pytest.raises(MambaCudaError, match="length\\(scale\\) differs from variate length 2"):   # amwg.jl:38-43
        e.set_scheme([dict(kind="amwg", nodes=[0], scale=[1.0, 2.0, 3.0])])
    one = Engine(tpl, 1); one.set_scheme(blocks); one.set_inits(inits); one.run(20, burnin=5)
    with pytest.raises(MambaCudaError, match="less than 2 chains"):   # gelmandiag.jl:6-7
        one.gelman()


# ---- the fused seeds/AMWG kernel (mamba.jl_b200/csrc/seeds_fast.cu) ------------------------------------
def test_seeds_fast_matches_oracle_and_generic(oracle):
    # dispatch happens inside mcu_run when the scheme is [AMWG(alphas), AMWG(b), AMWG(s2)] on the seeds template
    g, o, eng, _ = run_pair(oracle, "seeds_amwg", 64, 400, 200, 4, seed=21, force_generic=False)
    tied = assert_same_run(g, o, 400, 200, 4)
    assert (g[2][:, 0] == 400).all() and (o[2][:, 0] == 400).all()           # m of block 0 (adapt=:all)
    keep = [c for c in range(64) if c not in tied]
    np.testing.assert_array_equal(g[2][keep, 6:10], o[2][keep, 6:10])        # alpha accept counters are integers
    g2, o2, _, _ = run_pair(oracle, "seeds_amwg", 64, 400, 200, 4, seed=21, force_generic=True)
    tied2 = assert_same_run(g2, o2, 400, 200, 4)
    assert_same_chains(g[0], g2[0], skip=set(tied) | set(tied2))


def test_seeds_fast_burnin_adaptation_and_restart(oracle):
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("seeds_amwg")
    blocks = [dict(b, adapt="burnin") for b in blocks]
    a = Engine(tpl, 32, seed=9); a.set_scheme(blocks); a.set_inits(inits, jitter_sd=0.1)
    full = a.run(300, burnin=120, thin=3)
    b = Engine(tpl, 32, seed=9); b.set_scheme(blocks); b.set_inits(inits, jitter_sd=0.1)
    p1 = b.run(130, burnin=120, thin=3); p2 = b.run(170, burnin=120, thin=3)
    assert p1.shape[0] == 3
    np.testing.assert_array_equal(np.concatenate([p1, p2], axis=0), full)
    orc = oracle.Oracle(tpl); orc.set_scheme([helpers.oracle_block(x) for x in blocks])
    out_o, st_o, tune_o, marg = orc.run(32, inits, 300, burnin=120, thin=3, seed=9, jitter_sd=0.1, nthreads=4, margins=True)
    st, tune, _ = a.get_state()
    assert_same_run((full, st, tune), (out_o, st_o, tune_o, marg), 300, 120, 3)
    assert (tune[:, 0] == 120).all() and (tune[:, 1] == 0).all()             # adaptation stopped after burn-in


def test_seeds_fast_posterior_within_3_mcse_of_reference(oracle):
    # doc/examples/seeds.rst:37-56 (AMM+AMWG+AMWG, 2 x 12,500, burnin 2,500, thin 2): mean [MCSE]
    ref = {"alpha0": (-0.556154341, 0.0101730837), "alpha1": (0.088700176, 0.0128300598), "alpha2": (1.310728093, 0.0153996801),
           "alpha12": (-0.746440855, 0.0251658152), "s2": (0.085705306, 0.0080848189)}
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("seeds_amwg")
    # same inits, iterations, burn-in and thinning as the reference run, 4096 chains instead of 2 (the chain
    # mixes slowly — ESS 145 of 10,000 for s2 in the reference — so a shorter burn-in is visibly biased)
    eng = Engine(tpl, 4096, seed=2024); eng.set_scheme(blocks); eng.set_inits(inits)
    eng.run(12500, burnin=2500, thin=2, store=False, out=False)
    ss = eng.summary_streaming()
    names = eng.names(1)
    for j, nm in enumerate(names):
        mean, mcse_ref = ref[nm]
        tol = 3.0 * np.hypot(mcse_ref, ss[j, 3])
        assert abs(ss[j, 0] - mean) < tol, (nm, ss[j, 0], mean, tol)
    psrf = eng.gelman(0.05, True)
    assert (psrf[:, 0] < 1.2).all()   # doc/tutorial.rst:321 rule of thumb


# ---- GLM / NUTS tick engine (mamba.jl_b200/csrc/glm_nuts.cu) ------------------------------------------------
def glm_pair(oracle, N, d, n_chains, seed):
    from mambacuda.engine import Engine
    X, y, _ = helpers.glm_data(N=N, d=d, seed=seed)
    eng = Engine("glm", n_chains, seed=seed)
    eng.set_data("X", X); eng.set_data("y", y)
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    orc = oracle.Oracle("glm", glm_d=d)
    orc.set_data("X", X); orc.set_data("y", y)
    orc.set_scheme([dict(kind=4, nodes=[0], max_depth=10)])
    inits = 0.1 * np.random.default_rng(seed).normal(size=(n_chains, d))
    return eng, orc, inits


def glm_resync(oracle, N, d, n_chains, seed, iters, burnin, run_kw):
    eng, orc, inits = glm_pair(oracle, N, d, n_chains, seed)
    compared, ties = helpers.resync_audit(eng, orc, np.arange(n_chains), inits, iters, burnin, seed, 0.0, 1e-7, 1e-7, run_kw=run_kw, nthreads=4)
    assert len(ties) <= max(1, compared // 200), ties
    return eng


def test_glm_tick_engine_matches_generic_kernel_and_oracle(oracle):
    # every chain is a resumable state machine fed by a shared gradient pass; it must request exactly the
    # gradients the reference's recursion does (same draws, same post-order merges, same dual averaging):
    # every step of 20 adaptive + 8 non-adaptive iterations from the oracle's state, tick engine (FP64 gradient kernel) and generic kernel
    eng = glm_resync(oracle, 400, 8, 32, 13, 28, 20, dict(glm_reference=True))
    _, tune, _ = eng.get_state()
    assert tune[:, 7].max() >= 4                                 # trees deeper than one doubling
    glm_resync(oracle, 400, 8, 32, 13, 28, 20, dict(force_generic=True))


def test_glm_tick_engine_fixed_stepsize_and_restart(oracle):
    glm_resync(oracle, 300, 6, 16, 17, 40, 0, dict(glm_reference=True))
    eng, orc, inits = glm_pair(oracle, 300, 6, 16, seed=17)
    eng.set_inits(inits)
    full = eng.run(40, burnin=0, thin=2, glm_reference=True)
    eng.set_inits(inits)
    p1 = eng.run(14, burnin=0, thin=2, glm_reference=True); p2 = eng.run(26, burnin=0, thin=2, glm_reference=True)
    np.testing.assert_allclose(np.concatenate([p1, p2], axis=0), full, rtol=1e-12)


def test_glm_posterior_recovers_coefficients(oracle):
    from mambacuda.engine import Engine
    X, y, beta = helpers.glm_data(N=4000, d=6, seed=5)
    eng = Engine("glm", 128, seed=1)
    eng.set_data("X", X); eng.set_data("y", y)
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    eng.set_inits(np.zeros((1, 6)), jitter_sd=0.1)
    out = eng.run(300, burnin=150, thin=1)
    mean = out.mean(axis=(0, 2)); sd = out.std(axis=(0, 2))
    assert np.all(np.abs(mean - beta) < 5 * sd)
    assert (eng.gelman(0.05, False)[:, 0] < 1.1).all()


@pytest.mark.parametrize("N,d,C", [(1000, 10, 128), (5000, 100, 256), (777, 37, 100)])
def test_glm_tensor_core_gradient_matches_fp64_kernel(oracle, N, d, C):
    # north_star: logpdf and gradient within 1e-5 relative of the reference arithmetic (here: the FP64 kernel, which
    # itself matches the oracle to 1e-12 in test_glm_density_and_gradient)
    from mambacuda.engine import Engine
    X, y, _ = helpers.glm_data(N=N, d=d, seed=3)
    eng = Engine("glm", C, seed=1)
    eng.set_data("X", X); eng.set_data("y", y)
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    beta = np.random.default_rng(4).normal(scale=0.5 / np.sqrt(d), size=(C, d))
    lp0, g0 = eng.glm_gradient(beta, impl=0)
    lp1, g1 = eng.glm_gradient(beta, impl=1)
    eta = beta @ X.T
    lp_np = (y * eta - np.logaddexp(0, eta)).sum(axis=1)
    g_np = (y - 1 / (1 + np.exp(-eta))) @ X
    np.testing.assert_allclose(lp0, lp_np, rtol=1e-12)
    np.testing.assert_allclose(g0, g_np, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(lp1, lp0, rtol=1e-5)
    gscale = np.abs(g0).max(axis=1, keepdims=True)
    assert np.max(np.abs(g1 - g0) / gscale) < 1e-5


def test_glm_tensor_core_gradient_with_columns_outside_the_fp16_range(oracle):
    # The tensor-core pass holds X as fp16 hi + lo pairs: a column in the millions would overflow (65,504) and one around 1e-7 would lose
    # its lo term.  Such columns are packed times a power of two and the factor is undone exactly on Theta and on the gradient, so the
    # 1e-5 bound of north_star holds per column whatever the units of the covariates.
    from mambacuda.engine import Engine
    N, d, C = 4000, 12, 128
    X, y, _ = helpers.glm_data(N=N, d=d, seed=11)
    colfac = np.ones(d); colfac[1] = 3.7e6; colfac[2] = 2.9e-7; colfac[5] = 1.0e5; colfac[7] = 4.4e-4
    Xs = X * colfac
    eng = Engine("glm", C, seed=1)
    eng.set_data("X", Xs); eng.set_data("y", y)
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    beta = np.random.default_rng(12).normal(scale=0.5 / np.sqrt(d), size=(C, d)) / colfac      # eta stays O(1)
    lp0, g0 = eng.glm_gradient(beta, impl=0)
    lp1, g1 = eng.glm_gradient(beta, impl=1)
    eta = beta @ Xs.T
    np.testing.assert_allclose(lp0, (y * eta - np.logaddexp(0, eta)).sum(axis=1), rtol=1e-12)
    assert np.all(np.isfinite(lp1)) and np.all(np.isfinite(g1))
    np.testing.assert_allclose(lp1, lp0, rtol=1e-5)
    # per column: the gradient of column j carries the units of that column
    gscale = np.abs(g0 / colfac).max(axis=1, keepdims=True)
    assert np.max(np.abs((g1 - g0) / colfac) / gscale) < 1e-5


def test_glm_tick_engine_with_tensor_core_gradient_is_statistically_equivalent(oracle):
    from mambacuda.engine import Engine
    X, y, beta = helpers.glm_data(N=4000, d=6, seed=5)
    res = []
    for ref in (True, False):
        eng = Engine("glm", 256, seed=1)
        eng.set_data("X", X); eng.set_data("y", y)
        eng.set_scheme([dict(kind="nuts", nodes=[0])])
        eng.set_inits(np.zeros((1, 6)), jitter_sd=0.1)
        eng.run(300, burnin=150, thin=1, store=False, out=False, glm_reference=ref)
        res.append(eng.summary_streaming())
    a, b = res
    assert np.all(np.abs(a[:, 0] - b[:, 0]) < 4 * np.hypot(a[:, 3], b[:, 3]) + 1e-3)     # means within MCSE
    np.testing.assert_allclose(a[:, 1], b[:, 1], rtol=0.1)                                # posterior SDs


# ---- magnesium: bounded (Uniform / truncated) priors and the two-sided link (transformdistribution.jl:6-48) ---------------------------
def test_magnesium_trajectories_match_oracle(oracle):
    # the script's own scheme (doc/examples/magnesium.jl:99-102): AMWG(mu) proposes on logit((mu + 10) / 20); the Slice blocks work on the
    # constrained scale where proposals outside [0, 1] / [0, 50] have density -Inf and shrink the interval
    g, o, _, _ = run_pair(oracle, "magnesium", 16, 300, 150, 2)
    assert_same_run(g, o, 300, 150, 2)
    out = g[0]
    assert out.shape[1] == 12 and (out[:, :6, :] > 0).all() and (out[:, 6:, :] > 0).all()


def test_magnesium_steps_on_the_link_scale_match_oracle(oracle):
    # every block on the link scale, incl. NUTS over (priors, mu): log links, two-sided links and their Jacobians in the gradient
    resync(oracle, "magnesium_transformed", 16, 40, 30, seed=5, rtol=1e-7, tie=1e-7)


def test_magnesium_densities_and_predict_match_oracle(oracle):
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme("magnesium_transformed")
    eng = Engine(tpl, 64, seed=3); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.5)
    st, _, _ = eng.get_state()                                   # jittered on the link scale: inside every support
    orc = oracle.Oracle(tpl); orc.set_scheme([helpers.oracle_block(b) for b in blocks])
    _, st_o, _ = orc.run(64, inits, 1, burnin=0, thin=1, seed=3, jitter_sd=0.5, partial=True)
    rng = np.random.default_rng(2)
    for b in range(4):
        np.testing.assert_allclose(eng.logpdf(b, st), orc.logpdf(b, st), rtol=RTOL_LP, atol=1e-12)
        x = np.stack([orc.unlist(b, s) for s in st]) + rng.normal(scale=0.1, size=(64, orc.unlist(b, st[0]).size))
        np.testing.assert_allclose(eng.logpdf(b, st, x), orc.logpdf(b, st, x), rtol=RTOL_LP, atol=1e-12)
    nn, nf = eng.factor_counts()
    for mask in (1 << nn, 1 << (nn + 1), (1 << nf) - 1, 1, 2, 8):
        np.testing.assert_allclose(eng.logpdf_nodes(mask, st), orc.logpdf_nodes(mask, st), rtol=RTOL_LP, atol=1e-12)
    g = eng.predict(st, stream_id=5); o = orc.predict(st, 3, stream_id=5)
    assert g.shape == o.shape == (64, 96) and (g == o).mean() > 0.999


def test_magnesium_posterior_matches_published_table(oracle):
    # doc/examples/magnesium.rst:45-57 (2 x 12,500, burnin 2,500, thin 2): mean, MCSE, SD
    from mambacuda.engine import Engine
    ref = {"tau[1]": (0.55098858, 0.0221, 0.358), "tau[2]": (1.11557619, 0.0238, 0.589), "tau[3]": (0.83211110, 0.0223, 0.491), "tau[4]": (0.47864203, 0.0136, 0.263),
           "tau[5]": (0.48624861, 0.0215, 0.354), "tau[6]": (0.56841884, 0.0059, 0.189), "OR[1]": (0.47784058, 0.0067, 0.154), "OR[2]": (0.42895913, 0.0081, 0.322),
           "OR[3]": (0.43118350, 0.0064, 0.183), "OR[4]": (0.47587697, 0.0065, 0.139), "OR[5]": (0.48545299, 0.0084, 0.146), "OR[6]": (0.44554385, 0.0054, 0.141)}
    tpl, blocks, inits = helpers.scheme("magnesium")
    eng = Engine(tpl, 512, seed=14); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.02)
    eng.run(8000, burnin=2500, thin=2, store=False, out=False)
    summ = eng.summary_streaming(); names = eng.names(1)
    for nm, (mean, mcse_ref, sd) in ref.items():
        j = names.index(nm)
        assert abs(summ[j, 0] - mean) < 3 * np.hypot(mcse_ref, summ[j, 3]) + 0.02 * sd, (nm, summ[j, 0], mean)
    codes = eng.link_codes(True)        # tau, OR are Logical columns > 0 (tau[4..6] and the ORs also exceed 1 somewhere): log link
    assert (codes[:3] == 1).all()
    assert (eng.gelman(0.05, True)[:, 0] < 1.1).all()


# ---- diagnostics over several handles: the packed two-round protocol and its NCCL transport (include/mambacuda.h) --------------------
def _combine_round1(bufs, p):
    b = np.stack(bufs)
    return np.concatenate([b[:, :p].min(axis=0), b[:, p:2 * p].max(axis=0), b[:, 2 * p:].sum(axis=0)])


def test_diag_global_equals_gelman_and_summary_and_the_two_handle_protocol(oracle):
    from mambacuda.engine import Engine, diag_finish
    tpl, blocks, inits = helpers.scheme("surgical_amwg")          # monitored: identity, log and Logical (logit heuristic) columns
    whole = Engine(tpl, 64, seed=5); whole.set_scheme(blocks); whole.set_inits(inits, jitter_sd=0.05)
    out = whole.run(1200, burnin=200, thin=2)
    for transform in (False, True):
        psrf, summ, codes = whole.diag_global(0.05, transform)
        np.testing.assert_allclose(psrf, whole.gelman(0.05, transform), rtol=1e-10)
        np.testing.assert_allclose(summ, whole.summary_streaming(), rtol=1e-10)
        np.testing.assert_array_equal(codes, whole.link_codes(transform))
        np.testing.assert_allclose(psrf, oracle.gelmandiag(out, 0.05, [{0: 0, 1: 1, 2: -1}[int(c)] for c in codes] if transform else None), rtol=1e-7)
    np.testing.assert_allclose(summ, oracle.summarystats(out, 0, 100), rtol=1e-8)
    # the same chains on two handles (global chain ids 0..39 and 40..63): buffers combined by the host, as any transport would
    parts = []
    for off, n in ((0, 40), (40, 24)):
        e = Engine(tpl, n, seed=5, chain_offset=off); e.set_scheme(blocks); e.set_inits(inits, jitter_sd=0.05)
        e.run(1200, burnin=200, thin=2, store=False, out=False, wait=False)      # MCU_RUN_ASYNC: both handles are queued before either is waited for
        parts.append(e)
    for e in parts:
        e.wait()
    p = whole.dims()[1]
    r1 = _combine_round1([e.diag_round1() for e in parts], p)
    r2 = np.sum([e.diag_round2(True, r1) for e in parts], axis=0)
    psrf2, summ2, codes2 = diag_finish(parts[0].n_kept(), parts[0].monitor_links(), 0.05, True, r1, r2)
    np.testing.assert_allclose(psrf2, psrf, rtol=1e-9)
    np.testing.assert_allclose(summ2, summ, rtol=1e-9)
    np.testing.assert_array_equal(codes2, codes)


def test_async_run_and_get_samples(oracle):
    from mambacuda.engine import Engine, MambaCudaError
    tpl, blocks, inits = helpers.scheme("seeds_amwg")
    a = Engine(tpl, 48, seed=3); a.set_scheme(blocks); a.set_inits(inits, jitter_sd=0.1)
    want = a.run(200, burnin=100, thin=4)
    b = Engine(tpl, 48, seed=3); b.set_scheme(blocks); b.set_inits(inits, jitter_sd=0.1)
    assert b.run(200, burnin=100, thin=4, wait=False) is None
    b.wait()
    np.testing.assert_array_equal(b.samples(), want)
    assert b.last_kernel_ms() > 0
    into = np.empty(want.shape, order="F")
    assert b.samples(into) is into and np.array_equal(into, want)
    c = Engine(tpl, 4, seed=3); c.set_scheme(blocks); c.set_inits(inits)
    c.run(50, burnin=10, thin=1, store=False, out=False)
    with pytest.raises(MambaCudaError, match="no stored samples"):
        c.samples()


def _nccl_rank(rank, world, tmp):
    import os, sys, time
    sys.path[:0] = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mamba.jl_b200"), os.path.dirname(os.path.abspath(__file__))]
    import numpy as np
    import helpers
    from mambacuda.engine import Engine, comm_unique_id
    idf = os.path.join(tmp, "nccl_id")
    if rank == 0:
        with open(idf + ".tmp", "wb") as f:
            f.write(comm_unique_id())
        os.replace(idf + ".tmp", idf)                     # the id travels by whatever messaging the host language has (here: a file)
    while not os.path.exists(idf):
        time.sleep(0.01)
    uid = open(idf, "rb").read()
    tpl, blocks, inits = helpers.scheme("surgical_amwg")
    n = 40 if rank == 0 else 24
    e = Engine(tpl, n, seed=5, chain_offset=0 if rank == 0 else 40, device=rank); e.set_scheme(blocks); e.set_inits(inits, jitter_sd=0.05)
    e.comm_init(rank, world, uid)
    assert e.comm_size() == (rank, world)
    e.run(1200, burnin=200, thin=2, store=False, out=False)
    psrf, summ, codes = e.diag_global(0.05, True)
    np.savez(os.path.join(tmp, f"r{rank}.npz"), psrf=psrf, summ=summ, codes=codes)


def test_nccl_diag_global_over_two_gpus(oracle, tmp_path):
    # mcu_comm_init + mcu_diag_global: both protocol rounds on the device, all-reduced by NCCL inside libmambacuda (one rank per GPU)
    from mambacuda import _lib
    from mambacuda.engine import Engine
    if _lib.lib().mcu_device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_nccl_rank, args=(2, str(tmp_path)), nprocs=2, join=True)
    tpl, blocks, inits = helpers.scheme("surgical_amwg")
    whole = Engine(tpl, 64, seed=5); whole.set_scheme(blocks); whole.set_inits(inits, jitter_sd=0.05)
    whole.run(1200, burnin=200, thin=2, store=False, out=False)
    psrf, summ, codes = whole.diag_global(0.05, True)
    for r in range(2):
        got = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_allclose(got["psrf"], psrf, rtol=1e-9)
        np.testing.assert_allclose(got["summ"], summ, rtol=1e-9)
        np.testing.assert_array_equal(got["codes"], codes)


# ---- oxford (244 state elements) and epil (303): the largest examples of the corpus, on the generic kernel ------------------------------
@pytest.mark.parametrize("name,iters", [("oxford", 80), ("oxford_componentwise", 120), ("epil", 80), ("epil_componentwise", 120)])
def test_oxford_and_epil_trajectories_match_oracle(oracle, name, iters):
    # the scripts' own schemes (120- / 236-dimensional multivariate Slice blocks: doc/examples/oxford.jl:97-100, epil.jl:126-130) and
    # componentwise schemes (AMWG / univariate Slice over the same nodes: local term updates)
    g, o, _, _ = run_pair(oracle, name, 8, iters, iters // 2, 2)
    assert_same_run(g, o, iters, iters // 2, 2)


@pytest.mark.parametrize("name", ["oxford_componentwise", "epil_componentwise"])
def test_oxford_and_epil_posteriors_are_close_to_the_published_tables(oracle, name):
    # doc/examples/oxford.rst / epil.rst: single poorly mixed runs (published ESS 104-268 of 10,000 / 12,500 draws, PSRF between our own
    # chains of the scripts' schemes 1.1-1.5 after 7,500 iterations), so the tables are matched within the published SDs, not within MCSE;
    # the componentwise schemes mix faster and run at one warp-instruction per term instead of one block evaluation per shrinkage step
    from mambacuda.engine import Engine
    tpl, blocks, inits = helpers.scheme(name)
    eng = Engine(tpl, 256, seed=15); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.02)
    eng.run(6000, burnin=2000, thin=2, store=False, out=False)
    summ = eng.summary_streaming(); names = eng.names(1)
    pub = {"oxford": {"beta2": (0.005477119, 0.0035675748), "beta1": (-0.043336269, 0.0161754258), "alpha": (0.565784774, 0.0630050896), "s2": (0.026238992, 0.0307989154)},
           "epil": {"s2_b": (0.13523750, 0.031819272), "s2_b1": (0.24911885, 0.073166731), "alpha_V4": (-0.09287934, 0.083666872), "alpha_Age": (0.45830900, 0.394536219),
                    "alpha_BT": (0.24217000, 0.190566444), "alpha_Trt": (-0.75931393, 0.397734236), "alpha_Base": (0.91104974, 0.135354470), "alpha0": (-1.35617079, 1.313240197)}}[tpl]
    for nm, (mean, sd) in pub.items():
        j = names.index(nm)
        assert abs(summ[j, 0] - mean) < (2.0 if nm.startswith("s2") else 1.0) * sd, (nm, summ[j, 0], mean, sd)
    assert np.isfinite(eng.gelman(0.05, True)).all()


# ---- the north_star shim on the BASELINE models: both sides consume the same host-supplied uniform stream ----------------------------
@pytest.mark.parametrize("name,iters,burnin,per_iter", [("seeds_amwg", 150, 75, 90), ("seeds_amm", 120, 60, 90), ("rats_slice_amwg", 60, 30, 400), ("pumps_gibbs_amwg", 150, 75, 120)])
def test_external_stream_on_the_baseline_models(oracle, name, iters, burnin, per_iter):
    # mcu_set_rng_mode(EXTERNAL): draws are read sequentially from u[chain][...] (a normal takes two, cosine branch) in the order the
    # reference's sampler consumes rand() / randn() — what a Julia shim that overrides rand / randn hands over.  Schemes with a fused
    # kernel run on the generic kernel in this mode (the fused kernels address the counter-based stream by position).
    eng, orc, inits = make_pair(oracle, name, 8)
    u = np.random.default_rng(7).uniform(size=(8, per_iter * iters))
    eng.set_external_stream(u)
    eng.set_inits(inits)
    out_g = eng.run(iters, burnin=burnin, thin=1)
    st_g, tune_g, _ = eng.get_state()
    out_o, st_o, tune_o, marg = orc.run(8, inits, iters, burnin=burnin, thin=1, ext_u=u, margins=True)
    assert_same_run((out_g, st_g, tune_g), (out_o, st_o, tune_o, marg), iters, burnin, 1, tune_rtol=1e-4 if "amm" in name else 1e-6)


def test_external_stream_nuts_steps_and_exhaustion(oracle):
    # NUTS on the shim stream (rats, 62-dimensional block; the number of draws per iteration depends on the tree): decisions and trajectories of
    # the first iterations agree (nutsepsilon, two adaptive and two non-adaptive iterations; longer horizons are covered step by step by the resync
    # tests because dual averaging amplifies rounding from one iteration to the next), and a stream that runs dry is an error, not a silently
    # degenerate chain
    from mambacuda.engine import MambaCudaError
    eng, orc, inits = make_pair(oracle, "rats_nuts_slice", 8)
    tpl, blocks, _ = helpers.scheme("rats_nuts_slice")
    ob = [helpers.oracle_block(b) for b in blocks]; ob[0]["max_depth"] = 10
    orc.set_scheme(ob)
    u = np.random.default_rng(9).uniform(size=(8, 60000))
    eng.set_external_stream(u); eng.set_inits(inits)
    out_g = eng.run(4, burnin=2, thin=1)
    st_g, tune_g, _ = eng.get_state()
    out_o, st_o, tune_o, marg = orc.run(8, inits, 4, burnin=2, thin=1, ext_u=u, margins=True)
    assert_same_run((out_g, st_g, tune_g), (out_o, st_o, tune_o, marg), 4, 2, 1, rtol=1e-6, tune_rtol=1e-5)
    short = np.random.default_rng(9).uniform(size=(8, 50))
    eng.set_external_stream(short); eng.set_inits(inits)
    with pytest.raises(MambaCudaError, match="external uniform stream exhausted"):
        eng.run(6, burnin=4, thin=1)


def test_streaming_mpsrf_equals_gelmandiag_on_the_stored_draws(oracle):
    # gelmandiag(c; mpsrf = true) (gelmandiag.jl:49-55) without the draws: within-chain covariances are streamed as Welford co-moments (raw scale
    # and the nodes' own link scale), reduced over chains with the packed protocol; compared with the host-array entry point on the stored draws
    from mambacuda import api
    from mambacuda.engine import Engine, diag_finish
    for name, transform in (("seeds_amwg", False), ("seeds_amwg", True), ("pumps_slice", True), ("rats_slice_amwg", False)):
        tpl, blocks, inits = helpers.scheme(name)
        eng = Engine(tpl, 48, seed=6); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
        out = eng.run(900, burnin=300, thin=2, force_generic=(name == "pumps_slice"), mpsrf=True)
        psrf, summ, codes, mv = eng.diag_global(0.05, transform, mpsrf=True)
        want = api._chains_gelman(out, 0.05, codes if transform else None, True)
        np.testing.assert_allclose(psrf, want[:-1], rtol=1e-7)
        np.testing.assert_allclose(mv, want[-1, 0], rtol=1e-7)
        # the same through two handles and the host-carried protocol
        parts = []
        for off, n in ((0, 20), (20, 28)):
            e = Engine(tpl, n, seed=6, chain_offset=off); e.set_scheme(blocks); e.set_inits(inits, jitter_sd=0.05)
            e.run(900, burnin=300, thin=2, store=False, out=False, force_generic=(name == "pumps_slice"), mpsrf=True)
            parts.append(e)
        p = eng.dims()[1]
        r1 = _combine_round1([e.diag_round1() for e in parts], p)
        r2 = np.sum([e.diag_round2(transform, r1) for e in parts], axis=0)
        _, _, _, mv2 = diag_finish(parts[0].n_kept(), parts[0].monitor_links(), 0.05, transform, r1, r2, mpsrf=True)
        np.testing.assert_allclose(mv2, mv, rtol=1e-9)
    # rats with transform: alpha0 is a Logical column whose link is resolved by the heuristic to log — no streamed co-moments on that scale
    tpl, blocks, inits = helpers.scheme("rats_slice_amwg")
    eng = Engine(tpl, 16, seed=6); eng.set_scheme(blocks); eng.set_inits(inits, jitter_sd=0.05)
    eng.run(400, burnin=100, thin=2, store=False, out=False, mpsrf=True)
    assert np.isnan(eng.diag_global(0.05, True, mpsrf=True)[3])
    assert not np.isnan(eng.diag_global(0.05, False, mpsrf=True)[3])
    eng.run(100, burnin=100, thin=2, store=False, out=False)              # a run without MCU_RUN_MPSRF: the co-moments no longer cover every kept draw
    assert np.isnan(eng.diag_global(0.05, False, mpsrf=True)[3])
