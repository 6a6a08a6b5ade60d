"""Shared fixtures for the parity tests: schemes, initial values and comparison helpers."""
import numpy as np

# state record layouts (include/mambacuda.h):
#  line : beta[2], s2                 seeds: alpha0, alpha1, alpha2, alpha12, s2, b[21]
#  rats : mu_alpha, mu_beta, s2_alpha, s2_beta, s2_c, alpha[30], beta[30]
#  pumps: alpha, beta, theta[10]      glm  : beta[d]

SEEDS_INITS = np.zeros((2, 26)); SEEDS_INITS[0, 4] = 0.01; SEEDS_INITS[1, 4] = 1.0   # doc/examples/seeds.jl:60-65
RATS_INITS = np.array([[150, 10, 1, 1, 1] + [250] * 30 + [6] * 30,                   # doc/examples/rats.jl:100-108
                       [15, 1, 10, 10, 10] + [20] * 30 + [0.6] * 30], dtype=float)
LINE_INITS = np.array([[0.3, -0.2, 1.5], [1.0, 0.5, 0.7], [-0.5, 1.2, 3.0]])


SURGICAL_INITS = np.array([[0.0, 1.0] + [0.1] * 12, [1.0, 10.0] + [0.5] * 12])                       # doc/examples/surgical.jl:47-50


DYES_INITS = np.array([[1.0, 1500.0, 1.0] + [1500.0] * 6, [10.0, 3000.0, 10.0] + [3000.0] * 6])       # doc/examples/dyes.jl:51-56 (state order s2_between, theta, s2_within, mu)


SALM_INITS = np.array([[10.0, 0.0, 0.0, 0.0] + [0.0] * 18, [1.0, 0.01, 1.0, 1.0] + [0.0] * 18])          # doc/examples/salm.jl:56-61 (state order s2, gamma, beta, alpha, lambda)
EQUIV_INITS = np.array([[1.0, 1.0, 0.0, 0.0, 0.0] + [0.0] * 20, [10.0, 10.0, 10.0, 10.0, 10.0] + [0.0] * 20])   # doc/examples/equiv.jl:79-84 (s2_2, s2_1, pi, phi, mu, delta)


BLOCKER_INITS = np.array([[1.0, 0.0, 0.0] + [0.0] * 44, [10.0, 2.0, 2.0] + [2.0] * 44])                 # doc/examples/blocker.jl:73-78 (s2, d, delta_new, mu, delta)


STACKS_INITS = np.array([[10.0, 0.0, 0.0, 0.0, 10.0], [1.0, 1.0, 1.0, 1.0, 1.0]])                       # doc/examples/stacks.jl:98-101 (beta0, beta[3], s2)


# doc/examples/magnesium.jl:86-95 (state order priors[6], mu[6], theta[6 x 8], pc[6 x 8])
MAGNESIUM_INITS = np.array([[1, 1, 1, 0.5, 0.5, 1] + [-0.5] * 6 + [0.0] * 48 + [0.5] * 48, [1, 1, 1, 0.5, 0.5, 1] + [0.5] * 6 + [0.0] * 48 + [0.5] * 48], dtype=float)


OXFORD_INITS = np.array([[0, 0, 0, 1] + [0] * 240, [1, 1, 1, 10] + [0] * 240], dtype=float)       # doc/examples/oxford.jl:86-93 (alpha, beta1, beta2, s2, b[120], mu[120])
EPIL_INITS = np.array([[0] * 6 + [1, 1] + [0] * 295, [1] * 6 + [10, 10] + [0] * 295], dtype=float)  # doc/examples/epil.jl:115-122 (a0, coefficients, s2_b1, s2_b, b1[59], b[236])


def pumps_inits(seed=1):
    rng = np.random.default_rng(seed)   # doc/examples/pumps.jl:43-49 draws theta from Gamma
    return np.array([[1.0, 1.0] + list(rng.gamma(1.0, 1.0, 10)), [10.0, 10.0] + list(rng.gamma(10.0, 0.1, 10))])


SCHEMES = {
    # AMWG(beta) + Slice(s2, transform)
    "line_amwg_slice": ("line", [dict(kind="amwg", nodes=[0], scale=1.0, adapt="burnin"),
                                 dict(kind="slice_multi", nodes=[1], scale=5.0, transform=1)], LINE_INITS),
    # doc/tutorial/line.jl:48-49 scheme1
    "line_nuts_slice": ("line", [dict(kind="nuts", nodes=[0]), dict(kind="slice_multi", nodes=[1], scale=3.0)], LINE_INITS),
    # doc/tutorial/line.jl:52 scheme2
    "line_nuts_all": ("line", [dict(kind="nuts", nodes=[0, 1])], LINE_INITS),
    "line_nuts_fd": ("line", [dict(kind="nuts", nodes=[0, 1], grad="forward")], LINE_INITS),
    "line_rwm": ("line", [dict(kind="rwm", nodes=[0, 1], scale=[0.5, 0.2, 0.8])], LINE_INITS),
    "line_rwm_unif": ("line", [dict(kind="rwm", nodes=[0, 1], scale=0.6, proposal="symuniform")], LINE_INITS),
    "line_rwm_tri": ("line", [dict(kind="rwm", nodes=[0, 1], scale=0.6, proposal="symtriangular")], LINE_INITS),
    # doc/tutorial/line.jl:54 scheme3 = [Gibbs_beta, Gibbs_s2]: the tutorial's user-defined conjugate samplers
    "line_gibbs": ("line", [dict(kind="gibbs", nodes=[0]), dict(kind="gibbs", nodes=[1])], LINE_INITS),
    "line_rwm_cos": ("line", [dict(kind="rwm", nodes=[0, 1], scale=0.8, proposal="cosine")], LINE_INITS),
    "line_rwm_epa": ("line", [dict(kind="rwm", nodes=[0, 1], scale=0.8, proposal="epanechnikov")], LINE_INITS),
    "line_rwm_biw": ("line", [dict(kind="rwm", nodes=[0, 1], scale=0.9, proposal="biweight")], LINE_INITS),
    "line_rwm_trw": ("line", [dict(kind="rwm", nodes=[0, 1], scale=1.0, proposal="triweight")], LINE_INITS),
    "line_mala": ("line", [dict(kind="mala", nodes=[0, 1], epsilon=0.08)], LINE_INITS),
    "line_mala_sigma": ("line", [dict(kind="mala", nodes=[0, 1], epsilon=0.08,
                                      scale=np.array([[1.0, 0.2, 0.0], [0.2, 0.5, 0.1], [0.0, 0.1, 2.0]]))], LINE_INITS),
    "line_hmc": ("line", [dict(kind="hmc", nodes=[0, 1], epsilon=0.05, L=8)], LINE_INITS),
    "line_hmc_sigma": ("line", [dict(kind="hmc", nodes=[0, 1], epsilon=0.05, L=5,
                                     scale=np.array([[1.0, 0.2, 0.0], [0.2, 0.5, 0.1], [0.0, 0.1, 2.0]]))], LINE_INITS),
    # small initial proposals: every early move is accepted, so the adapted covariance is full rank by the time it is
    # first used (m > 2n, amm.jl:75-77); a rank-deficient Sigma makes cholfact(.., Val{true})'s rank decision — and the
    # trajectory — depend on rounding noise in the reference itself
    "line_amm": ("line", [dict(kind="amm", nodes=[0, 1], scale=0.0005 * np.eye(3))], LINE_INITS),
    "line_slice_uni": ("line", [dict(kind="slice_uni", nodes=[0, 1], scale=[1.0, 1.0, 2.0], transform=1)], LINE_INITS),
    # SURVEY.md §8d config 2 scheme A (north_star "AMWG on seeds")
    "seeds_amwg": ("seeds", [dict(kind="amwg", nodes=[0, 1, 2, 3], scale=0.1), dict(kind="amwg", nodes=[5], scale=0.01),
                             dict(kind="amwg", nodes=[4], scale=0.1)], SEEDS_INITS),
    # doc/examples/seeds.jl:69-71 scheme B (reference-faithful)
    "seeds_amm": ("seeds", [dict(kind="amm", nodes=[0, 1, 2, 3], scale=0.01 * np.eye(4)), dict(kind="amwg", nodes=[5], scale=0.01),
                            dict(kind="amwg", nodes=[4], scale=0.1)], SEEDS_INITS),
    # doc/examples/rats.jl:112-116
    "rats_slice_amwg": ("rats", [dict(kind="slice_multi", nodes=[4], scale=10.0), dict(kind="amwg", nodes=[5], scale=100.0),
                                 dict(kind="slice_uni", nodes=[0, 2], scale=[100.0, 10.0]), dict(kind="amwg", nodes=[6], scale=1.0),
                                 dict(kind="slice_uni", nodes=[1, 3], scale=1.0)], RATS_INITS),
    # NUTS(alpha, beta, mu_alpha, mu_beta) + Slice(s2_c, s2_alpha, s2_beta; univariate)
    "rats_nuts_slice": ("rats", [dict(kind="nuts", nodes=[5, 6, 0, 1]), dict(kind="slice_uni", nodes=[4, 2, 3], scale=[10.0, 10.0, 1.0])], RATS_INITS),
    # doc/examples/dyes.jl:60-73: four schemes on one model (scheme4: RWM(theta) with a Cosine proposal)
    "dyes_nuts_slice": ("dyes", [dict(kind="nuts", nodes=[3, 1]), dict(kind="slice_multi", nodes=[2, 0], scale=1000.0)], DYES_INITS),
    "dyes_mala_slice": ("dyes", [dict(kind="mala", nodes=[1], epsilon=50.0), dict(kind="mala", nodes=[3], epsilon=50.0, scale=np.eye(6)),
                                 dict(kind="slice_multi", nodes=[2, 0], scale=1000.0)], DYES_INITS),
    "dyes_hmc_slice": ("dyes", [dict(kind="hmc", nodes=[1], epsilon=10.0, L=5), dict(kind="hmc", nodes=[3], epsilon=10.0, L=5, scale=np.eye(6)),
                                dict(kind="slice_multi", nodes=[2, 0], scale=1000.0)], DYES_INITS),
    "dyes_rwm_slice": ("dyes", [dict(kind="rwm", nodes=[1], scale=50.0, proposal="cosine"), dict(kind="rwm", nodes=[3], scale=50.0),
                                dict(kind="slice_multi", nodes=[2, 0], scale=1000.0)], DYES_INITS),
    # doc/examples/salm.jl:63-64: Slice([:alpha, :beta, :gamma], [1.0, 1.0, 0.1]), AMWG([:lambda, :s2], 0.1)
    "salm_slice_amwg": ("salm", [dict(kind="slice_multi", nodes=[3, 2, 1], scale=[1.0, 1.0, 0.1]), dict(kind="amwg", nodes=[4, 0], scale=0.1)], SALM_INITS),
    # doc/examples/blocker.jl:84-86: AMWG(:mu, 0.1), AMWG([:delta, :delta_new], 0.1), Slice([:d, :s2], 1.0)
    "blocker_amwg_slice": ("blocker", [dict(kind="amwg", nodes=[3], scale=0.1), dict(kind="amwg", nodes=[4, 2], scale=0.1),
                                       dict(kind="slice_multi", nodes=[1, 0], scale=1.0)], BLOCKER_INITS),
    "blocker_nuts_slice": ("blocker", [dict(kind="nuts", nodes=[3, 4, 2]), dict(kind="slice_multi", nodes=[1, 0], scale=1.0)], BLOCKER_INITS),
    # doc/examples/stacks.jl:104-105: NUTS([:beta0, :beta]), Slice(:s2, 1.0)
    "stacks_nuts_slice": ("stacks", [dict(kind="nuts", nodes=[0, 1]), dict(kind="slice_multi", nodes=[2], scale=1.0)], STACKS_INITS),
    "stacks_amwg": ("stacks", [dict(kind="amwg", nodes=[0, 1], scale=1.0), dict(kind="amwg", nodes=[2], scale=0.5)], STACKS_INITS),
    # doc/examples/equiv.jl:89-91: NUTS(:delta), Slice([:mu, :phi, :pi], 1.0), Slice([:s2_1, :s2_2], 1.0, Univariate)
    "equiv_nuts_slice": ("equiv", [dict(kind="nuts", nodes=[5]), dict(kind="slice_multi", nodes=[4, 3, 2], scale=1.0),
                                   dict(kind="slice_uni", nodes=[1, 0], scale=1.0)], EQUIV_INITS),
    "equiv_amwg": ("equiv", [dict(kind="amwg", nodes=[5], scale=0.1), dict(kind="amwg", nodes=[4, 3, 2], scale=0.1), dict(kind="amwg", nodes=[1, 0], scale=0.5)], EQUIV_INITS),
    # doc/examples/surgical.jl:54-55: NUTS(:b), Slice([:mu, :s2], 1.0)
    "surgical_nuts_slice": ("surgical", [dict(kind="nuts", nodes=[2]), dict(kind="slice_multi", nodes=[0, 1], scale=1.0)], SURGICAL_INITS),
    "surgical_amwg": ("surgical", [dict(kind="amwg", nodes=[2], scale=0.3), dict(kind="amwg", nodes=[0, 1], scale=0.3)], SURGICAL_INITS),
    # doc/examples/magnesium.jl:99-102: AMWG(:theta, 0.1), AMWG(:mu, 0.1) [Uniform(-10, 10): two-sided link], Slice(:pc, 0.25, Univariate),
    # Slice(:priors, [1.0, 5.0, 5.0, 0.25, 0.25, 5.0], Univariate)
    "magnesium": ("magnesium", [dict(kind="amwg", nodes=[2], scale=0.1), dict(kind="amwg", nodes=[1], scale=0.1), dict(kind="slice_uni", nodes=[3], scale=0.25),
                                dict(kind="slice_uni", nodes=[0], scale=[1.0, 5.0, 5.0, 0.25, 0.25, 5.0])], MAGNESIUM_INITS),
    # every block on the link scale: log for priors[1] / priors[6], two-sided logit for the Uniform priors, mu and pc
    "magnesium_transformed": ("magnesium", [dict(kind="amwg", nodes=[2], scale=0.1), dict(kind="amwg", nodes=[1, 0], scale=0.2),
                                            dict(kind="slice_uni", nodes=[3], scale=1.0, transform=1), dict(kind="nuts", nodes=[0, 1])], MAGNESIUM_INITS),
    # doc/examples/oxford.jl:97-100: AMWG([:alpha, :beta1, :beta2], 1.0), Slice(:s2, 1.0), Slice(:mu, 1.0), Slice(:b, 1.0) — 120-dimensional multivariate slices
    "oxford": ("oxford", [dict(kind="amwg", nodes=[0, 1, 2], scale=1.0), dict(kind="slice_multi", nodes=[3], scale=1.0), dict(kind="slice_multi", nodes=[5], scale=1.0),
                          dict(kind="slice_multi", nodes=[4], scale=1.0)], OXFORD_INITS),
    "oxford_componentwise": ("oxford", [dict(kind="amwg", nodes=[0, 1, 2], scale=0.5), dict(kind="slice_uni", nodes=[3], scale=1.0, transform=1), dict(kind="amwg", nodes=[5], scale=0.5),
                                        dict(kind="slice_uni", nodes=[4], scale=1.0)], OXFORD_INITS),
    # doc/examples/epil.jl:126-130: AMWG([:a0, :alpha_Base, :alpha_Trt, :alpha_BT, :alpha_Age, :alpha_V4], 0.1), Slice(:b1, 0.5), Slice(:b, 0.5), Slice([:s2_b1, :s2_b], 1.0)
    "epil": ("epil", [dict(kind="amwg", nodes=[0, 1, 2, 3, 4, 5], scale=0.1), dict(kind="slice_multi", nodes=[8], scale=0.5), dict(kind="slice_multi", nodes=[9], scale=0.5),
                      dict(kind="slice_multi", nodes=[6, 7], scale=1.0)], EPIL_INITS),
    "epil_componentwise": ("epil", [dict(kind="amwg", nodes=[0, 1, 2, 3, 4, 5], scale=0.1), dict(kind="amwg", nodes=[8], scale=0.3), dict(kind="slice_uni", nodes=[9], scale=0.5),
                                    dict(kind="amwg", nodes=[6, 7], scale=0.5)], EPIL_INITS),
    # doc/examples/pumps.jl:52-53
    "pumps_slice": ("pumps", [dict(kind="slice_uni", nodes=[0, 1], scale=1.0), dict(kind="slice_uni", nodes=[2], scale=1.0)], None),
    "pumps_amwg_nuts": ("pumps", [dict(kind="amwg", nodes=[0, 1], scale=0.5), dict(kind="nuts", nodes=[2])], None),
    # BASELINE.json configs[4] / SURVEY.md §8d config 5: Gibbs(theta), Gibbs(beta), AMWG(alpha)
    "pumps_gibbs_amwg": ("pumps", [dict(kind="gibbs", nodes=[2]), dict(kind="gibbs", nodes=[1]), dict(kind="amwg", nodes=[0], scale=1.0)], None),
}


def oracle_block(b):
    """helpers scheme dict -> pyoracle.make_desc kwargs (ints instead of names)."""
    from mambacuda import _lib
    o = dict(b)
    if "adapt" in o and isinstance(o["adapt"], str):
        o["adapt"] = _lib.ADAPT[o["adapt"]]
    if "proposal" in o and isinstance(o["proposal"], str):
        o["proposal"] = _lib.PROPOSAL[o["proposal"]]
    if "grad" in o and isinstance(o["grad"], str):
        o["grad"] = _lib.GRAD[o["grad"]]
    return o


def scheme(name):
    tpl, blocks, inits = SCHEMES[name]
    if inits is None:
        inits = pumps_inits()
    return tpl, blocks, np.array(inits, dtype=float)


def glm_data(N=200, d=8, seed=1):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, d)); X[:, 0] = 1.0
    beta = rng.normal(size=d) / np.sqrt(d)
    y = (rng.uniform(size=N) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(float)
    return X, y, beta


# ---- audited comparison of a device run with the oracle's (no tolerated fraction of diverging chains) -------------------------
TIE = 1e-9   # a decision is a "threshold tie" when |lhs - rhs| < TIE * max(1, |rhs|) (oracle/samplers.hpp note_margin)


def audit_divergence(g, o, margins, kept_iters, iter0, rtol=1e-8, atol=1e-10, tune_rtol=1e-6, tie=TIE, int_tune_cols=()):
    """g, o = (out [kept x p x C], state [C x D], tune [C x T]) of the device and of the oracle on the same stream;
    margins [C x iters] = the oracle's smallest decision margin of iterations iter0 + 1 .. iter0 + iters;
    kept_iters = the (absolute, 1-based) iteration of every row of `out`.

    Every chain must reproduce the oracle to `rtol`.  A chain that does not is accepted ONLY if the oracle took a decision within
    `tie` of its threshold inside the window of iterations in which the first difference appears (the two sides evaluate the same
    density in a different floating-point order, so only such a decision can legitimately go the other way); from there on the two
    trajectories are different realisations and are not compared.  Integer tune columns (accept counters, m) of agreeing chains
    must be equal.  Returns (number of chains that agree throughout, list of (chain, first differing iteration window, margin))."""
    out_g, st_g, tune_g = g
    out_o, st_o, tune_o = o
    C = st_g.shape[0]
    iters = margins.shape[1]
    kept_iters = np.asarray(kept_iters, dtype=np.int64)
    assert out_g.shape == out_o.shape and out_g.shape[0] == kept_iters.size
    assert not np.isnan(out_g).any() and not np.isnan(st_g).any()
    ties, bad = [], []
    for c in range(C):
        rows_ok = np.all(np.isclose(out_g[:, :, c], out_o[:, :, c], rtol=rtol, atol=atol), axis=1) if kept_iters.size else np.ones(0, bool)
        end_ok = np.allclose(st_g[c], st_o[c], rtol=rtol, atol=atol) and \
            np.allclose(tune_g[c], tune_o[c], rtol=tune_rtol, atol=tune_rtol * 1e-2, equal_nan=True)
        if rows_ok.all() and end_ok:
            for j in int_tune_cols:
                assert tune_g[c, j] == tune_o[c, j], f"chain {c}: integer tune column {j} differs ({tune_g[c, j]} vs {tune_o[c, j]})"
            continue
        if not rows_ok.all():
            r = int(np.argmin(rows_ok))
            lo = kept_iters[r - 1] if r > 0 else iter0
            hi = kept_iters[r]
        else:
            lo = kept_iters[-1] if kept_iters.size else iter0
            hi = iter0 + iters
        w = margins[c, lo - iter0:hi - iter0]
        m = float(w.min()) if w.size else np.inf
        (ties if m < tie else bad).append((c, (int(lo) + 1, int(hi)), m))
    assert not bad, (f"{len(bad)} of {C} chains part from the oracle where no decision was within {tie:g} of its threshold "
                     f"(chain, iterations, smallest margin): {bad[:8]}")
    return C - len(ties), ties


def resync_audit(eng, orc, ids, inits, iters, burnin, seed, jitter_sd, rtol, tie, run_kw=None, nthreads=None, tune_rtol=None):
    """Per-step parity (north_star: same state + same stream => same decisions and the same next state).  At every iteration the
    sampled device chains `ids` are put at the oracle's state / tune records, both sides take ONE iteration, and the results must
    agree to `rtol` unless the oracle's smallest decision margin of that iteration is below `tie`.  The other chains of the launch
    simply keep running from their own states.  Returns (steps compared, [(chain, iteration, margin) of the ties])."""
    import os
    run_kw = run_kw or {}
    nthreads = nthreads or os.cpu_count() or 4
    ids = np.asarray(ids, dtype=np.int64)
    tune_rtol = tune_rtol or rtol * 10
    ties, compared = [], 0
    prev_o = prev_t = st = tune = None
    for i in range(1, iters + 1):
        if i == 1:      # iteration 1 starts from the (jittered) inits on both sides; the tune records are created there (sampler.jl:40-45)
            eng.set_inits(inits, jitter_sd=jitter_sd)
            _, st_o, tune_o, marg = orc.run(0, inits, 1, burnin=burnin, thin=1, seed=seed, jitter_sd=jitter_sd, chain_ids=ids,
                                            nthreads=nthreads, margins=True, store=False, partial=True)
        else:
            st[ids] = prev_o; tune[ids] = prev_t
            eng.set_state(st, tune, i - 1)
            _, st_o, tune_o, marg = orc.run(0, prev_o, 1, burnin=burnin, thin=1, seed=seed, chain_ids=ids, iter0=i - 1, tune_in=prev_t,
                                            nthreads=nthreads, margins=True, store=False)
        eng.run(1, burnin=burnin, thin=1, store=False, out=False, partial=True, **run_kw)
        st, tune, it = eng.get_state()
        assert it == i
        for k, c in enumerate(ids):
            same = np.allclose(st[c], st_o[k], rtol=rtol, atol=1e-9) and np.allclose(tune[c], tune_o[k], rtol=tune_rtol, atol=1e-9, equal_nan=True)
            if not same:
                assert marg[k, 0] < tie, (f"iteration {i}, chain {c}: device and oracle differ after one step from a common state and no "
                                          f"decision was within {tie:g} of its threshold (smallest margin {marg[k, 0]:.3g}; "
                                          f"max state diff {np.max(np.abs(st[c] - st_o[k])):.3g})")
                ties.append((int(c), i, float(marg[k, 0])))
        compared += len(ids)
        prev_o, prev_t = st_o, tune_o
    return compared, ties
