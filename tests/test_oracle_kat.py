"""CPU tests that pin the oracle: the reference holds no assert of its own, so the anchors are
(i) the closed-form log posterior / gradient of the line model written out in the reference's
doc/samplers/amwg.jl:17-25 and doc/samplers/nuts.jl:17-31, (ii) hand-derivable node values,
(iii) scipy for the third-party special functions and distributions, (iv) Random123's published Philox vectors."""
import numpy as np
import pytest
import scipy.special as sp
import scipy.stats as st

import helpers


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stream_contract(oracle):
    # two streams per block update (0 = rand, 1 = randn), counter = (k >> 1, iter, chain, block | kind << 16 | stream << 24);
    # one Philox block gives two draws: uniforms (w0,w1) / (w2,w3), normals rad cos(ang) / rad sin(ang)
    seed, chain, it, block = 0x123456789, 7, 3, 2
    key = [seed & 0xffffffff, seed >> 32]
    u53 = lambda hi, lo: ((hi << 21) | (lo >> 11)) * 2.0 ** -53
    d = oracle.draws(seed, chain, it, block, [0, 1, 0, 0, 1, 1])    # u0 n0 u1 u2 n1 n2 — interleaving does not matter
    wu0 = oracle.philox([0, it, chain, block], key); wu1 = oracle.philox([1, it, chain, block], key)
    wn0 = oracle.philox([0, it, chain, block | (1 << 24)], key); wn1 = oracle.philox([1, it, chain, block | (1 << 24)], key)
    assert d[0] == u53(wu0[0], wu0[1]) and d[2] == u53(wu0[2], wu0[3]) and d[3] == u53(wu1[0], wu1[1])
    rad0 = np.sqrt(-2 * np.log(1 - u53(wn0[0], wn0[1]))); ang0 = 2 * np.pi * u53(wn0[2], wn0[3])
    rad1 = np.sqrt(-2 * np.log(1 - u53(wn1[0], wn1[1]))); ang1 = 2 * np.pi * u53(wn1[2], wn1[3])
    assert d[1] == pytest.approx(rad0 * np.cos(ang0), rel=1e-14, abs=1e-15)
    assert d[4] == pytest.approx(rad0 * np.sin(ang0), rel=1e-14, abs=1e-15)
    assert d[5] == pytest.approx(rad1 * np.cos(ang1), rel=1e-14, abs=1e-15)
    u = oracle.draws(1, 0, 1, 0, [0] * 20000)
    z = oracle.draws(1, 0, 1, 0, [1] * 20000)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03 and st.kstest(z, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 0.03 and abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 0.03


def test_line_closed_form_known_answers(oracle):
    # SURVEY App. D; formulas from doc/samplers/amwg.jl:17-25 and doc/samplers/nuts.jl:25-29
    assert oracle.line_logf([0.0, 0.0, 0.0]) == pytest.approx(-26.501, abs=1e-12)
    assert oracle.line_logf([0.6, 0.8, 0.0]) == pytest.approx(-0.8015, abs=1e-12)
    v, g = oracle.line_logf([1.0, 0.5, -0.5], grad=True)
    assert v == pytest.approx(-1.84312610383344, rel=1e-13)
    np.testing.assert_allclose(g, [4.12080318, 17.31107334, 0.5920011], rtol=1e-8)
    _, g = oracle.line_logf([0.0, 0.0, 0.0], grad=True)
    np.testing.assert_allclose(g, [15.0, 53.0, 24.0], rtol=1e-13)


def test_line_model_density_equals_reference_closed_form_up_to_a_constant(oracle):
    # the model-based block densities (logpdf! over the DAG) must differ from the doc's stand-alone logf by a constant
    o = oracle.Oracle("line")
    o.set_scheme([dict(kind="nuts", nodes=[0, 1])])
    rng = np.random.default_rng(0)
    th = rng.normal(size=(50, 3))
    states = np.column_stack([th[:, 0], th[:, 1], np.exp(th[:, 2])])
    lp = o.logpdf(0, states)
    ref = np.array([oracle.line_logf(t) for t in th])
    diff = lp - ref
    np.testing.assert_allclose(diff, diff[0], rtol=0, atol=1e-10)
    # the constant is the sum of the normalising constants dropped in the doc: 5 + 2 normal factors and the IG prior
    const = -3.5 * np.log(2 * np.pi) - np.log(1000.0) + 0.001 * np.log(0.001) - sp.gammaln(0.001)
    assert diff[0] == pytest.approx(const, abs=1e-10)
    lpg, g = o.gradlogpdf(0, states, mode=0)
    gref = np.array([oracle.line_logf(t, grad=True)[1] for t in th])
    np.testing.assert_allclose(g, gref, rtol=1e-10, atol=1e-10)


def test_node_level_known_answers(oracle):
    # SURVEY App. D (hand-derivable; computed with scipy)
    o = oracle.Oracle("line")
    o.set_scheme([dict(kind="amwg", nodes=[0], scale=1.0), dict(kind="slice_multi", nodes=[1], scale=1.0, transform=1)])
    s = np.array([0.6, 0.8, 1.0])
    lik, pb, ps2 = -5.3946926660233645, -8.746132345391482, -6.9150866406628335
    assert o.logpdf(0, s)[0] == pytest.approx(pb + lik, rel=1e-14)
    assert o.logpdf(1, s)[0] == pytest.approx(ps2 + lik, rel=1e-14)
    s = np.array([0.0, 0.0, 2.0])
    assert o.logpdf(0, s)[0] == pytest.approx(-8.745632345391481 + -19.577560617423224, rel=1e-14)
    assert o.logpdf(1, s)[0] == pytest.approx(-6.915279787843396 + -19.577560617423224, rel=1e-14)
    # seeds at the reference inits (doc/examples/seeds.jl:60-65)
    o = oracle.Oracle("seeds")
    o.set_scheme([dict(kind="amwg", nodes=[0, 1, 2, 3], scale=0.1), dict(kind="amwg", nodes=[5], scale=0.01), dict(kind="amwg", nodes=[4], scale=0.1)])
    binom, apri = -87.83175502138508, -31.306775248747236
    s = helpers.SEEDS_INITS[0]
    assert o.logpdf(0, s)[0] == pytest.approx(apri + binom, rel=1e-13)
    assert o.logpdf(1, s)[0] == pytest.approx(29.05657775557683 + binom, rel=1e-13)
    assert o.logpdf(2, s)[0] == pytest.approx(-7.009481470476847 + 29.05657775557683, rel=1e-13)
    s = helpers.SEEDS_INITS[1]
    assert o.logpdf(2, s)[0] == pytest.approx(-6.9150866406628335 + -19.29770919729813, rel=1e-13)
    # pumps at alpha = beta = 1, theta = y / t
    o = oracle.Oracle("pumps")
    o.set_scheme([dict(kind="slice_uni", nodes=[0, 1], scale=1.0), dict(kind="slice_uni", nodes=[2], scale=1.0)])
    y = np.array([5, 1, 5, 14, 3, 19, 1, 1, 4, 22.0]); t = np.array([94.3, 15.7, 62.9, 126, 5.24, 31.4, 1.05, 1.05, 2.1, 10.5])
    s = np.concatenate([[1.0, 1.0], y / t])
    assert o.logpdf(0, s)[0] == pytest.approx(-1.0 + -3.252712651734206 + -7.3896954340746515, rel=1e-13)
    assert o.logpdf(1, s)[0] == pytest.approx(-7.3896954340746515 + -16.717612879429662, rel=1e-13)


def test_distribution_formulas_against_scipy(oracle):
    # the Distributions.jl log densities the oracle restates, via block densities of the templates
    rng = np.random.default_rng(5)
    o = oracle.Oracle("pumps")
    o.set_scheme([dict(kind="slice_uni", nodes=[0], scale=1.0), dict(kind="slice_uni", nodes=[1], scale=1.0), dict(kind="slice_uni", nodes=[2], scale=1.0)])
    y = np.array([5, 1, 5, 14, 3, 19, 1, 1, 4, 22.0]); t = np.array([94.3, 15.7, 62.9, 126, 5.24, 31.4, 1.05, 1.05, 2.1, 10.5])
    for _ in range(10):
        a, b = rng.gamma(2.0, 1.0, 2); th = rng.gamma(1.0, 1.0, 10)
        s = np.concatenate([[a, b], th])
        lth = st.gamma.logpdf(th, a, scale=1 / b).sum()
        assert o.logpdf(0, s)[0] == pytest.approx(st.expon.logpdf(a) + lth, rel=1e-12)
        assert o.logpdf(1, s)[0] == pytest.approx(st.gamma.logpdf(b, 0.1, scale=1.0) + lth, rel=1e-12)
        assert o.logpdf(2, s)[0] == pytest.approx(lth + st.poisson.logpmf(y, th * t).sum(), rel=1e-12)
    o = oracle.Oracle("seeds")
    o.set_scheme([dict(kind="amwg", nodes=[5], scale=0.01), dict(kind="amwg", nodes=[4], scale=0.1)])
    r = np.array([10, 23, 23, 26, 17, 5, 53, 55, 32, 46, 10, 8, 10, 8, 23, 0, 3, 22, 15, 32, 3.0])
    n = np.array([39, 62, 81, 51, 39, 6, 74, 72, 51, 79, 13, 16, 30, 28, 45, 4, 12, 41, 30, 51, 7.0])
    x1 = np.array([0.0] * 11 + [1.0] * 10); x2 = np.array([0.0] * 5 + [1.0] * 6 + [0.0] * 5 + [1.0] * 5)
    for _ in range(10):
        al = rng.normal(size=4); s2 = rng.gamma(1.0, 0.5); b = rng.normal(scale=np.sqrt(s2), size=21)
        s = np.concatenate([al, [s2], b])
        eta = al[0] + al[1] * x1 + al[2] * x2 + al[3] * x1 * x2 + b
        lb = st.norm.logpdf(b, 0, np.sqrt(s2)).sum()
        assert o.logpdf(0, s)[0] == pytest.approx(lb + st.binom.logpmf(r, n, sp.expit(eta)).sum(), rel=1e-11)
        assert o.logpdf(1, s)[0] == pytest.approx(st.invgamma.logpdf(s2, 0.001, scale=0.001) + np.log(s2) + lb, rel=1e-11)


def test_out_of_support_and_early_exit(oracle):
    # logpdf_sub returns -Inf outside the support (distributionstruct.jl:138-140); logpdf! stops at the first
    # non-finite partial sum (simulation.jl:64,84) so the targets' closures never see sqrt(negative)
    o = oracle.Oracle("rats")
    tpl, blocks, inits = helpers.scheme("rats_slice_amwg")
    o.set_scheme([helpers.oracle_block(b) for b in blocks])
    assert np.isneginf(o.logpdf(0, inits[0], np.array([[-1.0]])))[0]          # s2_c < 0, Slice on the constrained scale
    assert np.isneginf(o.logpdf(2, inits[0], np.array([[150.0, -3.0]])))[0]   # (mu_alpha, s2_alpha < 0)
    assert np.isfinite(o.logpdf(2, inits[0], np.array([[150.0, 3.0]])))[0]


@pytest.mark.parametrize("name", ["line_nuts_all", "seeds_amwg", "rats_nuts_slice", "pumps_amwg_nuts", "rats_slice_amwg"])
def test_analytic_gradient_against_finite_differences(oracle, name):
    # the hand-derived joint gradients (engine mode) against the reference's Calculus.gradient restatement
    tpl, blocks, inits = helpers.scheme(name)
    o = oracle.Oracle(tpl)
    o.set_scheme([helpers.oracle_block(b) for b in blocks])
    rng = np.random.default_rng(2)
    st_ = np.repeat(inits, 4, axis=0)
    st_ = st_ * np.exp(rng.normal(scale=0.05, size=st_.shape)) + (st_ == 0) * rng.normal(scale=0.3, size=st_.shape)
    for b in range(len(blocks)):
        lp, ga = o.gradlogpdf(b, st_, mode=0)
        _, gc = o.gradlogpdf(b, st_, mode=2)
        _, gf = o.gradlogpdf(b, st_, mode=1)
        scale = 1.0 + np.abs(ga)
        assert np.all(np.abs(gc - ga) / scale < 1e-6 * np.maximum(1, np.abs(lp))[:, None] ** 0.5 + 1e-5)
        assert np.all(np.abs(gf - ga) / scale < 1e-6 * np.maximum(1, np.abs(lp))[:, None] + 1e-3)


def test_glm_template(oracle):
    X, y, _ = helpers.glm_data(N=150, d=6)
    o = oracle.Oracle("glm", glm_d=6)
    o.set_data("X", X); o.set_data("y", y)
    o.set_scheme([dict(kind=4, nodes=[0])])
    be = np.random.default_rng(1).normal(size=(5, 6))
    lp, g = o.gradlogpdf(0, be, mode=0)
    eta = be @ X.T
    ref = (y * eta - np.logaddexp(0, eta)).sum(axis=1) + st.norm.logpdf(be, 0, np.sqrt(1000.0)).sum(axis=1)
    np.testing.assert_allclose(lp, ref, rtol=1e-12)
    gref = (y - sp.expit(eta)) @ X - be / 1000.0
    np.testing.assert_allclose(g, gref, rtol=1e-11, atol=1e-11)


def test_special_functions_against_scipy(oracle):
    L = oracle.lib()
    for x in [0.001, 0.3, 1.0, 2.5, 10.0, 123.4]:
        assert L.orc_digamma(x) == pytest.approx(sp.digamma(x), rel=1e-13, abs=1e-13)
    for d1, d2, q in [(1, 10.5, 0.975), (2, 3000.0, 0.975), (7, 47.3, 0.9), (1, 1e7, 0.975), (63, 250000.0, 0.975), (3, 0.7, 0.5)]:
        assert L.orc_fquantile(q, d1, d2) == pytest.approx(st.f.ppf(q, d1, d2), rel=1e-7)   # PSRF is rounded to 3 dp
    assert L.orc_fquantile(0.975, 4.0, float("inf")) == pytest.approx(st.chi2.ppf(0.975, 4) / 4, rel=1e-9)


def _np_gelman(c, alpha=0.05):
    """Independent numpy restatement of gelmandiag.jl:5-47 for checking the oracle."""
    n, p, m = c.shape
    out = np.empty((p, 2))
    for j in range(p):
        x = c[:, j, :]
        s2 = x.var(axis=0, ddof=1); psibar = x.mean(axis=0)
        w = s2.mean(); b = n * psibar.var(ddof=1)
        var_w = s2.var(ddof=1) / m; var_b = 2 * b * b / (m - 1)
        cov = lambda a, bb: np.cov(a, bb, ddof=1)[0, 1]
        var_wb = n / m * (cov(s2, psibar ** 2) - 2 * psibar.mean() * cov(s2, psibar))
        V = (n - 1) / n * w + (m + 1) / (m * n) * b
        var_V = ((n - 1) ** 2 * var_w + ((m + 1) / m) ** 2 * var_b + 2 * (n - 1) * (m + 1) / m * var_wb) / n ** 2
        df = 2 * V * V / var_V; W_df = 2 * w * w / var_w
        corr = (df + 3) / (df + 1); Rf = (n - 1) / n; Rr = (m + 1) / (m * n) * b / w
        out[j, 0] = np.sqrt(corr * (Rf + Rr))
        out[j, 1] = np.sqrt(corr * (Rf + Rr * st.f.ppf(1 - alpha / 2, m - 1, W_df)))
    return out


def ar1_chains(n, p, m, rho, seed):
    rng = np.random.default_rng(seed)
    c = np.empty((n, p, m))
    for j in range(p):
        e = rng.normal(size=(n, m))
        x = np.empty((n, m)); x[0] = e[0]
        for i in range(1, n):
            x[i] = rho * x[i - 1] + np.sqrt(1 - rho * rho) * e[i]
        c[:, j, :] = (j + 1) * x + 3.0 * j + 0.1 * rng.normal(size=m)
    return c


def test_gelmandiag_against_numpy(oracle):
    c = ar1_chains(400, 3, 5, 0.6, 1)
    np.testing.assert_allclose(oracle.gelmandiag(c), _np_gelman(c), rtol=1e-9)
    with pytest.raises(ValueError, match="less than 2 chains"):
        oracle.gelmandiag(c[:, :, :1])
    # link(c): log for all-positive columns, logit for (0,1) columns (chains.jl:237-246)
    cp = np.exp(c * 0.1); cp[:, 2, :] = 1 / (1 + np.exp(-c[:, 2, :] * 0.1))
    lk = cp.copy(); lk[:, :2, :] = np.log(cp[:, :2, :]); lk[:, 2, :] = np.log(cp[:, 2, :] / (1 - cp[:, 2, :]))
    np.testing.assert_allclose(oracle.gelmandiag(cp, linkcode=[-1, -1, -1]), _np_gelman(lk), rtol=1e-9)


def test_summarystats_against_numpy(oracle):
    c = ar1_chains(450, 2, 3, 0.5, 2)
    ss = oracle.summarystats(c, 0, 100)
    for j in range(2):
        x = c[:, j, :].T.ravel()     # vec(x): chain-major
        mb = x[: (x.size // 100) * 100].reshape(-1, 100).mean(axis=1)     # batches straddle chains when 100 does not divide n
        mcse = mb.std(ddof=1) / np.sqrt(mb.size)
        np.testing.assert_allclose(ss[j], [x.mean(), x.std(ddof=1), x.std(ddof=1) / np.sqrt(x.size), mcse,
                                           min((x.std(ddof=1) / mcse) ** 2, 450)], rtol=1e-10)
    si = oracle.summarystats(c, 1)
    x = c[:, 0, :].T.ravel(); z = x - x.mean(); N = x.size
    ac = lambda k: (z[: N - k] * z[k:]).sum() / N
    G = ac(0) + ac(1); val = -ac(0) + 2 * G
    for i in range(1, (N - 2) // 2 + 1):
        G = min(G, ac(2 * i) + ac(2 * i + 1))
        if not G > 0:
            break
        val += 2 * G
    assert si[0, 3] == pytest.approx(np.sqrt(val / N), rel=1e-10)


def test_sampler_draw_order_is_the_references(oracle):
    # AMWG consumes n normals up front, then one uniform per component (amwg.jl:102-107).  External stream: a normal
    # consumes 2 entries, so one iteration of a 3-component block reads 3 x (ua, ub) then 3 uniforms.
    MODE = np.array([[0.6, 0.8, 1.0]])   # near the posterior mode: a 5-sigma proposal is always worse
    o = oracle.Oracle("line")
    o.set_scheme([dict(kind="amwg", nodes=[0, 1], scale=1.0)])
    far = [1 - 1e-6, 0.0]            # box_muller -> sqrt(-2 log 1e-6) = 5.26: a proposal far in the tail
    u = np.tile(np.concatenate([np.tile(far, 3), [0.999999999] * 3]), 5)[None, :]
    out, fin, tune = o.run(1, MODE, 5, ext_u=u)
    np.testing.assert_array_equal(fin[0, :2], MODE[0, :2])      # every proposal rejected, v restored exactly
    assert tune[0, 0] == 5 and (tune[0, 5:8] == 0).all()
    # uniforms = 0 accept (almost) everything: the accept counters fill up and beta[1] moved by count x 5.26
    u2 = np.tile(np.concatenate([np.tile(far, 3), [0.0] * 3]), 5)[None, :]
    _, fin2, tune2 = o.run(1, MODE, 5, ext_u=u2)
    assert (tune2[0, 5:8] >= 4).all()       # (u = 0 rejects only when exp(delta) underflows to 0)
    z = np.sqrt(-2 * np.log(1e-6))
    np.testing.assert_allclose(fin2[0, 0], MODE[0, 0] + tune2[0, 5] * z, rtol=1e-9)
    # misordered stream (uniforms first) gives a different result: the order is observable
    u3 = np.tile(np.concatenate([[0.0] * 3, np.tile(far, 3)]), 5)[None, :]
    _, fin3, _ = o.run(1, MODE, 5, ext_u=u3)
    assert not np.allclose(fin3, fin2)


def test_engine_bookkeeping(oracle):
    # thinning rule i > burnin && (i - burnin) % thin == 0 (mcmc.jl:76-78) and the [kept x p x chains] layout
    tpl, blocks, inits = helpers.scheme("line_amwg_slice")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, fin, _ = o.run(3, inits, 50, burnin=11, thin=4, seed=3)
    assert out.shape == ((50 - 11) // 4, 3, 3) and not np.isnan(out).any()
    out1, fin1, _ = o.run(3, inits, 47, burnin=11, thin=4, seed=3)       # last kept iteration is 47
    np.testing.assert_array_equal(out[-1], np.column_stack([fin1[:, 0], fin1[:, 1], fin1[:, 2]]).T)
    with pytest.raises(RuntimeError, match="burnin is greater than or equal to iters"):
        o.run(1, inits, 10, burnin=10)
    # chains are independent of how they are sharded (global chain id keys the stream)
    a, _, _ = o.run(4, inits, 30, seed=9)
    b, _, _ = o.run(2, inits, 30, seed=9, chain_offset=2)
    np.testing.assert_array_equal(a[:, :, 2:], b)
    c, _, _ = o.run(4, inits, 30, seed=9, nthreads=3)
    np.testing.assert_array_equal(a, c)


def test_node_logpdf_is_the_sum_of_scipy_densities(oracle):
    """logpdf(m, nodekeys) (src/model/simulation.jl:60-67, used by dic / logpdf(mc, ...) in src/output/modelstats.jl): the observed
    node of the line model (y ~ MvNormal(xmat * beta, sqrt(s2)), doc/tutorial/line.jl:6-12) and of pumps (y[i] ~ Poisson(theta[i] t[i]),
    doc/examples/pumps.jl:14-20), and every node of pumps, against scipy.stats."""
    import helpers
    import scipy.stats as st
    rng = np.random.default_rng(11)
    tpl, blocks, inits = helpers.scheme("line_amwg_slice")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    S = np.column_stack([rng.normal(size=(6, 2)), rng.gamma(2.0, 1.0, size=6)])
    x = np.arange(1.0, 6.0); y = np.array([1.0, 3, 3, 3, 5])
    want_y = np.array([st.norm.logpdf(y, s[0] + s[1] * x, np.sqrt(s[2])).sum() for s in S])
    np.testing.assert_allclose(o.logpdf_nodes(0b100, S), want_y, rtol=1e-12)
    want_all = want_y + np.array([st.norm.logpdf(s[:2], 0, np.sqrt(1000.0)).sum() + st.invgamma.logpdf(s[2], 0.001, scale=0.001) for s in S])
    np.testing.assert_allclose(o.logpdf_nodes(0b111, S), want_all, rtol=1e-11)
    tpl, blocks, inits = helpers.scheme("pumps_slice")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    yp = np.array([5, 1, 5, 14, 3, 19, 1, 1, 4, 22.0]); t = np.array([94.3, 15.7, 62.9, 126, 5.24, 31.4, 1.05, 1.05, 2.1, 10.5])   # pumps.jl:4-9
    S = np.column_stack([rng.gamma(2.0, 0.5, size=5), rng.gamma(2.0, 0.5, size=5), rng.gamma(2.0, 0.3, size=(5, 10))])
    want_y = np.array([st.poisson.logpmf(yp, s[2:] * t).sum() for s in S])
    np.testing.assert_allclose(o.logpdf_nodes(0b1000, S), want_y, rtol=1e-12)
    want_all = want_y + np.array([st.expon.logpdf(s[0], scale=1.0) + st.gamma.logpdf(s[1], 0.1, scale=1.0)
                                  + st.gamma.logpdf(s[2:], s[0], scale=1.0 / s[1]).sum() for s in S])
    np.testing.assert_allclose(o.logpdf_nodes(0b1111, S), want_all, rtol=1e-11)


def test_rwm_proposal_kernels_have_the_right_distribution(oracle):
    """RWM proposals (src/samplers/rwm.jl:65-71 draws rand(proposal(0, 1)); SymDistributionType, src/distributions/extensions.jl:51-53):
    support [-1, 1], variance and CDF of the Cosine, Epanechnikov, Biweight and Triweight kernels (Kolmogorov-Smirnov against the closed forms)."""
    import scipy.stats as st
    cdf = {3: lambda z: 0.5 * (1 + z + np.sin(np.pi * z) / np.pi),
           4: lambda z: 0.5 + 0.75 * (z - z ** 3 / 3),
           5: lambda z: 0.5 + (15.0 / 16.0) * (z - 2 * z ** 3 / 3 + z ** 5 / 5),
           6: lambda z: 0.5 + (35.0 / 32.0) * (z - z ** 3 + 3 * z ** 5 / 5 - z ** 7 / 7)}
    var = {3: 1.0 / 3.0 - 2.0 / np.pi ** 2, 4: 0.2, 5: 1.0 / 7.0, 6: 1.0 / 9.0}
    for code in (3, 4, 5, 6):
        z = oracle.rwm_draws(code, 99, 200000)
        assert z.min() >= -1.0 and z.max() <= 1.0
        assert abs(z.mean()) < 4 * np.sqrt(var[code] / z.size)
        np.testing.assert_allclose(z.var(), var[code], rtol=0.02)
        assert st.kstest(z, cdf[code]).pvalue > 1e-3


def test_predict_draws_are_the_quantile_functions_of_the_observed_nodes(oracle):
    """predict(mc) (src/output/modelstats.jl:63-96) on the engine's stream contract (seed, chain = stream id, iteration = record index,
    block 0, kind 15): a discrete element is the inverse CDF of ONE uniform (sequential search from 0 == scipy's ppf), a Laplace element
    its closed-form quantile, a Normal element mu + sigma z with z from the normal stream."""
    import helpers
    import scipy.stats as st
    rng = np.random.default_rng(21)
    seed, stream = 77, 3
    # pumps: y[i] ~ Poisson(theta[i] t[i])
    tpl, blocks, inits = helpers.scheme("pumps_slice")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    S = np.column_stack([rng.gamma(2.0, 0.5, size=6), rng.gamma(2.0, 0.5, size=6), rng.gamma(2.0, 0.4, size=(6, 10))])
    t = np.array([94.3, 15.7, 62.9, 126, 5.24, 31.4, 1.05, 1.05, 2.1, 10.5])
    got = o.predict(S, seed, stream_id=stream)
    for i in range(6):
        u = oracle.draws(seed, stream, i, 0, [0] * 10, kind=15)
        np.testing.assert_array_equal(got[i], st.poisson.ppf(u, S[i, 2:] * t))
    # surgical: r[i] ~ Binomial(n[i], invlogit(b[i]))
    tpl, blocks, inits = helpers.scheme("surgical_amwg")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    S = np.column_stack([rng.normal(-2.5, 0.2, 5), rng.gamma(2.0, 0.1, 5), rng.normal(-2.5, 0.5, (5, 12))])
    n = np.array([47, 148, 119, 810, 211, 196, 148, 215, 207, 97, 256, 360])
    got = o.predict(S, seed, stream_id=stream)
    for i in range(5):
        u = oracle.draws(seed, stream, i, 0, [0] * 12, kind=15)
        want = st.binom.ppf(u, n, 1 / (1 + np.exp(-S[i, 2:])))
        assert (got[i] == want).mean() >= 11 / 12          # a uniform within rounding of a CDF step may fall on either side
    # stacks: y[i] ~ Laplace(mu[i], s2)
    tpl, blocks, inits = helpers.scheme("stacks_amwg")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    S = np.column_stack([rng.normal(17.5, 1, 4), rng.normal(0, 2, (4, 3)), rng.gamma(6, 0.45, 4)])
    got = o.predict(S, seed, stream_id=stream)
    for i in range(4):
        u = oracle.draws(seed, stream, i, 0, [0] * 21, kind=15)
        resid = st.laplace.ppf(u, 0.0, S[i, 4])
        # y_rep - resid = mu[i] = beta0 + z . beta: the same vector for any stream
        other = o.predict(S[i:i + 1], seed, stream_id=stream + 1)[0] - st.laplace.ppf(oracle.draws(seed, stream + 1, 0, 0, [0] * 21, kind=15), 0.0, S[i, 4])
        np.testing.assert_allclose(got[i] - resid, other, rtol=1e-10, atol=1e-10)
    # line: y[i] ~ Normal(beta1 + beta2 x[i], sqrt(s2)) through the normal stream
    tpl, blocks, inits = helpers.scheme("line_amwg_slice")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    S = np.column_stack([rng.normal(size=(3, 2)), rng.gamma(2.0, 1.0, 3)])
    got = o.predict(S, seed, stream_id=stream)
    x = np.arange(1.0, 6.0)
    for i in range(3):
        z = oracle.draws(seed, stream, i, 0, [1] * 5, kind=15)
        np.testing.assert_allclose(got[i], S[i, 0] + S[i, 1] * x + np.sqrt(S[i, 2]) * z, rtol=1e-13)
