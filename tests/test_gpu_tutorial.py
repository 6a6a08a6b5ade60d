"""The reference's tutorial (doc/tutorial/line.jl, output in doc/tutorial.rst:380-570) run through the host-side mirror of its
interface on the device engine: model, scheme, mcmc, convergence diagnostics, posterior summaries, DIC, subsetting, file I/O and
restart — the calls a user of the reference makes, in the order the tutorial makes them.  The published tables are the
reference's own output (3 chains): posterior summaries are compared within Monte Carlo error."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LINE = dict(x=[1, 2, 3, 4, 5], y=[1, 3, 3, 3, 5])


@pytest.fixture(scope="module")
def sim1():
    from mambacuda import api
    model = api.Model("line")                                                # doc/tutorial/line.jl:5-25
    api.setsamplers(model, [api.NUTS("beta"), api.Slice("s2", 3.0)])         # scheme1, line.jl:48-49
    rng = np.random.default_rng(123)
    inits = [dict(beta=rng.normal(0, 1, 2), s2=rng.gamma(1.0, 1.0)) for _ in range(3)]   # line.jl:82-89
    return api.mcmc(model, LINE, inits, 10000, burnin=250, thin=2, chains=3)  # line.jl:95


def test_tutorial_output_shape_and_summary(sim1):
    from mambacuda import api
    assert sim1.header() == "Iterations = 252:10000\nThinning interval = 2\nChains = 1,2,3\nSamples per chain = 4875\n"   # tutorial.rst:427-430
    assert sim1.names == ["beta[1]", "beta[2]", "s2"] and sim1.value.shape == (4875, 3, 3)
    (ss, names, cols), (qq, _, qcols) = api.describe(sim1)
    ref = np.array([[0.5971183, 1.14894446, 0.016925598], [0.8017036, 0.34632566, 0.004793345], [1.2203777, 2.00876760, 0.101798287]])
    for j in range(3):   # tutorial.rst:432-436: mean within 4 combined MCSE; SD of beta within 10 % (s2 is heavy-tailed: SD not compared)
        assert abs(ss[j, 0] - ref[j, 0]) < 4 * np.hypot(ref[j, 2], ss[j, 3]), (names[j], ss[j], ref[j])
    np.testing.assert_allclose(ss[:2, 1], ref[:2, 1], rtol=0.1)
    np.testing.assert_allclose(ss[:, 2], ss[:, 1] / np.sqrt(3 * 4875), rtol=1e-12)             # naive SE: stats.jl:88
    assert (ss[:, 4] <= 4875).all()                                                            # ESS capped at the iterations: stats.jl:91
    qref = np.array([[-1.74343373, 0.026573102, 0.59122696, 1.1878720, 2.8308472], [0.12168742, 0.628297573, 0.80357822, 0.9719441, 1.5051573],
                     [0.17091385, 0.383671702, 0.65371989, 1.2206381, 6.0313970]])
    np.testing.assert_allclose(qq[:2, 1:4], qref[:2, 1:4], atol=0.06)                          # quartiles and median, tutorial.rst:438-442
    np.testing.assert_allclose(qq[2, 1:4], qref[2, 1:4], rtol=0.12)                            # s2: ESS ~ 400 in the reference's own run
    assert qcols == ["2.5%", "25.0%", "50.0%", "75.0%", "97.5%"]


def test_tutorial_convergence_diagnostics(sim1):
    from mambacuda import api
    psrf, names, cols = api.gelmandiag(sim1, mpsrf=True, transform=True)     # line.jl:105, tutorial.rst:388-393: all 1.00x
    assert names == ["beta[1]", "beta[2]", "s2", "Multivariate"] and cols == ["PSRF", "97.5%"]
    assert (psrf[:, 0] < 1.02).all() and np.isnan(psrf[3, 1])
    np.testing.assert_allclose(api.gelmandiag(sim1, transform=True)[0], psrf[:3], atol=1.1e-3)   # device streaming moments vs host array
    z, _, zc = api.gewekediag(sim1)                                           # line.jl:108
    assert z.shape == (3, 2, 3) and zc == ["Z-score", "p-value"] and ((z[:, 1, :] >= 0) & (z[:, 1, :] <= 1)).all()
    h, _, hc = api.heideldiag(sim1)                                           # line.jl:111, tutorial.rst: burn-in 251, all tests passed for beta
    assert h.shape == (3, 6, 3) and (h[:2, 1, :] == 1).all() and (h[:2, 0, :] >= 251).all()
    r, _, rc = api.rafterydiag(sim1)                                          # line.jl:114, tutorial.rst: thinning 2, Nmin 3746
    assert r.shape == (3, 5, 3) and (r[:, 0, :] >= 2).all() and np.allclose(r[:, 3, :], 3746.0)


def test_tutorial_posterior_statistics(sim1):
    from mambacuda import api
    iv, _, cols = api.hpd(sim1)                                               # tutorial.rst:447-450
    np.testing.assert_allclose(iv[:2], [[-1.75436235, 2.8109571], [0.09721501, 1.4733163]], atol=0.35)   # 3 chains, heavy tails: the end points move by ~0.2 between runs
    cm, _, _ = api.cor(sim1)                                                  # tutorial.rst:455-458
    assert abs(cm[0, 1] + 0.905245029) < 0.02 and np.allclose(np.diag(cm), 1.0) and abs(cm[0, 2]) < 0.1
    ac, _, lags = api.autocor(sim1)                                           # tutorial.rst:463-466: lags are in iterations (x thinning)
    assert lags == ["Lag 2", "Lag 10", "Lag 20", "Lag 100"] and ac.shape == (3, 4, 3)
    assert (ac[2, 0, :] > 0.7).all() and (np.abs(ac[:2, 1:, :]) < 0.15).all()                  # s2 mixes slowly, beta does not
    cr, names, _ = api.changerate(sim1)                                       # tutorial.rst:481-485: beta 0.844 (NUTS keeps the point), s2 1.0
    assert names[-1] == "Multivariate" and abs(cr[0] - 0.844) < 0.04 and cr[0] == cr[1] and cr[2] == 1.0 and cr[3] == 1.0
    d, rows, cols = api.dic(sim1)                                             # tutorial.rst:490-492: pD 13.83 / 1.166, pV 22.62 / 5.56
    assert rows == ["pD", "pV"] and cols == ["DIC", "Effective Parameters"]
    assert abs(d[0, 0] - 13.828540) < 1.0 and abs(d[0, 1] - 1.1661193) < 0.6 and 2.0 < d[1, 1] < 12.0
    assert abs(d[0, 0] - d[0, 1] - (13.828540 - 1.1661193)) < 0.6                               # mean deviance


def test_tutorial_deviance_matches_the_closed_form(sim1):
    # -2 logpdf(y | beta, s2) of the line model (doc/tutorial/line.jl:6-12: y ~ MvNormal(xmat * beta, sqrt(s2))) at every kept draw
    from mambacuda import api
    lp = api.logpdf(sim1, "y")
    assert lp.names == ["logpdf"] and lp.value.shape == (4875, 1, 3)
    x, y = np.array(LINE["x"], float), np.array(LINE["y"], float)
    b1, b2, s2 = sim1.value[:, 0, :], sim1.value[:, 1, :], sim1.value[:, 2, :]
    res = y[None, None, :] - b1[..., None] - b2[..., None] * x[None, None, :]
    want = -0.5 * 5 * np.log(2 * np.pi * s2) - 0.5 * (res ** 2).sum(axis=-1) / s2
    np.testing.assert_allclose(lp.value[:, 0, :], want, rtol=1e-12)


def test_posterior_predictive_draws(sim1):
    # predict(sim1) (src/output/modelstats.jl:63-96): y_rep[i] ~ Normal(beta1 + beta2 x_i, sqrt(s2)) at every kept draw
    from mambacuda import api
    pp = api.predict(sim1)
    assert pp.names == ["y[1]", "y[2]", "y[3]", "y[4]", "y[5]"] and pp.value.shape == (4875, 5, 3) and pp.header() == sim1.header()
    x = np.array(LINE["x"], float)
    mu = sim1.value[:, 0, :][:, None, :] + sim1.value[:, 1, :][:, None, :] * x[None, :, None]
    z = (pp.value - mu) / np.sqrt(sim1.value[:, 2, :])[:, None, :]          # standard normal if the draws are right
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert np.abs(np.median(pp.value, axis=(0, 2)) - (0.6 + 0.8 * x)).max() < 0.15   # posterior predictive centre = fitted line
    with pytest.raises(api.ArgumentError, match="nodekeys are not all observed"):
        api.predict(sim1, "beta")


def test_tutorial_subsetting(sim1):
    from mambacuda import api
    sim = sim1[range(1000, 5001), ["beta[1]", "beta[2]"], None]              # line.jl:143, tutorial.rst:509-512
    assert sim.header() == "Iterations = 1000:5000\nThinning interval = 2\nChains = 1,2,3\nSamples per chain = 2001\n"
    np.testing.assert_array_equal(sim.value, sim1.value[374:2375, :2, :])
    ss, names, _ = api.summarystats(sim)
    assert names == ["beta[1]", "beta[2]"] and abs(ss[0, 0] - 0.6) < 0.15
    assert sim1[:, "beta", [1]].names == ["beta[1]", "beta[2]"]              # node keys select all of a node's elements


def test_tutorial_file_io_and_restart(sim1, tmp_path):
    from mambacuda import api
    f = str(tmp_path / "sim1.npz")
    api.write(f, sim1)                                                        # line.jl:148-149
    back = api.read(f, api.ModelChains)
    np.testing.assert_array_equal(back.value, sim1.value)
    assert back.header() == sim1.header() and back.model.iter == 10000
    sim = api.mcmc(back, 5000)                                                # line.jl:153, tutorial.rst:549-552
    assert sim.header() == "Iterations = 252:15000\nThinning interval = 2\nChains = 1,2,3\nSamples per chain = 7375\n"
    np.testing.assert_array_equal(sim.value[:4875], sim1.value)
    live = api.mcmc(sim1, 5000)                                               # the same restart on the handle that produced sim1
    np.testing.assert_array_equal(live.value, sim.value)
    ss, _, _ = api.summarystats(sim)
    ref = np.array([[0.59655228, 0.014053505], [0.80144540, 0.003954871], [1.18366563, 0.070481708]])   # tutorial.rst:556-558
    for j in range(3):
        assert abs(ss[j, 0] - ref[j, 0]) < 4 * np.hypot(ref[j, 1], ss[j, 3])
    with pytest.raises(api.ArgumentError, match="chain is missing its last iteration"):        # mcmc.jl:5-6
        api.mcmc(sim1[range(252, 9001), None, None], 100)


def test_tutorial_schemes_2_and_3():
    # line.jl:51-54, 97-102: scheme2 = [NUTS([:beta, :s2])], scheme3 = [Gibbs_beta, Gibbs_s2] — the user-defined conjugate samplers, which
    # the device engine provides as registered full conditionals (api.Gibbs); same posterior as scheme1 (tutorial.rst:432-436)
    from mambacuda import api
    rng = np.random.default_rng(123)
    inits = [dict(beta=rng.normal(0, 1, 2), s2=rng.gamma(1.0, 1.0)) for _ in range(3)]
    for scheme in ([api.NUTS(["beta", "s2"])], [api.Gibbs("beta"), api.Gibbs("s2")]):
        model = api.Model("line")
        api.setsamplers(model, scheme)
        sim = api.mcmc(model, LINE, inits, 10000, burnin=250, thin=2, chains=3)
        ss, names, _ = api.summarystats(sim)
        assert abs(ss[0, 0] - 0.5971183) < 4 * np.hypot(0.016925598, ss[0, 3]) and abs(ss[1, 0] - 0.8017036) < 4 * np.hypot(0.004793345, ss[1, 3])
        assert abs(ss[2, 0] - 1.2203777) < 4 * np.hypot(0.101798287, ss[2, 3]) + 0.1
    assert api.changerate(sim)[0][3] == 1.0          # a Gibbs sweep always moves


def test_store_false_keeps_only_streaming_moments():
    from mambacuda import api
    model = api.Model("line")
    api.setsamplers(model, [api.AMWG("beta", 1.0), api.Slice("s2", 5.0, transform=True)])
    rng = np.random.default_rng(5)
    inits = [dict(beta=rng.normal(0, 1, 2), s2=rng.gamma(1.0, 1.0)) for _ in range(64)]
    sim = api.mcmc(model, LINE, inits, 6000, burnin=1000, thin=1, chains=64, store=False)
    assert sim.value.shape == (0, 3, 64)
    ss, _, _ = api.summarystats(sim)
    assert abs(ss[0, 0] - 0.5971) < 0.06 and abs(ss[1, 0] - 0.8017) < 0.02
    assert (api.gelmandiag(sim, transform=True)[0][:, 0] < 1.05).all()
    with pytest.raises(api.ArgumentError, match="no stored samples"):
        api.summarystats(sim, etype="imse")
