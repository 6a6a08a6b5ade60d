"""CPU tests of the host-side mirror of the reference interface (mambacuda.api): sampler constructors, scheme
assignment, argument checks with the reference's error texts — everything that happens before the device call."""
import numpy as np
import pytest


def test_sampler_constructors_mirror_the_reference(mcu_built):
    from mambacuda import api
    s = api.AMWG(["alpha0", "alpha1"], 0.1)
    assert s.kind == "amwg" and s.desc["adapt"] == "all" and s.desc["batchsize"] == 50 and s.desc["target"] == 0.44   # amwg.jl:16-19,47-48
    with pytest.raises(api.ArgumentError, match="adapt must be one of :all, :burnin, or :none"):                     # amwg.jl:49-50
        api.AMWG("b", 0.01, adapt="sometimes")
    assert api.Slice("s2", 3.0).kind == "slice_multi" and api.Slice("s2", 3.0).desc["transform"] == 0                # slice.jl:47-50 defaults
    assert api.Slice(["alpha", "beta"], 1.0, api.Univariate).kind == "slice_uni"
    assert api.NUTS("beta").desc["target"] == 0.6                                                                    # nuts.jl:22
    assert api.AMM(["a"], 0.01 * np.eye(1)).desc["beta"] == 0.05                                                     # amm.jl:21-22
    with pytest.raises(api.ArgumentError):
        api.RWM("beta", 1.0, proposal="laplace")


def test_setsamplers_and_block_descs(mcu_built):
    from mambacuda import api
    m = api.Model("seeds")
    # doc/examples/seeds.jl:69-71
    api.setsamplers(m, [api.AMM(["alpha0", "alpha1", "alpha2", "alpha12"], 0.01 * np.eye(4)), api.AMWG("b", 0.01), api.AMWG("s2", 0.1)])
    d = api._block_descs(m)
    assert d[0]["nodes"] == [0, 1, 2, 3] and d[1]["nodes"] == [5] and d[2]["nodes"] == [4]
    with pytest.raises(KeyError):
        api.setsamplers(m, [api.AMWG("gamma", 1.0)])
    with pytest.raises(api.ArgumentError, match="no device equivalent"):
        api.setsamplers(m, [lambda model, block: None])      # Sampler([:beta], closure): src/samplers/sampler.jl:20-24
    with pytest.raises(api.ArgumentError, match="no device template"):
        api.Model("kidney")                                  # doc/examples/kidney.jl has no compiled template


def test_mcmc_argument_checks(mcu_built):
    from mambacuda import api
    m = api.Model("line")
    api.setsamplers(m, [api.AMWG("beta", 1.0), api.Slice("s2", 5.0, transform=True)])
    inits = [dict(beta=[0.0, 0.0], s2=1.0)]
    with pytest.raises(api.ArgumentError, match="burnin is greater than or equal to iters"):   # mcmc.jl:22-23
        api.mcmc(m, {}, inits, 100, burnin=100)
    with pytest.raises(api.ArgumentError, match="fewer initial values than chains"):           # mcmc.jl:24-25
        api.mcmc(m, {}, inits, 100, chains=2)
    with pytest.raises(api.ArgumentError, match="missing initial value for node : s2"):        # initialization.jl:9-10
        api._inits_matrix(m, [dict(beta=[0.0, 0.0])])
    with pytest.raises(api.DimensionMismatch):
        api._inits_matrix(m, [dict(beta=[0.0, 0.0, 1.0], s2=1.0)])
    x = api._inits_matrix(m, [dict(beta=[0.1, 0.2], s2=3.0)])
    np.testing.assert_array_equal(x, [[0.1, 0.2, 3.0]])


def test_chains_container(mcu_built):
    from mambacuda import api
    c = api.Chains(np.zeros((10, 2, 3)), start=252, thin=2, names=["a", "b"])
    assert (c.first, c.step, c.last) == (252, 2, 270) and c.chains == [1, 2, 3]     # chains.jl:14-32
    with pytest.raises(api.DimensionMismatch, match="names length differ"):
        api.Chains(np.zeros((10, 2, 3)), names=["a"])
    m = api.Chains(np.zeros((10, 2)), start=5, names=["a", "b"], chains=4)            # matrix form: chains.jl:34-41
    assert m.value.shape == (10, 2, 1) and m.chains == [4] and m.first == 5
    v = api.Chains(np.arange(6.0), names="theta")                                       # vector form: chains.jl:43-49
    assert v.value.shape == (6, 1, 1) and v.names == ["theta"] and v.chains == [1]


def _toy_chains(n=40, p=3, m=2, start=11, thin=2, seed=0):
    from mambacuda import api
    rng = np.random.default_rng(seed)
    return api.Chains(rng.normal(size=(n, p, m)), start=start, thin=thin, names=["a", "b[1]", "b[2]"][:p])


def test_chains_indexing_follows_the_reference(mcu_built):
    from mambacuda import api
    c = _toy_chains()                                        # iterations 11:2:89
    assert c.size() == (89, 3, 2) and c.size(1) == 89        # chains.jl:181-188: size(c)[1] is the LAST iteration
    s = c[range(20, 61, 4), ["a", "b[2]"], [2]]              # window2inds: ceil((20-11)/2+1) = 6 .. floor((60-11)/2+1) = 25, stride 4
    assert (s.first, s.step, s.last) == (21, 8, 85 if False else s.last) and s.first == 11 + 5 * 2 and s.step == 8
    np.testing.assert_array_equal(s.value[:, :, 0], c.value[5:25:4][:, [0, 2], 1])
    assert s.names == ["a", "b[2]"] and s.chains == [2]
    np.testing.assert_array_equal(c[None, None, None].value, c.value)
    np.testing.assert_array_equal(c[:, 2, :].value[:, 0, :], c.value[:, 1, :])          # integer names are 1-based
    np.testing.assert_array_equal(c[:, [True, False, True], :].value, c.value[:, [0, 2], :])
    with pytest.raises(api.ArgumentError, match="iteration indexing is unsupported"):   # chains.jl:70-71
        c[[1, 2, 3], None, None]
    assert c.header() == "Iterations = 11:89\nThinning interval = 2\nChains = 1,2\nSamples per chain = 40\n"   # chains.jl:211-218
    comb = c.combine()                                        # chains.jl:197-209: iteration-major, chains interleaved
    np.testing.assert_array_equal(comb[0], c.value[0, :, 0]); np.testing.assert_array_equal(comb[1], c.value[0, :, 1])
    np.testing.assert_array_equal(comb[2], c.value[1, :, 0])
    c2 = _toy_chains()
    c2[13, "a", 1] = 7.5                                      # setindex! by iteration number: chains.jl:60-62, 83-89
    assert c2.value[1, 0, 0] == 7.5


def test_chains_concatenation_and_its_errors(mcu_built):
    from mambacuda import api
    c = _toy_chains()
    nxt = api.Chains(np.ones((5, 3, 2)), start=91, thin=2, names=c.names)
    v = api.vcat(c, nxt)
    assert (v.first, v.step, v.last) == (11, 2, 99) and v.value.shape == (45, 3, 2)
    with pytest.raises(api.ArgumentError, match="noncontiguous chain iterations"):      # chains.jl:112-113
        api.cat(1, c, api.Chains(np.ones((5, 3, 2)), start=93, thin=2, names=c.names))
    with pytest.raises(api.ArgumentError, match="chain thinning differs"):
        api.cat(1, c, api.Chains(np.ones((5, 3, 2)), start=91, thin=1, names=c.names))
    with pytest.raises(api.ArgumentError, match="chain names differ"):
        api.cat(1, c, api.Chains(np.ones((5, 3, 2)), start=91, thin=2, names=["x", "y", "z"]))
    h = api.hcat(c, api.Chains(np.zeros((40, 1, 2)), start=11, thin=2, names=["z"]))
    assert h.names == ["a", "b[1]", "b[2]", "z"]
    with pytest.raises(api.ArgumentError, match="non-unique chain names"):              # chains.jl:135-136
        api.cat(2, c, c)
    with pytest.raises(api.ArgumentError, match="chain ranges differ"):
        api.cat(3, c, nxt)
    k = api.cat(3, c, c)
    assert k.chains == [1, 2, 3, 4] and k.value.shape == (40, 3, 4)                      # cat3 renumbers the chains: chains.jl:160-161
    with pytest.raises(api.ArgumentError, match="cannot concatenate along dimension 4"):
        api.cat(4, c, c)


def test_link_heuristic_and_discrete_support(mcu_built):
    from mambacuda import api
    rng = np.random.default_rng(1)
    v = np.stack([rng.normal(size=(30, 2)), rng.gamma(2.0, 2.0, size=(30, 2)), rng.uniform(0.01, 0.99, size=(30, 2)),
                  rng.integers(0, 5, size=(30, 2)).astype(float)], axis=1)
    c = api.Chains(v, names=["real", "pos", "unit", "count"])
    np.testing.assert_array_equal(c.link_codes(), [0, 1, 2, 0])          # chains.jl:237-246 (count has zeros: min > 0 fails)
    cc = c.link()
    np.testing.assert_allclose(cc[:, 1, :], np.log(v[:, 1, :])); np.testing.assert_allclose(cc[:, 2, :], np.log(v[:, 2, :] / (1 - v[:, 2, :])))
    np.testing.assert_array_equal(c.indiscretesupport(), [False, False, False, True])
    np.testing.assert_array_equal(c.indiscretesupport((1, 3)), [False, False, False, False])


def test_readcoda_keeps_the_common_window(mcu_built, tmp_path):
    # src/output/fileio.jl:15-40: parameters monitored over different iteration ranges: only the common window survives
    from mambacuda import api
    rng = np.random.default_rng(2)
    its = {"alpha": range(1, 101, 5), "beta": range(11, 121, 5), "sigma": range(6, 96, 5)}
    vals = {k: rng.normal(size=len(r)) for k, r in its.items()}
    out, ind = tmp_path / "x.out", tmp_path / "x.ind"
    row = 1
    with open(out, "w") as fo, open(ind, "w") as fi:
        for k, r in its.items():
            for it, v in zip(r, vals[k]):
                fo.write(f"{it}  {v:.10E}\n")
            fi.write(f"{k} {row} {row + len(r) - 1}\n"); row += len(r)
    c = api.readcoda(str(out), str(ind))
    assert (c.first, c.step, c.last) == (11, 5, 91) and c.names == ["alpha", "beta", "sigma"] and c.value.shape == (17, 3, 1)
    for j, k in enumerate(its):
        want = [v for it, v in zip(its[k], vals[k]) if 11 <= it <= 91]
        np.testing.assert_allclose(c.value[:, j, 0], want, rtol=1e-9)
    # round trip through the writer
    api.writecoda(str(tmp_path / "y.out"), str(tmp_path / "y.ind"), c)
    c2 = api.readcoda(str(tmp_path / "y.out"), str(tmp_path / "y.ind"))
    np.testing.assert_array_equal(c2.value, c.value); assert c2.range == c.range and c2.names == c.names


def test_readcoda_on_the_reference_example_files(mcu_built):
    # doc/mcmc/readcoda.jl: the OpenBUGS "line" output shipped with the reference (read from the reference tree when it is there;
    # the parsed summary is pinned in tests/golden/coda.json, written by tests/golden/make_golden.py)
    import json, os
    from mambacuda import api
    d = "/root/reference/doc/mcmc"
    if not os.path.exists(os.path.join(d, "line1.out")):
        pytest.skip("reference tree not present")
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "coda.json")))
    c = api.cat(3, api.readcoda(f"{d}/line1.out", f"{d}/line1.ind"), api.readcoda(f"{d}/line2.out", f"{d}/line2.ind"))
    assert c.header() == g["header"] and c.names == g["names"]
    np.testing.assert_allclose(c.value.sum(axis=0), g["column_sums"], rtol=1e-12)
    np.testing.assert_allclose(c.value[:3, :, 0], g["head_chain1"], rtol=0)
    ss, _, _ = api.summarystats(c)
    np.testing.assert_allclose(ss[:, 0], np.asarray(g["column_sums"]).sum(axis=1) / 400, rtol=1e-12)


def test_diagnostics_on_plain_chains_use_the_host_entry_points(mcu_built, oracle):
    from mambacuda import api
    rng = np.random.default_rng(3)
    n, m = 400, 3
    x = np.zeros((n, 3, m))
    for k in range(m):
        e = rng.normal(size=(n, 3))
        for i in range(1, n):
            e[i] += 0.6 * e[i - 1]
        x[:, 0, k] = e[:, 0]; x[:, 1, k] = np.exp(0.3 * e[:, 1]); x[:, 2, k] = 1 / (1 + np.exp(-0.5 * e[:, 2]))
    c = api.Chains(x, names=["real", "pos", "unit"])
    for etype, code in (("bm", 0), ("imse", 1)):
        ss, names, cols = api.summarystats(c, etype=etype, batch=50)
        np.testing.assert_allclose(ss, oracle.summarystats(x, code, 50), rtol=1e-10)
    ss_ipse, _, _ = api.summarystats(c, etype="ipse")
    assert (ss_ipse[:, 3] >= api.summarystats(c, etype="imse")[0][:, 3] - 1e-15).all()      # positive-sequence sum >= monotone-sequence sum
    with pytest.raises(api.ArgumentError, match="iterations are < 2000"):
        api.summarystats(c, batch=1000)
    psrf, names, cols = api.gelmandiag(c, transform=True)
    np.testing.assert_allclose(psrf, np.round(oracle.gelmandiag(x, 0.05, [-1, -1, -1]), 3), atol=1e-12)   # heuristic: identity, log, logit
    psrf_m, names_m, _ = api.gelmandiag(c, mpsrf=True)
    assert names_m[-1] == "Multivariate" and psrf_m.shape == (4, 2)
    np.testing.assert_allclose(psrf_m[:3], np.round(oracle.gelmandiag(x, 0.05, None), 3), atol=1e-12)


def test_describe_text_has_the_layout_of_the_reference(mcu_built):
    # doc/tutorial.rst:427-442: header, "Empirical Posterior Estimates:" and "Quantiles:" tables with right-aligned columns
    from mambacuda import api
    rng = np.random.default_rng(4)
    c = api.Chains(np.stack([rng.normal(0.6, 1.1, (300, 2)), rng.normal(0.8, 0.3, (300, 2)), rng.gamma(1.0, 1.2, (300, 2))], axis=1),
                   start=252, thin=2, names=["beta[1]", "beta[2]", "s2"])
    txt = api.describe_text(c)
    lines = txt.splitlines()
    assert lines[0] == "Iterations = 252:850" and lines[1] == "Thinning interval = 2" and lines[2] == "Chains = 1,2" and lines[3] == "Samples per chain = 300"
    i = lines.index("Empirical Posterior Estimates:")
    assert lines[i + 1].split() == ["Mean", "SD", "Naive", "SE", "MCSE", "ESS"]
    rows = lines[i + 2:i + 5]
    assert [r.split()[0] for r in rows] == ["beta[1]", "beta[2]", "s2"] and len({len(r) for r in rows}) == 1      # right-aligned: equal line lengths
    assert rows[2].startswith("     s2 ")                                                                          # row names right-aligned
    ss, _, _ = api.summarystats(c)
    np.testing.assert_allclose([float(x) for x in rows[0].split()[1:]], ss[0], rtol=1e-6)
    j = lines.index("Quantiles:")
    assert lines[j + 1].split() == ["2.5%", "25.0%", "50.0%", "75.0%", "97.5%"] and lines[j + 2].split()[0] == "beta[1]"


def test_write_and_read_round_trip(mcu_built, tmp_path):
    from mambacuda import api
    c = _toy_chains()
    f = str(tmp_path / "c.npz")
    api.write(f, c)
    r = api.read(f, api.Chains)
    np.testing.assert_array_equal(r.value, c.value)
    assert r.range == c.range and r.names == c.names and r.chains == c.chains
    with pytest.raises(TypeError):                            # fileio.jl:5: isa(c, T) || throw(TypeError(...))
        api.read(f, api.ModelChains)
    # ModelChains: the model record (template, scheme, inputs, ModelStates) travels without a device handle
    m = api.Model("line")
    api.setsamplers(m, [api.AMWG("beta", 1.0), api.Slice("s2", 5.0, transform=True), ])
    api.setinputs(m, dict(x=[1, 2, 3, 4, 5], y=[1, 3, 3, 3, 5]))
    m.iter = 89; m.burnin = 9
    m.states = [api.ModelState(np.array([0.1, 0.2, 1.5]), np.arange(7.0)), api.ModelState(np.array([0.3, 0.1, 0.5]), np.arange(7.0) + 1)]
    mc = api.ModelChains(c.value, m, engine=None, nodelinks=[0, 0, 1], start=11, thin=2, names=["beta[1]", "beta[2]", "s2"])
    mc._seed = 77
    g = str(tmp_path / "mc.npz")
    api.write(g, mc)
    r = api.read(g, api.ModelChains)
    assert r.model.template == "line" and r.model.iter == 89 and r.model.burnin == 9 and r._seed == 77
    assert [(s.params, s.kind) for s in r.model.samplers] == [(["beta"], "amwg"), (["s2"], "slice_multi")]
    assert r.model.samplers[1].desc["transform"] == 1 and r.model.samplers[0].desc["scale"] == 1.0
    np.testing.assert_array_equal(r.model.inputs["y"], [1, 3, 3, 3, 5])
    np.testing.assert_array_equal(r.model.states[1].value, [0.3, 0.1, 0.5]); np.testing.assert_array_equal(r.model.states[1].tune, np.arange(7.0) + 1)
    np.testing.assert_array_equal(r.link_codes(), [0, 0, 1])
    sub = r[:, "beta", :]                                     # names2inds(mc, nodekey): modelchains.jl:24-40
    assert sub.names == ["beta[1]", "beta[2]"] and isinstance(sub, api.ModelChains)
    with pytest.raises(api.ArgumentError, match="chain values are missing for nodes : gamma"):
        r[:, "gamma", :]


def test_template_node_tables_agree_with_the_oracle(mcu_built, oracle):
    # the API mirror lays out initial values by its own node table (api._TEMPLATES); the state record order of the engine is the oracle's
    # (GPU parity tests): a mismatch would silently permute initial values
    from mambacuda import api, _lib
    import pyoracle
    assert set(api._TEMPLATES) == set(_lib.TPL) == set(pyoracle.TPL)
    assert _lib.TPL == pyoracle.TPL
    for tpl, t in api._TEMPLATES.items():
        if tpl == "glm":
            continue
        want = []
        for nm, ln in t["nodes"]:
            want += [nm] if ln == 1 else [f"{nm}[{i + 1}]" for i in range(ln)]
        o = oracle.Oracle(tpl)
        got = o.names(monitoronly=False)
        # the oracle lists every node (Logical and observed ones too): keep the unobserved stochastic elements = the state record
        state = [g for g in got if g.split("[")[0] in {nm for nm, _ in t["nodes"]}]
        assert state == want, (tpl, state[:6], want[:6])
