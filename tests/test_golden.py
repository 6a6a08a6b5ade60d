"""Golden-vector tests.  tests/golden/*.json are formula-level goldens produced by tests/golden/make_golden.py from an
independent scipy.stats restatement of logpdf!(m, x, block, transform) per block and of gelmandiag / summarystats, on the
data sets parsed from the reference's own example scripts (the reference cannot run here: SURVEY.md §8c).
CPU: the oracle must reproduce them; GPU (-m gpu): the CUDA path must reproduce them through the C ABI."""
import json
import os

import numpy as np
import pytest

import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# golden key -> (scheme blocks that expose the block, index of the block inside that scheme)
BLOCKS = {
    "line": {
        "beta": ([dict(kind="amwg", nodes=[0], scale=1.0), dict(kind="slice_multi", nodes=[1], scale=1.0, transform=1)], 0),
        "s2_transformed": ([dict(kind="amwg", nodes=[0], scale=1.0), dict(kind="slice_multi", nodes=[1], scale=1.0, transform=1)], 1),
        "s2_constrained": ([dict(kind="slice_multi", nodes=[1], scale=1.0, transform=0)], 0),
        "beta_s2_transformed": ([dict(kind="nuts", nodes=[0, 1])], 0),
    },
    "seeds": {
        "alpha": ([dict(kind="amwg", nodes=[0, 1, 2, 3], scale=0.1), dict(kind="amwg", nodes=[5], scale=0.01), dict(kind="amwg", nodes=[4], scale=0.1)], 0),
        "b": ([dict(kind="amwg", nodes=[0, 1, 2, 3], scale=0.1), dict(kind="amwg", nodes=[5], scale=0.01), dict(kind="amwg", nodes=[4], scale=0.1)], 1),
        "s2_transformed": ([dict(kind="amwg", nodes=[0, 1, 2, 3], scale=0.1), dict(kind="amwg", nodes=[5], scale=0.01), dict(kind="amwg", nodes=[4], scale=0.1)], 2),
    },
    "rats": {
        "s2_c": ([dict(kind="slice_multi", nodes=[4], scale=10.0)], 0),
        "alpha": ([dict(kind="amwg", nodes=[5], scale=100.0)], 0),
        "mu_alpha_s2_alpha": ([dict(kind="slice_uni", nodes=[0, 2], scale=[100.0, 10.0])], 0),
        "beta": ([dict(kind="amwg", nodes=[6], scale=1.0)], 0),
        "mu_beta_s2_beta": ([dict(kind="slice_uni", nodes=[1, 3], scale=1.0)], 0),
        "nuts_alpha_beta_mu": ([dict(kind="nuts", nodes=[5, 6, 0, 1]), dict(kind="slice_uni", nodes=[4, 2, 3], scale=[10.0, 10.0, 1.0])], 0),
        "slice_s2c_s2a_s2b": ([dict(kind="nuts", nodes=[5, 6, 0, 1]), dict(kind="slice_uni", nodes=[4, 2, 3], scale=[10.0, 10.0, 1.0])], 1),
    },
    "surgical": {
        "b": ([dict(kind="nuts", nodes=[2]), dict(kind="slice_multi", nodes=[0, 1], scale=1.0)], 0),
        "mu_s2_constrained": ([dict(kind="nuts", nodes=[2]), dict(kind="slice_multi", nodes=[0, 1], scale=1.0)], 1),
        "mu_s2_transformed": ([dict(kind="amwg", nodes=[0, 1], scale=0.3)], 0),
    },
    "dyes": {
        "nuts_mu_theta": ([dict(kind="nuts", nodes=[3, 1]), dict(kind="slice_multi", nodes=[2, 0], scale=1000.0)], 0),
        "slice_s2w_s2b": ([dict(kind="nuts", nodes=[3, 1]), dict(kind="slice_multi", nodes=[2, 0], scale=1000.0)], 1),
        "theta": ([dict(kind="mala", nodes=[1], epsilon=50.0)], 0),
        "mu": ([dict(kind="hmc", nodes=[3], epsilon=10.0, L=5)], 0),
    },
    "salm": {
        "slice_alpha_beta_gamma": ([dict(kind="slice_multi", nodes=[3, 2, 1], scale=[1.0, 1.0, 0.1]), dict(kind="amwg", nodes=[4, 0], scale=0.1)], 0),
        "amwg_lambda_s2_transformed": ([dict(kind="slice_multi", nodes=[3, 2, 1], scale=[1.0, 1.0, 0.1]), dict(kind="amwg", nodes=[4, 0], scale=0.1)], 1),
    },
    "equiv": {
        "nuts_delta": ([dict(kind="nuts", nodes=[5]), dict(kind="slice_multi", nodes=[4, 3, 2], scale=1.0), dict(kind="slice_uni", nodes=[1, 0], scale=1.0)], 0),
        "slice_mu_phi_pi": ([dict(kind="nuts", nodes=[5]), dict(kind="slice_multi", nodes=[4, 3, 2], scale=1.0), dict(kind="slice_uni", nodes=[1, 0], scale=1.0)], 1),
        "slice_s2_1_s2_2": ([dict(kind="nuts", nodes=[5]), dict(kind="slice_multi", nodes=[4, 3, 2], scale=1.0), dict(kind="slice_uni", nodes=[1, 0], scale=1.0)], 2),
    },
    "blocker": {
        "amwg_mu": ([dict(kind="amwg", nodes=[3], scale=0.1), dict(kind="amwg", nodes=[4, 2], scale=0.1), dict(kind="slice_multi", nodes=[1, 0], scale=1.0)], 0),
        "amwg_delta_delta_new": ([dict(kind="amwg", nodes=[3], scale=0.1), dict(kind="amwg", nodes=[4, 2], scale=0.1), dict(kind="slice_multi", nodes=[1, 0], scale=1.0)], 1),
        "slice_d_s2": ([dict(kind="amwg", nodes=[3], scale=0.1), dict(kind="amwg", nodes=[4, 2], scale=0.1), dict(kind="slice_multi", nodes=[1, 0], scale=1.0)], 2),
    },
    "stacks": {
        "nuts_beta0_beta": ([dict(kind="nuts", nodes=[0, 1]), dict(kind="slice_multi", nodes=[2], scale=1.0)], 0),
        "slice_s2": ([dict(kind="nuts", nodes=[0, 1]), dict(kind="slice_multi", nodes=[2], scale=1.0)], 1),
    },
    "pumps": {
        "alpha_beta_constrained": ([dict(kind="slice_uni", nodes=[0, 1], scale=1.0), dict(kind="slice_uni", nodes=[2], scale=1.0)], 0),
        "theta_constrained": ([dict(kind="slice_uni", nodes=[0, 1], scale=1.0), dict(kind="slice_uni", nodes=[2], scale=1.0)], 1),
        "alpha_beta_transformed": ([dict(kind="amwg", nodes=[0, 1], scale=0.5), dict(kind="nuts", nodes=[2])], 0),
        "theta_transformed": ([dict(kind="amwg", nodes=[0, 1], scale=0.5), dict(kind="nuts", nodes=[2])], 1),
    },
}


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "block_logpdf.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gold_extra():
    with open(os.path.join(GOLD, "block_logpdf_extra.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gold_more():
    with open(os.path.join(GOLD, "block_logpdf_more.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gold_diag():
    with open(os.path.join(GOLD, "diagnostics.json")) as f:
        return json.load(f)


def _oracle_blocks(blocks):
    import helpers
    return [helpers.oracle_block(b) for b in blocks]


def test_golden_files_are_committed_with_their_generator():
    for f in ("block_logpdf.json", "diagnostics.json", "make_golden.py", "block_logpdf_more.json", "make_golden_more.py", "coda.json", "make_coda_golden.py"):
        assert os.path.exists(os.path.join(GOLD, f))


@pytest.mark.parametrize("tpl", ["line", "seeds", "rats", "pumps"])
def test_oracle_block_densities_match_golden(oracle, gold, tpl):
    S = np.array(gold["blocks"][tpl]["states"])
    for key, (blocks, bi) in BLOCKS[tpl].items():
        o = oracle.Oracle(tpl)
        o.set_scheme(_oracle_blocks(blocks))
        np.testing.assert_allclose(o.logpdf(bi, S), gold["blocks"][tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=f"{tpl}/{key}")


@pytest.mark.parametrize("tpl", ["surgical", "dyes"])
def test_oracle_extra_template_block_densities_match_golden(oracle, gold_extra, tpl):
    S = np.array(gold_extra["blocks"][tpl]["states"])
    for key, (blocks, bi) in BLOCKS[tpl].items():
        o = oracle.Oracle(tpl)
        o.set_scheme(_oracle_blocks(blocks))
        np.testing.assert_allclose(o.logpdf(bi, S), gold_extra["blocks"][tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)


@pytest.mark.parametrize("tpl", ["salm", "equiv", "blocker", "stacks"])
def test_oracle_salm_equiv_block_densities_match_golden(oracle, gold_more, tpl):
    S = np.array(gold_more["blocks"][tpl]["states"])
    for key, (blocks, bi) in BLOCKS[tpl].items():
        o = oracle.Oracle(tpl)
        o.set_scheme(_oracle_blocks(blocks))
        np.testing.assert_allclose(o.logpdf(bi, S), gold_more["blocks"][tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)
    o = oracle.Oracle(tpl); o.set_scheme(_oracle_blocks(next(iter(BLOCKS[tpl].values()))[0]))
    nn = {"salm": 5, "equiv": 6, "blocker": 5, "stacks": 3}[tpl]
    for q, key in enumerate(["rc", "rt"] if tpl == "blocker" else ["y"]):                                          # observed nodes, one at a time
        np.testing.assert_allclose(o.logpdf_nodes(1 << (nn + q), S), gold_more["blocks"][tpl]["logpdf"][key], rtol=1e-11)
    # analytic gradient of the joint against central differences of the block density that holds every parameter node
    allb = [dict(kind="nuts", nodes=list(range(nn)))]
    o = oracle.Oracle(tpl); o.set_scheme(_oracle_blocks(allb))
    lp_a, g_a = o.gradlogpdf(0, S, mode=0)
    lp_c, g_c = o.gradlogpdf(0, S, mode=2)
    np.testing.assert_allclose(g_a, g_c, rtol=2e-5, atol=1e-4 * np.abs(g_c).max())   # (stacks: |y - mu| is piecewise linear, no kink within 6e-6 of these states)
    if tpl == "stacks":   # the monitored columns are all Logical nodes: b, b0, sigma, outlier[1, 3, 4, 21]
        still = [dict(kind="rwm", nodes=[0, 1, 2], scale=0.0)]          # a random walk of step 0: one "iteration" records unlist(m, true) at S
        o = oracle.Oracle(tpl); o.set_scheme(_oracle_blocks(still))
        out, st, _ = o.run(len(S), S, 1, burnin=0, thin=1, seed=1)
        assert o.names() == ["b[1]", "b[2]", "b[3]", "b0", "sigma", "outlier[1]", "outlier[3]", "outlier[4]", "outlier[21]"]
        np.testing.assert_array_equal(st, S)
        np.testing.assert_allclose(out[0].T, gold_more["blocks"]["stacks"]["monitor"], rtol=1e-12)


def test_oracle_glm_density_and_gradient_match_golden(oracle, gold):
    g = gold["blocks"]["glm"]
    X, y, B = np.array(g["X"]), np.array(g["y"]), np.array(g["states"])
    o = oracle.Oracle("glm", glm_d=X.shape[1])
    o.set_data("X", X); o.set_data("y", y)
    o.set_scheme([dict(kind=4, nodes=[0])])
    lp, gr = o.gradlogpdf(0, B, mode=0)
    np.testing.assert_allclose(lp, g["logpdf"]["beta"], rtol=1e-12)
    np.testing.assert_allclose(gr, g["grad"]["beta"], rtol=1e-10, atol=1e-10)


@pytest.fixture(scope="module")
def gold_fam():
    with open(os.path.join(GOLD, "glm_family.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["poisson", "normal"])
def test_oracle_glm_family_matches_golden(oracle, gold_fam, name):
    g = gold_fam[name]
    X, y, B = np.array(g["X"]), np.array(g["y"]), np.array(g["states"])
    o = oracle.Oracle("glm", glm_d=X.shape[1])
    o.set_data("X", X); o.set_data("y", y); o.set_data("family", np.array([float(g["family"])])); o.set_data("sigma", np.array([g["sigma"]]))
    o.set_scheme([dict(kind=4, nodes=[0])])
    lp, gr = o.gradlogpdf(0, B, mode=0)
    np.testing.assert_allclose(lp, g["logpdf"], rtol=1e-12)
    np.testing.assert_allclose(gr, g["grad"], rtol=1e-10, atol=1e-10)


def test_oracle_diagnostics_match_golden(oracle, gold_diag):
    c = np.array(gold_diag["chains"])
    np.testing.assert_allclose(oracle.gelmandiag(c), gold_diag["gelmandiag_alpha_0.05"], rtol=1e-9)
    np.testing.assert_allclose(oracle.gelmandiag(c, linkcode=[0, 0, 1]), gold_diag["gelmandiag_log_last_column"], rtol=1e-9)
    np.testing.assert_allclose(oracle.summarystats(c), gold_diag["summarystats_bm100"], rtol=1e-10)


def test_abi_host_finalisation_matches_golden(mcu_built, gold_diag):
    # mcu_gelman_from_moments / mcu_summary_from_sums are the host half of gelmandiag / summarystats behind the C ABI (no device
    # needed): feed them the cross-chain sums of the golden chains, as the all-reduce would deliver them
    import ctypes as C
    from mambacuda import _lib
    L = _lib.lib()
    c = np.array(gold_diag["chains"])
    n, p, m = c.shape
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    mean_c = c.mean(axis=0); var_c = c.var(axis=0, ddof=1)                  # p x m
    center = np.ascontiguousarray(np.stack([mean_c.mean(axis=1), var_c.mean(axis=1)], axis=1))
    d = mean_c - center[:, :1]; e = var_c - center[:, 1:]
    sums = np.ascontiguousarray(np.stack([np.full(p, float(m)), d.sum(1), (d * d).sum(1), e.sum(1), (e * e).sum(1), (e * d).sum(1), (e * d * d).sum(1)], axis=1))
    psrf = np.empty((p, 2))
    assert L.mcu_gelman_from_moments(C.c_int64(n), p, dp(center), dp(sums), C.c_double(0.05), dp(psrf)) == 0
    np.testing.assert_allclose(psrf, gold_diag["gelmandiag_alpha_0.05"], rtol=1e-9)
    # streaming summary sums: {C, sum mean_c, sum M2_c, sum (mean_c - c1)^2, sum nb_c, sum nb_c bmean_c, sum bM2_c, sum nb_c (bmean_c - c2)^2}
    nb = n // 100
    bm = c.reshape(nb, 100, p, m).mean(axis=1)                               # nb x p x m batch means (100 | 400: batches stay inside a chain)
    bmean_c = bm.mean(axis=0); bM2_c = ((bm - bmean_c) ** 2).sum(axis=0)
    M2_c = ((c - mean_c) ** 2).sum(axis=0)
    ctr = np.ascontiguousarray(np.stack([mean_c.mean(axis=1), bmean_c.mean(axis=1)], axis=1))
    s8 = np.ascontiguousarray(np.stack([np.full(p, float(m)), mean_c.sum(1), M2_c.sum(1), ((mean_c - ctr[:, :1]) ** 2).sum(1), np.full(p, float(nb * m)),
                                        (nb * bmean_c).sum(1), bM2_c.sum(1), (nb * (bmean_c - ctr[:, 1:]) ** 2).sum(1)], axis=1))
    out = np.empty((p, 5))
    L.mcu_summary_from_sums(C.c_int64(n), p, dp(ctr), dp(s8), dp(out))
    np.testing.assert_allclose(out, gold_diag["summarystats_bm100"], rtol=1e-9)


def test_embedded_data_equal_the_reference_scripts(oracle, gold):
    # the golden file carries the data parsed from doc/examples/*.jl; the engine / oracle defaults must be the same numbers:
    # a density at a fixed state that matches to 1e-11 (above) already implies it; here the seeds known answer of SURVEY App. D
    d = gold["data"]["seeds"]
    assert len(d["r"]) == 21 and sum(d["r"]) == 424 and sum(d["n"]) == 831
    assert len(gold["data"]["rats"]["y"]) == 150 and gold["data"]["rats"]["xbar"] == 22.0
    assert gold["data"]["pumps"]["t"][4] == 5.24


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tpl", ["line", "seeds", "rats", "pumps"])
def test_gpu_block_densities_match_golden(gold, tpl):
    from mambacuda.engine import Engine
    S = np.array(gold["blocks"][tpl]["states"])
    for key, (blocks, bi) in BLOCKS[tpl].items():
        eng = Engine(tpl, 4)
        eng.set_scheme(blocks)
        np.testing.assert_allclose(eng.logpdf(bi, S), gold["blocks"][tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=f"{tpl}/{key}")
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tpl", ["surgical", "dyes"])
def test_gpu_extra_template_block_densities_match_golden(gold_extra, tpl):
    from mambacuda.engine import Engine
    S = np.array(gold_extra["blocks"][tpl]["states"])
    for key, (blocks, bi) in BLOCKS[tpl].items():
        eng = Engine(tpl, 4)
        eng.set_scheme(blocks)
        np.testing.assert_allclose(eng.logpdf(bi, S), gold_extra["blocks"][tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tpl", ["salm", "equiv", "blocker", "stacks"])
def test_gpu_salm_equiv_block_densities_match_golden(gold_more, tpl):
    from mambacuda.engine import Engine
    S = np.array(gold_more["blocks"][tpl]["states"])
    for key, (blocks, bi) in BLOCKS[tpl].items():
        eng = Engine(tpl, 4)
        eng.set_scheme(blocks)
        np.testing.assert_allclose(eng.logpdf(bi, S), gold_more["blocks"][tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)
        nn, nf = eng.factor_counts()
        for q, okey in enumerate(["rc", "rt"] if tpl == "blocker" else ["y"]):
            np.testing.assert_allclose(eng.logpdf_nodes(1 << (nn + q), S), gold_more["blocks"][tpl]["logpdf"][okey], rtol=1e-11)
        eng.close()
    if tpl == "stacks":   # Logical monitored columns at fixed states (a random walk of step 0 records unlist(m, true))
        eng = Engine(tpl, len(S))
        eng.set_scheme([dict(kind="rwm", nodes=[0, 1, 2], scale=0.0)]); eng.set_inits(S)
        out = eng.run(1, burnin=0, thin=1)
        np.testing.assert_allclose(out[0].T, gold_more["blocks"]["stacks"]["monitor"], rtol=1e-12)
        eng.close()


@pytest.mark.gpu
def test_gpu_glm_density_and_gradient_match_golden(gold):
    from mambacuda.engine import Engine
    g = gold["blocks"]["glm"]
    X, y, B = np.array(g["X"]), np.array(g["y"]), np.array(g["states"])
    eng = Engine("glm", 4)
    eng.set_data("X", X); eng.set_data("y", y)
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    lp, gr = eng.gradlogpdf(0, B, X.shape[1])
    np.testing.assert_allclose(lp, g["logpdf"]["beta"], rtol=1e-12)
    np.testing.assert_allclose(gr, g["grad"]["beta"], rtol=1e-10, atol=1e-10)
    # the tensor-core pass on the same states (12 requested positions padded to the handle's chain count): 1e-5 (north_star)
    eng2 = Engine("glm", 12)
    eng2.set_data("X", X); eng2.set_data("y", y)
    eng2.set_scheme([dict(kind="nuts", nodes=[0])])
    lp_tc, g_tc = eng2.glm_gradient(B, impl=1)
    prior = -0.5 * (B ** 2).sum(axis=1) / 1000.0 - 0.5 * X.shape[1] * np.log(2 * np.pi * 1000.0)
    np.testing.assert_allclose(lp_tc + prior, g["logpdf"]["beta"], rtol=1e-5)
    gl = np.array(g["grad"]["beta"]) + B / 1000.0
    assert np.max(np.abs(g_tc - gl) / np.abs(gl).max(axis=1, keepdims=True)) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["poisson", "normal"])
def test_gpu_glm_family_matches_golden(gold_fam, name):
    # CUDA-core FP64 path to 1e-12, tensor-core path (Poisson / Normal epilogues of glm_tc_kernel) to north_star's 1e-5
    from mambacuda.engine import Engine
    g = gold_fam[name]
    X, y, B = np.array(g["X"]), np.array(g["y"]), np.array(g["states"])
    d = X.shape[1]
    eng = Engine("glm", 12)
    eng.set_data("X", X); eng.set_data("y", y); eng.set_data("family", np.array([float(g["family"])])); eng.set_data("sigma", np.array([g["sigma"]]))
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    lp, gr = eng.gradlogpdf(0, B, d)
    np.testing.assert_allclose(lp, g["logpdf"], rtol=1e-12)
    np.testing.assert_allclose(gr, g["grad"], rtol=1e-10, atol=1e-10)
    prior = -0.5 * (B ** 2).sum(axis=1) / 1000.0 - 0.5 * d * np.log(2 * np.pi * 1000.0)
    gl = np.array(g["grad"]) + B / 1000.0
    for impl in (0, 1):
        lp_k, g_k = eng.glm_gradient(B, impl=impl)
        np.testing.assert_allclose(lp_k + prior, g["logpdf"], rtol=1e-12 if impl == 0 else 1e-5)
        assert np.max(np.abs(g_k - gl) / np.abs(gl).max(axis=1, keepdims=True)) < (1e-10 if impl == 0 else 1e-5)


@pytest.mark.gpu
def test_gpu_poisson_glm_nuts_recovers_coefficients():
    # NUTS through the tick engine with the tensor-core Poisson epilogue
    from mambacuda.engine import Engine
    rng = np.random.default_rng(12)
    N, d = 5000, 5
    X = rng.normal(scale=0.5, size=(N, d)); X[:, 0] = 1.0
    beta = np.array([0.3, -0.5, 0.8, 0.2, -0.1])
    y = rng.poisson(np.exp(X @ beta)).astype(float)
    eng = Engine("glm", 128, seed=3)
    eng.set_data("X", X); eng.set_data("y", y); eng.set_data("family", np.array([1.0]))
    eng.set_scheme([dict(kind="nuts", nodes=[0])])
    eng.set_inits(np.zeros((1, d)), jitter_sd=0.1)
    eng.run(300, burnin=150, thin=1, store=False, out=False)
    summ = eng.summary_streaming()
    assert np.all(np.abs(summ[:, 0] - beta) < 5 * summ[:, 1] + 0.01)
    assert (eng.gelman(0.05, False)[:, 0] < 1.1).all()


@pytest.mark.gpu
def test_gpu_diagnostics_match_the_golden_formulas_on_sampled_chains():
    # device reductions + host finalisation (mcu_gelman, mcu_summarystats, streaming form) against the independent numpy
    # restatement of gelmandiag.jl / stats.jl / mcse.jl in tests/golden/make_golden.py, on chains the engine sampled itself
    import importlib.util
    import helpers
    from mambacuda.engine import Engine
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    tpl, blocks, inits = helpers.scheme("seeds_amwg")
    eng = Engine(tpl, 6, seed=11)
    eng.set_scheme(blocks); eng.set_inits(inits)
    out = eng.run(1400, burnin=600, thin=2)            # 400 kept draws x 5 monitored x 6 chains
    np.testing.assert_allclose(eng.gelman(0.05, False), mg.gelmandiag(out), rtol=1e-8)
    ref = mg.summarystats(out)
    np.testing.assert_allclose(eng.summarystats("bm", 100), ref, rtol=1e-9)
    np.testing.assert_allclose(eng.summary_streaming(), ref, rtol=1e-8)   # 100 | 400: streaming batches = the reference's


# ------------------------------------------------------------------------- chain post-processing behind the C ABI (CPU)
def _np_hpd(x, alpha):
    x = np.sort(x); n = x.size; m = max(1, int(np.ceil(alpha * n)))
    a, b = x[:m], x[n - m:]
    i = int(np.argmin(b - a))
    return a[i], b[i]


def _np_autocor(x, lag):
    z = x - x.mean()
    return float((z[:x.size - lag] * z[lag:]).sum() / (z * z).sum())


def _np_mpsrf(c):
    n, p, m = c.shape
    W = np.mean([np.cov(c[:, :, k], rowvar=False) for k in range(m)], axis=0)
    B = n * np.cov(c.mean(axis=0).T, rowvar=False)
    lam = np.max(np.linalg.eigvals(np.linalg.solve(W, B)).real)
    return (n - 1) / n + (m + 1) / (m * n) * lam


def test_abi_chain_postprocessing_matches_numpy(mcu_built, gold_diag):
    # quantile / hpd / autocor / changerate / gelmandiag(mpsrf = true) of src/output/stats.jl and gelmandiag.jl:49-55 through the
    # host-array entry points (mcu_chains_*), against numpy restatements of the same formulas
    from mambacuda import api
    c = np.array(gold_diag["chains"])
    n, p, m = c.shape
    ch = api.Chains(c, start=3, thin=2, names=["a", "b", "c"])
    q, names, labels = api.quantile(ch)
    ref = np.array([np.quantile(c[:, j, :].reshape(-1), [0.025, 0.25, 0.5, 0.75, 0.975]) for j in range(p)])   # numpy default = Julia's type 7
    np.testing.assert_allclose(q, ref, rtol=1e-13)
    assert labels[0] == "2.5%" and names == ["a", "b", "c"]
    h, _, _ = api.hpd(ch, alpha=0.05)
    np.testing.assert_allclose(h, [_np_hpd(c[:, j, :].T.reshape(-1), 0.05) for j in range(p)], rtol=1e-13)
    ac, _, lab = api.autocor(ch, lags=[1, 5], relative=True)       # index lags 2, 10 on the stored series (stats.jl:5-6)
    assert lab == ["Lag 2", "Lag 10"] and ac.shape == (p, 2, m)
    for j in range(p):
        for k in range(m):
            np.testing.assert_allclose(ac[j, :, k], [_np_autocor(c[:, j, k], 2), _np_autocor(c[:, j, k], 10)], rtol=1e-11)
    with pytest.raises(api.ArgumentError):
        api.autocor(ch, lags=[1, 3], relative=False)                # "lags do not correspond to thinning interval"
    # changerate: a chain array with repeated values
    rng = np.random.default_rng(0)
    z = np.cumsum(rng.integers(0, 2, size=(50, 2, 3)), axis=0).astype(float)
    cr, nm, _ = api.changerate(api.Chains(z))
    d = np.diff(z, axis=0) != 0
    np.testing.assert_allclose(cr, np.round(np.append(d.mean(axis=(0, 2)), d.any(axis=1).mean()), 3))
    assert nm[-1] == "Multivariate"
    # gelmandiag on the materialised array incl. the multivariate PSRF
    g = api._chains_gelman(c, 0.05, None, True)
    np.testing.assert_allclose(g[:p], gold_diag["gelmandiag_alpha_0.05"], rtol=1e-9)
    assert g[p, 0] == pytest.approx(_np_mpsrf(c), rel=1e-9) and np.isnan(g[p, 1])
    g2 = api._chains_gelman(c, 0.05, [0, 0, 1], True)
    np.testing.assert_allclose(g2[:p], gold_diag["gelmandiag_log_last_column"], rtol=1e-9)
    cl = c.copy(); cl[:, 2, :] = np.log(cl[:, 2, :])
    assert g2[p, 0] == pytest.approx(_np_mpsrf(cl), rel=1e-9)
    with pytest.raises(api.ArgumentError):
        api._chains_gelman(c[:, :, :1], 0.05, None, False)          # less than 2 chains


# ------------------------------------------------------------------------- geweke / heidel / raftery behind the C ABI (CPU)
def _np_mcse_imse(x):
    n = x.size; z = x - x.mean()
    ac = lambda k: float((z[:n - k] * z[k:]).sum() / n)
    Ghat = ac(0) + ac(1); value = -ac(0) + 2 * Ghat
    for i in range(1, (n - 2) // 2 + 1):
        Ghat = min(Ghat, ac(2 * i) + ac(2 * i + 1))
        if not Ghat > 0:
            break
        value += 2 * Ghat
    return np.sqrt(value / n)


def _np_geweke(x, first=0.1, last=0.5):
    from scipy.special import erf
    n = x.size
    rnd = lambda v: int(np.floor(v + 0.5))
    x1 = x[:rnd(first * n)]; x2 = x[rnd(n - last * n + 1) - 1:]
    z = (x1.mean() - x2.mean()) / np.sqrt(_np_mcse_imse(x1) ** 2 + _np_mcse_imse(x2) ** 2)
    return z, 1 - erf(abs(z) / np.sqrt(2))


def _np_pcramer(q):
    from scipy.special import kv, gamma
    from math import factorial
    p = 0.0
    for k in range(4):
        c1 = 4.0 * k + 1.0; c2 = c1 ** 2 / (16.0 * q)
        p += gamma(k + 0.5) / factorial(k) * np.sqrt(c1) * np.exp(-c2) * kv(0.25, c2)
    return p / (np.pi ** 1.5 * np.sqrt(q))


def _np_heidel(x, alpha=0.05, eps=0.1, start=1):
    from scipy.special import erfinv
    n = x.size; delta = int(0.10 * n)
    y = x[int(n / 2) - 1:]
    S0 = y.size * _np_mcse_imse(y) ** 2
    i, pvalue, converged, ybar = 1, 1.0, False, np.nan
    while i < n / 2:
        y = x[i - 1:]; m = y.size; ybar = y.mean()
        B = np.cumsum(y) - ybar * np.arange(1, m + 1)
        I = ((B * B) / (m * S0)).sum() / m
        pvalue = 1 - _np_pcramer(I); converged = pvalue > alpha
        if converged:
            break
        i += delta
    hw = np.sqrt(2) * erfinv(1 - alpha) * _np_mcse_imse(y)
    return [i + start - 2, float(converged), pvalue, ybar, hw, float(hw / abs(ybar) <= eps)]


def _np_raftery(x, q=0.025, r=0.005, s=0.95, eps=0.001, start=1, step=1):
    from scipy.special import erfinv
    nx = x.size; phi = np.sqrt(2) * erfinv(s)
    nmin = int(np.ceil(q * (1 - q) * (phi / r) ** 2))
    if nmin > nx:
        return [np.nan, np.nan, np.nan, nmin, np.nan]
    dichot = (x <= np.quantile(x, q)).astype(int)
    kthin, bic = 0, 1.0
    while bic >= 0:
        kthin += 1
        test = dichot[::kthin]; nt = test.size
        temp = test[:nt - 2] + 2 * test[1:nt - 1] + 4 * test[2:]
        tr = np.bincount(temp, minlength=8).reshape(2, 2, 2, order="F").astype(float)
        g2 = 0.0
        for i1 in range(2):
            for i2 in range(2):
                for i3 in range(2):
                    tt = tr[i1, i2, i3]
                    if tt > 0:
                        fitted = tr[:, i2, i3].sum() * tr[i1, i2, :].sum() / tr[:, i2, :].sum()
                        g2 += 2 * tt * np.log(tt / fitted)
        bic = g2 - 2 * np.log(nt - 2.0)
    tf = np.bincount(test[:nt - 1] + 2 * test[1:], minlength=4).astype(float)
    al = tf[2] / (tf[0] + tf[2]); be = tf[1] / (tf[1] + tf[3])
    kt = kthin * step
    m = np.log(eps * (al + be) / max(al, be)) / np.log(abs(1 - al - be))
    burnin = kt * np.ceil(m) + start - 1
    nn = ((2 - al - be) * al * be * phi ** 2) / (r ** 2 * (al + be) ** 3)
    total = burnin + kt * np.ceil(nn)
    return [kt, burnin, total, nmin, total / nmin]


def test_abi_convergence_diagnostics_match_numpy(mcu_built, gold_diag):
    # gewekediag / heideldiag / rafterydiag (src/output/{gewekediag,heideldiag,rafterydiag}.jl) through mcu_chains_geweke / _heidel /
    # _raftery against numpy / scipy.special restatements (besselk, erfinv from scipy)
    from mambacuda import api
    c = np.array(gold_diag["chains"])
    n, p, m = c.shape
    ch = api.Chains(c, start=251, thin=2, names=["a", "b", "c"])
    g, _, lab = api.gewekediag(ch)
    assert lab == ["Z-score", "p-value"] and g.shape == (p, 2, m)
    h, _, _ = api.heideldiag(ch)
    for j in range(p):
        for k in range(m):
            z, pv = _np_geweke(c[:, j, k])
            assert g[j, 0, k] == pytest.approx(round(z, 3), abs=1.1e-3) and g[j, 1, k] == pytest.approx(pv, abs=2e-4)
            ref = _np_heidel(c[:, j, k], start=251)
            np.testing.assert_allclose(h[j, [0, 1, 3, 4, 5], k], np.array(ref)[[0, 1, 3, 4, 5]], rtol=1e-9)
            assert h[j, 2, k] == pytest.approx(ref[2], abs=1e-4)
    with pytest.raises(api.ArgumentError, match="overlap"):
        api.gewekediag(ch, first=0.6, last=0.5)
    # raftery needs long chains (nmin = 3746 at the defaults): AR(1) series with thinning-dependent dichotomised dependence
    rng = np.random.default_rng(5)
    L = 6000
    z = np.zeros((L, 2, 2))
    e = rng.normal(size=(L, 2, 2))
    for t in range(1, L):
        z[t] = np.array([0.2, 0.85])[None, :, None] * z[t - 1] + e[t]
    zc = api.Chains(z, start=1001, thin=3)
    r, _, lab = api.rafterydiag(zc)
    assert lab[0] == "Thinning" and r.shape == (2, 5, 2)
    for j in range(2):
        for k in range(2):
            np.testing.assert_allclose(r[j, :, k], _np_raftery(z[:, j, k], start=1001, step=3), rtol=1e-9)
    assert r[1, 0, 0] >= r[0, 0, 0]                       # the strongly autocorrelated column needs more thinning
    short, _, _ = api.rafterydiag(api.Chains(z[:500]))     # fewer than nmin samples: NaN rows, nmin reported
    assert np.isnan(short[0, 0, 0]) and short[0, 3, 0] == 3746


# ---- link layer (src/distributions/transformdistribution.jl:6-93) and the magnesium template that exercises it -----------------------
MAG_BLOCKS = {   # golden key -> (scheme, block index)
    "amwg_theta": ("magnesium", 0), "amwg_mu_transformed": ("magnesium", 1), "slice_pc": ("magnesium", 2), "slice_priors": ("magnesium", 3),
    "slice_pc_transformed": ("magnesium_transformed", 2), "priors_mu_transformed": ("magnesium_transformed", 3),
}


@pytest.fixture(scope="module")
def gold_links():
    with open(os.path.join(GOLD, "links.json")) as f:
        return json.load(f)


def test_oracle_link_functions_match_golden(oracle, gold_links):
    for a, b, x, lk, jac in gold_links["links"]:
        if b is not None:      # two-sided: Uniform(a, b); the unit interval also through Beta (UnitDistribution, :83-93)
            kinds = [("uniform", a, b, 0.0, 0.0)] + ([("beta", 2.0, 3.0, 0.0, 0.0)] if (a, b) == (0.0, 1.0) else [])
        else:                  # lower bound only: a Normal truncated to [a, Inf); a = 0 also through the PositiveDistribution union (:66-78)
            kinds = [("truncnormal", 0.3, 2.0, a, np.inf)] + ([("invgamma", 0.001, 0.001, 0.0, 0.0), ("gamma", 2.0, 3.0, 0.0, 0.0), ("exponential", 1.0, 0.0, 0.0, 0.0)] if a == 0.0 else [])
        for kind, p1, p2, lo, hi in kinds:
            got = oracle.link(kind, p1, p2, x, lo, hi)
            np.testing.assert_allclose(got[0], lk, rtol=1e-12, atol=1e-14, err_msg=f"link {kind} {a} {b} {x}")
            np.testing.assert_allclose(got[1], x, rtol=1e-9, atol=1e-12 * max(1.0, abs(a)), err_msg=f"invlink {kind} {a} {b} {x}")
            np.testing.assert_allclose(got[2], jac, rtol=1e-12, atol=1e-14, err_msg=f"jacobian {kind} {a} {b} {x}")
    # unbounded and discrete distributions: identity, no Jacobian (RealDistribution :53-61, fallbacks distributionstruct.jl:84,104,136)
    for kind, p1, p2 in (("normal", 0.0, 2.0), ("laplace", 1.0, 2.0), ("binomial", 10.0, 0.3), ("poisson", 3.0, 0.0)):
        np.testing.assert_array_equal(oracle.link(kind, p1, p2, 3.0), [3.0, 3.0, 0.0])
    # the densities the new links come with, against scipy
    import scipy.stats as st
    for x in (0.0, 0.2, 7.0, 50.0, 50.1, -0.1):
        np.testing.assert_allclose(oracle.udist_logpdf("uniform", 0.0, 50.0, x), st.uniform.logpdf(x, 0, 50), rtol=1e-14)
    for x in (0.01, 0.4, 0.99):
        np.testing.assert_allclose(oracle.udist_logpdf("beta", 2.5, 0.7, x), st.beta.logpdf(x, 2.5, 0.7), rtol=1e-12)
    for x in (-0.5, 0.0, 0.3, 4.0):
        np.testing.assert_allclose(oracle.udist_logpdf("truncnormal", 0.0, 1.3, x, 0.0, np.inf), st.truncnorm.logpdf(x, 0, np.inf, loc=0, scale=1.3), rtol=1e-13)
        np.testing.assert_allclose(oracle.udist_logpdf("truncnormal", 1.0, 2.0, x, -0.25, 3.0), st.truncnorm.logpdf(x, (-0.25 - 1) / 2, (3 - 1) / 2, loc=1, scale=2), rtol=1e-13)


def test_oracle_magnesium_block_densities_match_golden(oracle, gold_links):
    S = np.array(gold_links["states"])
    for key, (scheme, bi) in MAG_BLOCKS.items():
        tpl, blocks, _ = helpers.scheme(scheme)
        o = oracle.Oracle(tpl); o.set_scheme(_oracle_blocks(blocks))
        np.testing.assert_allclose(o.logpdf(bi, S), gold_links["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)
    tpl, blocks, _ = helpers.scheme("magnesium_transformed")
    o = oracle.Oracle(tpl); o.set_scheme(_oracle_blocks(blocks))
    for q, key in enumerate(["rcx", "rtx"]):
        np.testing.assert_allclose(o.logpdf_nodes(1 << (4 + q), S), gold_links["logpdf"][key], rtol=1e-11)
    # unlist(block, true): priors and mu on their link scales (log / two-sided logit), and the way back
    np.testing.assert_allclose(np.stack([o.unlist(3, s) for s in S]), gold_links["unlist_priors_mu"], rtol=1e-11, atol=1e-13)
    # monitored Logical columns tau[6], OR[6] at fixed states (a random walk of step 0 records unlist(m, true))
    o2 = oracle.Oracle(tpl); o2.set_scheme(_oracle_blocks([dict(kind="rwm", nodes=[2], scale=0.0)]))
    out, st_, _ = o2.run(len(S), S, 1, burnin=0, thin=1, seed=1)
    assert o2.names() == [f"tau[{i}]" for i in range(1, 7)] + [f"OR[{i}]" for i in range(1, 7)]
    np.testing.assert_allclose(out[0].T, gold_links["monitor"], rtol=1e-12)
    # analytic gradient on the link scale (chain rule through the two-sided link) against central differences of the transformed density
    for bi in (1, 3):
        lp_a, g_a = o.gradlogpdf(bi, S, mode=0)
        lp_c, g_c = o.gradlogpdf(bi, S, mode=2)
        np.testing.assert_allclose(g_a, g_c, rtol=2e-5, atol=1e-5 * np.abs(g_c).max())
    # out of the support: -Inf on the constrained scale (distributionstruct.jl:138-140)
    tplc, blocksc, _ = helpers.scheme("magnesium")
    oc = oracle.Oracle(tplc); oc.set_scheme(_oracle_blocks(blocksc))
    x = np.array(S[0][:6]); x[1] = 50.5
    assert np.isneginf(oc.logpdf(3, S[:1], x[None, :]))
    x = np.array(S[0][60:]); x[7] = -0.01
    assert np.isneginf(oc.logpdf(2, S[:1], x[None, :]))


@pytest.mark.gpu
def test_gpu_magnesium_block_densities_links_and_gradients_match_golden(oracle, gold_links):
    from mambacuda.engine import Engine
    S = np.array(gold_links["states"])
    for key, (scheme, bi) in MAG_BLOCKS.items():
        tpl, blocks, _ = helpers.scheme(scheme)
        eng = Engine(tpl, 4); eng.set_scheme(blocks)
        np.testing.assert_allclose(eng.logpdf(bi, S), gold_links["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)
        eng.close()
    tpl, blocks, _ = helpers.scheme("magnesium_transformed")
    eng = Engine(tpl, len(S)); eng.set_scheme(blocks)
    nn, nf = eng.factor_counts()
    assert (nn, nf) == (4, 6)
    for q, key in enumerate(["rcx", "rtx"]):
        np.testing.assert_allclose(eng.logpdf_nodes(1 << (nn + q), S), gold_links["logpdf"][key], rtol=1e-11)
    # explicit block vectors on the link scale: invlink on the device must land on the golden states' densities
    x = np.array(gold_links["unlist_priors_mu"])
    np.testing.assert_allclose(eng.logpdf(3, S, x), gold_links["logpdf"]["priors_mu_transformed"], rtol=1e-10, atol=1e-9)
    o = oracle.Oracle(tpl); o.set_scheme(_oracle_blocks(blocks))
    for bi, k in ((1, 12), (3, 12), (2, 48)):
        if blocks[bi]["kind"] == "slice_uni":
            continue
        lp_g, g_g = eng.gradlogpdf(bi, S, k)
        lp_o, g_o = o.gradlogpdf(bi, S, mode=0)
        np.testing.assert_allclose(lp_g, lp_o, rtol=1e-12)
        np.testing.assert_allclose(g_g, g_o, rtol=1e-10, atol=1e-10)
    eng2 = Engine(tpl, len(S)); eng2.set_scheme([dict(kind="rwm", nodes=[2], scale=0.0)]); eng2.set_inits(S)
    out = eng2.run(1, burnin=0, thin=1)
    np.testing.assert_allclose(out[0].T, gold_links["monitor"], rtol=1e-12)
    # jitter on the link scale stays inside every support (two-sided links map back into (a, b))
    eng2.set_inits(S, jitter_sd=2.0)
    st_, _, _ = eng2.get_state()
    assert (st_[:, 1:3] > 0).all() and (st_[:, 1:3] < 50).all() and (st_[:, 3:5] > 0).all() and (st_[:, 3:5] < 1).all()
    assert (np.abs(st_[:, 6:12]) < 10).all() and (st_[:, 60:] > 0).all() and (st_[:, 60:] < 1).all() and (st_[:, [0, 5]] > 0).all()


# ---- oxford and epil: 244 / 303 unobserved elements per chain (doc/examples/oxford.jl, doc/examples/epil.jl) -------------------------
LARGE_BLOCKS = {
    "oxford": {"amwg_alpha_beta1_beta2": ("oxford", 0), "slice_s2": ("oxford", 1), "slice_mu": ("oxford", 2), "slice_b": ("oxford", 3),
               "s2_transformed": ("oxford_componentwise", 1)},
    "epil": {"amwg_coefficients": ("epil", 0), "slice_b1": ("epil", 1), "slice_b": ("epil", 2), "slice_s2_b1_s2_b": ("epil", 3)},
}
LARGE_OBS = {"oxford": ["r0", "r1"], "epil": ["y"]}


@pytest.fixture(scope="module")
def gold_large():
    with open(os.path.join(GOLD, "block_logpdf_large.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("tpl", ["oxford", "epil"])
def test_oracle_oxford_epil_block_densities_match_golden(oracle, gold_large, tpl):
    S = np.array(gold_large[tpl]["states"])
    for key, (scheme, bi) in LARGE_BLOCKS[tpl].items():
        t, blocks, _ = helpers.scheme(scheme)
        o = oracle.Oracle(t); o.set_scheme(_oracle_blocks(blocks))
        np.testing.assert_allclose(o.logpdf(bi, S), gold_large[tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)
    t, blocks, _ = helpers.scheme(tpl)
    o = oracle.Oracle(t); o.set_scheme(_oracle_blocks(blocks))
    nn = {"oxford": 6, "epil": 10}[tpl]
    for q, key in enumerate(LARGE_OBS[tpl]):
        np.testing.assert_allclose(o.logpdf_nodes(1 << (nn + q), S), gold_large[tpl]["logpdf"][key], rtol=1e-11)
    # analytic gradient of the joint against central differences of the density of one block that holds every parameter node
    o = oracle.Oracle(t); o.set_scheme(_oracle_blocks([dict(kind="nuts", nodes=list(range(min(nn, 8))))] if nn <= 8 else
                                                      [dict(kind="nuts", nodes=list(range(8))), dict(kind="nuts", nodes=[8, 9])]))
    for bi in range(1 if nn <= 8 else 2):
        _, g_a = o.gradlogpdf(bi, S, mode=0)
        _, g_c = o.gradlogpdf(bi, S, mode=2)
        np.testing.assert_allclose(g_a, g_c, rtol=2e-5, atol=1e-5 * np.abs(g_c).max())
    if tpl == "epil":   # monitored Logical column alpha0 (epil.jl:85-91) at fixed states
        o2 = oracle.Oracle(t); o2.set_scheme(_oracle_blocks([dict(kind="rwm", nodes=[0], scale=0.0)]))
        out, _, _ = o2.run(len(S), S, 1, burnin=0, thin=1, seed=1)
        assert o2.names() == ["alpha_Base", "alpha_Trt", "alpha_BT", "alpha_Age", "alpha_V4", "alpha0", "s2_b1", "s2_b"]
        np.testing.assert_allclose(out[0, 5, :], gold_large["epil"]["logpdf"]["alpha0"], rtol=1e-12)
        np.testing.assert_allclose(out[0, :5, :].T, S[:, 1:6], rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("tpl", ["oxford", "epil"])
def test_gpu_oxford_epil_block_densities_match_golden(oracle, gold_large, tpl):
    from mambacuda.engine import Engine, MambaCudaError
    S = np.array(gold_large[tpl]["states"])
    for key, (scheme, bi) in LARGE_BLOCKS[tpl].items():
        t, blocks, _ = helpers.scheme(scheme)
        eng = Engine(t, 4); eng.set_scheme(blocks)
        np.testing.assert_allclose(eng.logpdf(bi, S), gold_large[tpl]["logpdf"][key], rtol=1e-11, atol=1e-9, err_msg=key)
        eng.close()
    t, blocks, _ = helpers.scheme(tpl)
    eng = Engine(t, len(S)); eng.set_scheme(blocks)
    nn, nf = eng.factor_counts()
    for q, key in enumerate(LARGE_OBS[tpl]):
        np.testing.assert_allclose(eng.logpdf_nodes(1 << (nn + q), S), gold_large[tpl]["logpdf"][key], rtol=1e-11)
    # the analytic gradient entry point works at this state size too (a batched density call, not a sampler)
    o = oracle.Oracle(t); o.set_scheme(_oracle_blocks(blocks))
    k = o.unlist(0, S[0]).size
    lp_g, g_g = eng.gradlogpdf(0, S, k)
    lp_o, g_o = o.gradlogpdf(0, S, mode=0)
    np.testing.assert_allclose(lp_g, lp_o, rtol=1e-12)
    np.testing.assert_allclose(g_g, g_o, rtol=1e-10, atol=1e-9)
    if tpl == "epil":
        e2 = Engine(t, len(S)); e2.set_scheme([dict(kind="rwm", nodes=[0], scale=0.0)]); e2.set_inits(S)
        out = e2.run(1, burnin=0, thin=1)
        np.testing.assert_allclose(out[0, 5, :], gold_large["epil"]["logpdf"]["alpha0"], rtol=1e-12)
    # gradient-based samplers are not compiled for templates of this size: a clear error, no silent fallback
    with pytest.raises(MambaCudaError, match="not compiled for this template"):
        eng.set_scheme([dict(kind="nuts", nodes=[0])])
