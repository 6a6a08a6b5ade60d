"""Generates tests/golden/*.json — golden vectors for the hot path, computed WITHOUT the oracle and WITHOUT the engine.

Run in the build container only (it reads the reference tree):  python tests/golden/make_golden.py

What it restates, in plain numpy / scipy.stats:
  * the data sets of the four example templates, parsed from the reference's own scripts
    (doc/tutorial/line.jl:69-73, doc/examples/seeds.jl:4-12, doc/examples/rats.jl:4-45, doc/examples/pumps.jl:4-9),
    so the fixtures also pin the data embedded in the engine and in the oracle;
  * logpdf!(m, x, block, transform) (src/model/simulation.jl:77-90): prior terms of the block's own nodes (plus the
    log-Jacobian of the link when the block samples on the transformed scale, src/distributions/transformdistribution.jl:34-48,
    61, 75-78) + the log-densities of the block's stochastic targets, every term through scipy.stats distributions
    parameterised as Distributions.jl does (Normal(mu, sd), InverseGamma(shape, scale), Gamma(shape, scale),
    Exponential(scale), Binomial, Poisson, Bernoulli);
  * gelmandiag (src/output/gelmandiag.jl:5-59) and summarystats / mcse_bm / ESS (src/output/stats.jl:85-94,
    src/output/mcse.jl:10-19) on a seeded chain array, straight from the formulas (scipy.stats.f.ppf for the F quantile).
The reference itself cannot run here (Julia 0.5), so these are formula-level goldens: they pin the
oracle and the CUDA path against an independent implementation, not against Julia output.
"""
import json
import os
import re

import numpy as np
import scipy.stats as st

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _nums(text):
    return [float(v) for v in re.findall(r"-?\d+\.?\d*(?:[eE]-?\d+)?", text)]


def _field(src, name):
    m = re.search(r":" + name + r"\s*=>\s*\[(.*?)\]", src, re.S)
    return np.array(_nums(m.group(1)))


def reference_data():
    seeds = open(f"{REF}/doc/examples/seeds.jl").read()
    pumps = open(f"{REF}/doc/examples/pumps.jl").read()
    rats = open(f"{REF}/doc/examples/rats.jl").read()
    line = open(f"{REF}/doc/tutorial/line.jl").read()
    surgical = open(f"{REF}/doc/examples/surgical.jl").read()
    dyes = open(f"{REF}/doc/examples/dyes.jl").read()
    d = {
        "seeds": {k: _field(seeds, k) for k in ("r", "n", "x1", "x2")},
        "pumps": {k: _field(pumps, k) for k in ("y", "t")},
        "line": {k: _field(line, k) for k in ("x", "y")},
        "surgical": {k: _field(surgical, k) for k in ("r", "n")},
        "dyes": {"y": _field(dyes, "y"), "batch": np.repeat(np.arange(6), 5)},                                # dyes.jl:16
    }
    y = _field(rats, "y"); x = _field(rats, "x")
    assert y.size == 150 and x.size == 5
    d["rats"] = {"y": y, "rat": np.repeat(np.arange(30), 5), "Xm": np.tile(x, 30) - x.mean(), "xbar": x.mean()}   # rats.jl:37-45
    return d


# ---- Distributions.jl parameterisations -------------------------------------------------------------------------
def normal(x, mu, sd): return st.norm.logpdf(x, mu, sd)
def invgamma(x, a, scale): return st.invgamma.logpdf(x, a, scale=scale)
def gamma(x, a, scale): return st.gamma.logpdf(x, a, scale=scale)
def exponential(x, scale): return st.expon.logpdf(x, scale=scale)
def invlogit(e): return 1.0 / (np.exp(-e) + 1.0)   # src/utils.jl:64


def line_blocks(D, s):
    b, s2 = s[:2], s[2]
    mu = b[0] + b[1] * D["x"]
    lik = normal(D["y"], mu, np.sqrt(s2)).sum()
    pb = normal(b, 0, np.sqrt(1000.0)).sum()
    ps2 = invgamma(s2, 0.001, 0.001)
    return {"beta": pb + lik, "s2_transformed": ps2 + np.log(s2) + lik, "s2_constrained": ps2 + lik,
            "beta_s2_transformed": pb + ps2 + np.log(s2) + lik}


def seeds_blocks(D, s):
    al, s2, b = s[:4], s[4], s[5:]
    eta = al[0] + al[1] * D["x1"] + al[2] * D["x2"] + al[3] * D["x1"] * D["x2"] + b
    lik = st.binom.logpmf(D["r"], D["n"], invlogit(eta)).sum()
    pa = normal(al, 0, 1000.0).sum()
    pb = normal(b, 0, np.sqrt(s2)).sum()
    return {"alpha": pa + lik, "b": pb + lik, "s2_transformed": invgamma(s2, 0.001, 0.001) + np.log(s2) + pb}


def rats_blocks(D, s):
    mua, mub, s2a, s2b, s2c = s[:5]
    al, be = s[5:35], s[35:65]
    mu = al[D["rat"]] + be[D["rat"]] * D["Xm"]
    lik = normal(D["y"], mu, np.sqrt(s2c)).sum()
    pal = normal(al, mua, np.sqrt(s2a)).sum(); pbe = normal(be, mub, np.sqrt(s2b)).sum()
    pmu_a = normal(mua, 0, 1000.0); pmu_b = normal(mub, 0, 1000.0)
    ig = lambda v: invgamma(v, 0.001, 0.001)
    return {"s2_c": ig(s2c) + lik, "alpha": pal + lik, "mu_alpha_s2_alpha": pmu_a + ig(s2a) + pal, "beta": pbe + lik,
            "mu_beta_s2_beta": pmu_b + ig(s2b) + pbe,
            "nuts_alpha_beta_mu": pmu_a + pmu_b + pal + pbe + lik,
            "slice_s2c_s2a_s2b": ig(s2c) + ig(s2a) + ig(s2b) + pal + pbe + lik}


def pumps_blocks(D, s):
    a, b, th = s[0], s[1], s[2:]
    lik = st.poisson.logpmf(D["y"], th * D["t"]).sum()
    pth = gamma(th, a, 1.0 / b).sum()
    pa = exponential(a, 1.0); pb = gamma(b, 0.1, 1.0)
    return {"alpha_beta_constrained": pa + pb + pth, "theta_constrained": pth + lik,
            "alpha_beta_transformed": pa + pb + np.log(a) + np.log(b) + pth, "theta_transformed": pth + np.log(th).sum() + lik}


def surgical_blocks(D, s):
    mu, s2, b = s[0], s[1], s[2:]
    lik = st.binom.logpmf(D["r"], D["n"], invlogit(b)).sum()
    pb = normal(b, mu, np.sqrt(s2)).sum()
    pmu = normal(mu, 0, 1000.0); ps2 = invgamma(s2, 0.001, 0.001)
    return {"b": pb + lik, "mu_s2_constrained": pmu + ps2 + pb, "mu_s2_transformed": pmu + ps2 + np.log(s2) + pb}


def dyes_blocks(D, s):
    s2b, th, s2w, mu = s[0], s[1], s[2], s[3:]
    lik = normal(D["y"], mu[D["batch"]], np.sqrt(s2w)).sum()
    pmu = normal(mu, th, np.sqrt(s2b)).sum()
    pth = normal(th, 0, 1000.0); ig = lambda v: invgamma(v, 0.001, 0.001)
    return {"nuts_mu_theta": pth + pmu + lik, "slice_s2w_s2b": ig(s2w) + ig(s2b) + pmu + lik, "theta": pth + pmu, "mu": pmu + lik}


def glm_block(X, y, beta):
    eta = X @ beta
    p = invlogit(eta)
    lik = np.where(y == 1, np.log(p), np.log1p(-p)).sum()   # Bernoulli(p): SURVEY.md App. B
    return normal(beta, 0, np.sqrt(1000.0)).sum() + lik, X.T @ (y - p) - beta / 1000.0


def glm_family_block(X, y, beta, family, sigma=1.0):
    """The GLM family of north_star: 1 = Poisson / log link, 2 = Normal / identity link with known sd."""
    eta = X @ beta
    prior = normal(beta, 0, np.sqrt(1000.0)).sum()
    if family == 1:
        return prior + st.poisson.logpmf(y, np.exp(eta)).sum(), X.T @ (y - np.exp(eta)) - beta / 1000.0
    return prior + normal(y, eta, sigma).sum(), X.T @ ((y - eta) / sigma ** 2) - beta / 1000.0


def states(rng, tpl, n):
    if tpl == "line":
        return np.column_stack([rng.normal(0.5, 1, n), rng.normal(0.8, 0.5, n), rng.gamma(2, 0.7, n)])
    if tpl == "seeds":
        return np.column_stack([rng.normal(0, 0.7, (n, 4)), rng.gamma(2, 0.1, n), rng.normal(0, 0.4, (n, 21))])
    if tpl == "rats":
        return np.column_stack([rng.normal(240, 10, n), rng.normal(6, 1, n), rng.gamma(3, 60, n), rng.gamma(3, 0.1, n), rng.gamma(3, 12, n),
                                rng.normal(240, 15, (n, 30)), rng.normal(6, 0.6, (n, 30))])
    if tpl == "pumps":
        return np.column_stack([rng.gamma(2, 0.5, n), rng.gamma(2, 0.5, n), rng.gamma(1.5, 0.6, (n, 10))])
    if tpl == "dyes":
        return np.column_stack([rng.gamma(2, 1500, n), rng.normal(1525, 20, n), rng.gamma(3, 900, n), rng.normal(1525, 40, (n, 6))])
    if tpl == "surgical":
        return np.column_stack([rng.normal(-2.5, 0.3, n), rng.gamma(2, 0.1, n), rng.normal(-2.5, 0.5, (n, 12))])
    raise ValueError(tpl)


# ---- gelmandiag / summarystats (src/output/gelmandiag.jl:5-59, stats.jl:85-94, mcse.jl:10-19) ---------------------
def gelmandiag(c, alpha=0.05):
    n, p, m = c.shape
    S2 = np.array([np.cov(c[:, :, k], rowvar=False).reshape(p, p) for k in range(m)])
    W = S2.mean(axis=0)
    psibar = c.mean(axis=0).T                                   # m x p
    B = n * np.cov(psibar, rowvar=False).reshape(p, p)
    w, b = np.diag(W), np.diag(B)
    s2 = np.array([np.diag(S2[k]) for k in range(m)])           # m x p
    psibar2 = psibar.mean(axis=0)
    var_w = s2.var(axis=0, ddof=1) / m
    var_b = 2 * b ** 2 / (m - 1)
    cov1 = np.array([np.cov(s2[:, j], psibar[:, j] ** 2)[0, 1] for j in range(p)])
    cov2 = np.array([np.cov(s2[:, j], psibar[:, j])[0, 1] for j in range(p)])
    var_wb = n / m * (cov1 - 2 * psibar2 * cov2)
    V = (n - 1) / n * w + (m + 1) / (m * n) * b
    var_V = ((n - 1) ** 2 * var_w + ((m + 1) / m) ** 2 * var_b + 2 * (n - 1) * (m + 1) / m * var_wb) / n ** 2
    df = 2 * V ** 2 / var_V
    B_df = m - 1
    W_df = 2 * w ** 2 / var_w
    psrf = np.sqrt((df + 3) / (df + 1) * ((n - 1) / n + (m + 1) / (m * n) * b / w))
    q = st.f.ppf(1 - alpha / 2, B_df, W_df)
    upper = np.sqrt((df + 3) / (df + 1) * ((n - 1) / n + (m + 1) / (m * n) * b / w * q))
    return np.column_stack([psrf, upper])


def summarystats(c, batch=100):
    n, p, m = c.shape
    out = np.empty((p, 5))
    for j in range(p):
        x = c[:, j, :].T.reshape(-1)                            # vec(): chain-major
        N = x.size
        sd = x.std(ddof=1)
        nb = N // batch
        bm = x[:nb * batch].reshape(nb, batch).mean(axis=1)
        mcse = bm.std(ddof=1) / np.sqrt(nb)                     # sem of the batch means
        out[j] = [x.mean(), sd, sd / np.sqrt(N), mcse, min((sd / mcse) ** 2, n)]
    return out


def ar1_chains(rng, n, p, m):
    c = np.empty((n, p, m))
    for j in range(p):
        rho = 0.3 + 0.2 * j
        e = rng.normal(size=(n, m))
        x = np.zeros((n, m)); x[0] = e[0]
        for t in range(1, n):
            x[t] = rho * x[t - 1] + np.sqrt(1 - rho ** 2) * e[t]
        c[:, j, :] = (j + 1.0) * x + 3.0 * j + 0.15 * rng.normal(size=m)   # chain-specific offsets: PSRF > 1
    c[:, p - 1, :] = np.exp(0.3 * c[:, p - 1, :])                              # a positive column
    return c


def main():
    D = reference_data()
    rng = np.random.default_rng(20261018)
    out = {"_about": "formula-level golden vectors (scipy.stats restatement of logpdf! per block); see make_golden.py",
           "data": {k: {kk: np.asarray(vv).tolist() for kk, vv in v.items()} for k, v in D.items() if k not in ("surgical", "dyes")}, "blocks": {}}
    fn = {"line": line_blocks, "seeds": seeds_blocks, "rats": rats_blocks, "pumps": pumps_blocks}
    for tpl in ("line", "seeds", "rats", "pumps"):
        S = states(rng, tpl, 12)
        vals = [fn[tpl](D[tpl], s) for s in S]
        out["blocks"][tpl] = {"states": S.tolist(), "logpdf": {k: [float(v[k]) for v in vals] for k in vals[0]}}
    # GLM: synthetic data of a small shape, committed with the fixture
    N, d = 64, 7
    X = rng.normal(size=(N, d)); X[:, 0] = 1.0
    y = (rng.uniform(size=N) < invlogit(X @ (rng.normal(size=d) / np.sqrt(d)))).astype(float)
    B = rng.normal(scale=0.6, size=(12, d))
    lg = [glm_block(X, y, b) for b in B]
    out["blocks"]["glm"] = {"X": X.tolist(), "y": y.tolist(), "states": B.tolist(),
                            "logpdf": {"beta": [float(v[0]) for v in lg]}, "grad": {"beta": [v[1].tolist() for v in lg]}}
    with open(os.path.join(HERE, "block_logpdf.json"), "w") as f:
        json.dump(out, f)
    # diagnostics (drawn before the GLM-family fixtures below so that adding fixtures leaves earlier files unchanged)
    c = ar1_chains(rng, 400, 3, 4)
    # GLM family: Poisson / log and Normal / identity on small synthetic designs
    fam = {"_about": "GLM family golden vectors (scipy.stats), see make_golden.py"}
    N, d = 96, 6
    Xf = rng.normal(scale=0.5, size=(N, d)); Xf[:, 0] = 1.0
    bt = rng.normal(size=d) / np.sqrt(d)
    yp = rng.poisson(np.exp(Xf @ bt)).astype(float)
    yn = Xf @ bt + 0.7 * rng.normal(size=N)
    Bf = rng.normal(scale=0.4, size=(12, d))
    for name, family, yy, sg in (("poisson", 1, yp, 1.0), ("normal", 2, yn, 0.7)):
        lg = [glm_family_block(Xf, yy, b, family, sg) for b in Bf]
        fam[name] = {"family": family, "sigma": sg, "X": Xf.tolist(), "y": yy.tolist(), "states": Bf.tolist(),
                     "logpdf": [float(v[0]) for v in lg], "grad": [v[1].tolist() for v in lg]}
    with open(os.path.join(HERE, "glm_family.json"), "w") as f:
        json.dump(fam, f)
    # templates added after the first fixture files (own RNG stream, so earlier files do not change)
    rng2 = np.random.default_rng(20261019)
    extra = {"_about": "block_logpdf fixtures of templates added later; same construction as block_logpdf.json",
             "data": {"surgical": {k: v.tolist() for k, v in D["surgical"].items()}}, "blocks": {}}
    S = states(rng2, "surgical", 12)
    vals = [surgical_blocks(D["surgical"], s) for s in S]
    extra["blocks"]["surgical"] = {"states": S.tolist(), "logpdf": {k: [float(v[k]) for v in vals] for k in vals[0]}}
    extra["data"]["dyes"] = {k: np.asarray(v).tolist() for k, v in D["dyes"].items()}
    S = states(rng2, "dyes", 12)
    vals = [dyes_blocks(D["dyes"], s) for s in S]
    extra["blocks"]["dyes"] = {"states": S.tolist(), "logpdf": {k: [float(v[k]) for v in vals] for k in vals[0]}}
    with open(os.path.join(HERE, "block_logpdf_extra.json"), "w") as f:
        json.dump(extra, f)
    diag = {"_about": "gelmandiag / summarystats golden values from the formulas of src/output/{gelmandiag,stats,mcse}.jl; see make_golden.py",
            "chains_shape": list(c.shape), "chains": c.tolist(), "gelmandiag_alpha_0.05": gelmandiag(c).tolist(),
            "gelmandiag_log_last_column": gelmandiag(np.concatenate([c[:, :2, :], np.log(c[:, 2:, :])], axis=1)).tolist(),
            "summarystats_bm100": summarystats(c).tolist()}
    with open(os.path.join(HERE, "diagnostics.json"), "w") as f:
        json.dump(diag, f)
    print("wrote", os.path.join(HERE, "block_logpdf.json"), os.path.join(HERE, "diagnostics.json"))


if __name__ == "__main__":
    main()
