"""Writes tests/golden/block_logpdf_more.json: formula-level golden vectors (scipy.stats, no oracle, no engine) for the templates
added after block_logpdf{,_extra}.json — salm, equiv and blocker (doc/examples/{salm,equiv,blocker}.jl).  Same construction as
make_golden.py: data parsed from the reference's own scripts, logpdf!(m, x, block, transform) (src/model/simulation.jl:77-90) per
sampling block of the scripts' schemes, plus the observed-node log density (the deviance term of dic, src/output/modelstats.jl:3-13).
Run in the build container (reads /root/reference):  python tests/golden/make_golden_more.py"""
import json
import os
import re

import numpy as np
import scipy.stats as st

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def nums(text):
    return [float(v) for v in re.findall(r"-?\d+\.?\d*(?:[eE]-?\d+)?", text)]


def normal(x, mu, sd): return st.norm.logpdf(x, mu, sd)
def invgamma(x, a, scale): return st.invgamma.logpdf(x, a, scale=scale)


def salm_data():
    src = open(f"{REF}/doc/examples/salm.jl").read()
    y = np.array(nums(re.search(r":y => reshape\(\s*\[(.*?)\]", src, re.S).group(1)))        # reshape(v, 3, 6): column-major, plate fastest
    x = np.array(nums(re.search(r":x => \[(.*?)\]", src, re.S).group(1)))
    assert y.size == 18 and x.size == 6
    return {"y": y, "x": x}


def equiv_data():
    src = open(f"{REF}/doc/examples/equiv.jl").read()
    group = np.array(nums(re.search(r":group => \[(.*?)\]", src, re.S).group(1)))
    ymat = np.array(nums(re.search(r":y =>\s*\[(.*?)\]", src, re.S).group(1))).reshape(10, 2)   # matrix literal: rows = subjects
    return {"group": group, "y": ymat.flatten(order="F")}                                      # unlist order: column-major


def salm_blocks(D, s):      # state: s2, gamma, beta, alpha, lambda[18]
    s2, gamma, beta, alpha, lam = s[0], s[1], s[2], s[3], s[4:]
    x = np.repeat(D["x"], 3)
    mu = np.exp(alpha + beta * np.log(x + 10) + gamma * x + lam)
    lik = st.poisson.logpmf(D["y"], mu).sum()
    pl = normal(lam, 0, np.sqrt(s2)).sum()
    pabc = normal(np.array([alpha, beta, gamma]), 0, 1000.0).sum()
    ps2 = invgamma(s2, 0.001, 0.001)
    return {"slice_alpha_beta_gamma": pabc + lik,                       # Slice([:alpha, :beta, :gamma], ...): salm.jl:63
            "amwg_lambda_s2_transformed": pl + ps2 + np.log(s2) + lik,  # AMWG([:lambda, :s2], 0.1): salm.jl:64
            "y": lik}


def equiv_blocks(D, s):     # state: s2_2, s2_1, pi, phi, mu, delta[20]
    s22, s21, pi, phi, mu, dl = s[0], s[1], s[2], s[3], s[4], s[5:]
    T = np.concatenate([D["group"], 3 - D["group"]])
    j = np.repeat([1, 2], 10)
    m = mu + (-1.0) ** (T - 1) * phi / 2 + (-1.0) ** (j - 1) * pi / 2 + dl
    lik = normal(D["y"], m, np.sqrt(s21)).sum()
    pd = normal(dl, 0, np.sqrt(s22)).sum()
    ig = lambda v: invgamma(v, 0.001, 0.001)
    return {"nuts_delta": pd + lik,                                                   # NUTS(:delta): equiv.jl:89
            "slice_mu_phi_pi": normal(np.array([mu, phi, pi]), 0, 1000.0).sum() + lik, # Slice([:mu, :phi, :pi], 1.0): equiv.jl:90
            "slice_s2_1_s2_2": ig(s21) + ig(s22) + pd + lik,                           # Slice([:s2_1, :s2_2], 1.0, Univariate): equiv.jl:91
            "y": lik}


def blocker_data():
    src = open(f"{REF}/doc/examples/blocker.jl").read()
    return {k: np.array(nums(re.search(r":" + k + r" =>\s*\[(.*?)\]", src, re.S).group(1))) for k in ("rt", "nt", "rc", "nc")}


def blocker_blocks(D, s):   # state: s2, d, delta_new, mu[22], delta[22]
    s2, d, dn, mu, dl = s[0], s[1], s[2], s[3:25], s[25:47]
    invlogit = lambda e: 1.0 / (np.exp(-e) + 1.0)
    lc = st.binom.logpmf(D["rc"], D["nc"], invlogit(mu)).sum()
    lt = st.binom.logpmf(D["rt"], D["nt"], invlogit(mu + dl)).sum()
    pmu = normal(mu, 0, 1000.0).sum(); pdl = normal(dl, d, np.sqrt(s2)).sum(); pdn = normal(dn, d, np.sqrt(s2))
    return {"amwg_mu": pmu + lc + lt,                                              # AMWG(:mu, 0.1): blocker.jl:84
            "amwg_delta_delta_new": pdl + pdn + lt,                                # AMWG([:delta, :delta_new], 0.1): blocker.jl:85
            "slice_d_s2": normal(d, 0, 1000.0) + invgamma(s2, 0.001, 0.001) + pdn + pdl,   # Slice([:d, :s2], 1.0): blocker.jl:86
            "rc": lc, "rt": lt}


def stacks_data():
    src = open(f"{REF}/doc/examples/stacks.jl").read()
    y = np.array(nums(re.search(r":y => \[(.*?)\]", src, re.S).group(1)))
    x = np.array(nums(re.search(r":x =>\s*\[(.*?)\]", src, re.S).group(1))).reshape(21, 3)
    return {"y": y, "x": x}


def stacks_blocks(D, s):    # state: beta0, beta[3], s2
    b0, be, s2 = s[0], s[1:4], s[4]
    z = (D["x"] - D["x"].mean(axis=0)) / D["x"].std(axis=0, ddof=1)                # stacks.jl:32-37
    lik = st.laplace.logpdf(D["y"], b0 + z @ be, s2).sum()                          # Laplace(mu[i], s2): stacks.jl:45
    pb = normal(np.concatenate([[b0], be]), 0, 1000.0).sum()
    ps2 = invgamma(s2, 0.001, 0.001)
    return {"nuts_beta0_beta": pb + lik, "slice_s2": ps2 + lik, "y": lik}            # NUTS([:beta0, :beta]), Slice(:s2, 1.0): stacks.jl:104-105


def stacks_monitor(D, s):   # b[3], b0, sigma, outlier[1, 3, 4, 21]: stacks.jl:68-88
    b0, be, s2 = s[0], s[1:4], s[4]
    mx, sx = D["x"].mean(axis=0), D["x"].std(axis=0, ddof=1)
    z = (D["x"] - mx) / sx
    b = be / sx
    sigma = np.sqrt(2.0) * s2
    out = (np.abs((D["y"] - (b0 + z @ be)) / sigma) > 2.5).astype(float)
    return np.concatenate([b, [b0 - b @ mx, sigma], out[[0, 2, 3, 20]]])


def main():
    rng = np.random.default_rng(20261020)
    D = {"salm": salm_data(), "equiv": equiv_data(), "blocker": blocker_data(), "stacks": stacks_data()}
    n = 12
    S = {"salm": np.column_stack([rng.gamma(2, 0.05, n), rng.normal(-0.001, 0.0005, n), rng.normal(0.35, 0.1, n), rng.normal(2.0, 0.3, n),
                                  rng.normal(0, 0.25, (n, 18))]),
         "equiv": np.column_stack([rng.gamma(2, 0.01, n), rng.gamma(2, 0.01, n), rng.normal(-0.2, 0.1, n), rng.normal(0, 0.1, n),
                                   rng.normal(1.44, 0.05, n), rng.normal(0, 0.1, (n, 20))])}
    rng_b = np.random.default_rng(20261021)
    S["blocker"] = np.column_stack([rng_b.gamma(2, 0.01, n), rng_b.normal(-0.25, 0.06, n), rng_b.normal(-0.25, 0.15, n), rng_b.normal(-2.2, 0.4, (n, 22)),
                                    rng_b.normal(-0.25, 0.15, (n, 22))])
    rng_s = np.random.default_rng(20261022)
    S["stacks"] = np.column_stack([rng_s.normal(17.5, 1.0, n), rng_s.normal(0, 2.0, (n, 3)) + [7.7, 2.4, -1.0], rng_s.gamma(6, 0.45, n)])
    fn = {"salm": salm_blocks, "equiv": equiv_blocks, "blocker": blocker_blocks, "stacks": stacks_blocks}
    out = {"_about": "block_logpdf fixtures for salm and equiv; see make_golden_more.py",
           "data": {k: {kk: vv.tolist() for kk, vv in v.items()} for k, v in D.items()}, "blocks": {}}
    for tpl in ("salm", "equiv", "blocker", "stacks"):
        vals = [fn[tpl](D[tpl], s) for s in S[tpl]]
        out["blocks"][tpl] = {"states": S[tpl].tolist(), "logpdf": {k: [float(v[k]) for v in vals] for k in vals[0]}}
    out["blocks"]["stacks"]["monitor"] = [stacks_monitor(D["stacks"], s).tolist() for s in S["stacks"]]
    with open(os.path.join(HERE, "block_logpdf_more.json"), "w") as f:
        json.dump(out, f)
    print("wrote block_logpdf_more.json")


if __name__ == "__main__":
    main()
