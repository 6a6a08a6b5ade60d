"""Writes tests/golden/block_logpdf_large.json: formula-level golden vectors (scipy.stats / numpy — no oracle, no engine) for the two
examples with 240-300 unobserved elements per chain, doc/examples/oxford.jl and doc/examples/epil.jl: logpdf!(m, x, block, transform)
(src/model/simulation.jl:77-90) of every sampling block of the scripts' own schemes, the observed-node log densities, and the monitored
Logical column alpha0 of epil.  Data parsed from the reference's own scripts.
Run in the build container (reads /root/reference):  python tests/golden/make_golden_large.py"""
import json
import os
import re

import numpy as np
import scipy.special as sp
import scipy.stats as st

REF = "/root/reference/doc/examples"
HERE = os.path.dirname(os.path.abspath(__file__))


def arr(src, name):
    m = re.search(r":%s =>\s*\[(.*?)\]" % name, src, re.S)
    return np.array([float(v) for v in re.findall(r"-?\d+\.?\d*", m.group(1))])


def oxford_data():
    src = open(f"{REF}/oxford.jl").read()
    return {k: arr(src, k) for k in ("r1", "n1", "r0", "n0", "year")}


def epil_data():
    src = open(f"{REF}/epil.jl").read()
    D = {k: arr(src, k) for k in ("Trt", "Base", "Age", "V4")}
    D["y"] = arr(src, "y").reshape(59, 4)            # the matrix literal: one patient per row
    return D


def normal(x, mu, sd): return st.norm.logpdf(x, mu, sd)
def ig(x): return st.invgamma.logpdf(x, 0.001, scale=0.001)


def oxford_blocks(D, s):     # state: alpha, beta1, beta2, s2, b[120], mu[120]
    al, b1, b2, s2, b, mu = s[0], s[1], s[2], s[3], s[4:124], s[124:244]
    yr = D["year"]
    lr0 = st.binom.logpmf(D["r0"], D["n0"], sp.expit(mu)).sum()
    lr1 = st.binom.logpmf(D["r1"], D["n1"], sp.expit(mu + al + b1 * yr + b2 * (yr ** 2 - 22.0) + b)).sum()
    pb = normal(b, 0, np.sqrt(s2)).sum()
    return {"amwg_alpha_beta1_beta2": normal(np.array([al, b1, b2]), 0, 1000.0).sum() + lr1,     # AMWG([:alpha, :beta1, :beta2], 1.0): oxford.jl:97
            "slice_s2": ig(s2) + pb,                                                                 # Slice(:s2, 1.0): :98 (constrained scale)
            "slice_mu": normal(mu, 0, 1000.0).sum() + lr0 + lr1,                                     # Slice(:mu, 1.0): :99
            "slice_b": pb + lr1,                                                                     # Slice(:b, 1.0): :100
            "s2_transformed": ig(s2) + np.log(s2) + pb,
            "r0": lr0, "r1": lr1}


def epil_cov(D):             # epil.jl:25-30
    lb = np.log(D["Base"] / 4); trt = D["Trt"]; bt = lb * trt; la = np.log(D["Age"]); v4 = D["V4"]
    return lb, trt, bt, la, v4


def epil_blocks(D, s):       # state: a0, alpha_Base, alpha_Trt, alpha_BT, alpha_Age, alpha_V4, s2_b1, s2_b, b1[59], b[59 x 4 column-major]
    a0, aB, aT, aBT, aA, aV, s2b1, s2b = s[:8]
    b1 = s[8:67]; b = s[67:303].reshape(59, 4, order="F")
    lb, trt, bt, la, v4 = epil_cov(D)
    eta = (a0 + aB * (lb - lb.mean()) + aT * (trt - trt.mean()) + aBT * (bt - bt.mean()) + aA * (la - la.mean()))[:, None] + aV * (v4 - v4.mean())[None, :] + b1[:, None] + b
    lik = st.poisson.logpmf(D["y"], np.exp(eta)).sum()
    pb1 = normal(b1, 0, np.sqrt(s2b1)).sum(); pb = normal(b, 0, np.sqrt(s2b)).sum()
    alpha0 = a0 - aB * lb.mean() - aT * trt.mean() - aBT * bt.mean() - aA * la.mean() - aV * v4.mean()      # epil.jl:85-91
    return {"amwg_coefficients": normal(s[:6], 0, 100.0).sum() + lik,      # AMWG([:a0, :alpha_Base, ...], 0.1): epil.jl:126-127
            "slice_b1": pb1 + lik,                                         # Slice(:b1, 0.5): :128
            "slice_b": pb + lik,                                           # Slice(:b, 0.5): :129
            "slice_s2_b1_s2_b": ig(s2b1) + ig(s2b) + pb1 + pb,            # Slice([:s2_b1, :s2_b], 1.0): :130
            "y": lik, "alpha0": alpha0}


def main():
    rng = np.random.default_rng(20261024)
    n = 8
    Dox, Dep = oxford_data(), epil_data()
    Sox = np.column_stack([rng.normal(0.55, 0.1, n), rng.normal(-0.04, 0.02, n), rng.normal(0.005, 0.004, n), rng.gamma(2, 0.02, n),
                           rng.normal(0, 0.15, (n, 120)), rng.normal(-2.0, 0.5, (n, 120))])
    Sep = np.column_stack([rng.normal(1.6, 0.2, n), rng.normal(0.9, 0.15, n), rng.normal(-0.8, 0.3, n), rng.normal(0.25, 0.2, n), rng.normal(0.45, 0.3, n),
                           rng.normal(-0.1, 0.08, n), rng.gamma(3, 0.08, n), rng.gamma(3, 0.045, n), rng.normal(0, 0.45, (n, 59)), rng.normal(0, 0.35, (n, 236))])
    vo = [oxford_blocks(Dox, s) for s in Sox]
    ve = [epil_blocks(Dep, s) for s in Sep]
    out = {"_about": "block_logpdf fixtures for oxford and epil; see make_golden_large.py",
           "oxford": {"states": Sox.tolist(), "logpdf": {k: [float(v[k]) for v in vo] for k in vo[0]}},
           "epil": {"states": Sep.tolist(), "logpdf": {k: [float(v[k]) for v in ve] for k in ve[0]}}}
    with open(os.path.join(HERE, "block_logpdf_large.json"), "w") as f:
        json.dump(out, f)
    print("wrote block_logpdf_large.json")


if __name__ == "__main__":
    main()
