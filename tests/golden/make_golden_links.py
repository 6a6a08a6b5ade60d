"""Writes tests/golden/links.json: formula-level golden vectors (scipy.stats / numpy only — no oracle, no engine) for the link layer
of the reference, src/distributions/transformdistribution.jl:

  * link / invlink / log-Jacobian of every support shape the generic TransformDistribution methods distinguish (:6-48): two-sided
    [a, b] -> logit((x - a) / (b - a)) with Jacobian log((x - a)(b - x) / (b - a)); lower bound only -> log(x - a); the unit interval
    of UnitDistribution (:83-93) -> logit(x), Jacobian log(x (1 - x));
  * logpdf!(m, x, block, transform) (src/model/simulation.jl:77-90) of every sampling block of doc/examples/magnesium.jl — the example
    whose parameter nodes carry Uniform(a, b) and truncated priors — on the constrained scale (the script's Slice blocks) and on the
    link scale (its AMWG(:mu) block: Uniform(-10, 10) under the two-sided link), data parsed from the reference's own script.

Run in the build container (reads /root/reference):  python tests/golden/make_golden_links.py"""
import json
import os
import re

import numpy as np
import scipy.special as sp
import scipy.stats as st

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ints(name, src):
    return np.array([float(v) for v in re.search(r":%s => \[(.*?)\]" % name, src, re.S).group(1).split(",")])


def magnesium_data():
    src = open(f"{REF}/doc/examples/magnesium.jl").read()
    D = {k: ints(k, src) for k in ("rt", "nt", "rc", "nc")}
    s2 = 1 / (D["rt"] + 0.5) + 1 / (D["nt"] - D["rt"] + 0.5) + 1 / (D["rc"] + 0.5) + 1 / (D["nc"] - D["rc"] + 0.5)     # magnesium.jl:13-16
    D["s2_0"] = np.array([1 / np.mean(1 / s2)])                                                                    # :17
    return D


def jac_bounded(x, a, b):
    return np.log((x - a) * (b - x) / (b - a))


def magnesium_blocks(D, s):
    """state: priors[6], mu[6], theta[6 x 8] column-major, pc[6 x 8] column-major"""
    pr, mu = s[:6], s[6:12]
    theta = s[12:60].reshape(6, 8, order="F"); pc = s[60:108].reshape(6, 8, order="F")
    s2_0 = D["s2_0"][0]
    tau = np.array([np.sqrt(pr[0]), np.sqrt(pr[1]), pr[2], np.sqrt(s2_0 * (1 / pr[3] - 1)), np.sqrt(s2_0) * (1 / pr[4] - 1), np.sqrt(pr[5])])   # :60-70
    sd6 = np.sqrt(s2_0 / sp.erf(0.75))
    lp_priors = np.array([st.invgamma.logpdf(pr[0], 0.001, scale=0.001), st.uniform.logpdf(pr[1], 0, 50), st.uniform.logpdf(pr[2], 0, 50),
                          st.uniform.logpdf(pr[3], 0, 1), st.uniform.logpdf(pr[4], 0, 1), st.truncnorm.logpdf(pr[5], 0, np.inf, loc=0, scale=sd6)])
    jac_priors = np.array([np.log(pr[0]), jac_bounded(pr[1], 0, 50), jac_bounded(pr[2], 0, 50), jac_bounded(pr[3], 0, 1), jac_bounded(pr[4], 0, 1), np.log(pr[5])])
    lp_mu = st.uniform.logpdf(mu, -10, 20)
    jac_mu = jac_bounded(mu, -10, 10)
    lp_theta = st.norm.logpdf(theta, mu[:, None], tau[:, None]).sum()
    lp_pc = st.uniform.logpdf(pc, 0, 1).sum()
    jac_pc = jac_bounded(pc, 0, 1).sum()
    rcx = st.binom.logpmf(D["rc"][None, :], D["nc"][None, :], pc).sum()
    pt = sp.expit(theta + sp.logit(pc))
    rtx = st.binom.logpmf(D["rt"][None, :], D["nt"][None, :], pt).sum()
    return {
        "amwg_theta": lp_theta + rtx,                                        # AMWG(:theta, 0.1): magnesium.jl:99 (Normal: identity link)
        "amwg_mu_transformed": lp_mu.sum() + jac_mu.sum() + lp_theta,        # AMWG(:mu, 0.1): :100, on logit((mu + 10) / 20)
        "slice_pc": lp_pc + rcx + rtx,                                       # Slice(:pc, 0.25, Univariate): :101, constrained scale
        "slice_priors": lp_priors.sum() + lp_theta,                          # Slice(:priors, [...], Univariate): :102, constrained scale
        "slice_pc_transformed": lp_pc + jac_pc + rcx + rtx,                  # the same blocks sampled on the link scale
        "priors_mu_transformed": lp_priors.sum() + jac_priors.sum() + lp_mu.sum() + jac_mu.sum() + lp_theta,
        "rcx": rcx, "rtx": rtx,
        "monitor": np.concatenate([tau, np.exp(mu)]).tolist(),               # tau[6], OR[6] = exp(mu): :56-70
        # unlist(block, transform = true) (src/samplers/sampler.jl:113-115): the block vector on the link scale
        "unlist_priors_mu": np.concatenate([[np.log(pr[0]), sp.logit(pr[1] / 50), sp.logit(pr[2] / 50), sp.logit(pr[3]), sp.logit(pr[4]), np.log(pr[5])],
                                            sp.logit((mu + 10) / 20)]).tolist(),
    }


def main():
    rng = np.random.default_rng(20261023)
    D = magnesium_data()
    n = 12
    S = np.column_stack([rng.gamma(2, 0.2, n), rng.uniform(0.05, 3, n), rng.uniform(0.05, 2, n), rng.uniform(0.05, 0.95, n), rng.uniform(0.05, 0.95, n), rng.gamma(2, 0.2, n),
                         rng.normal(-0.7, 0.4, (n, 6)), rng.normal(-0.7, 0.5, (n, 48)), rng.beta(2, 20, (n, 48))])
    vals = [magnesium_blocks(D, s) for s in S]
    out = {"_about": "link-layer fixtures (transformdistribution.jl:6-93) and the block densities of doc/examples/magnesium.jl; see make_golden_links.py",
           "data": {k: v.tolist() for k, v in D.items()}, "states": S.tolist(),
           "logpdf": {k: [float(v[k]) for v in vals] for k in vals[0] if k not in ("monitor", "unlist_priors_mu")},
           "monitor": [v["monitor"] for v in vals], "unlist_priors_mu": [v["unlist_priors_mu"] for v in vals]}
    # scalar link table: (a, b, x) -> link(x), invlink(link(x)), log-Jacobian
    tab = []
    for a, b in ((0.0, 1.0), (0.0, 50.0), (-10.0, 10.0), (2.5, 2.75)):
        for x in a + (b - a) * np.array([1e-6, 0.01, 0.3, 0.5, 0.9, 1 - 1e-9]):
            tab.append([a, b, float(x), float(sp.logit((x - a) / (b - a))), float(np.log((x - a) * (b - x) / (b - a)))])
    for a in (0.0, -3.0, 7.5):
        for x in a + np.array([1e-9, 0.2, 1.0, 40.0]):
            tab.append([a, None, float(x), float(np.log(x - a)), float(np.log(x - a))])
    out["links"] = tab
    with open(os.path.join(HERE, "links.json"), "w") as f:
        json.dump(out, f)
    print("wrote links.json")


if __name__ == "__main__":
    main()
