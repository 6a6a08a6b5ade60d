"""Statistical pin of the oracle: posterior means of the reference's published runs (doc/*.rst tables; single
MersenneTwister realisations that cannot be replayed) must be within 3 MCSE, MCSE = sqrt(ref^2 + ours^2)."""
import numpy as np
import pytest

import helpers


def within_3_mcse(ss, names, ref):
    for nm, (mean, mcse_ref) in ref.items():
        j = names.index(nm)
        tol = 3.0 * np.hypot(mcse_ref, ss[j, 3])
        assert abs(ss[j, 0] - mean) < tol, (nm, ss[j, 0], mean, tol)


def test_line_standalone_amwg_slice(oracle):
    # doc/examples/line_amwg_slice.jl:35-43, table doc/examples/line_amwg_slice.rst:26-31 (1 x 10,000)
    ref = {"b0": (0.64401798, 0.060725564), "b1": (0.78985612, 0.017888106), "s2": (1.20785292, 0.062566344)}
    outs = [oracle.standalone_line(4, 10000, 1000, seed=s) for s in range(8)]
    c = np.stack(outs, axis=2)
    ss = oracle.summarystats(c, 0, 100)
    within_3_mcse(ss, ["b0", "b1", "s2"], ref)


def test_line_standalone_nuts_and_slice(oracle):
    # doc/samplers/nuts.jl:34-43 and doc/samplers/slice.jl:31-40: same posterior as the tutorial table
    # (doc/tutorial.rst:432-436: beta1 0.5971 [0.0169], beta2 0.8017 [0.0048], s2 1.2204 [0.1018])
    ref = {"b0": (0.5971183, 0.016925598), "b1": (0.8017036, 0.004793345), "s2": (1.2203777, 0.101798287)}
    for which, n, burn in ((1, 5000, 1000), (2, 5000, 0), (3, 5000, 0)):
        c = np.stack([oracle.standalone_line(which, n, burn, seed=s)[burn:] for s in range(4)], axis=2)
        within_3_mcse(oracle.summarystats(c, 0, 100), ["b0", "b1", "s2"], ref)


def test_line_standalone_amm_hmc_mala_rwm(oracle):
    # the remaining stand-alone sampler demos of the reference, with the scripts' own settings, on the closed-form log posterior they all share
    # (doc/samplers/amm.jl:28-35 AMMVariate(eye(3)), hmc.jl:37-51 HMC(0.1, 50) without / with Sigma = eye(3), mala.jl:37-50 MALA(0.1) without / with
    # Sigma, rwm.jl:28-35 RWM([0.5, 0.25, 1.0], SymUniform)): the .rst pages publish no output, so they are pinned to the tutorial table of the same
    # posterior (doc/tutorial.rst:432-436)
    ref = {"b0": (0.5971183, 0.016925598), "b1": (0.8017036, 0.004793345), "s2": (1.2203777, 0.101798287)}
    for which, n, burn, drop in ((5, 5000, 1000, 1000), (6, 5000, 0, 200), (7, 5000, 0, 200), (8, 20000, 0, 2000), (9, 20000, 0, 2000), (10, 20000, 0, 2000)):
        c = np.stack([oracle.standalone_line(which, n, burn, seed=100 + s)[drop:] for s in range(6)], axis=2)
        within_3_mcse(oracle.summarystats(c, 0, 100), ["b0", "b1", "s2"], ref)


def test_line_model_based_nuts_slice(oracle):
    # doc/tutorial/line.jl:48-49,99: scheme1 = [NUTS(:beta), Slice(:s2, 3.0)], 3 x 10,000, burnin 250, thin 2
    ref = {"beta[1]": (0.5971183, 0.016925598), "beta[2]": (0.8017036, 0.004793345), "s2": (1.2203777, 0.101798287)}
    tpl, blocks, inits = helpers.scheme("line_nuts_slice")
    blocks = [dict(b, grad="forward") if b["kind"] == "nuts" else b for b in blocks]   # the reference differentiates numerically
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(6, inits, 10000, burnin=250, thin=2, seed=11, nthreads=6)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_seeds_reference_scheme(oracle):
    # doc/examples/seeds.jl:69-75 (AMM + AMWG + AMWG, 2 x 12,500, burnin 2,500, thin 2), table doc/examples/seeds.rst:43-48
    ref = {"alpha0": (-0.556154341, 0.0101730837), "alpha1": (0.088700176, 0.0128300598), "alpha2": (1.310728093, 0.0153996801),
           "alpha12": (-0.746440855, 0.0251658152), "s2": (0.085705306, 0.0080848189)}
    tpl, blocks, inits = helpers.scheme("seeds_amm")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 12500, burnin=2500, thin=2, seed=1, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_rats_reference_scheme(oracle):
    # doc/examples/rats.jl:112-116 (intended run 2 x 10,000, burnin 2,500, thin 2), table doc/examples/rats.rst:42-46
    ref = {"s2_c": (37.2543133, 0.2337982327), "mu_beta": (6.1830663, 0.0017921615), "alpha0": (106.6259925, 0.0526804390)}
    tpl, blocks, inits = helpers.scheme("rats_slice_amwg")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 10000, burnin=2500, thin=2, seed=2, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_line_mala_scheme(oracle):
    # MALA(:beta, :s2 jointly on the transformed scale; mala.jl:67-86) targets the tutorial posterior (doc/tutorial.rst:432-436)
    ref = {"beta[1]": (0.5971183, 0.016925598), "beta[2]": (0.8017036, 0.004793345)}
    tpl, blocks, inits = helpers.scheme("line_mala_sigma")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 30000, burnin=2000, thin=2, seed=5, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_dyes_reference_scheme(oracle):
    # doc/examples/dyes.jl:60-61 (NUTS([mu, theta]) + Slice([s2_within, s2_between], 1000), 2 x 10,000, burnin 2,500, thin 2), doc/examples/dyes.rst
    ref = {"theta": (1526.7186, 0.37724897), "s2_within": (2887.5853, 76.89117959), "mu[1]": (1511.4798, 0.52158448),
           "mu[3]": (1552.6742, 0.70276515), "mu[5]": (1578.6636, 1.29216105), "mu[6]": (1487.1934, 1.23710390)}
    tpl, blocks, inits = helpers.scheme("dyes_nuts_slice")
    ob = [helpers.oracle_block(b) for b in blocks]; ob[0]["max_depth"] = 10
    o = oracle.Oracle(tpl); o.set_scheme(ob)
    out, _, _ = o.run(8, inits, 10000, burnin=2500, thin=2, seed=4, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_salm_reference_scheme(oracle):
    # doc/examples/salm.jl:63-69 (Slice([alpha, beta, gamma], [1, 1, 0.1]) + AMWG([lambda, s2], 0.1), 2 x 10,000, burnin 2,500, thin 2),
    # table doc/examples/salm.rst:43-47
    # The multivariate slice block mixes slowly (the published run has ESS 93-185 for alpha, beta, gamma, so its batch-means MCSE is
    # itself unreliable): the published means are matched within 0.75 of the published SD, and a 4x longer run must agree with itself
    # across two seeds within 3 MCSE.
    ref = {"s2": (0.0690769709, 0.04304237136), "gamma": (-0.0011250515, 0.00034536546), "beta": (0.3543443166, 0.07160779229),
           "alpha": (2.0100584321, 0.26156942610)}
    tpl, blocks, inits = helpers.scheme("salm_slice_amwg")
    runs = []
    for seed in (6, 16):
        o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
        out, _, _ = o.run(8, inits, 40000, burnin=10000, thin=2, seed=seed, nthreads=8)
        runs.append(oracle.summarystats(out, 0, 100))
    names = o.names()
    for nm, (mean, sd) in ref.items():
        j = names.index(nm)
        assert abs(runs[0][j, 0] - mean) < 0.75 * sd, (nm, runs[0][j, 0], mean)
        assert abs(runs[0][j, 0] - runs[1][j, 0]) < 3.0 * np.hypot(runs[0][j, 3], runs[1][j, 3]) + 0.02 * sd, (nm, runs[0][j], runs[1][j])


def test_equiv_reference_scheme(oracle):
    # doc/examples/equiv.jl:89-96 (NUTS(delta) + Slice([mu, phi, pi], 1.0) + Slice([s2_1, s2_2], 1.0, Univariate), 2 x 12,500, burnin 2,500,
    # thin 2), table doc/examples/equiv.rst:43-50
    ref = {"s2_2": (0.0173121833, 0.0007329722), "s2_1": (0.0184397014, 0.0005689492), "pi": (-0.1874240524, 0.0032257037),
           "phi": (-0.0035569545, 0.0035141650), "theta": (1.0002921934, 0.0036227671), "equiv": (0.9751, 0.0036666529), "mu": (1.4387396416, 0.0013735876)}
    tpl, blocks, inits = helpers.scheme("equiv_nuts_slice")
    ob = [helpers.oracle_block(b) for b in blocks]; ob[0]["max_depth"] = 10
    o = oracle.Oracle(tpl); o.set_scheme(ob)
    out, _, _ = o.run(8, inits, 12500, burnin=2500, thin=2, seed=7, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_blocker_reference_scheme(oracle):
    # doc/examples/blocker.jl:84-92 (AMWG(mu) + AMWG([delta, delta_new]) + Slice([d, s2], 1.0), 2 x 10,000, burnin 2,500, thin 2),
    # table doc/examples/blocker.rst:46-49
    ref = {"s2": (0.01822186, 0.0014150714), "d": (-0.25563567, 0.0040205781), "delta_new": (-0.25005767, 0.0050219145)}
    tpl, blocks, inits = helpers.scheme("blocker_amwg_slice")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 10000, burnin=2500, thin=2, seed=8, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_stacks_reference_scheme(oracle):
    # doc/examples/stacks.jl:104-110 (NUTS([beta0, beta]) + Slice(s2, 1.0), 2 x 10,000, burnin 2,500, thin 2), table doc/examples/stacks.rst:42-51
    ref = {"b[1]": (0.836863707, 0.0027601754), "b[2]": (0.744454449, 0.0065756939), "b[3]": (-0.116648437, 0.0015143922),
           "b0": (-38.776564595, 0.0979006137), "sigma": (3.487643717, 0.0279025494), "outlier[1]": (0.042666667, 0.0029490162),
           "outlier[4]": (0.298, 0.0089200654), "outlier[21]": (0.6064, 0.0113877443)}
    tpl, blocks, inits = helpers.scheme("stacks_nuts_slice")
    ob = [helpers.oracle_block(b) for b in blocks]; ob[0]["max_depth"] = 10
    o = oracle.Oracle(tpl); o.set_scheme(ob)
    out, _, _ = o.run(8, inits, 10000, burnin=2500, thin=2, seed=9, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_line_gibbs_scheme(oracle):
    # doc/tutorial/line.jl:54,101-102: scheme3 = [Gibbs_beta, Gibbs_s2] targets the tutorial posterior (doc/tutorial.rst:432-436)
    ref = {"beta[1]": (0.5971183, 0.016925598), "beta[2]": (0.8017036, 0.004793345), "s2": (1.2203777, 0.101798287)}
    tpl, blocks, inits = helpers.scheme("line_gibbs")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(6, inits, 10000, burnin=250, thin=2, seed=12, nthreads=6)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_surgical_reference_scheme(oracle):
    # doc/examples/surgical.jl:54-60 (NUTS(b) + Slice([mu, s2], 1.0), 2 x 10,000, burnin 2,500, thin 2), table doc/examples/surgical.rst
    ref = {"mu": (-2.550263247, 0.00352027397), "pop_mean": (0.073062651, 0.00022880854), "s2": (0.183080212, 0.00629499754),
           "p[1]": (0.053571675, 0.00059140521), "p[4]": (0.059863573, 0.00033190971), "p[8]": (0.122296440, 0.00086456417),
           "p[12]": (0.068534503, 0.00015162331)}
    tpl, blocks, inits = helpers.scheme("surgical_nuts_slice")
    ob = [helpers.oracle_block(b) for b in blocks]; ob[0]["max_depth"] = 10
    o = oracle.Oracle(tpl); o.set_scheme(ob)
    out, _, _ = o.run(8, inits, 10000, burnin=2500, thin=2, seed=4, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_pumps_gibbs_amwg_scheme(oracle):
    # BASELINE.json configs[4]: Gibbs(theta) + Gibbs(beta) + AMWG(alpha) targets the same posterior as the reference's Slice scheme
    ref = {"beta": (0.93036099, 0.01824153419), "alpha": (0.69679849, 0.00722593007), "theta[1]": (0.05991674, 0.00032725274),
           "theta[2]": (0.10125873, 0.00129985769), "theta[5]": (0.59971611, 0.00585119652), "theta[10]": (1.98475207, 0.00912748779)}
    tpl, blocks, inits = helpers.scheme("pumps_gibbs_amwg")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 10000, burnin=2500, thin=2, seed=3, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


def test_pumps_reference_scheme(oracle):
    # doc/examples/pumps.jl:52-57 (2 x 10,000, burnin 2,500, thin 2), table doc/examples/pumps.rst:43-56
    ref = {"beta": (0.93036099, 0.01824153419), "alpha": (0.69679849, 0.00722593007), "theta[1]": (0.05991674, 0.00032725274),
           "theta[2]": (0.10125873, 0.00129985769), "theta[5]": (0.59971611, 0.00585119652), "theta[7]": (0.86767451, 0.02858200254),
           "theta[9]": (1.55721556, 0.03109274798), "theta[10]": (1.98475207, 0.00912748779)}
    tpl, blocks, inits = helpers.scheme("pumps_slice")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 10000, burnin=2500, thin=2, seed=3, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), ref)


MAGNESIUM_TABLE = {   # doc/examples/magnesium.rst:45-57: mean, MCSE
    "tau[1]": (0.55098858, 0.0221132365), "tau[2]": (1.11557619, 0.0237788755), "tau[3]": (0.83211110, 0.0222839957), "tau[4]": (0.47864203, 0.0135868530),
    "tau[5]": (0.48624861, 0.0215005369), "tau[6]": (0.56841884, 0.0058505056), "OR[1]": (0.47784058, 0.0066922017), "OR[2]": (0.42895913, 0.0081170895),
    "OR[3]": (0.43118350, 0.0064385836), "OR[4]": (0.47587697, 0.0064893426), "OR[5]": (0.48545299, 0.0083912319), "OR[6]": (0.44554385, 0.0053818401)}


def test_magnesium_reference_scheme(oracle):
    # doc/examples/magnesium.jl:99-107 (AMWG(theta) + AMWG(mu) [Uniform(-10, 10) on the two-sided link] + Slice(pc) + Slice(priors), 2 x 12,500,
    # burnin 2,500, thin 2), table doc/examples/magnesium.rst:45-57; here 8 chains x 5,000 (the interpreted 108-element model is slow)
    tpl, blocks, inits = helpers.scheme("magnesium")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 5000, burnin=1500, thin=2, seed=21, nthreads=8)
    within_3_mcse(oracle.summarystats(out, 0, 100), o.names(), MAGNESIUM_TABLE)


# doc/examples/oxford.rst / epil.rst: mean, SD of single, poorly mixed runs (published ESS 104-268 of 10,000 / 12,500 draws; our own chains of the
# same schemes still show PSRF 1.1-1.5 between chains after 7,500 iterations), so the oracle is matched within the published SDs, not within MCSE
OXFORD_TABLE = {"beta2": (0.005477119, 0.0035675748), "beta1": (-0.043336269, 0.0161754258), "alpha": (0.565784774, 0.0630050896), "s2": (0.026238992, 0.0307989154)}
EPIL_TABLE = {"s2_b": (0.13523750, 0.031819272), "s2_b1": (0.24911885, 0.073166731), "alpha_V4": (-0.09287934, 0.083666872), "alpha_Age": (0.45830900, 0.394536219),
              "alpha_BT": (0.24217000, 0.190566444), "alpha_Trt": (-0.75931393, 0.397734236), "alpha_Base": (0.91104974, 0.135354470), "alpha0": (-1.35617079, 1.313240197)}


def within_published_sd(ss, names, ref):
    for nm, (mean, sd) in ref.items():
        j = names.index(nm)
        assert abs(ss[j, 0] - mean) < (2.0 if nm.startswith("s2") else 1.0) * sd, (nm, ss[j, 0], mean, sd)


def test_oxford_reference_scheme(oracle):
    # doc/examples/oxford.jl:97-106 (AMWG + three multivariate Slice blocks, 2 x 12,500, burnin 2,500, thin 2); here 8 chains x 7,500
    tpl, blocks, inits = helpers.scheme("oxford")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 7500, burnin=2500, thin=2, seed=31, nthreads=8)
    within_published_sd(oracle.summarystats(out, 0, 100), o.names(), OXFORD_TABLE)


def test_epil_reference_scheme(oracle):
    # doc/examples/epil.jl:126-136 (AMWG + Slice(b1) + Slice(b) + Slice([s2_b1, s2_b]), 2 x 15,000, burnin 2,500, thin 2); here 8 chains x 9,000
    tpl, blocks, inits = helpers.scheme("epil")
    o = oracle.Oracle(tpl); o.set_scheme([helpers.oracle_block(b) for b in blocks])
    out, _, _ = o.run(8, inits, 9000, burnin=2500, thin=2, seed=32, nthreads=8)
    within_published_sd(oracle.summarystats(out, 0, 100), o.names(), EPIL_TABLE)
