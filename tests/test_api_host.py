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
        api.RWM("beta", 1.0, proposal="cosine")


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
        api.Model("epil")                                    # doc/examples/epil.jl has no compiled template


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
