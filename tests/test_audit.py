"""The audited comparison used by the GPU parity tests (helpers.audit_divergence) and the oracle features behind it
(decision margins, restart from ModelStates, explicit global chain ids), exercised oracle-against-oracle on the CPU."""
import numpy as np
import pytest

import helpers


def run(oracle, name, ids, iters, burnin, thin, seed=11, **kw):
    tpl, blocks, inits = helpers.scheme(name)
    orc = oracle.Oracle(tpl)
    ob = [helpers.oracle_block(b) for b in blocks]
    for b in ob:
        if b["kind"] == "nuts":
            b["max_depth"] = 10
    orc.set_scheme(ob)
    return orc, inits, orc.run(0, inits, iters, burnin=burnin, thin=thin, seed=seed, jitter_sd=0.05, chain_ids=np.asarray(ids), margins=True, **kw)


def kept(iters, burnin, thin):
    return [i for i in range(1, iters + 1) if i > burnin and (i - burnin) % thin == 0]


@pytest.mark.parametrize("name", ["seeds_amwg", "rats_slice_amwg", "pumps_gibbs_amwg", "rats_nuts_slice", "line_amm"])
def test_restart_and_scattered_ids_reproduce_one_run(oracle, name):
    ids = [3, 1000, 77777]
    orc, inits, (out, st, tune, marg) = run(oracle, name, ids, 40, 15, 2)
    # chain 1000 alone, as part of another call
    o1 = orc.run(1, inits, 40, burnin=15, thin=2, seed=11, jitter_sd=0.05, chain_offset=1000)
    np.testing.assert_array_equal(o1[0][:, :, 0], out[:, :, 1])
    # mcmc(mc, iters) restart (mcmc.jl:3-16): 17 + 23 iterations from the stored ModelStates == 40 iterations
    a = orc.run(0, inits, 17, burnin=15, thin=2, seed=11, jitter_sd=0.05, chain_ids=np.asarray(ids), margins=True)
    b = orc.run(0, a[1], 23, burnin=15, thin=2, seed=11, chain_ids=np.asarray(ids), iter0=17, tune_in=a[2], margins=True)
    np.testing.assert_array_equal(np.concatenate([a[0], b[0]], axis=0), out)
    np.testing.assert_array_equal(b[1], st)
    np.testing.assert_array_equal(np.nan_to_num(b[2]), np.nan_to_num(tune))
    np.testing.assert_array_equal(np.concatenate([a[3], b[3]], axis=1), marg)
    assert np.isfinite(marg).all() and (marg > 0).all()


def test_audit_accepts_identical_runs_and_only_explained_divergence(oracle):
    ids = list(range(8))
    iters, burnin, thin = 60, 20, 4
    orc, inits, (out, st, tune, marg) = run(oracle, "seeds_amwg", ids, iters, burnin, thin)
    o = (out, st, tune)
    n, ties = helpers.audit_divergence((out.copy(), st.copy(), tune.copy()), o, marg, kept(iters, burnin, thin), 0)
    assert n == 8 and ties == []
    # a chain whose kept samples differ from row 3 on, with no near-tie in that window: rejected
    bad = out.copy(); bad[3:, 0, 5] += 1e-3
    with pytest.raises(AssertionError, match="part from the oracle"):
        helpers.audit_divergence((bad, st.copy(), tune.copy()), o, marg, kept(iters, burnin, thin), 0)
    # the same difference IS accepted when the oracle took a decision at rounding distance from its threshold in that window
    m2 = marg.copy(); m2[5, kept(iters, burnin, thin)[3] - 2] = 1e-13          # iteration kept[3] - 1, inside (kept[2], kept[3]]
    n, ties = helpers.audit_divergence((bad, st.copy(), tune.copy()), o, m2, kept(iters, burnin, thin), 0)
    assert n == 7 and ties[0][0] == 5 and ties[0][1] == (kept(iters, burnin, thin)[2] + 1, kept(iters, burnin, thin)[3])
    # ... but not when the near-tie lies outside the window of the first difference
    m3 = marg.copy(); m3[5, 2] = 1e-13
    with pytest.raises(AssertionError, match="part from the oracle"):
        helpers.audit_divergence((bad, st.copy(), tune.copy()), o, m3, kept(iters, burnin, thin), 0)
    # a difference that only shows in the final state is looked for after the last kept row
    st2 = st.copy(); st2[2, 7] += 1e-4
    with pytest.raises(AssertionError, match="part from the oracle"):
        helpers.audit_divergence((out.copy(), st2, tune.copy()), o, marg, kept(iters, burnin, thin), 0)
    # integer tune columns of agreeing chains must be equal, not close
    t2 = tune.copy(); t2[1, 6] += 1e-9
    with pytest.raises(AssertionError, match="integer tune column"):
        helpers.audit_divergence((out.copy(), st.copy(), t2), o, marg, kept(iters, burnin, thin), 0, int_tune_cols=[6])
