"""CPU tests of the host-side finalisation inside libmambacuda.so (no device needed) and of the multi-rank
all-reduce logic (gloo, world_size 2): PSRF / summary statistics from per-rank partial sums must equal the
reference computation (oracle) over all chains."""
import os
import sys

import numpy as np
import pytest

from test_oracle_kat import ar1_chains

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyLocal:
    """What a rank's Engine offers for diagnostics, computed with numpy from that rank's chains [n x p x m_local]
    (the device computes the same per-chain moments with Welford updates and reduces them with kernels)."""

    def __init__(self, chains, monlink):
        import ctypes as C
        from mambacuda import _lib
        self.c = chains
        self.monlink = monlink
        self.L = _lib.lib()
        self.C = C

    def minmax(self):
        return np.stack([self.c.min(axis=(0, 2)), self.c.max(axis=(0, 2))], axis=1)

    def link_codes(self, transform, minmax):
        codes = []
        for j, ml in enumerate(self.monlink):
            c = 0
            if transform:
                c = 1 if ml == 1 or (ml == -1 and minmax[j, 0] > 0) else 0
            codes.append(c)
        return np.array(codes)

    def moments(self, codes, center):
        n, p, m = self.c.shape
        sums = np.zeros((p, 7))
        for j in range(p):
            x = np.log(self.c[:, j, :]) if codes is not None and codes[j] else self.c[:, j, :]
            d = x.mean(axis=0) - (center[j, 0] if center is not None else 0.0)
            e = x.var(axis=0, ddof=1) - (center[j, 1] if center is not None else 0.0)
            sums[j] = [m, d.sum(), (d * d).sum(), e.sum(), (e * e).sum(), (e * d).sum(), (e * d * d).sum()]
        return sums, n

    def gelman_from_moments(self, n_kept, center, sums, alpha=0.05):
        C = self.C
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        p = sums.shape[0]
        psrf = np.empty((p, 2))
        center = np.ascontiguousarray(center); sums = np.ascontiguousarray(sums)
        assert self.L.mcu_gelman_from_moments(int(n_kept), p, dp(center), dp(sums), float(alpha), dp(psrf)) == 0
        return psrf

    def summary_sums(self, center):
        n, p, m = self.c.shape
        nb = n // 100
        sums = np.zeros((p, 8))
        for j in range(p):
            x = self.c[:, j, :]
            mean = x.mean(axis=0); M2 = ((x - mean) ** 2).sum(axis=0)
            bm = x[: nb * 100].reshape(nb, 100, m).mean(axis=1)
            bmean = bm.mean(axis=0); bM2 = ((bm - bmean) ** 2).sum(axis=0)
            c1, c2 = (center[j] if center is not None else (0.0, 0.0))
            sums[j] = [m, mean.sum(), M2.sum(), ((mean - c1) ** 2).sum(), nb * m, (nb * bmean).sum(), bM2.sum(), (nb * (bmean - c2) ** 2).sum()]
        return sums

    # ---- the packed two-round protocol (mamba.jl_b200/csrc/diagproto.hpp), as the device kernels compute it ----------------
    def _scales(self, j):
        x = self.c[:, j, :]
        with np.errstate(invalid="ignore", divide="ignore"):
            return [x, np.log(x), np.log(x / (1 - x)) if self.monlink[j] == -1 else np.zeros_like(x)]

    def diag_round1(self):
        n, p, m = self.c.shape
        nb = n // 100
        mn, mx, sums = np.empty(p), np.empty(p), np.zeros((p, 9))
        for j in range(p):
            x = self.c[:, j, :]
            mn[j], mx[j] = x.min(), x.max()
            sc = self._scales(j)
            bmean = x[: nb * 100].reshape(nb, 100, m).mean(axis=(0, 1))
            sums[j] = [m] + [v for y in sc for v in (y.mean(axis=0).sum(), y.var(axis=0, ddof=1).sum())] + [nb * m, (nb * bmean).sum()]
        return np.concatenate([mn, mx, sums.ravel()])

    def diag_round2(self, transform, r1):
        n, p, m = self.c.shape
        nb = n // 100
        out = np.zeros((p, 15))
        s1 = r1[2 * p:].reshape(p, 9)
        for j in range(p):
            ml = self.monlink[j]
            code = 0
            if transform:
                code = 1 if ml == 1 else ((2 if r1[p + j] < 1 else 1) if (ml == -1 and r1[j] > 0) else 0)
            c1, c2 = s1[j, 1 + 2 * code] / s1[j, 0], s1[j, 2 + 2 * code] / s1[j, 0]
            k1, k2 = s1[j, 1] / s1[j, 0], (s1[j, 8] / s1[j, 7] if s1[j, 7] > 0 else 0.0)
            y = self._scales(j)[code]
            d = y.mean(axis=0) - c1; e = y.var(axis=0, ddof=1) - c2
            x = self.c[:, j, :]
            mean = x.mean(axis=0); M2 = ((x - mean) ** 2).sum(axis=0)
            bm = x[: nb * 100].reshape(nb, 100, m).mean(axis=1)
            bmean = bm.mean(axis=0); bM2 = ((bm - bmean) ** 2).sum(axis=0)
            out[j] = [m, d.sum(), (d * d).sum(), e.sum(), (e * e).sum(), (e * d).sum(), (e * d * d).sum(),
                      m, mean.sum(), M2.sum(), ((mean - k1) ** 2).sum(), nb * m, (nb * bmean).sum(), bM2.sum(), (nb * (bmean - k2) ** 2).sum()]
        # pair sums for the multivariate PSRF: { Σ cov_k(i, j), Σ d_i d_j } on the raw scale (transform = false) or the nodes' own link scale
        pairs = []
        ys = [np.log(self.c[:, j, :]) if (transform and self.monlink[j] == 1) else self.c[:, j, :] for j in range(p)]
        for i in range(p):
            for j in range(i + 1, p):
                ci = ys[i] - ys[i].mean(axis=0); cj = ys[j] - ys[j].mean(axis=0)
                cov = (ci * cj).sum(axis=0) / (n - 1)
                pairs += [cov.sum(), (self._d(i, transform, r1) * self._d(j, transform, r1)).sum()]
        return np.concatenate([out.ravel(), np.array(pairs)])

    def _d(self, j, transform, r1):
        p = self.c.shape[1]
        s1 = r1[2 * p:].reshape(p, 9)
        ml = self.monlink[j]
        code = 0
        if transform:
            code = 1 if ml == 1 else ((2 if r1[p + j] < 1 else 1) if (ml == -1 and r1[j] > 0) else 0)
        return self._scales(j)[code].mean(axis=0) - s1[j, 1 + 2 * code] / s1[j, 0]

    def diag_finish(self, alpha, transform, r1, r2, mpsrf=False):
        from mambacuda.engine import diag_finish
        return diag_finish(self.c.shape[0], self.monlink, alpha, transform, r1, r2, mpsrf=mpsrf)

    def summary_from_sums(self, n_kept, center, sums):
        C = self.C
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        p = sums.shape[0]
        out = np.empty((p, 5))
        center = np.ascontiguousarray(center); sums = np.ascontiguousarray(sums)
        assert self.L.mcu_summary_from_sums(int(n_kept), p, dp(center), dp(sums), dp(out)) == 0
        return out


def make_chains(m=6, seed=4):
    c = ar1_chains(300, 3, m, 0.5, seed)
    c[:, 1, :] = np.exp(0.2 * c[:, 1, :])          # a positive column (log link)
    c[:, 2, :] = 100.0 + c[:, 2, :]                # a far-from-zero column: exercises the centring
    return c


def make_chains4(m=6, seed=4):
    c3 = make_chains(m, seed)
    u = 1.0 / (1.0 + np.exp(-ar1_chains(300, 1, m, 0.5, seed + 1)))     # a Logical column inside (0, 1): logit link (chains.jl:241-243)
    return np.concatenate([c3, u], axis=1)


def test_gelman_from_moments_equals_reference_gelmandiag(oracle, mcu_built):
    from mambacuda import distributed as mdist
    c = make_chains()
    for transform, linkcode in ((False, None), (True, [0, 1, -1])):
        local = NumpyLocal(c, [0, 1, -1])
        got = mdist.global_gelman(local, 0.05, transform)
        want = oracle.gelmandiag(c, 0.05, linkcode)
        np.testing.assert_allclose(got, want, rtol=1e-9)


def test_summary_from_sums_equals_reference_summarystats(oracle, mcu_built):
    from mambacuda import distributed as mdist
    c = make_chains()      # 300 kept per chain: batches of 100 never straddle chains
    got = mdist.global_summary(NumpyLocal(c, [0, 1, -1]))
    want = oracle.summarystats(c, 0, 100)
    np.testing.assert_allclose(got, want, rtol=1e-9)


def test_packed_protocol_equals_reference_diagnostics(oracle, mcu_built):
    # one rank: gelmandiag(transform) with identity / log / heuristic-log / heuristic-logit columns, and summarystats, in one pass
    from mambacuda import distributed as mdist
    c = make_chains4()
    monlink = [0, 1, -1, -1]
    for transform, linkcode in ((False, None), (True, [0, 1, -1, -1])):
        psrf, summ, codes = mdist.global_diagnostics(NumpyLocal(c, monlink), 0.05, transform)
        np.testing.assert_allclose(psrf, oracle.gelmandiag(c, 0.05, linkcode), rtol=1e-9)
        np.testing.assert_allclose(summ, oracle.summarystats(c, 0, 100), rtol=1e-9)
        assert list(codes) == ([0, 1, 1, 2] if transform else [0, 0, 0, 0])
    with pytest.raises(ValueError, match="less than 2 chains"):
        mdist.global_diagnostics(NumpyLocal(c[:, :, :1], monlink), 0.05, False)
    # multivariate PSRF from the pair sums of round 2 (gelmandiag.jl:49-55) against the host-array entry point on the draws themselves
    from mambacuda import api
    local = NumpyLocal(c[:, :3, :], [0, 1, 0])
    for transform, codes in ((False, None), (True, [0, 1, 0])):
        b1 = local.diag_round1(); r2 = local.diag_round2(transform, b1)
        psrf, summ, cd, mv = local.diag_finish(0.05, transform, b1, r2, mpsrf=True)
        want = api._chains_gelman(c[:, :3, :], 0.05, codes, True)
        np.testing.assert_allclose(psrf, want[:-1], rtol=1e-9)
        np.testing.assert_allclose(mv, want[-1, 0], rtol=1e-9)
    # a Logical column resolved to log / logit by the heuristic has no streamed co-moments: NaN, not a wrong number
    b1 = NumpyLocal(c, monlink).diag_round1(); r2 = NumpyLocal(c, monlink).diag_round2(True, b1)
    assert np.isnan(NumpyLocal(c, monlink).diag_finish(0.05, True, b1, r2, mpsrf=True)[3])


def test_f_quantile_of_the_product_against_scipy(mcu_built):
    # exercised through mcu_gelman_from_moments: upper limit = sqrt(corr * (Rf + Rr * qf(0.975, m-1, W_df)))
    import scipy.stats as st
    c = make_chains(m=4, seed=9)
    local = NumpyLocal(c, [0, 0, 0])
    s0, n = local.moments(None, None)
    center = np.stack([s0[:, 1] / s0[:, 0], s0[:, 3] / s0[:, 0]], axis=1)
    s1, _ = local.moments(None, center)
    psrf = local.gelman_from_moments(n, center, s1)
    for j in range(3):
        x = c[:, j, :]
        s2 = x.var(axis=0, ddof=1); w = s2.mean(); var_w = s2.var(ddof=1) / 4
        W_df = 2 * w * w / var_w
        m = 4
        b = n * x.mean(axis=0).var(ddof=1)
        Rr = (m + 1) / (m * n) * b / w
        ratio = (psrf[j, 1] ** 2 / (psrf[j, 0] ** 2) * ((n - 1) / n + Rr) - (n - 1) / n) / Rr
        assert ratio == pytest.approx(st.f.ppf(0.975, m - 1, W_df), rel=1e-7)


def _rank_main(rank, world, port, tmp):
    sys.path[:0] = [os.path.join(ROOT, "mamba.jl_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mambacuda import distributed as mdist
    c = make_chains4(m=6)
    mine = c[:, :, rank * 3:(rank + 1) * 3]            # contiguous chain shards, as the engine shards them
    if rank == 1:
        mine = c[:, :, 3:5]                            # uneven shards: 3 + 2 chains ... and the sixth chain is left out on purpose below
    local = NumpyLocal(mine, [0, 1, -1, -1])
    psrf, summ, _ = mdist.global_diagnostics(local, 0.05, True)
    np.save(os.path.join(tmp, f"psrf{rank}.npy"), psrf)
    np.save(os.path.join(tmp, f"summ{rank}.npy"), summ)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process(oracle, mcu_built, tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    c = make_chains4(m=6)[:, :, :5]
    want_psrf = oracle.gelmandiag(c, 0.05, [0, 1, -1, -1])
    want_summ = oracle.summarystats(c, 0, 100)
    for r in range(2):
        np.testing.assert_allclose(np.load(tmp_path / f"psrf{r}.npy"), want_psrf, rtol=1e-9)
        np.testing.assert_allclose(np.load(tmp_path / f"summ{r}.npy"), want_summ, rtol=1e-9)
