"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (mamba.jl_b200/mambacuda) never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TPL = {"line": 0, "seeds": 1, "rats": 2, "pumps": 3, "glm": 4, "surgical": 5, "dyes": 6, "salm": 7, "equiv": 8, "blocker": 9, "stacks": 10, "magnesium": 11, "oxford": 12, "epil": 13}
KIND = {"amwg": 0, "slice_uni": 1, "slice_multi": 2, "rwm": 3, "nuts": 4, "hmc": 5, "amm": 6, "gibbs": 7, "mala": 8}
MAX_BLOCK_NODES = 8


class BlockDesc(C.Structure):
    """Mirror of mcu_block_desc (include/mambacuda.h)."""
    _fields_ = [
        ("kind", C.c_int32), ("n_nodes", C.c_int32), ("nodes", C.c_int32 * MAX_BLOCK_NODES),
        ("transform", C.c_int32), ("adapt", C.c_int32), ("batchsize", C.c_int32),
        ("proposal", C.c_int32), ("L", C.c_int32), ("grad", C.c_int32), ("max_depth", C.c_int32),
        ("n_scale", C.c_int32), ("target", C.c_double), ("epsilon", C.c_double),
        ("beta", C.c_double), ("amm_scale", C.c_double), ("scale", C.POINTER(C.c_double)),
    ]


def make_desc(kind, nodes, scale=None, transform=None, adapt=0, batchsize=0, proposal=0, L=0,
              grad=0, max_depth=0, target=0.0, epsilon=0.0, beta=0.0, amm_scale=0.0):
    """Build a BlockDesc; returns (desc, keepalive)."""
    d = BlockDesc()
    d.kind = KIND[kind] if isinstance(kind, str) else int(kind)
    nodes = list(nodes)
    d.n_nodes = len(nodes)
    for i, n in enumerate(nodes):
        d.nodes[i] = n
    if transform is None:
        transform = 0 if d.kind in (1, 2) else 1
    d.transform = int(transform)
    d.adapt, d.batchsize, d.proposal, d.L, d.grad, d.max_depth = adapt, batchsize, proposal, L, grad, max_depth
    d.target, d.epsilon, d.beta, d.amm_scale = target, epsilon, beta, amm_scale
    keep = None
    if scale is not None:
        keep = np.ascontiguousarray(np.atleast_1d(np.asarray(scale, dtype=np.float64)).ravel(order="F"))
        d.n_scale = keep.size
        d.scale = keep.ctypes.data_as(C.POINTER(C.c_double))
    return d, keep


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "mambacuda.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_set_data.argtypes = [C.c_void_p, C.c_char_p, dp, C.c_int64]
        L.orc_set_scheme.argtypes = [C.c_void_p, C.c_int, C.POINTER(BlockDesc)]
        L.orc_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_names.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_logpdf.argtypes = [C.c_void_p, C.c_int, C.c_int64, dp, dp, dp]
        L.orc_logpdf_nodes.argtypes = [C.c_void_p, C.c_uint32, C.c_int64, dp, dp]
        L.orc_predict.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, dp, dp, C.POINTER(C.c_int64)]
        L.orc_gradlogpdf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, dp, dp, dp, dp]
        L.orc_unlist.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.orc_tune_size.restype = C.c_int64
        L.orc_tune_size.argtypes = [C.c_void_p]
        L.orc_kept.restype = C.c_int64
        L.orc_kept.argtypes = [C.c_int64] * 3
        L.orc_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, dp, C.c_int64, C.c_double,
                              C.c_int64, C.c_int64, C.c_int64, dp, dp, dp, dp, C.c_int64, C.c_int]
        L.orc_kept2.restype = C.c_int64
        L.orc_kept2.argtypes = [C.c_int64] * 4
        L.orc_run2.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_uint64, dp, C.c_int64, C.c_double,
                               C.c_int64, dp, C.c_int64, C.c_int64, C.c_int64, dp, dp, dp, dp, dp, C.c_int64, C.c_int, C.c_int]
        L.orc_gelmandiag.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_int), dp]
        L.orc_summarystats.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int64, dp]
        L.orc_fquantile.restype = C.c_double
        L.orc_fquantile.argtypes = [C.c_double] * 3
        L.orc_digamma.restype = C.c_double
        L.orc_digamma.argtypes = [C.c_double]
        L.orc_philox.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_draws.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int), dp]
        L.orc_line_logf.restype = C.c_double
        L.orc_line_logf.argtypes = [dp, dp]
        L.orc_standalone_line.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_int64, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One model template + sampling scheme evaluated by the CPU restatement."""

    def __init__(self, template, glm_d=0):
        self.L = lib()
        self.tid = TPL[template] if isinstance(template, str) else int(template)
        self.h = self.L.orc_create(self.tid, glm_d)
        if not self.h:
            raise RuntimeError("orc_create failed")
        self._keep = []

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def set_data(self, name, arr):
        a = _f64(np.asarray(arr)).ravel()
        self._chk(self.L.orc_set_data(self.h, name.encode(), _dp(a), a.size))

    def set_scheme(self, blocks):
        """blocks: list of dicts of make_desc kwargs."""
        arr = (BlockDesc * len(blocks))()
        self._keep = []
        for i, b in enumerate(blocks):
            d, keep = make_desc(**b)
            arr[i] = d
            self._keep.append(keep)
        self._chk(self.L.orc_set_scheme(self.h, len(blocks), arr))

    def dims(self):
        D, p = C.c_int(), C.c_int()
        self.L.orc_dims(self.h, C.byref(D), C.byref(p))
        return D.value, p.value

    def names(self, monitoronly=True):
        buf = C.create_string_buffer(1 << 16)
        self.L.orc_names(self.h, int(monitoronly), buf, len(buf))
        return buf.value.decode().split("\n")

    def logpdf(self, block, state, x=None):
        state = _f64(np.atleast_2d(state)); x = _f64(None if x is None else np.atleast_2d(x))
        lp = np.empty(state.shape[0])
        self._chk(self.L.orc_logpdf(self.h, block, state.shape[0], _dp(state), _dp(x), _dp(lp)))
        return lp

    def logpdf_nodes(self, mask, state):
        state = _f64(np.atleast_2d(state))
        lp = np.empty(state.shape[0])
        self._chk(self.L.orc_logpdf_nodes(self.h, C.c_uint32(int(mask)), C.c_int64(state.shape[0]), _dp(state), _dp(lp)))
        return lp

    def predict(self, state, seed, stream_id=0):
        state = _f64(np.atleast_2d(state))
        n = C.c_int64()
        self._chk(self.L.orc_predict(self.h, seed, stream_id, 0, None, None, C.byref(n)))
        out = np.empty((state.shape[0], n.value))
        self._chk(self.L.orc_predict(self.h, seed, stream_id, state.shape[0], _dp(state), _dp(out), C.byref(n)))
        return out

    def gradlogpdf(self, block, state, x=None, mode=0):
        state = _f64(np.atleast_2d(state)); x = _f64(None if x is None else np.atleast_2d(x))
        k = self.unlist(block, state[0]).size
        lp = np.empty(state.shape[0]); g = np.empty((state.shape[0], k))
        self._chk(self.L.orc_gradlogpdf(self.h, block, mode, state.shape[0], _dp(state), _dp(x), _dp(lp), _dp(g)))
        return lp, g

    def unlist(self, block, state):
        state = _f64(state); out = np.empty(4096)
        k = self.L.orc_unlist(self.h, block, _dp(state), _dp(out))
        return out[:k].copy()

    def tune_size(self):
        return self.L.orc_tune_size(self.h)

    def run(self, n_chains, inits, iters, burnin=0, thin=1, seed=123, chain_offset=0, jitter_sd=0.0,
            ext_u=None, nthreads=1, store=True, chain_ids=None, iter0=0, tune_in=None, margins=False, partial=False):
        """mcmc() for n_chains chains.  chain_ids: explicit global chain ids (scattered samples of a large run);
        iter0 > 0: restart — inits holds one state record per chain and tune_in their tune records;
        margins=True also returns [n_chains x iters], the smallest decision margin of every iteration (samplers.hpp);
        partial=True: one segment of a longer run (burn-in may extend past it)."""
        inits = _f64(np.atleast_2d(inits))
        D, p = self.dims()
        assert inits.shape[1] == D
        if chain_ids is not None:
            chain_ids = np.ascontiguousarray(chain_ids, dtype=np.int64)
            n_chains = chain_ids.size
        kept = self.L.orc_kept2(iter0, iters, burnin, thin)
        out = np.full((kept, p, n_chains), np.nan, order="F") if store else None
        final = np.empty((n_chains, D))
        nt = self.tune_size()
        tune = np.zeros((n_chains, max(nt, 1)))
        tin = None
        if tune_in is not None and nt > 0:
            tin = _f64(np.atleast_2d(tune_in)); assert tin.shape == (n_chains, nt)
        marg = np.full((n_chains, iters), np.inf) if margins else None
        npc = 0
        if ext_u is not None:
            ext_u = _f64(np.atleast_2d(ext_u)); npc = ext_u.shape[1]
        ids = None if chain_ids is None else chain_ids.ctypes.data_as(C.POINTER(C.c_int64))
        rc = self.L.orc_run2(self.h, n_chains, chain_offset, ids, seed, _dp(inits), inits.shape[0], jitter_sd, iter0, _dp(tin),
                             iters, burnin, thin, _dp(out), _dp(final), _dp(tune), _dp(marg), _dp(ext_u), npc, nthreads, int(partial))
        self._chk(rc)
        if margins:
            return out, final, tune[:, :nt], marg
        return out, final, tune[:, :nt]


def gelmandiag(chains, alpha=0.05, linkcode=None):
    L = lib()
    c = np.asfortranarray(chains, dtype=np.float64)
    n, p, m = c.shape
    psrf = np.empty((p, 2))
    lc = None
    if linkcode is not None:
        lc = (C.c_int * p)(*[int(v) for v in linkcode])
    rc = L.orc_gelmandiag(_dp(c), n, p, m, alpha, lc, _dp(psrf))
    if rc != 0:
        raise ValueError("less than 2 chains supplied to gelman diagnostic")
    return psrf


def summarystats(chains, etype=0, batch=100):
    L = lib()
    c = np.asfortranarray(chains, dtype=np.float64)
    n, p, m = c.shape
    out = np.empty((p, 5))
    L.orc_summarystats(_dp(c), n, p, m, etype, batch, _dp(out))
    return out


DKIND = {"normal": 1, "invgamma": 2, "gamma": 3, "exponential": 4, "binomial": 5, "poisson": 6, "bernoulli": 7, "laplace": 8, "uniform": 9, "beta": 10, "truncnormal": 11}


def link(kind, a, b, x, lo=0.0, hi=0.0):
    """(link(x), invlink(link(x)), log-Jacobian) of one univariate distribution: transformdistribution.jl:6-93."""
    L = lib()
    L.orc_link.argtypes = [C.c_int] + [C.c_double] * 5 + [C.POINTER(C.c_double)]
    out = np.empty(3)
    L.orc_link(DKIND[kind], a, b, lo, hi, x, _dp(out))
    return out


def udist_logpdf(kind, a, b, x, lo=0.0, hi=0.0):
    L = lib()
    L.orc_udist_logpdf.restype = C.c_double
    L.orc_udist_logpdf.argtypes = [C.c_int] + [C.c_double] * 5
    return L.orc_udist_logpdf(DKIND[kind], a, b, lo, hi, x)


def rwm_draws(proposal, seed, n):
    L = lib()
    out = np.empty(n)
    L.orc_rwm_draws.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.POINTER(C.c_double)]
    L.orc_rwm_draws(int(proposal), int(seed), int(n), _dp(out))
    return out


def philox(ctr, key):
    L = lib()
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    L.orc_philox(c, k, o)
    return list(o)


def draws(seed, chain, it, block, kinds, kind=0):
    L = lib()
    n = len(kinds)
    out = np.empty(n)
    L.orc_draws(seed, chain, it, block, kind, n, (C.c_int * n)(*kinds), _dp(out))
    return out


def line_logf(x, grad=False):
    L = lib()
    x = _f64(x)
    g = np.empty(3) if grad else None
    v = L.orc_line_logf(_dp(x), _dp(g))
    return (v, g) if grad else v


def standalone_line(which, n, burnin, seed=123):
    L = lib()
    out = np.empty((n, 3), order="F")
    rc = L.orc_standalone_line(which, seed, n, burnin, _dp(out))
    assert rc == 0
    return out
