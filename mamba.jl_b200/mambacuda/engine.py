"""Engine — one handle of libmambacuda.so: all chains of one model template on one GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import BlockDesc, MambaCudaError


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def make_desc(kind, nodes, scale=None, transform=None, adapt="all", batchsize=0, proposal="normal", L=0,
              grad="analytic", max_depth=0, target=0.0, epsilon=0.0, beta=0.0, amm_scale=0.0):
    """Fill a mcu_block_desc; returns (desc, keepalive-array)."""
    d = BlockDesc()
    d.kind = _lib.KIND[kind] if isinstance(kind, str) else int(kind)
    nodes = list(nodes)
    if not 1 <= len(nodes) <= _lib.MAX_BLOCK_NODES:
        raise ValueError("a sampling block names 1..8 nodes")
    d.n_nodes = len(nodes)
    for i, n in enumerate(nodes):
        d.nodes[i] = int(n)
    if transform is None:
        transform = 0 if d.kind in (1, 2) else 1
    d.transform = int(bool(transform))
    d.adapt = _lib.ADAPT[adapt] if isinstance(adapt, str) else int(adapt)
    d.batchsize = int(batchsize)
    d.proposal = _lib.PROPOSAL[proposal] if isinstance(proposal, str) else int(proposal)
    d.L = int(L)
    d.grad = _lib.GRAD[grad] if isinstance(grad, str) else int(grad)
    d.max_depth = int(max_depth)
    d.target, d.epsilon, d.beta, d.amm_scale = float(target), float(epsilon), float(beta), float(amm_scale)
    keep = None
    if scale is not None:
        keep = np.ascontiguousarray(np.atleast_1d(np.asarray(scale, dtype=np.float64)).ravel(order="F"))
        d.n_scale = keep.size
        d.scale = keep.ctypes.data_as(C.POINTER(C.c_double))
    return d, keep


class Engine:
    def __init__(self, template, n_chains, seed=123, chain_offset=0, device=0):
        self.L = _lib.lib()
        self.h = C.c_void_p()
        tid = _lib.TPL[template] if isinstance(template, str) else int(template)
        rc = self.L.mcu_create(tid, int(n_chains), int(chain_offset), int(device), int(seed), C.byref(self.h))
        if rc != 0:
            msg = self.L.mcu_last_error(None).decode()
            self.h = None
            raise MambaCudaError(rc, msg)
        self.template = tid
        self.n_chains = int(n_chains)
        self.seed = int(seed)
        self.chain_offset = int(chain_offset)
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.mcu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise MambaCudaError(rc, self.L.mcu_last_error(self.h).decode())

    # ---- model -------------------------------------------------------------------------------
    def set_data(self, name, arr):
        a = _f64(np.asarray(arr))
        dims = (C.c_int64 * max(a.ndim, 1))(*(a.shape if a.ndim else (1,)))
        self._chk(self.L.mcu_set_data(self.h, name.encode(), max(a.ndim, 1), dims, _dp(a)))

    def set_scheme(self, blocks):
        arr = (BlockDesc * len(blocks))()
        self._keep = []
        for i, b in enumerate(blocks):
            d, keep = make_desc(**b)
            arr[i] = d
            self._keep.append(keep)
        self._chk(self.L.mcu_set_scheme(self.h, len(blocks), arr))
        self.n_blocks = len(blocks)

    def dims(self):
        D, p, nn = C.c_int(), C.c_int(), C.c_int()
        self._chk(self.L.mcu_dims(self.h, C.byref(D), C.byref(p), C.byref(nn)))
        return D.value, p.value, nn.value

    def names(self, which=1):
        buf = C.create_string_buffer(1 << 16)
        self.L.mcu_names(self.h, which, buf, len(buf))
        return buf.value.decode().split("\n")

    def tune_size(self):
        n = C.c_int64()
        self._chk(self.L.mcu_tune_size(self.h, C.byref(n)))
        return n.value

    # ---- chains ------------------------------------------------------------------------------
    def set_inits(self, x, jitter_sd=0.0):
        x = _f64(np.atleast_2d(x))
        D = self.dims()[0]
        if x.shape[1] != D:
            raise ValueError(f"initial values have {x.shape[1]} entries per record, model state has {D}")
        self._chk(self.L.mcu_set_inits(self.h, _dp(x), x.shape[0], float(jitter_sd)))

    def kept(self, first_iter, iters, burnin, thin):
        return self.L.mcu_kept(first_iter, iters, burnin, thin)

    def run(self, iters, burnin=0, thin=1, store=True, out=True, force_generic=False, glm_reference=False, partial=False, wait=True, mpsrf=False):
        """Returns the [kept × p × chains] block (Fortran order) or None when out=False.
        partial=True: this call is one segment of a longer run (MCU_RUN_PARTIAL).
        wait=False: MCU_RUN_ASYNC — the call returns once the run is queued; finish it with wait() (then samples()).
        mpsrf=True: MCU_RUN_MPSRF — also stream the within-chain covariances (multivariate PSRF without the draws)."""
        if not wait:
            out = False
        _, p, _ = self.dims()
        it0 = self.iter()
        kept = self.kept(it0, iters, burnin, thin)
        flags = (0 if store else _lib.RUN_NO_STORE) | (_lib.RUN_FORCE_GENERIC if force_generic else 0) | \
            (_lib.RUN_GLM_REFERENCE if glm_reference else 0) | (_lib.RUN_PARTIAL if partial else 0) | (0 if wait else _lib.RUN_ASYNC) | (_lib.RUN_MPSRF if mpsrf else 0)
        self._last_kept = kept
        arr = None
        if out:
            arr = np.full((kept, p, self.n_chains), np.nan, order="F")
        self._chk(self.L.mcu_run(self.h, int(iters), int(burnin), int(thin), _dp(arr) if out and kept > 0 else None, flags))
        return arr

    def wait(self):
        self._chk(self.L.mcu_wait(self.h))

    def samples(self, into=None):
        """ModelChains.value of the last run ([kept × p × chains], Fortran order); `into`: a preallocated (e.g. pinned) array."""
        _, p, _ = self.dims()
        arr = into if into is not None else np.empty((self._last_kept, p, self.n_chains), order="F")
        self._chk(self.L.mcu_get_samples(self.h, _dp(arr)))
        return arr

    def iter(self):
        it = C.c_int64()
        rc = self.L.mcu_get_state(self.h, None, None, C.byref(it))
        return it.value if rc == 0 else 0

    def get_state(self):
        D = self.dims()[0]
        nt = self.tune_size()
        v = np.empty((self.n_chains, D))
        t = np.empty((self.n_chains, max(nt, 1)))
        it = C.c_int64()
        self._chk(self.L.mcu_get_state(self.h, _dp(v), _dp(t) if nt else None, C.byref(it)))
        return v, t[:, :nt], it.value

    def set_state(self, values, tune, it):
        v = _f64(values)
        t = _f64(tune) if tune is not None and tune.size else None
        self._chk(self.L.mcu_set_state(self.h, _dp(v), _dp(t), int(it)))

    def set_external_stream(self, u):
        if u is None:
            self._chk(self.L.mcu_set_rng_mode(self.h, 0, None, 0))
            return
        u = _f64(np.atleast_2d(u))
        assert u.shape[0] == self.n_chains
        self._chk(self.L.mcu_set_rng_mode(self.h, 1, _dp(u), u.shape[1]))

    # ---- densities ---------------------------------------------------------------------------
    def logpdf(self, block, state, x=None):
        state = _f64(np.atleast_2d(state))
        x = _f64(None if x is None else np.atleast_2d(x))
        lp = np.empty(state.shape[0])
        self._chk(self.L.mcu_logpdf(self.h, block, state.shape[0], _dp(state), _dp(x), _dp(lp)))
        return lp

    def factor_counts(self):
        nn, nf = C.c_int(), C.c_int()
        self._chk(self.L.mcu_factor_counts(self.h, C.byref(nn), C.byref(nf)))
        return nn.value, nf.value

    def factor_parents(self, f):
        m = C.c_uint32()
        self._chk(self.L.mcu_factor_parents(self.h, int(f), C.byref(m)))
        return m.value

    def logpdf_nodes(self, mask, state):
        """logpdf(mc, nodekeys): src/output/modelstats.jl:16-58 — sum of the selected node densities at each state record."""
        state = _f64(np.atleast_2d(state))
        lp = np.empty(state.shape[0])
        self._chk(self.L.mcu_logpdf_nodes(self.h, int(mask), state.shape[0], _dp(state), _dp(lp)))
        return lp

    def predict(self, state, stream_id=0):
        """predict(mc): src/output/modelstats.jl:63-96 — one draw of the observed node at each state record → [B × len(y)]."""
        state = _f64(np.atleast_2d(state))
        n = C.c_int64()
        self._chk(self.L.mcu_predict(self.h, 0, None, 0, None, C.byref(n)))
        out = np.empty((state.shape[0], n.value))
        self._chk(self.L.mcu_predict(self.h, state.shape[0], _dp(state), int(stream_id), _dp(out), C.byref(n)))
        return out

    def gradlogpdf(self, block, state, k, x=None, mode="analytic"):
        state = _f64(np.atleast_2d(state))
        x = _f64(None if x is None else np.atleast_2d(x))
        lp = np.empty(state.shape[0])
        g = np.empty((state.shape[0], k))
        m = _lib.GRAD[mode] if isinstance(mode, str) else int(mode)
        self._chk(self.L.mcu_gradlogpdf(self.h, block, m, state.shape[0], _dp(state), _dp(x), _dp(lp), _dp(g)))
        return lp, g

    def glm_gradient(self, beta, impl=1):
        """Likelihood logf and gradient of the GLM template for all chains in one pass over X."""
        beta = _f64(np.atleast_2d(beta))
        assert beta.shape[0] == self.n_chains
        lp = np.empty(self.n_chains); g = np.empty_like(beta)
        self._chk(self.L.mcu_glm_gradient(self.h, int(impl), _dp(beta), _dp(lp), _dp(g)))
        return lp, g

    # ---- diagnostics -------------------------------------------------------------------------
    def minmax(self):
        p = self.dims()[1]
        mm = np.empty((p, 2))
        self._chk(self.L.mcu_minmax(self.h, _dp(mm)))
        return mm

    def link_codes(self, transform, minmax=None):
        p = self.dims()[1]
        codes = (C.c_int * p)()
        mm = _f64(minmax)
        self._chk(self.L.mcu_link_codes(self.h, int(bool(transform)), _dp(mm), codes))
        return np.array(list(codes), dtype=np.int32)

    def node_links(self):
        """Static part of link(c::ModelChains) (modelchains.jl:57-76) per monitored column: 0 identity, 1 log (the node's own
        link), -1 = not a stochastic node: the data-dependent heuristic of chains.jl:237-246 applies."""
        p = self.dims()[1]
        codes = self.link_codes(True, minmax=np.tile([0.5, 0.6], (p, 1)))
        return np.where(codes == 2, -1, codes).astype(np.int32)

    def moments(self, codes=None, center=None):
        p = self.dims()[1]
        sums = np.empty((p, 7))
        n = C.c_int64()
        cc = None if codes is None else (C.c_int * p)(*[int(c) for c in codes])
        ctr = _f64(center)
        self._chk(self.L.mcu_moments(self.h, cc, _dp(ctr), _dp(sums), C.byref(n)))
        return sums, n.value

    def gelman_from_moments(self, n_kept, center, sums, alpha=0.05):
        p = sums.shape[0]
        psrf = np.empty((p, 2))
        center = _f64(center); sums = _f64(sums)
        rc = self.L.mcu_gelman_from_moments(int(n_kept), p, _dp(center), _dp(sums), float(alpha), _dp(psrf))
        if rc != 0:
            raise ValueError("less than 2 chains supplied to gelman diagnostic")
        return psrf

    def gelman(self, alpha=0.05, transform=False):
        p = self.dims()[1]
        psrf = np.empty((p, 2))
        self._chk(self.L.mcu_gelman(self.h, float(alpha), int(bool(transform)), _dp(psrf)))
        return psrf

    def summarystats(self, etype="bm", batch=100):
        p = self.dims()[1]
        out = np.empty((p, 5))
        self._chk(self.L.mcu_summarystats(self.h, {"bm": 0, "imse": 1}[etype], int(batch), _dp(out)))
        return out

    def summary_sums(self, center=None):
        p = self.dims()[1]
        sums = np.empty((p, 8))
        ctr = _f64(center)
        self._chk(self.L.mcu_summary_sums(self.h, _dp(ctr), _dp(sums)))
        return sums

    def summary_from_sums(self, n_kept, center, sums):
        p = sums.shape[0]
        out = np.empty((p, 5))
        center = _f64(center); sums = _f64(sums)
        self.L.mcu_summary_from_sums(int(n_kept), p, _dp(center), _dp(sums), _dp(out))
        return out

    def summary_streaming(self):
        p = self.dims()[1]
        out = np.empty((p, 5))
        self._chk(self.L.mcu_summary_streaming(self.h, _dp(out)))
        return out

    # ---- diagnostics over the chains of several handles / GPUs (include/mambacuda.h, csrc/diagproto.hpp) -------------------
    def monitor_links(self):
        p = self.dims()[1]
        ml = (C.c_int * p)()
        self._chk(self.L.mcu_monitor_links(self.h, ml))
        return np.array(list(ml), dtype=np.int32)

    def n_kept(self):
        n = C.c_int64()
        self._chk(self.L.mcu_n_kept(self.h, C.byref(n)))
        return n.value

    def diag_round1(self):
        p = self.dims()[1]
        buf = np.empty(11 * p)
        self._chk(self.L.mcu_diag_round1(self.h, _dp(buf)))
        return buf

    def diag_round2(self, transform, reduced1):
        p = self.dims()[1]
        n1, n2 = C.c_int(), C.c_int()
        self.L.mcu_diag_sizes(p, C.byref(n1), C.byref(n2))
        r1 = _f64(reduced1); buf = np.empty(n2.value)
        self._chk(self.L.mcu_diag_round2(self.h, int(bool(transform)), _dp(r1), _dp(buf)))
        return buf

    def diag_finish(self, alpha, transform, reduced1, reduced2, mpsrf=False):
        return diag_finish(self.n_kept(), self.monitor_links(), alpha, transform, reduced1, reduced2, mpsrf=mpsrf)

    def comm_init(self, rank, nranks, unique_id):
        """Join the NCCL communicator of `unique_id` (bytes from comm_unique_id() on rank 0) as `rank` of `nranks`."""
        self._chk(self.L.mcu_comm_init(self.h, int(rank), int(nranks), bytes(unique_id)))

    def comm_size(self):
        r, n = C.c_int(), C.c_int()
        self._chk(self.L.mcu_comm_size(self.h, C.byref(r), C.byref(n)))
        return r.value, n.value

    def diag_global(self, alpha=0.05, transform=False, mpsrf=False):
        """gelmandiag + streaming summarystats over every rank of the handle's communicator (this handle alone without one):
        (psrf [p x 2], summary [p x 5], link codes [p]) and, with mpsrf=True, the multivariate PSRF from the streamed within-chain
        covariances as a fourth element; device-resident, NCCL all-reduces, one synchronisation."""
        p = self.dims()[1]
        psrf = np.empty((p, 2)); summ = np.empty((p, 5)); codes = (C.c_int * p)(); mv = C.c_double(float("nan"))
        self._chk(self.L.mcu_diag_global(self.h, float(alpha), int(bool(transform)), _dp(psrf), _dp(summ), codes, C.byref(mv) if mpsrf else None))
        out = (psrf, summ, np.array(list(codes), dtype=np.int32))
        return out + (mv.value,) if mpsrf else out

    def fp64_peak_tflops(self):
        return self.L.mcu_fp64_peak_tflops(self.h)

    def work_count(self):
        """(gradient evaluations of the fused gradient-based paths, GLM ticks) since the handle was created; also sets
        self.glm_pass_slots and self.nuts_cap_hits (mcu_work_count)."""
        out = (C.c_uint64 * 4)()
        self._chk(self.L.mcu_work_count(self.h, out))
        self.glm_pass_slots = int(out[2]); self.nuts_cap_hits = int(out[3])
        return int(out[0]), int(out[1])

    def launch_count(self):
        return self.L.mcu_launch_count(self.h)

    def last_kernel_ms(self):
        return self.L.mcu_last_kernel_ms(self.h)


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0); hand the 128 bytes to the other ranks with the host language's own messaging."""
    buf = C.create_string_buffer(128)
    L = _lib.lib()
    rc = L.mcu_comm_unique_id(buf)
    if rc != 0:
        raise MambaCudaError(rc, L.mcu_last_error(None).decode())
    return buf.raw


def diag_finish(n_kept, monlink, alpha, transform, reduced1, reduced2, mpsrf=False):
    """Host arithmetic of the two-round protocol (no device): (psrf, summary, codes[, mpsrf]) from the all-reduced buffers."""
    L = _lib.lib()
    p = len(monlink)
    ml = (C.c_int * p)(*[int(v) for v in monlink])
    r1 = _f64(reduced1); r2 = _f64(reduced2)
    psrf = np.empty((p, 2)); summ = np.empty((p, 5)); codes = (C.c_int * p)(); mv = C.c_double(float("nan"))
    rc = L.mcu_diag_finish(int(n_kept), p, float(alpha), ml, int(bool(transform)), _dp(r1), _dp(r2), _dp(psrf), _dp(summ), codes, C.byref(mv) if mpsrf else None)
    if rc == _lib.ERR_ARG:
        raise ValueError("less than 2 chains supplied to gelman diagnostic")
    if rc != 0:
        raise MambaCudaError(rc, "mcu_diag_finish failed")
    out = (psrf, summ, np.array(list(codes), dtype=np.int32))
    return out + (mv.value,) if mpsrf else out
