"""Host-side mirror of the reference's interface for the sampler hot path.

The reference is Julia; its toolchain is absent here, so the host side above the C ABI is written in
Python with the same names, argument meaning and error behaviour (mamba.jl_b200/julia/MambaCUDA.jl is
the Julia shim a maintainer would load instead).  What is mirrored (paths in the reference tree):

    Model(; nodes...)              src/model/model.jl:5-27        → Model.template("seeds") etc.: the node
                                                                     closures of the four example scripts and
                                                                     the GLM are compiled device templates
    AMWG / Slice / RWM / NUTS / HMC / AMM   src/samplers/*.jl sampler constructors → Sampler records
    setsamplers!(m, scheme)        src/model/initialization.jl:42-48
    mcmc(m, data, inits, iters; burnin, thin, chains)   src/model/mcmc.jl:19-33   → ModelChains
    mcmc(mc, iters)                src/model/mcmc.jl:3-16 (restart)
    gelmandiag(c; alpha, transform)  src/output/gelmandiag.jl:3-60
    summarystats(c; etype)         src/output/stats.jl:85-94
    Chains/ModelChains             src/Mamba.jl:172-185, src/output/chains.jl

User-defined closure samplers (Sampler(params, f), src/samplers/sampler.jl:20-24) have no device
equivalent and the engine has no CPU fallback: they raise ArgumentError (ValueError here).
"""
import numpy as np

from . import _lib
from .engine import Engine

Univariate, Multivariate = "Univariate", "Multivariate"


class ArgumentError(ValueError):
    """Julia's ArgumentError."""


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch."""


# node tables of the device templates: name -> (node id, length, monitored)
_TEMPLATES = {
    "line": dict(nodes=[("beta", 2), ("s2", 1)], inputs=["x", "y"]),
    "seeds": dict(nodes=[("alpha0", 1), ("alpha1", 1), ("alpha2", 1), ("alpha12", 1), ("s2", 1), ("b", 21)], inputs=["r", "n", "x1", "x2"]),
    "rats": dict(nodes=[("mu_alpha", 1), ("mu_beta", 1), ("s2_alpha", 1), ("s2_beta", 1), ("s2_c", 1), ("alpha", 30), ("beta", 30)],
                 inputs=["y", "rat", "Xm", "xbar"]),
    "pumps": dict(nodes=[("alpha", 1), ("beta", 1), ("theta", 10)], inputs=["y", "t"]),
    "surgical": dict(nodes=[("mu", 1), ("s2", 1), ("b", 12)], inputs=["r", "n"]),
    "dyes": dict(nodes=[("s2_between", 1), ("theta", 1), ("s2_within", 1), ("mu", 6)], inputs=["y", "batch"]),
    "glm": dict(nodes=[("beta", None)], inputs=["X", "y"]),
}


class Sampler:
    """Sampler{T}(params, eval, tune, targets): src/Mamba.jl:119-124.  `eval` lives on the device."""

    def __init__(self, params, kind, **desc):
        if isinstance(params, str):
            params = [params]
        self.params = list(params)
        self.kind = kind
        self.desc = desc
        self.tune = None   # filled after a run: per-chain tune record slice

    def __repr__(self):
        return f"Sampler({self.params}, {self.kind})"


def _check_adapt(adapt):
    if adapt not in ("all", "burnin", "none"):
        raise ArgumentError("adapt must be one of :all, :burnin, or :none")   # amwg.jl:49-50, amm.jl:47-48


def AMWG(params, sigma, adapt="all", batchsize=50, target=0.44):   # src/samplers/amwg.jl:47-61
    _check_adapt(adapt)
    return Sampler(params, "amwg", scale=sigma, adapt=adapt, batchsize=batchsize, target=target)


def Slice(params, width, form=Multivariate, transform=False):      # src/samplers/slice.jl:47-58
    if form not in (Univariate, Multivariate):
        raise ArgumentError("form must be Univariate or Multivariate")
    return Sampler(params, "slice_uni" if form == Univariate else "slice_multi", scale=width, transform=int(bool(transform)))


def RWM(params, scale, proposal="normal"):                         # src/samplers/rwm.jl:49-58
    if proposal not in _lib.PROPOSAL:
        raise ArgumentError(f"proposal {proposal} has no device implementation (normal, symuniform, symtriangular)")
    return Sampler(params, "rwm", scale=scale, proposal=proposal)


def NUTS(params, dtype="analytic", target=0.6, epsilon=0.0, max_depth=10):   # src/samplers/nuts.jl:47-56
    if dtype not in _lib.GRAD:
        raise ArgumentError("dtype must be :forward, :central or :analytic")
    return Sampler(params, "nuts", grad=dtype, target=target, epsilon=epsilon, max_depth=max_depth)


def HMC(params, epsilon, L, Sigma=None, dtype="analytic"):         # src/samplers/hmc.jl:47-65
    if dtype not in _lib.GRAD:
        raise ArgumentError("dtype must be :forward, :central or :analytic")
    d = dict(epsilon=epsilon, L=L, grad=dtype)
    if Sigma is not None:
        d["scale"] = np.asarray(Sigma, dtype=float)
    return Sampler(params, "hmc", **d)


def MALA(params, epsilon, Sigma=None, dtype="analytic"):            # src/samplers/mala.jl:43-65
    if dtype not in _lib.GRAD:
        raise ArgumentError("dtype must be :forward, :central or :analytic")
    d = dict(epsilon=epsilon, grad=dtype)
    if Sigma is not None:
        d["scale"] = np.asarray(Sigma, dtype=float)
    return Sampler(params, "mala", **d)


def AMM(params, Sigma, adapt="all", beta=0.05, scale=2.38):        # src/samplers/amm.jl:45-59
    _check_adapt(adapt)
    return Sampler(params, "amm", scale=np.asarray(Sigma, dtype=float), adapt=adapt, beta=beta, amm_scale=scale)


def Gibbs(params):
    """The device counterpart of a user-defined Gibbs sampler `Sampler(params, (args...) -> rand(full conditional))`
    (src/samplers/sampler.jl:20-24; tutorial's Gibbs_beta / Gibbs_s2): an exact draw from the block's full conditional, for the
    node sets the template registers a conjugate form for (pumps: [:theta], [:beta]); anything else raises at setsamplers time."""
    return Sampler(params, "gibbs")


class ModelState:
    """src/Mamba.jl:152-155"""

    def __init__(self, value, tune):
        self.value = value
        self.tune = tune


class Model:
    """A model whose node closures are one of the compiled device templates."""

    def __init__(self, template, iter=0, burnin=0, samplers=()):
        if template not in _TEMPLATES:
            raise ArgumentError(f"no device template named {template}; available: {sorted(_TEMPLATES)}")
        self.template = template
        self.iter = iter
        self.burnin = burnin
        self.samplers = []
        self.states = []
        self.inputs = {}
        self.hasinputs = False
        self.hasinits = False
        self.glm_d = None
        if samplers:
            setsamplers(self, samplers)

    @staticmethod
    def from_template(name):
        return Model(name)

    def node_ids(self):
        return {nm: i for i, (nm, _) in enumerate(_TEMPLATES[self.template]["nodes"])}

    def node_len(self, name):
        for nm, ln in _TEMPLATES[self.template]["nodes"]:
            if nm == name:
                return self.glm_d if ln is None else ln
        raise KeyError(name)

    def keys(self, ntype="all"):   # src/model/model.jl:58-72 (subset)
        t = _TEMPLATES[self.template]
        if ntype in ("stochastic", "dependent", "block"):
            return [n for n, _ in t["nodes"]]
        if ntype in ("input", "independent"):
            return list(t["inputs"])
        return [n for n, _ in t["nodes"]] + list(t["inputs"])

    def state_dim(self):
        return sum(self.node_len(n) for n, _ in _TEMPLATES[self.template]["nodes"])


def setsamplers(model, scheme):
    """setsamplers!(m, scheme): src/model/initialization.jl:42-48"""
    ids = model.node_ids()
    out = []
    for s in scheme:
        if not isinstance(s, Sampler):
            raise ArgumentError("user-defined closure samplers have no device equivalent (no CPU fallback)")
        for p in s.params:
            if p not in ids:
                raise KeyError(f"unknown node {p}")
        out.append(s)
    model.samplers = out
    return model


def setinputs(model, inputs):
    """setinputs!(m, inputs): src/model/initialization.jl:30-40"""
    need = [k for k in _TEMPLATES[model.template]["inputs"] if model.template == "glm"]
    for key in need:
        if key not in inputs:
            raise ArgumentError(f"missing inputs for node : {key}")
    model.inputs = {k: np.asarray(v, dtype=float) for k, v in inputs.items() if k in _TEMPLATES[model.template]["inputs"]}
    if model.template == "glm":
        model.glm_d = int(model.inputs["X"].shape[1])
    model.hasinputs = True
    return model


def _inits_matrix(model, inits):
    """Vector{Dict} → [n × D] records in the template's state order (setinits!: initialization.jl:3-28)."""
    rows = []
    for d in inits:
        rec = []
        for nm, _ in _TEMPLATES[model.template]["nodes"]:
            if nm not in d:
                raise ArgumentError(f"missing initial value for node : {nm}")   # initialization.jl:9-10
            v = np.atleast_1d(np.asarray(d[nm], dtype=float)).ravel()
            if v.size != model.node_len(nm):
                raise DimensionMismatch(f"incompatible initial value for node : {nm}")
            rec.extend(v.tolist())
        rows.append(rec)
    return np.array(rows, dtype=float)


def _block_descs(model):
    ids = model.node_ids()
    blocks = []
    for s in model.samplers:
        d = dict(kind=s.kind, nodes=[ids[p] for p in s.params])
        d.update(s.desc)
        blocks.append(d)
    return blocks


class Chains:
    """src/Mamba.jl:172-177: value [iters × params × chains], range, names, chains."""

    def __init__(self, value, start=1, thin=1, names=None, chains=None):
        value = np.asarray(value, dtype=float)
        if value.ndim != 3:
            raise DimensionMismatch("value must be iterations x parameters x chains")
        n, p, m = value.shape
        self.value = value
        self.range = range(start, start + thin * n, thin)
        self.names = list(names) if names is not None else [f"Param{i + 1}" for i in range(p)]
        if len(self.names) != p:
            raise DimensionMismatch("size(value, 2) and names length differ")     # chains.jl:20-21
        self.chains = list(chains) if chains is not None else list(range(1, m + 1))
        if len(self.chains) != m:
            raise DimensionMismatch("size(value, 3) and chains length differ")    # chains.jl:26-27

    @property
    def first(self):
        return self.range.start

    @property
    def step(self):
        return self.range.step

    @property
    def last(self):
        return self.range[-1] if len(self.range) else self.range.start - self.range.step


class ModelChains(Chains):
    """src/Mamba.jl:179-185"""

    def __init__(self, value, model, engine=None, **kw):
        super().__init__(value, **kw)
        self.model = model
        self.engine = engine


def mcmc(model, *args, burnin=0, thin=1, chains=1, verbose=False, seed=123, device=0, store=True):
    """mcmc(m, inputs, inits, iters; burnin, thin, chains) and the restart form mcmc(mc, iters)
    (src/model/mcmc.jl:3-33).  The whole chains x iterations loop is one call into libmambacuda."""
    if isinstance(model, ModelChains):
        return _restart(model, *args)
    inputs, inits, iters = args
    if not iters > burnin:
        raise ArgumentError("burnin is greater than or equal to iters")    # mcmc.jl:22-23
    if not len(inits) >= chains:
        raise ArgumentError("fewer initial values than chains")            # mcmc.jl:24-25
    if not model.samplers:
        raise ArgumentError("no samplers set: call setsamplers first")
    import copy
    mm = copy.deepcopy(model)                                               # mcmc.jl:27
    setinputs(mm, inputs)
    x = _inits_matrix(mm, inits[:chains])
    eng = Engine(mm.template, chains, seed=seed, device=device)
    for k, v in mm.inputs.items():
        eng.set_data(k, v)
    eng.set_scheme(_block_descs(mm))
    eng.set_inits(x)
    mm.burnin = burnin
    value = eng.run(iters, burnin=burnin, thin=thin, store=store)
    return _wrap(mm, eng, value, burnin + thin, thin, chains)


def _wrap(mm, eng, value, start, thin, chains):
    vals, tune, it = eng.get_state()
    mm.iter = it
    mm.states = [ModelState(vals[k].copy(), tune[k].copy()) for k in range(chains)]    # mcmc.jl:56,82
    mm.hasinits = True
    return ModelChains(value, mm, engine=eng, start=start, thin=thin, names=eng.names(1), chains=list(range(1, chains + 1)))


def _restart(mc, iters):
    thin = mc.step
    if mc.last != (mc.model.iter // thin) * thin:
        raise ArgumentError("chain is missing its last iteration")          # mcmc.jl:5-6
    eng = mc.engine
    value = eng.run(iters, burnin=mc.model.burnin, thin=thin)
    mc2 = _wrap(mc.model, eng, value, mc.last + thin, thin, len(mc.chains))
    return ModelChains(np.concatenate([mc.value, mc2.value], axis=0), mc2.model, engine=eng, start=mc.first, thin=thin,
                       names=mc.names, chains=mc.chains)


def gelmandiag(c, alpha=0.05, mpsrf=False, transform=False):
    """gelmandiag(c; alpha, mpsrf, transform): src/output/gelmandiag.jl:3-60 (PSRF and 97.5% columns, rounded to 3 dp)."""
    if len(c.chains) < 2:
        raise ArgumentError("less than 2 chains supplied to gelman diagnostic")   # gelmandiag.jl:6-7
    if mpsrf:   # needs the p x p within / between covariances: computed on the materialised array (mcu_chains_gelman)
        codes = c.engine.link_codes(transform) if (transform and getattr(c, "engine", None) is not None) else None
        psrf = _chains_gelman(c.value, alpha, codes, True)
        return np.round(psrf, 3), c.names + ["Multivariate"], ["PSRF", f"{100 * (1 - alpha / 2)}%"]
    psrf = c.engine.gelman(alpha, transform)
    return np.round(psrf, 3), c.names, ["PSRF", f"{100 * (1 - alpha / 2)}%"]


def _value_f(c):
    v = np.asfortranarray(c.value if isinstance(c, Chains) else c, dtype=np.float64)
    if v.ndim != 3:
        raise DimensionMismatch("value must be iterations x parameters x chains")
    return v


def _dp(a):
    import ctypes as C
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _chains_gelman(value, alpha, codes, mpsrf):
    import ctypes as C
    v = _value_f(value); n, p, m = v.shape
    out = np.empty((p + int(mpsrf), 2))
    cc = None if codes is None else (C.c_int * p)(*[int(x) for x in codes])
    if _lib.lib().mcu_chains_gelman(_dp(v), n, p, m, float(alpha), cc, int(mpsrf), _dp(out)) != 0:
        raise ArgumentError("less than 2 chains supplied to gelman diagnostic")
    return out


def quantile(c, q=(0.025, 0.25, 0.5, 0.75, 0.975)):
    """quantile(c; q): src/output/stats.jl:74-83 → [p × length(q)]"""
    v = _value_f(c); n, p, m = v.shape
    qq = np.ascontiguousarray(q, dtype=np.float64); out = np.empty((p, qq.size))
    _lib.lib().mcu_chains_quantile(_dp(v), n, p, m, _dp(qq), qq.size, _dp(out))
    return out, c.names, [f"{100 * x}%" for x in qq]


def hpd(c, alpha=0.05):
    """hpd(c; alpha): src/output/stats.jl:52-72 → [p × 2]"""
    v = _value_f(c); n, p, m = v.shape
    out = np.empty((p, 2))
    if _lib.lib().mcu_chains_hpd(_dp(v), n, p, m, float(alpha), _dp(out)) != 0:
        raise ArgumentError("alpha must be in (0, 1)")
    return out, c.names, [f"{100 * (1 - alpha)}% Lower", f"{100 * (1 - alpha)}% Upper"]


def autocor(c, lags=(1, 5, 10, 50), relative=True):
    """autocor(c; lags, relative): src/output/stats.jl:3-13 → [p × length(lags) × chains]"""
    import ctypes as C
    lags = np.asarray(lags, dtype=np.int64)
    if relative:
        lags = lags * c.step
    elif np.any(lags % c.step != 0):
        raise ArgumentError("lags do not correspond to thinning interval")
    v = _value_f(c); n, p, m = v.shape
    out = np.empty((p, lags.size, m), order="F")
    lg = np.ascontiguousarray(lags)
    _lib.lib().mcu_chains_autocor(_dp(v), n, p, m, lg.ctypes.data_as(C.POINTER(C.c_int64)), lags.size, _dp(out))
    return out, c.names, [f"Lag {x}" for x in lags]


def changerate(c):
    """changerate(c): src/output/stats.jl:19-39 → [p + 1] rounded to 3 dp, last row "Multivariate"."""
    v = _value_f(c); n, p, m = v.shape
    out = np.empty(p + 1)
    _lib.lib().mcu_chains_changerate(_dp(v), n, p, m, _dp(out))
    return np.round(out, 3), c.names + ["Multivariate"], ["Change Rate"]


_ETYPE = {"bm": 0, "imse": 1, "ipse": 2}


def gewekediag(c, first=0.1, last=0.5, etype="imse", size=100):
    """gewekediag(c; first, last, etype): src/output/gewekediag.jl:3-31 → [p × 2 × chains] (Z-score rounded to 3 dp, p-value to 4)."""
    if not 0.0 < first < 1.0:
        raise ArgumentError("first is not in (0, 1)")
    if not 0.0 < last < 1.0:
        raise ArgumentError("last is not in (0, 1)")
    if first + last > 1.0:
        raise ArgumentError("first and last proportions overlap")
    if etype not in _ETYPE:
        raise ArgumentError(f"unsupported mcse method {etype}")
    v = _value_f(c); n, p, m = v.shape
    out = np.empty((p, 2, m), order="F")
    if _lib.lib().mcu_chains_geweke(_dp(v), n, p, m, float(first), float(last), _ETYPE[etype], int(size), _dp(out)) != 0:
        raise ArgumentError(f"iterations are < {2 * size} and batch size is > {n // 2}")
    out[:, 0, :] = np.round(out[:, 0, :], 3); out[:, 1, :] = np.round(out[:, 1, :], 4)
    return out, c.names, ["Z-score", "p-value"]


def heideldiag(c, alpha=0.05, eps=0.1, etype="imse", size=100):
    """heideldiag(c; alpha, eps, etype): src/output/heideldiag.jl:3-41 → [p × 6 × chains]."""
    if etype not in _ETYPE:
        raise ArgumentError(f"unsupported mcse method {etype}")
    v = _value_f(c); n, p, m = v.shape
    out = np.empty((p, 6, m), order="F")
    if _lib.lib().mcu_chains_heidel(_dp(v), n, p, m, float(alpha), float(eps), _ETYPE[etype], int(size), int(c.first), _dp(out)) != 0:
        raise ArgumentError(f"iterations are < {2 * size} and batch size is > {n // 2}")
    out[:, 2, :] = np.round(out[:, 2, :], 4)
    return out, c.names, ["Burn-in", "Stationarity", "p-value", "Mean", "Halfwidth", "Test"]


def rafterydiag(c, q=0.025, r=0.005, s=0.95, eps=0.001):
    """rafterydiag(c; q, r, s, eps): src/output/rafterydiag.jl:3-61 → [p × 5 × chains]."""
    v = _value_f(c); n, p, m = v.shape
    out = np.empty((p, 5, m), order="F")
    if _lib.lib().mcu_chains_raftery(_dp(v), n, p, m, float(q), float(r), float(s), float(eps), int(c.first), int(c.step), _dp(out)) != 0:
        raise ArgumentError("q and s must be in (0, 1), r positive")
    return out, c.names, ["Thinning", "Burn-in", "Total", "Nmin", "Dependence Factor"]


def describe(c, q=(0.025, 0.25, 0.5, 0.75, 0.975), etype="bm"):
    """describe(c): src/output/stats.jl:41-52 — summarystats + quantiles."""
    return summarystats(c, etype=etype), quantile(c, q=q)


def summarystats(c, etype="bm", batch=100):
    """summarystats(c; etype): src/output/stats.jl:85-94 → [p × 5] Mean, SD, Naive SE, MCSE, ESS."""
    if etype not in ("bm", "imse"):
        raise ArgumentError(f"unsupported mcse method {etype}")                   # mcse.jl:3-8
    return c.engine.summarystats(etype, batch), c.names, ["Mean", "SD", "Naive SE", "MCSE", "ESS"]
