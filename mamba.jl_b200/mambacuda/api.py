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
    "line": dict(nodes=[("beta", 2), ("s2", 1)], inputs=["x", "y"], outputs=["y"]),
    "seeds": dict(nodes=[("alpha0", 1), ("alpha1", 1), ("alpha2", 1), ("alpha12", 1), ("s2", 1), ("b", 21)], inputs=["r", "n", "x1", "x2"], outputs=["r"]),
    "rats": dict(nodes=[("mu_alpha", 1), ("mu_beta", 1), ("s2_alpha", 1), ("s2_beta", 1), ("s2_c", 1), ("alpha", 30), ("beta", 30)],
                 inputs=["y", "rat", "Xm", "xbar"], outputs=["y"]),
    "pumps": dict(nodes=[("alpha", 1), ("beta", 1), ("theta", 10)], inputs=["y", "t"], outputs=["y"]),
    "surgical": dict(nodes=[("mu", 1), ("s2", 1), ("b", 12)], inputs=["r", "n"], outputs=["r"]),
    "dyes": dict(nodes=[("s2_between", 1), ("theta", 1), ("s2_within", 1), ("mu", 6)], inputs=["y", "batch"], outputs=["y"]),
    "salm": dict(nodes=[("s2", 1), ("gamma", 1), ("beta", 1), ("alpha", 1), ("lambda", 18)], inputs=["y", "x"], outputs=["y"]),
    "equiv": dict(nodes=[("s2_2", 1), ("s2_1", 1), ("pi", 1), ("phi", 1), ("mu", 1), ("delta", 20)], inputs=["y", "group"], outputs=["y"]),
    "blocker": dict(nodes=[("s2", 1), ("d", 1), ("delta_new", 1), ("mu", 22), ("delta", 22)], inputs=["rc", "nc", "rt", "nt"], outputs=["rc", "rt"],
                    output_lens=[22, 22]),
    "stacks": dict(nodes=[("beta0", 1), ("beta", 3), ("s2", 1)], inputs=["y", "x"], outputs=["y"]),
    "magnesium": dict(nodes=[("priors", 6), ("mu", 6), ("theta", 48), ("pc", 48)], inputs=["rc", "nc", "rt", "nt"], outputs=["rcx", "rtx"],
                      output_lens=[48, 48]),
    "oxford": dict(nodes=[("alpha", 1), ("beta1", 1), ("beta2", 1), ("s2", 1), ("b", 120), ("mu", 120)], inputs=["r1", "n1", "r0", "n0", "year"], outputs=["r0", "r1"],
                   output_lens=[120, 120]),
    "epil": dict(nodes=[("a0", 1), ("alpha_Base", 1), ("alpha_Trt", 1), ("alpha_BT", 1), ("alpha_Age", 1), ("alpha_V4", 1), ("s2_b1", 1), ("s2_b", 1), ("b1", 59), ("b", 236)],
                 inputs=["y", "Trt", "Base", "Age", "V4"], outputs=["y"]),
    "glm": dict(nodes=[("beta", None)], inputs=["X", "y"], outputs=["y"]),
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
        raise ArgumentError(f"proposal {proposal} is not a SymDistributionType ({', '.join(_lib.PROPOSAL)})")   # extensions.jl:51-53
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
    node sets the template registers a conjugate form for (pumps: [:theta], [:beta]; line: [:beta], [:s2] — the tutorial's Gibbs_beta / Gibbs_s2);
    anything else raises at setsamplers time."""
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
    need = ["X", "y"] if model.template == "glm" else []       # the example templates carry the data of their scripts as defaults
    for key in need:
        if key not in inputs:
            raise ArgumentError(f"missing inputs for node : {key}")
    allowed = list(_TEMPLATES[model.template]["inputs"]) + (["family", "sigma"] if model.template == "glm" else [])
    model.inputs = {k: np.atleast_1d(np.asarray(v, dtype=float)) for k, v in inputs.items() if k in allowed}
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
            v = np.atleast_1d(np.asarray(d[nm], dtype=float)).ravel(order="F")   # matrices flatten column-major, as unlist does
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
        if value.ndim == 2:                    # Chains(value::Matrix; ...): one chain (chains.jl:34-41)
            value = value[:, :, None]
            if isinstance(chains, (int, np.integer)):
                chains = [int(chains)]
        elif value.ndim == 1:                  # Chains(value::Vector; names = "Param1"): one parameter, one chain (chains.jl:43-49)
            value = value[:, None, None]
            if isinstance(names, str):
                names = [names]
            if isinstance(chains, (int, np.integer)):
                chains = [int(chains)]
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

    # ---- indexing: c[window, names, chains] (src/output/chains.jl:51-99) ------------------------------------
    def _window2inds(self, window):
        """window2inds (chains.jl:73-80): an ITERATION range first:step:last → 0-based row slice of value."""
        n = self.value.shape[0]
        if window is None or window == slice(None):
            return 0, 1, n
        if isinstance(window, slice):    # iteration numbers, stop inclusive like a Julia range
            window = range(window.start if window.start is not None else self.first,
                           (window.stop if window.stop is not None else self.last) + 1, window.step or 1)
        if not isinstance(window, range):
            raise ArgumentError(f"{type(window).__name__} iteration indexing is unsupported")          # chains.jl:70-71
        if len(window) == 0:
            return 0, window.step, 0
        lo = (window[0] - self.first) / self.step + 1.0                                                 # @mapiters, chains.jl:64-68
        hi = (window[-1] - self.first) / self.step + 1.0
        a = max(int(np.ceil(lo)), 1)
        b = min(int(np.floor(hi)), n)
        return a - 1, window.step, b

    def _names2inds(self, names):
        """names2inds (chains.jl:91-98): integers (1-based like the reference), strings, booleans, or None for all."""
        p = self.value.shape[1]
        if names is None or (isinstance(names, slice) and names == slice(None)):
            return list(range(p))
        if isinstance(names, (str, int, np.integer)):
            names = [names]
        names = list(names)
        if names and all(isinstance(x, (bool, np.bool_)) for x in names):
            return [j for j, keep in enumerate(names) if keep]
        out = []
        for x in names:
            if isinstance(x, str):
                if x not in self.names:
                    raise KeyError(x)
                out.append(self.names.index(x))
            else:
                out.append(int(x) - 1)
        return out

    def _chains2inds(self, chains):
        m = self.value.shape[2]
        if chains is None or (isinstance(chains, slice) and chains == slice(None)):
            return list(range(m))
        if isinstance(chains, (int, np.integer)):
            chains = [chains]
        chains = list(chains)
        if chains and all(isinstance(x, (bool, np.bool_)) for x in chains):
            return [k for k, keep in enumerate(chains) if keep]
        return [int(k) - 1 for k in chains]

    def _subset(self, key):
        if not isinstance(key, tuple) or len(key) != 3:
            raise ArgumentError("chains are indexed as c[window, names, chains]")
        a, stride, b = self._window2inds(key[0])
        j = self._names2inds(key[1]); k = self._chains2inds(key[2])
        value = self.value[a:b:stride][:, j][:, :, k]
        return value, dict(start=self.first + a * self.step, thin=stride * self.step, names=[self.names[x] for x in j],
                           chains=[self.chains[x] for x in k])

    def __getitem__(self, key):                                         # getindex(c::Chains, window, names, chains): chains.jl:51-58
        value, kw = self._subset(key)
        return Chains(value, **kw)

    def __setitem__(self, key, value):                                  # setindex!: chains.jl:60-62 (iterations by number)
        iters, names, chains = key
        a, stride, b = self._window2inds(iters if not isinstance(iters, (int, np.integer)) else range(iters, iters + 1))
        j = self._names2inds(names); k = self._chains2inds(chains)
        self.value[np.ix_(range(a, b, stride), j, k)] = value

    def keys(self):                                                     # chains.jl:171-173
        return self.names

    def size(self, ind=None):                                           # chains.jl:181-188: (last iteration, parameters, chains)
        dims = (self.last, self.value.shape[1], self.value.shape[2])
        return dims if ind is None else dims[ind - 1]

    def header(self):                                                   # chains.jl:211-218
        return (f"Iterations = {self.first}:{self.last}\nThinning interval = {self.step}\n"
                f"Chains = {','.join(str(k) for k in self.chains)}\nSamples per chain = {len(self.range)}\n")

    def combine(self):                                                  # chains.jl:197-209: rows ordered iteration-major, chain fastest
        n, p, m = self.value.shape
        return np.ascontiguousarray(self.value.transpose(0, 2, 1).reshape(n * m, p))

    def link_codes(self):
        """link(c::AbstractChains) (chains.jl:237-246) as codes: log if every value of a column is > 0, logit if also < 1."""
        mn = self.value.min(axis=(0, 2)); mx = self.value.max(axis=(0, 2))
        return np.where(mn > 0.0, np.where(mx < 1.0, 2, 1), 0).astype(np.int32)

    def link(self):
        cc = self.value.copy()
        for j, code in enumerate(self.link_codes()):
            if code == 2:
                cc[:, j, :] = np.log(cc[:, j, :] / (1.0 - cc[:, j, :]))
            elif code == 1:
                cc[:, j, :] = np.log(cc[:, j, :])
        return cc

    def indiscretesupport(self, bounds=(0, np.inf)):                    # chains.jl:220-235
        v = self.value
        ok = (v == np.round(v)) & (v >= bounds[0]) & (v <= bounds[1])
        return ok.all(axis=(0, 2))


def cat(dim, c1, *args):
    """cat(dim, c1, args...) for chains (src/output/chains.jl:102-163): 1 = iterations, 2 = parameters, 3 = chains."""
    cs = (c1,) + args
    if dim == 1:
        rng = c1.range
        for c in args:
            if rng[-1] + rng.step != c.first:
                raise ArgumentError("noncontiguous chain iterations")
            if rng.step != c.step:
                raise ArgumentError("chain thinning differs")
            rng = range(rng.start, c.last + 1, rng.step)
        if not all(c.names == c1.names for c in args):
            raise ArgumentError("chain names differ")
        if not all(c.chains == c1.chains for c in args):
            raise ArgumentError("sets of chains differ")
        return Chains(np.concatenate([c.value for c in cs], axis=0), start=rng.start, thin=rng.step, names=c1.names, chains=c1.chains)
    if dim == 2:
        if not all(c.range == c1.range for c in args):
            raise ArgumentError("chain ranges differ")
        names = list(c1.names)
        for c in args:
            if set(names) & set(c.names):
                raise ArgumentError("non-unique chain names")
            names += c.names
        if not all(c.chains == c1.chains for c in args):
            raise ArgumentError("sets of chains differ")
        return Chains(np.concatenate([c.value for c in cs], axis=1), start=c1.first, thin=c1.step, names=names, chains=c1.chains)
    if dim == 3:
        if not all(c.range == c1.range for c in args):
            raise ArgumentError("chain ranges differ")
        if not all(c.names == c1.names for c in args):
            raise ArgumentError("chain names differ")
        return Chains(np.concatenate([c.value for c in cs], axis=2), start=c1.first, thin=c1.step, names=c1.names)   # chains renumbered 1..m
    raise ArgumentError(f"cannot concatenate along dimension {dim}")


def hcat(c1, *args):
    return cat(2, c1, *args)


def vcat(c1, *args):
    return cat(1, c1, *args)


def readcoda(output, index):
    """readcoda(output, index) (src/output/fileio.jl:15-40): CODA files as written by OpenBUGS.  `index` rows are
    (name, first row, last row) into `output`, whose rows are (iteration, value); only the iterations at which every
    parameter was monitored are kept."""
    out = np.loadtxt(output, dtype=np.float64, ndmin=2)
    names, firstind, lastind = [], [], []
    with open(index) as f:
        for line in f:
            t = line.split()
            if t:
                names.append(t[0]); firstind.append(int(t[1])); lastind.append(int(t[2]))
    firstind = np.array(firstind); lastind = np.array(lastind)
    firstiter = out[firstind - 1, 0]; lastiter = out[lastind - 1, 0]
    thin = int((lastiter[0] - firstiter[0]) / (lastind[0] - firstind[0]))
    w0, w1 = int(firstiter.max()), int(lastiter.min())
    window = range(w0, w1 + 1, thin)
    startind = firstind + (window[0] - firstiter) / thin
    stopind = lastind - (lastiter - window[-1]) / thin
    value = np.empty((len(window), len(names)))
    for i in range(len(names)):
        value[:, i] = out[int(startind[i]) - 1:int(stopind[i]), 1]
    return Chains(value[:, :, None], start=window[0], thin=thin, names=names)


def writecoda(output, index, c, chain=1):
    """The inverse of readcoda for one chain (the reference only reads the format): rows (iteration, value) per
    parameter, index rows (name, first, last)."""
    k = c.chains.index(chain)
    row = 1
    with open(output, "w") as fo, open(index, "w") as fi:
        for j, nm in enumerate(c.names):
            for it, v in zip(c.range, c.value[:, j, k]):
                fo.write(f"{it}\t{float(v)!r}\n")
            fi.write(f"{nm}\t{row}\t{row + len(c.range) - 1}\n")
            row += len(c.range)


def _jsonable(x):
    if isinstance(x, np.ndarray):
        return {"__ndarray__": x.tolist()}
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, dict):
        return {k: _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    return x


def _unjson(x):
    if isinstance(x, dict):
        if "__ndarray__" in x:
            return np.asarray(x["__ndarray__"], dtype=float)
        return {k: _unjson(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_unjson(v) for v in x]
    return x


def write(name, c):
    """write(name, c) (src/output/fileio.jl:9-11).  The reference serialises the Julia object; the engine-side format is a
    NumPy .npz archive (no pickled objects) with the same fields (value, range, names, chains) plus, for ModelChains, what
    mcmc(mc, iters) needs to restart: template, inputs, sampler records, burnin, iteration count, seed and every chain's
    ModelState (value + tune)."""
    import json
    d = dict(kind=np.array("ModelChains" if isinstance(c, ModelChains) else "Chains"), value=c.value, start=np.int64(c.first),
             thin=np.int64(c.step), names=np.array(c.names), chains=np.array(c.chains, dtype=np.int64))
    if isinstance(c, ModelChains):
        m = c.model
        seed = c.engine.seed if c.engine is not None else getattr(c, "_seed", 123)
        d["model"] = np.array(json.dumps(dict(template=m.template, iter=int(m.iter), burnin=int(m.burnin), seed=int(seed),
                                              samplers=[_jsonable([s.params, s.kind, s.desc]) for s in m.samplers])))
        for k, v in m.inputs.items():
            d["input_" + k] = np.asarray(v)
        if c._nodelinks is not None:
            d["nodelinks"] = c._nodelinks
        if m.states:
            d["state_value"] = np.stack([st.value for st in m.states])
            d["state_tune"] = np.stack([st.tune for st in m.states])
    with open(name, "wb") as f:
        np.savez(f, **d)


def read(name, T=None):
    """read(name, T) (src/output/fileio.jl:3-7): TypeError when the stored object is not a T.  A ModelChains read back has no
    live device handle; mcmc(mc, iters) builds one from the stored template, inputs, scheme and ModelStates and continues
    the same chains (same Philox streams: the counter is the iteration number)."""
    import json
    z = np.load(name, allow_pickle=False)
    kw = dict(start=int(z["start"]), thin=int(z["thin"]), names=[str(x) for x in z["names"]], chains=[int(k) for k in z["chains"]])
    if str(z["kind"]) == "ModelChains":
        md = json.loads(str(z["model"]))
        m = Model(md["template"], iter=md["iter"], burnin=md["burnin"])
        m.samplers = [Sampler(p, k, **_unjson(d)) for p, k, d in md["samplers"]]
        setinputs(m, {k[len("input_"):]: z[k] for k in z.files if k.startswith("input_")})
        if "state_value" in z.files:
            m.states = [ModelState(v.copy(), t.copy()) for v, t in zip(z["state_value"], z["state_tune"])]
        m.hasinits = bool(m.states)
        c = ModelChains(z["value"], m, engine=None, nodelinks=z["nodelinks"] if "nodelinks" in z.files else None, **kw)
        c._seed = md["seed"]
    else:
        c = Chains(z["value"], **kw)
    if T is not None and not isinstance(c, T):
        raise TypeError(f'read("{name}", {T.__name__}): stored object is a {type(c).__name__}')
    return c


class ModelChains(Chains):
    """src/Mamba.jl:179-185"""

    def __init__(self, value, model, engine=None, nodelinks=None, streaming=True, **kw):
        super().__init__(value, **kw)
        self.model = model
        self.engine = engine          # live device handle (restart continues on it)
        self._streaming = bool(streaming) and engine is not None   # its streaming moments describe exactly these draws
        self._nodelinks = None if nodelinks is None else np.asarray(nodelinks, dtype=np.int32)

    def _names2inds(self, names):     # names2inds(mc, nodekeys): modelchains.jl:24-40 — a node key selects all of its elements
        if isinstance(names, str):
            names = [names]
        if isinstance(names, (list, tuple)) and names and all(isinstance(x, str) for x in names):
            inds, missing = [], []
            nodes = set(self.model.keys("dependent"))
            for key in names:
                if key in self.names:
                    inds.append(self.names.index(key))
                elif key in nodes:
                    found = [j for j, nm in enumerate(self.names) if nm.startswith(key + "[")]
                    if found:
                        inds += found
                    else:
                        missing.append(key)
                else:
                    missing.append(key)
            if missing:
                raise ArgumentError("chain values are missing for nodes : " + ", ".join(missing))
            return inds
        return super()._names2inds(names)

    def __getitem__(self, key):       # getindex(mc::ModelChains, …): modelchains.jl:19-22
        value, kw = self._subset(key)
        j = self._names2inds(key[1])
        nl = None if self._nodelinks is None else self._nodelinks[j]
        mc = ModelChains(value, self.model, engine=None, nodelinks=nl, **kw)
        mc._seed = getattr(self, "_seed", self.engine.seed if self.engine is not None else 123)
        return mc

    def link_codes(self):             # link(c::ModelChains): node links for stochastic nodes, the heuristic for the rest
        h = super().link_codes()
        if self._nodelinks is None:
            return h
        return np.where(self._nodelinks >= 0, self._nodelinks, h).astype(np.int32)


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
    value = eng.run(iters, burnin=burnin, thin=thin, store=store, out=store, mpsrf=True)
    if value is None:   # store=False: only the streaming moments exist (what 1e6 chains allow); diagnostics use them
        value = np.empty((0, eng.dims()[1], chains))
    return _wrap(mm, eng, value, burnin + thin, thin, chains)


def _wrap(mm, eng, value, start, thin, chains):
    vals, tune, it = eng.get_state()
    mm.iter = it
    mm.states = [ModelState(vals[k].copy(), tune[k].copy()) for k in range(chains)]    # mcmc.jl:56,82
    mm.hasinits = True
    return ModelChains(value, mm, engine=eng, nodelinks=eng.node_links(), start=start, thin=thin, names=eng.names(1),
                       chains=list(range(1, chains + 1)))


def _restart(mc, iters):
    import copy
    thin = mc.step
    if mc.last != (mc.model.iter // thin) * thin:
        raise ArgumentError("chain is missing its last iteration")          # mcmc.jl:5-6
    mm = copy.deepcopy(mc.model)                                            # mm = deepcopy(mc.model): mcmc.jl:8
    eng = mc.engine
    streaming = eng is not None and getattr(mc, "_streaming", False)
    seed = getattr(mc, "_seed", eng.seed if eng is not None else 123)
    if eng is not None and eng.iter() != mm.iter:
        # the live handle has been advanced by an earlier mcmc(mc, n) from this same object: put it back at mc's own ModelStates
        eng.set_state(np.stack([st.value for st in mm.states]), np.stack([st.tune for st in mm.states]), mm.iter)
        streaming = False
    if eng is None:   # a ModelChains that came back from read(): rebuild the device handle at the stored ModelStates
        if not mm.states:
            raise ArgumentError("chain is missing its last iteration")
        eng = Engine(mm.template, len(mc.chains), seed=seed)
        for k, v in mm.inputs.items():
            eng.set_data(k, v)
        eng.set_scheme(_block_descs(mm))
        eng.set_state(np.stack([st.value for st in mm.states]), np.stack([st.tune for st in mm.states]), mm.iter)
    value = eng.run(iters, burnin=mm.burnin, thin=thin, mpsrf=True)
    mc2 = _wrap(mm, eng, value, mc.last + thin, thin, len(mc.chains))
    # the handle's streaming moments now cover the new draws as well: only the returned object may read them; the source keeps
    # the handle for density calls (dic / predict) but its diagnostics go back to its own materialised array
    mc._streaming = False
    mc._seed = seed
    out = ModelChains(np.concatenate([mc.value, mc2.value], axis=0), mc2.model, engine=eng, nodelinks=mc2._nodelinks, streaming=streaming,
                      start=mc.first, thin=thin, names=mc.names, chains=mc.chains)
    out._seed = seed
    return out


def gelmandiag(c, alpha=0.05, mpsrf=False, transform=False):
    """gelmandiag(c; alpha, mpsrf, transform): src/output/gelmandiag.jl:3-60 (PSRF and 97.5% columns, rounded to 3 dp)."""
    if len(c.chains) < 2:
        raise ArgumentError("less than 2 chains supplied to gelman diagnostic")   # gelmandiag.jl:6-7
    eng = getattr(c, "engine", None) if getattr(c, "_streaming", False) else None
    if mpsrf and eng is not None:   # the within-chain covariances are streamed on the device for up to 12 monitored columns (mcu_diag_global)
        psrf, _, _, mv = eng.diag_global(alpha, transform, mpsrf=True)
        if not np.isnan(mv):
            return np.round(np.vstack([psrf, [mv, np.nan]]), 3), c.names + ["Multivariate"], ["PSRF", f"{100 * (1 - alpha / 2)}%"]
    if mpsrf or eng is None:   # subsets / files have no device moments, and a heuristic log / logit link has no streamed co-moments:
        codes = None           # both are computed on the materialised array (mcu_chains_gelman)
        if transform:
            codes = eng.link_codes(True) if eng is not None else c.link_codes()
        psrf = _chains_gelman(c.value, alpha, codes, bool(mpsrf))
        return np.round(psrf, 3), c.names + (["Multivariate"] if mpsrf else []), ["PSRF", f"{100 * (1 - alpha / 2)}%"]
    psrf = eng.gelman(alpha, transform)
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


def cor(c):
    """cor(c): src/output/stats.jl:14-16 — correlation matrix of the pooled draws (combine(c))."""
    return np.corrcoef(c.combine(), rowvar=False), c.names, c.names


def _device_for(mc):
    """A live handle for model-based post-processing: the one the chain was sampled on, or one rebuilt from the stored model."""
    if mc.engine is not None:
        return mc.engine
    mm = mc.model
    eng = Engine(mm.template, 1, seed=getattr(mc, "_seed", 123))
    for k, v in mm.inputs.items():
        eng.set_data(k, v)
    return eng


def _factor_mask(mc, eng, nodekeys):
    """nodekeys → (factor bitmask, state-element names the selected densities read).  The second item is what getsimkeys
    (src/output/modelstats.jl:102-127) collects as relistkeys: the stochastic nodes on a path to the selected nodes."""
    m = mc.model
    nn, nf = eng.factor_counts()
    pnodes = [n for n, _ in _TEMPLATES[m.template]["nodes"]]
    outputs = _TEMPLATES[m.template]["outputs"]
    if nodekeys is None:
        nodekeys = pnodes + outputs                        # keys(mc.model, :stochastic): modelstats.jl:28-29
    if isinstance(nodekeys, str):
        nodekeys = [nodekeys]
    mask, need = 0, set()
    for key in nodekeys:
        if key in pnodes:
            f = pnodes.index(key)
        elif key in outputs:
            f = nn + outputs.index(key)
        else:
            raise KeyError(key)
        mask |= 1 << f
        par = eng.factor_parents(f) | ((1 << f) if f < nn else 0)
        need |= {pnodes[q] for q in range(nn) if (par >> q) & 1}
    return mask, [n for n in pnodes if n in need]


def _states_from_chains(mc, eng, neednodes):
    """Full state records [n * m x D] (chain-major rows, as vec(value[:, j, :])) with the needed nodes' elements taken from the
    monitored columns; raises the reference's error when a needed node was not monitored (modelchains.jl:31-37)."""
    snames = eng.names(0)
    n, _, m = mc.value.shape
    D = len(snames)
    base = mc.model.states[0].value if mc.model.states else np.ones(D)
    st = np.tile(np.asarray(base, dtype=float), (n * m, 1))
    missing, cols = [], []
    for node in neednodes:
        el = [e for e, nm in enumerate(snames) if nm == node or nm.startswith(node + "[")]
        if all(snames[e] in mc.names for e in el):
            cols += [(e, mc.names.index(snames[e])) for e in el]
        else:
            missing.append(node)
    if missing:
        raise ArgumentError("chain values are missing for nodes : " + ", ".join(missing))
    for e, j in cols:
        st[:, e] = mc.value[:, j, :].T.reshape(-1)
    return st, cols


def logpdf(mc, nodekeys=None):
    """logpdf(mc::ModelChains, nodekeys = keys(mc.model, :stochastic)): src/output/modelstats.jl:28-58 — the summed log density of
    the named stochastic nodes at every kept draw → a one-column ModelChains named "logpdf".  Evaluated on the device."""
    eng = _device_for(mc)
    mask, need = _factor_mask(mc, eng, nodekeys)
    st, _ = _states_from_chains(mc, eng, need)
    n, _, m = mc.value.shape
    lp = eng.logpdf_nodes(mask, st).reshape(m, n).T
    out = ModelChains(lp[:, None, :], mc.model, engine=None, start=mc.first, thin=mc.step, names=["logpdf"], chains=mc.chains)
    return out


def dic(mc):
    """dic(mc): src/output/modelstats.jl:3-13 → [[DIC_pD, pD], [DIC_pV, pV]] with rows ("pD", "pV") and columns
    ("DIC", "Effective Parameters"); the deviance is -2 logpdf of the observed nodes (keys(m, :output))."""
    eng = _device_for(mc)
    outputs = _TEMPLATES[mc.model.template]["outputs"]
    mask, need = _factor_mask(mc, eng, outputs)
    st, cols = _states_from_chains(mc, eng, need)
    Dev = -2.0 * eng.logpdf_nodes(mask, st)
    mean_state = st[:1].copy()
    for e, j in cols:
        mean_state[0, e] = mc.value[:, j, :].mean()                       # logpdf(mc, mean, nodekeys): modelstats.jl:16-25
    Dhat = -2.0 * eng.logpdf_nodes(mask, mean_state)[0]
    p = np.array([Dev.mean() - Dhat, 0.5 * Dev.var(ddof=1)])
    return np.column_stack([Dhat + 2.0 * p, p]), ["pD", "pV"], ["DIC", "Effective Parameters"]


def predict(mc, nodekeys=None, stream_id=0):
    """predict(mc, nodekeys = keys(m, :output)): src/output/modelstats.jl:61-96 — posterior predictive draws of the observed nodes,
    one per kept draw and chain → ModelChains named y[1], y[2], …  Drawn on the device."""
    outputs = _TEMPLATES[mc.model.template]["outputs"]
    if nodekeys is None:
        nodekeys = outputs
    if isinstance(nodekeys, str):
        nodekeys = [nodekeys]
    if not all(k in outputs for k in nodekeys):
        raise ArgumentError("nodekeys are not all observed Stochastic nodess : " + ", ".join(outputs))     # modelstats.jl:69-73 (sic)
    eng = _device_for(mc)
    _, need = _factor_mask(mc, eng, nodekeys)
    st, _ = _states_from_chains(mc, eng, need)
    n, _, m = mc.value.shape
    draws = eng.predict(st, stream_id)                                   # [n * m x L], rows chain-major; all observed nodes in order
    lens = _TEMPLATES[mc.model.template].get("output_lens", [draws.shape[1]])
    offs = np.concatenate([[0], np.cumsum(lens)])
    cols, names = [], []
    for key in nodekeys:
        q = outputs.index(key)
        cols += list(range(offs[q], offs[q + 1])); names += [f"{key}[{i + 1}]" for i in range(lens[q])]
    draws = draws[:, cols]
    L = draws.shape[1]
    value = draws.reshape(m, n, L).transpose(1, 2, 0)
    return ModelChains(value, mc.model, engine=None, start=mc.first, thin=mc.step, names=names, chains=mc.chains)


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


def format_summary(value, rownames, colnames):
    """Text layout of show(io, s::ChainSummary) (src/output/chainsummary.jl:50-84): right-aligned columns, one common fixed-point format
    per column (the reference delegates the digits to Showoff.showoff; here 8 significant digits of the column's largest entry), column
    names centred on the column widths, row names right-aligned."""
    v = np.atleast_2d(np.asarray(value, dtype=float))
    cols = []
    for j in range(v.shape[1]):
        col = v[:, j]
        fin = np.abs(col[np.isfinite(col)])
        big = fin.max() if fin.size else 0.0
        if big != 0.0 and (big >= 1e9 or big < 1e-5):
            cols.append([f"{x:.7e}" for x in col])
        else:
            dec = int(min(12, max(0, 8 - (int(np.floor(np.log10(big))) + 1 if big > 0 else 1))))
            cols.append(["NaN" if np.isnan(x) else f"{x:.{dec}f}" for x in col])
    rn = max(len(r) for r in rownames)
    wid = [1 + max(len(cn), max(len(s) for s in cstr)) for cn, cstr in zip(colnames, cols)]
    out = [" " * rn]
    for cn, w in zip(colnames, wid):
        nspace = w - len(cn) - 1
        nright = nspace >> 1
        out[0] += " " * (1 + nspace - nright) + cn + " " * nright
    for i, r in enumerate(rownames):
        out.append(" " * (rn - len(r)) + r + "".join(" " * (w - len(cstr[i])) + cstr[i] for w, cstr in zip(wid, cols)))
    return "\n".join(out) + "\n"


def describe_text(c, q=(0.025, 0.25, 0.5, 0.75, 0.975), etype="bm"):
    """What describe(io, c) prints (src/output/stats.jl:43-52): header, "Empirical Posterior Estimates:", "Quantiles:"."""
    (ss, names, cols), (qq, _, qcols) = describe(c, q=q, etype=etype)
    return (c.header() + "\nEmpirical Posterior Estimates:\n" + format_summary(ss, names, cols) + "\nQuantiles:\n" + format_summary(qq, names, qcols) + "\n")


def describe(c, q=(0.025, 0.25, 0.5, 0.75, 0.975), etype="bm"):
    """describe(c): src/output/stats.jl:41-52 — summarystats + quantiles."""
    return summarystats(c, etype=etype), quantile(c, q=q)


def summarystats(c, etype="bm", batch=100):
    """summarystats(c; etype): src/output/stats.jl:85-94 → [p × 5] Mean, SD, Naive SE, MCSE, ESS."""
    if etype not in _ETYPE:
        raise ArgumentError(f"unsupported mcse method {etype}")                   # mcse.jl:3-8
    cols = ["Mean", "SD", "Naive SE", "MCSE", "ESS"]
    eng = getattr(c, "engine", None) if getattr(c, "_streaming", False) else None
    if c.value.shape[0] == 0:   # run with store=False: batch means of size 100 accumulated on the device
        if eng is None or etype != "bm":
            raise ArgumentError("no stored samples: only the streaming batch-means summary is available")
        return eng.summary_streaming(), c.names, cols
    v = _value_f(c); n, p, m = v.shape
    out = np.empty((p, 5))
    if _lib.lib().mcu_chains_summarystats(_dp(v), n, p, m, _ETYPE[etype], int(batch), _dp(out)) != 0:
        raise ArgumentError(f"iterations are < {2 * batch} and batch size is > {n * m // 2}")     # mcse.jl:13-16
    return out, c.names, cols
