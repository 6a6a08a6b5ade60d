## MambaCUDA.jl — Julia shim that routes Mamba's mcmc() through libmambacuda.so.
##
## NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` (and the reference targets
## Julia 0.5, src/Mamba.jl:52-185 uses `type` / `immutable` / `typealias`).  The same C ABI is exercised
## from Python (mamba.jl_b200/mambacuda, tests/) — this file is the binding a Mamba maintainer adds.
## Syntax follows the reference's Julia 0.5 dialect.
##
## What it replaces: mcmc_master!'s `pmap2(mcmc_worker!, lsts)` (src/model/mcmc.jl:36-59).  Everything
## above that line — Model construction, setsamplers!, argument checks, the ModelChains result — stays
## Mamba's own code.

module MambaCUDA

using Mamba
import Mamba: Model, ModelChains, ModelState, Sampler, Chains,
              AMWGTune, SliceTune, RWMTune, NUTSTune, HMCTune, AMMTune, MALATune

const libmambacuda = get(ENV, "LIBMAMBACUDA", "libmambacuda.so")

const MCU_MAX_BLOCK_NODES = 8

## mirror of `mcu_block_desc` (include/mambacuda.h); isbits so it can be passed in a Vector
immutable BlockDesc
  kind::Int32
  n_nodes::Int32
  nodes::NTuple{8, Int32}
  transform::Int32
  adapt::Int32
  batchsize::Int32
  proposal::Int32
  L::Int32
  grad::Int32
  max_depth::Int32
  n_scale::Int32
  target::Float64
  epsilon::Float64
  beta::Float64
  amm_scale::Float64
  scale::Ptr{Float64}
end

## template library: node order of the state record (include/mambacuda.h, MCU_TPL_*)
const TEMPLATES = Dict(
  0 => [:beta, :s2],                                                      # doc/tutorial/line.jl
  1 => [:alpha0, :alpha1, :alpha2, :alpha12, :s2, :b],                    # doc/examples/seeds.jl
  2 => [:mu_alpha, :mu_beta, :s2_alpha, :s2_beta, :s2_c, :alpha, :beta],  # doc/examples/rats.jl
  3 => [:alpha, :beta, :theta],                                           # doc/examples/pumps.jl
  5 => [:mu, :s2, :b],                                                    # doc/examples/surgical.jl
  6 => [:s2_between, :theta, :s2_within, :mu],                            # doc/examples/dyes.jl
  7 => [:s2, :gamma, :beta, :alpha, :lambda],                             # doc/examples/salm.jl
  8 => [:s2_2, :s2_1, :pi, :phi, :mu, :delta],                            # doc/examples/equiv.jl
  9 => [:s2, :d, :delta_new, :mu, :delta],                                # doc/examples/blocker.jl
  10 => [:beta0, :beta, :s2],                                             # doc/examples/stacks.jl
  11 => [:priors, :mu, :theta, :pc],                                      # doc/examples/magnesium.jl (6 x 8 matrices column-major)
  12 => [:alpha, :beta1, :beta2, :s2, :b, :mu],                           # doc/examples/oxford.jl
  13 => [:a0, :alpha_Base, :alpha_Trt, :alpha_BT, :alpha_Age, :alpha_V4, :s2_b1, :s2_b, :b1, :b]   # doc/examples/epil.jl
)                                                                         # (4 = the GLM family: pass template=4 and inputs X, y)

function check(h::Ptr{Void}, rc::Cint)
  if rc != 0
    msg = unsafe_string(ccall((:mcu_last_error, libmambacuda), Cstring, (Ptr{Void},), h))
    rc == -1 ? throw(ArgumentError(msg)) :
    rc == -2 ? throw(DimensionMismatch(msg)) : error("libmambacuda: $msg")
  end
end

## A model matches a template when its unobserved stochastic nodes are exactly the template's nodes.
## (The node closures themselves cannot be inspected; the caller asserts the template with `template=`
## when the names alone are ambiguous.)
function matchtemplate(m::Model)
  sampled = Symbol[]
  for s in m.samplers
    append!(sampled, s.params)
  end
  for (id, nodes) in TEMPLATES
    if Set(sampled) == Set(nodes)
      return id
    end
  end
  throw(ArgumentError("model does not match a device template (no CPU fallback)"))
end

adaptcode(a::Symbol) = a == :all ? 0 : a == :burnin ? 1 : 2

## Build one descriptor from a Sampler produced by Mamba's own constructors.  The constructor
## arguments live in the sampler closure (src/samplers/amwg.jl:52-59 etc.); the shim's versions of the
## constructors below record them in `shimargs` keyed by the Sampler object.
const shimargs = ObjectIdDict()

function blockdesc(m::Model, s::Sampler, nodeids::Dict{Symbol, Int}, keep::Vector{Any})
  a = get(shimargs, s, Dict{Symbol, Any}())
  kind = isa(s.tune, AMWGTune) ? 0 :
         isa(s.tune, SliceTune{Univariate}) ? 1 :
         isa(s.tune, SliceTune{Multivariate}) ? 2 :
         isa(s.tune, RWMTune) ? 3 :
         isa(s.tune, NUTSTune) ? 4 :
         isa(s.tune, HMCTune) ? 5 :
         isa(s.tune, AMMTune) ? 6 :
         isa(s.tune, MALATune) ? 8 :
         get(a, :gibbs, false) ? 7 :    # user-defined sampler registered through Gibbs(...) below
         throw(ArgumentError("sampler $(typeof(s.tune)) has no device implementation (no CPU fallback)"))
  length(s.params) <= MCU_MAX_BLOCK_NODES || throw(ArgumentError("a block names at most 8 nodes"))
  nodes = zeros(Int32, 8)
  for (i, p) in enumerate(s.params)
    nodes[i] = nodeids[p]
  end
  scale = Float64[]
  if haskey(a, :scale)
    scale = vec(convert(Array{Float64}, collect(a[:scale])))   # column-major for Sigma matrices
    push!(keep, scale)
  end
  BlockDesc(kind, length(s.params), (nodes...),
            Int32(get(a, :transform, kind in (1, 2) ? false : true)),
            Int32(adaptcode(get(a, :adapt, :all))), Int32(get(a, :batchsize, 0)),
            Int32(get(a, :proposal, 0)), Int32(get(a, :L, 0)),
            Int32(get(a, :dtype, :analytic) == :forward ? 1 : get(a, :dtype, :analytic) == :central ? 2 : 0),
            Int32(get(a, :max_depth, 0)), Int32(length(scale)),
            Float64(get(a, :target, 0.0)), Float64(get(a, :epsilon, 0.0)),
            Float64(get(a, :beta, 0.0)), Float64(get(a, :amm_scale, 0.0)),
            isempty(scale) ? Ptr{Float64}(0) : pointer(scale))
end

## constructor wrappers: call Mamba's constructor, remember the arguments
function AMWG(params, sigma; adapt::Symbol=:all, batchsize::Integer=50, target::Real=0.44)
  s = Mamba.AMWG(params, sigma, adapt=adapt, batchsize=batchsize, target=target)
  shimargs[s] = Dict(:scale => sigma, :adapt => adapt, :batchsize => batchsize, :target => target); s
end
function Slice(params, width, F=Multivariate; transform::Bool=false)
  s = Mamba.Slice(params, width, F, transform=transform)
  shimargs[s] = Dict(:scale => width, :transform => transform); s
end
function RWM(params, scale; proposal=Normal)
  s = Mamba.RWM(params, scale, proposal=proposal)
  code = proposal == Normal ? 0 : proposal == SymUniform ? 1 : proposal == SymTriangularDist ? 2 :
         proposal == Cosine ? 3 : proposal == Epanechnikov ? 4 : proposal == Biweight ? 5 : proposal == Triweight ? 6 :
         throw(ArgumentError("proposal $proposal has no device implementation"))
  shimargs[s] = Dict(:scale => scale, :proposal => code); s
end
function NUTS(params; dtype::Symbol=:forward, target::Real=0.6, max_depth::Integer=10)
  s = Mamba.NUTS(params, dtype=dtype, target=target)
  shimargs[s] = Dict(:dtype => dtype, :target => target, :max_depth => max_depth); s
end
function HMC(params, epsilon::Real, L::Integer, Sigma=nothing; dtype::Symbol=:forward)
  s = Sigma == nothing ? Mamba.HMC(params, epsilon, L, dtype=dtype) : Mamba.HMC(params, epsilon, L, Sigma, dtype=dtype)
  d = Dict{Symbol, Any}(:epsilon => epsilon, :L => L, :dtype => dtype)
  Sigma == nothing || (d[:scale] = Sigma)
  shimargs[s] = d; s
end
function MALA(params, epsilon::Real, Sigma=nothing; dtype::Symbol=:forward)
  s = Sigma == nothing ? Mamba.MALA(params, epsilon, dtype=dtype) : Mamba.MALA(params, epsilon, Sigma, dtype=dtype)
  d = Dict{Symbol, Any}(:epsilon => epsilon, :dtype => dtype)
  Sigma == nothing || (d[:scale] = Sigma)
  shimargs[s] = d; s
end
## A user-defined Gibbs sampler Sampler(params, f) has no device equivalent in general (f is a Julia closure); for the
## node sets a template registers a conjugate full conditional for (pumps: [:theta], [:beta]; line: [:beta], [:s2]) the shim tags the sampler so
## that the block runs as MCU_GIBBS; `f` stays attached for the CPU path of stock Mamba.
function Gibbs(params, f::Function)
  s = Mamba.Sampler(params, f)
  shimargs[s] = Dict{Symbol, Any}(:gibbs => true, :transform => false); s
end
function AMM(params, Sigma; adapt::Symbol=:all, beta::Real=0.05, scale::Real=2.38)
  s = Mamba.AMM(params, Sigma, adapt=adapt, beta=beta, scale=scale)
  shimargs[s] = Dict(:scale => Sigma, :adapt => adapt, :beta => beta, :amm_scale => scale); s
end

## What a restart needs to rebuild the device handle: template id, the inputs that were uploaded, the block descriptors (the constructor
## arguments live in `shimargs` under the ORIGINAL Sampler objects, so the descriptors are built before the model is deep-copied) and the
## seed.  Keyed by the model object stored in the ModelChains that mcmc returns.
type RunInfo
  tid::Int
  inputs::Dict{Symbol, Any}
  descs::Vector{BlockDesc}
  keep::Vector{Any}
  seed::UInt64
  device::Int
end
const shiminfo = ObjectIdDict()

function openhandle(info::RunInfo, chains::Integer; offset::Integer=0, device::Integer=info.device)
  h = Ref{Ptr{Void}}(C_NULL)
  rc = ccall((:mcu_create, libmambacuda), Cint, (Cint, Int64, Int64, Cint, UInt64, Ptr{Ptr{Void}}),
             info.tid, chains, offset, device, info.seed, h)
  rc == 0 || error(unsafe_string(ccall((:mcu_last_error, libmambacuda), Cstring, (Ptr{Void},), C_NULL)))
  ## inputs (setinputs!, src/model/initialization.jl:30-40); integer data travel as Float64
  for (key, value) in info.inputs
    isa(value, AbstractArray) || continue
    x = convert(Array{Float64}, value)
    ## index inputs are 0-based offsets on the device (include/mambacuda.h); the scripts hold Julia's 1-based indices (rats.jl:42, dyes.jl:16)
    if (info.tid == 2 && key == :rat) || (info.tid == 6 && key == :batch)
      x = x - 1.0
    end
    dims = Int64[size(x)...]
    check(h[], ccall((:mcu_set_data, libmambacuda), Cint, (Ptr{Void}, Cstring, Cint, Ptr{Int64}, Ptr{Float64}),
                     h[], string(key), length(dims), dims, x))
  end
  check(h[], ccall((:mcu_set_scheme, libmambacuda), Cint, (Ptr{Void}, Cint, Ptr{BlockDesc}), h[], length(info.descs), info.descs))
  h[]
end

## run `iters` more iterations on an initialised handle and wrap the result exactly as mcmc_master! does (mcmc.jl:54-58)
function runhandle(h::Ptr{Void}, mm::Model, first::Integer, iters::Integer, thin::Integer, chains::Integer)
  D = Ref{Cint}(0); P = Ref{Cint}(0); NN = Ref{Cint}(0)
  check(h, ccall((:mcu_dims, libmambacuda), Cint, (Ptr{Void}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), h, D, P, NN))
  kept = ccall((:mcu_kept, libmambacuda), Int64, (Int64, Int64, Int64, Int64), mm.iter, iters, mm.burnin, thin)
  value = Array{Float64}(kept, P[], chains)            # == ModelChains.value layout, filled in place
  check(h, ccall((:mcu_run, libmambacuda), Cint, (Ptr{Void}, Int64, Int64, Int64, Ptr{Float64}, UInt32),
                 h, iters, mm.burnin, thin, value, 0))
  ## final ModelStates (mcmc.jl:56,82): values + tune records
  nt = Ref{Int64}(0)
  check(h, ccall((:mcu_tune_size, libmambacuda), Cint, (Ptr{Void}, Ptr{Int64}), h, nt))
  vals = Array{Float64}(D[], chains); tune = Array{Float64}(max(nt[], 1), chains); it = Ref{Int64}(0)
  check(h, ccall((:mcu_get_state, libmambacuda), Cint, (Ptr{Void}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                 h, vals, tune, it))
  mm.iter = it[]
  mm.states = ModelState[ModelState(vals[:, k], Any[tune[:, k]]) for k in 1:chains]
  buf = Vector{UInt8}(1 << 16)
  ccall((:mcu_names, libmambacuda), Cint, (Ptr{Void}, Cint, Ptr{UInt8}, Csize_t), h, 1, buf, length(buf))
  pnames = split(unsafe_string(pointer(buf)), '\n')
  ModelChains(Chains(value, start=first, thin=thin, names=AbstractString[pnames...]), mm)
end

## mcmc(model, inputs, inits, iters; burnin, thin, chains): same signature, checks and result as
## src/model/mcmc.jl:19-33; the chains x iterations loop is ONE call into the library.
function mcmc(m::Model, inputs::Dict{Symbol}, inits::Vector{Dict{Symbol, Any}}, iters::Integer;
              burnin::Integer=0, thin::Integer=1, chains::Integer=1, verbose::Bool=true,
              seed::Integer=123, device::Integer=0, template::Integer=-1)
  iters > burnin || throw(ArgumentError("burnin is greater than or equal to iters"))
  length(inits) >= chains || throw(ArgumentError("fewer initial values than chains"))

  tid = template >= 0 ? template : matchtemplate(m)
  nodes = TEMPLATES[tid]
  nodeids = Dict{Symbol, Int}([nodes[i] => i - 1 for i in 1:length(nodes)])
  keep = Any[]
  descs = BlockDesc[blockdesc(m, s, nodeids, keep) for s in m.samplers]   # on the caller's Sampler objects: see RunInfo

  mm = deepcopy(m)
  setinputs!(mm, inputs)
  setinits!(mm, inits[1:chains])          # validates the Dicts exactly as the reference does
  mm.burnin = burnin
  mm.iter = 0
  info = RunInfo(tid, inputs, descs, keep, UInt64(seed), device)

  h = openhandle(info, chains)
  try
    ## initial values → [D × chains] (record contiguous == column-major D × n); matrices flatten column-major, as unlist does
    x0 = hcat([vcat([vec(Float64[inits[k][key]...]) for key in nodes]...) for k in 1:chains]...)
    check(h, ccall((:mcu_set_inits, libmambacuda), Cint, (Ptr{Void}, Ptr{Float64}, Int64, Float64), h, x0, chains, 0.0))
    mc = runhandle(h, mm, burnin + thin, iters, thin, chains)
    shiminfo[mc.model] = info
    return mc
  finally
    ccall((:mcu_destroy, libmambacuda), Cint, (Ptr{Void},), h)
  end
end

## mcmc over several GPUs of the box from ONE Julia process (the reference farms chains out with pmap2(mcmc_worker!, lsts),
## src/model/mcmc.jl:48-52): chains are cut into contiguous shards, shard k runs on devices[k] with global chain ids
## offset .. offset + n - 1 (the Philox key is the global id: the samples do not depend on the sharding), every run is queued with
## MCU_RUN_ASYNC before any is waited for, and the shards' samples land side by side in ModelChains.value.
const MCU_RUN_NO_STORE = UInt32(1)
const MCU_RUN_ASYNC = UInt32(16)

function mcmc(m::Model, inputs::Dict{Symbol}, inits::Vector{Dict{Symbol, Any}}, iters::Integer, devices::Vector{Int};
              burnin::Integer=0, thin::Integer=1, chains::Integer=1, verbose::Bool=true, seed::Integer=123, template::Integer=-1)
  iters > burnin || throw(ArgumentError("burnin is greater than or equal to iters"))
  length(inits) >= chains || throw(ArgumentError("fewer initial values than chains"))
  tid = template >= 0 ? template : matchtemplate(m)
  nodes = TEMPLATES[tid]
  nodeids = Dict{Symbol, Int}([nodes[i] => i - 1 for i in 1:length(nodes)])
  keep = Any[]
  descs = BlockDesc[blockdesc(m, s, nodeids, keep) for s in m.samplers]
  mm = deepcopy(m)
  setinputs!(mm, inputs); setinits!(mm, inits[1:chains]); mm.burnin = burnin; mm.iter = 0
  info = RunInfo(tid, inputs, descs, keep, UInt64(seed), devices[1])
  G = length(devices)
  cuts = [div(k * chains, G) for k in 0:G]                       # chain c lives on GPU floor(c G / chains): SURVEY.md §8e
  x0 = hcat([vcat([vec(Float64[inits[k][key]...]) for key in nodes]...) for k in 1:chains]...)
  hs = Ptr{Void}[]
  try
    for k in 1:G
      n = cuts[k + 1] - cuts[k]
      h = openhandle(info, n, offset=cuts[k], device=devices[k]); push!(hs, h)
      check(h, ccall((:mcu_set_inits, libmambacuda), Cint, (Ptr{Void}, Ptr{Float64}, Int64, Float64), h, pointer(x0, size(x0, 1) * cuts[k] + 1), n, 0.0))
      check(h, ccall((:mcu_run, libmambacuda), Cint, (Ptr{Void}, Int64, Int64, Int64, Ptr{Float64}, UInt32), h, iters, burnin, thin, C_NULL, MCU_RUN_ASYNC))
    end
    D = Ref{Cint}(0); P = Ref{Cint}(0); NN = Ref{Cint}(0)
    check(hs[1], ccall((:mcu_dims, libmambacuda), Cint, (Ptr{Void}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), hs[1], D, P, NN))
    kept = ccall((:mcu_kept, libmambacuda), Int64, (Int64, Int64, Int64, Int64), 0, iters, burnin, thin)
    value = Array{Float64}(kept, P[], chains)
    nt = Ref{Int64}(0)
    check(hs[1], ccall((:mcu_tune_size, libmambacuda), Cint, (Ptr{Void}, Ptr{Int64}), hs[1], nt))
    vals = Array{Float64}(D[], chains); tune = Array{Float64}(max(nt[], 1), chains); it = Ref{Int64}(0)
    for k in 1:G
      h = hs[k]
      check(h, ccall((:mcu_wait, libmambacuda), Cint, (Ptr{Void},), h))
      check(h, ccall((:mcu_get_samples, libmambacuda), Cint, (Ptr{Void}, Ptr{Float64}), h, pointer(value, kept * P[] * cuts[k] + 1)))
      check(h, ccall((:mcu_get_state, libmambacuda), Cint, (Ptr{Void}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                     h, pointer(vals, D[] * cuts[k] + 1), pointer(tune, size(tune, 1) * cuts[k] + 1), it))
    end
    mm.iter = it[]
    mm.states = ModelState[ModelState(vals[:, k], Any[tune[:, k]]) for k in 1:chains]
    buf = Vector{UInt8}(1 << 16)
    ccall((:mcu_names, libmambacuda), Cint, (Ptr{Void}, Cint, Ptr{UInt8}, Csize_t), hs[1], 1, buf, length(buf))
    pnames = split(unsafe_string(pointer(buf)), '\n')
    mc = ModelChains(Chains(value, start=burnin + thin, thin=thin, names=AbstractString[pnames...]), mm)
    shiminfo[mc.model] = info
    return mc
  finally
    for h in hs
      ccall((:mcu_destroy, libmambacuda), Cint, (Ptr{Void},), h)
    end
  end
end

## gelmandiag / summarystats over the chains of several live handles WITHOUT gathering the samples (10^6 chains do not fit a Chains
## array): the packed two-round protocol of include/mambacuda.h, the O(p) buffers combined here in Julia.  Worker processes
## (addprocs, one per GPU) can instead join an NCCL communicator: mcu_comm_unique_id on one, the 128 bytes sent with remotecall,
## mcu_comm_init on each, then mcu_diag_global does both rounds on the devices.
function devicediagnostics(hs::Vector{Ptr{Void}}; alpha::Real=0.05, transform::Bool=false)
  P = Ref{Cint}(0)
  check(hs[1], ccall((:mcu_dims, libmambacuda), Cint, (Ptr{Void}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), hs[1], C_NULL, P, C_NULL))
  p = Int(P[])
  n1 = Ref{Cint}(0); n2 = Ref{Cint}(0)                             # 11p and 15p + p(p-1) (the pair sums of the multivariate PSRF, p <= 12)
  ccall((:mcu_diag_sizes, libmambacuda), Cint, (Cint, Ptr{Cint}, Ptr{Cint}), p, n1, n2)
  r1 = Array{Float64}(n1[], length(hs))
  for (k, h) in enumerate(hs)
    check(h, ccall((:mcu_diag_round1, libmambacuda), Cint, (Ptr{Void}, Ptr{Float64}), h, pointer(r1, n1[] * (k - 1) + 1)))
  end
  red1 = vcat(minimum(r1[1:p, :], 2), maximum(r1[p+1:2p, :], 2), sum(r1[2p+1:end, :], 2))[:]
  r2 = Array{Float64}(n2[], length(hs))
  for (k, h) in enumerate(hs)
    check(h, ccall((:mcu_diag_round2, libmambacuda), Cint, (Ptr{Void}, Cint, Ptr{Float64}, Ptr{Float64}), h, transform, red1, pointer(r2, n2[] * (k - 1) + 1)))
  end
  red2 = sum(r2, 2)[:]
  monlink = Array{Cint}(p); nkept = Ref{Int64}(0)
  check(hs[1], ccall((:mcu_monitor_links, libmambacuda), Cint, (Ptr{Void}, Ptr{Cint}), hs[1], monlink))
  check(hs[1], ccall((:mcu_n_kept, libmambacuda), Cint, (Ptr{Void}, Ptr{Int64}), hs[1], nkept))
  psrf = Array{Float64}(2, p); summ = Array{Float64}(5, p); codes = Array{Cint}(p); mpsrf = Ref{Float64}(NaN)
  rc = ccall((:mcu_diag_finish, libmambacuda), Cint,
             (Int64, Cint, Float64, Ptr{Cint}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}, Ptr{Float64}),
             nkept[], p, alpha, monlink, transform, red1, red2, psrf, summ, codes, mpsrf)   # mpsrf: NaN unless the runs streamed co-moments (MCU_RUN_MPSRF)
  rc == 0 || throw(ArgumentError("less than 2 chains supplied to gelman diagnostic"))
  round(psrf', 3), summ'          # gelmandiag rounds to 3 dp (gelmandiag.jl:59); summary columns: Mean, SD, Naive SE, MCSE, ESS
end

## mcmc(mc, iters): restart (src/model/mcmc.jl:3-16) — the handle is rebuilt at the stored ModelStates (values, tune records, iteration
## counter); the Philox counters are functions of the iteration number, so the continuation equals an uninterrupted run.
function mcmc(mc::ModelChains, iters::Integer; verbose::Bool=true)
  thin = step(mc)
  last(mc) == div(mc.model.iter, thin) * thin || throw(ArgumentError("chain is missing its last iteration"))
  haskey(shiminfo, mc.model) || throw(ArgumentError("this ModelChains was not produced by MambaCUDA.mcmc"))
  info = shiminfo[mc.model]
  mm = deepcopy(mc.model)
  chains = length(mc.chains)
  h = openhandle(info, chains)
  try
    vals = hcat([st.value for st in mm.states]...)
    tune = hcat([st.tune[1] for st in mm.states]...)
    check(h, ccall((:mcu_set_state, libmambacuda), Cint, (Ptr{Void}, Ptr{Float64}, Ptr{Float64}, Int64), h, vals, tune, mm.iter))
    mc2 = runhandle(h, mm, last(mc) + thin, iters, thin, chains)
    shiminfo[mc2.model] = info
    if mc2.names != mc.names
      mc2 = mc2[:, mc.names, :]
    end
    return ModelChains(vcat(mc, mc2), mc2.model)
  finally
    ccall((:mcu_destroy, libmambacuda), Cint, (Ptr{Void},), h)
  end
end

end # module
