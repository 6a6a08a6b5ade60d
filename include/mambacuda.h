/*
 * mambacuda.h — C ABI of libmambacuda.so, the B200 (sm_100a) batched MCMC engine that sits
 * behind Mamba.jl's `Sampler` / `mcmc()` API.
 *
 * The reference (jamesonquinn/Mamba.jl, pure Julia) has no FFI for this path; these entry
 * points are what a Julia `ccall` shim (mamba.jl_b200/julia/MambaCUDA.jl)
 * binds in place of `mcmc_master!`'s `pmap2(mcmc_worker!, lsts)` (src/model/mcmc.jl:36-59).
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference root).
 *
 * Conventions
 *  - Every pointer is a HOST pointer owned by the caller, read/written only during the call.
 *  - All floating point data is IEEE double (the reference is Float64 throughout,
 *    src/Mamba.jl:129-131,152-155,172-177).
 *  - Host matrices are "one record contiguous": `x[n][D]` in C order == Julia `Array{Float64}(D, n)`.
 *  - The thinned sample block written by mcu_run is Julia column-major
 *    `[kept × n_monitor × n_chains]` (iteration fastest), i.e. exactly `ModelChains.value`
 *    (src/Mamba.jl:172-185, src/output/chains.jl:5-32).
 *  - Return value 0 = success, negative = error code; text via mcu_last_error().
 *    Nothing throws or longjmps across the ABI.  Numerical trouble (-Inf / NaN log densities)
 *    is in-band, as in the reference (src/distributions/distributionstruct.jl:138-140).
 *  - A handle is bound to ONE CUDA device and is not thread-safe; distinct handles may be used
 *    from distinct threads/processes.  Chains are sharded across handles by
 *    (chain_offset, n_chains); the Philox key is the GLOBAL chain id, so results do not depend
 *    on how chains are sharded.
 *  - There is no CPU fallback: if no CUDA device is usable every compute entry point fails.
 */
#ifndef MAMBACUDA_H
#define MAMBACUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCU_ABI_VERSION 1

/* ---- error codes -------------------------------------------------------------------------- */
enum {
  MCU_OK = 0,
  MCU_ERR_ARG = -1,        /* ArgumentError in the reference (e.g. src/model/mcmc.jl:22-25)   */
  MCU_ERR_DIM = -2,        /* DimensionMismatch (src/model/simulation.jl:20-23)               */
  MCU_ERR_STATE = -3,      /* call order (inputs before inits: src/model/initialization.jl:4) */
  MCU_ERR_CUDA = -4,       /* CUDA runtime failure / no device                                */
  MCU_ERR_UNSUPPORTED = -5 /* no device template for the request (no CPU fallback)            */
};

/* ---- model templates: fixed library compiled to device functions ------- */
enum {
  MCU_TPL_LINE = 0,      /* doc/tutorial/line.jl:5-25      nodes: beta[2], s2                 */
  MCU_TPL_SEEDS = 1,     /* doc/examples/seeds.jl:16-56    nodes: alpha0, alpha1, alpha2, alpha12, s2, b[21] */
  MCU_TPL_RATS = 2,      /* doc/examples/rats.jl:49-97     nodes: mu_alpha, mu_beta, s2_alpha, s2_beta, s2_c, alpha[30], beta[30] */
  MCU_TPL_PUMPS = 3,     /* doc/examples/pumps.jl:12-39    nodes: alpha, beta, theta[10]      */
  MCU_TPL_GLM_LOGIT = 4, /* synthetic GLM family (inputs X, y, family, sigma)  nodes: beta[d]  (no reference file; closest doc/examples/seeds.jl) */
  MCU_TPL_SURGICAL = 5,  /* doc/examples/surgical.jl:11-43 nodes: mu, s2, b[12]; monitored mu, pop_mean, s2, p[12] */
  MCU_TPL_DYES = 6,      /* doc/examples/dyes.jl:22-47     nodes: s2_between, theta, s2_within, mu[6] (all monitored) */
  MCU_TPL_SALM = 7,      /* doc/examples/salm.jl:16-53     nodes: s2, gamma, beta, alpha, lambda[3x6]; monitored s2, gamma, beta, alpha */
  MCU_TPL_EQUIV = 8,     /* doc/examples/equiv.jl:25-75    nodes: s2_2, s2_1, pi, phi, mu, delta[10x2]; monitored s2_2, s2_1, pi, phi, theta, equiv, mu */
  MCU_TPL_BLOCKER = 9,   /* doc/examples/blocker.jl:22-69  nodes: s2, d, delta_new, mu[22], delta[22]; two observed nodes rc, rt; monitored s2, d, delta_new */
  MCU_TPL_STACKS = 10,   /* doc/examples/stacks.jl:41-94   nodes: beta0, beta[3], s2; Laplace likelihood; monitored (all Logical) b[3], b0, sigma, outlier[1,3,4,21] */
  MCU_TPL_MAGNESIUM = 11, /* doc/examples/magnesium.jl:21-82 nodes: priors[6], mu[6], theta[6x8], pc[6x8]; bounded (Uniform / truncated) priors: mu is sampled on
                             the two-sided link logit((x-a)/(b-a)) (src/distributions/transformdistribution.jl:6-48); monitored (Logical) tau[6], OR[6] */
  MCU_TPL_OXFORD = 12,    /* doc/examples/oxford.jl:31-82   nodes: alpha, beta1, beta2, s2, b[120], mu[120] (244 elements; AMWG / Slice / RWM blocks only) */
  MCU_TPL_EPIL = 13,      /* doc/examples/epil.jl:33-111    nodes: a0, alpha_Base, alpha_Trt, alpha_BT, alpha_Age, alpha_V4, s2_b1, s2_b, b1[59], b[59x4] (303 elements;
                             AMWG / Slice / RWM blocks only); monitored: the five coefficients, alpha0 (Logical), s2_b1, s2_b */
  MCU_N_TEMPLATES = 14
};

/* ---- sampler kinds (src/samplers/) ---------------------------------------------------------- */
enum {
  MCU_AMWG = 0,        /* src/samplers/amwg.jl:47-115  */
  MCU_SLICE_UNI = 1,   /* src/samplers/slice.jl:66-92  */
  MCU_SLICE_MULTI = 2, /* src/samplers/slice.jl:95-117 */
  MCU_RWM = 3,         /* src/samplers/rwm.jl:49-71    */
  MCU_NUTS = 4,        /* src/samplers/nuts.jl:47-205  */
  MCU_HMC = 5,         /* src/samplers/hmc.jl:47-111   */
  MCU_AMM = 6,         /* src/samplers/amm.jl:45-108   */
  MCU_MALA = 8,        /* src/samplers/mala.jl:43-86 (epsilon, optional Sigma in scale[k*k]; dtype in grad)   */
  MCU_GIBBS = 7        /* exact draw from the block's full conditional, where the template has a conjugate form — the device
                          counterpart of a user-defined Gibbs sampler Sampler([:theta], (theta, ...) -> rand(...))
                          (src/samplers/sampler.jl:20-24, doc/mcmc/sampler.rst; the tutorial's Gibbs_beta / Gibbs_s2 are of this kind).
                          pumps: nodes = [theta]: theta_i ~ Gamma(alpha + y_i, 1/(beta + t_i)); nodes = [beta]:
                          beta ~ Gamma(0.1 + N alpha, 1/(1 + sum theta)).  Gamma variates: Marsaglia-Tsang on the block's
                          Philox streams (for shape < 1: one uniform first, then per attempt normals until 1 + c x > 0 and
                          one uniform).  Other templates / node sets: MCU_ERR_UNSUPPORTED.                                    */
};

enum { MCU_ADAPT_ALL = 0, MCU_ADAPT_BURNIN = 1, MCU_ADAPT_NONE = 2 }; /* amwg.jl:47-56, amm.jl:45-55 */
enum { MCU_PROP_NORMAL = 0, MCU_PROP_SYMUNIFORM = 1, MCU_PROP_SYMTRIANGULAR = 2, MCU_PROP_COSINE = 3, MCU_PROP_EPANECHNIKOV = 4,
       MCU_PROP_BIWEIGHT = 5, MCU_PROP_TRIWEIGHT = 6 }; /* rwm.jl:12-13, distributions/extensions.jl:43-53 (SymDistributionType) */
enum { MCU_GRAD_ANALYTIC = 0, MCU_GRAD_FORWARD = 1, MCU_GRAD_CENTRAL = 2 };       /* nuts.jl:47 dtype; simulation.jl:47-51 */
enum { MCU_RNG_PHILOX = 0, MCU_RNG_EXTERNAL = 1 };
enum { MCU_ETYPE_BM = 0, MCU_ETYPE_IMSE = 1, MCU_ETYPE_IPSE = 2 };                                    /* src/output/mcse.jl:3-8 */

/* mcu_run flags */
#define MCU_RUN_NO_STORE 1u     /* do not keep thinned samples on the device, only streaming moments */
#define MCU_RUN_FORCE_GENERIC 2u /* never dispatch to a specialised (fused) kernel */
#define MCU_RUN_GLM_REFERENCE 4u /* GLM/NUTS tick engine: use the FP64 CUDA-core gradient kernel instead of the tensor-core one */
#define MCU_RUN_PARTIAL 8u       /* this call is one segment of a longer mcmc() run driven by the caller (burn-in may extend past it):
                                    skips the "burnin is greater than or equal to iters" check of mcmc.jl:22-23 */

#define MCU_RUN_ASYNC 16u        /* queue the run on the handle's stream and return (out must be NULL); complete it with mcu_wait.  One host thread can
                                    then drive a handle per GPU concurrently, as pmap2 drives a worker per chain (src/model/mcmc.jl:48-52, src/utils.jl:91-98).
                                    (The GLM / NUTS tick engine is host-driven and returns only when done.) */

#define MCU_RUN_MPSRF 32u        /* also stream the within-chain covariances of the monitored columns (Welford co-moments, p <= 12), so that the multivariate
                                    PSRF of gelmandiag(c; mpsrf = true) (src/output/gelmandiag.jl:49-55) is available without the draws (mcu_diag_global /
                                    mcu_diag_finish: mpsrf).  Off by default: p (p - 1) extra doubles are read and written per chain and kept draw.
                                    The MPSRF is NaN unless EVERY run since mcu_set_inits / mcu_set_state carried the flag. */

#define MCU_MAX_BLOCK_NODES 8

/*
 * One sampling block == one `Sampler(params, f, tune)` produced by the reference's sampler
 * constructors (src/samplers/sampler.jl:20-24).  The Julia shim fills it by inspecting the
 * Sampler object (type of `tune`, `params`, constructor arguments).
 */
typedef struct mcu_block_desc {
  int32_t kind;                       /* MCU_AMWG ...                                         */
  int32_t n_nodes;                    /* number of entries in nodes[]                         */
  int32_t nodes[MCU_MAX_BLOCK_NODES]; /* template node ids, in the order given to the sampler */
  int32_t transform;                  /* SamplingBlock(model, block, transform): AMWG/RWM/NUTS/HMC/AMM always 1; Slice: kwarg, default 0 (slice.jl:47-50) */
  int32_t adapt;                      /* MCU_ADAPT_*  (AMWG, AMM)                             */
  int32_t batchsize;                  /* AMWG, default 50 (amwg.jl:16-19); 0 → default        */
  int32_t proposal;                   /* RWM: MCU_PROP_*                                      */
  int32_t L;                          /* HMC: leapfrog steps                                  */
  int32_t grad;                       /* NUTS/HMC: MCU_GRAD_*                                 */
  int32_t max_depth;                  /* NUTS tree-depth cap; 0 → 10. (reference: unbounded, nuts.jl:106-124) */
  int32_t n_scale;                    /* entries behind `scale`: 1 (broadcast) or block dim k; AMM/HMC-with-Sigma: k*k */
  double target;                      /* AMWG 0.44 (amwg.jl:18) / NUTS 0.6 (nuts.jl:22); 0 → default */
  double epsilon;                     /* HMC step size; NUTS: <= 0 → nutsepsilon() heuristic (nuts.jl:192-205) */
  double beta;                        /* AMM mixing weight, default 0.05 (amm.jl:21); 0 → default */
  double amm_scale;                   /* AMM scale, default 2.38 (amm.jl:22); 0 → default      */
  const double* scale;                /* AMWG sigma / Slice width / RWM scale; AMM Sigma (k×k, column-major); HMC Sigma or NULL (=I) */
} mcu_block_desc;

typedef struct mcu_ctx* mcu_handle;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces deepcopy(model) + per-chain ModelState allocation: src/model/mcmc.jl:27-30.       */
int mcu_create(int template_id, int64_t n_chains, int64_t chain_offset, int device,
               uint64_t seed, mcu_handle* out);
int mcu_destroy(mcu_handle h);
/* Last error text for h (or for the failed mcu_create when h == NULL). */
const char* mcu_last_error(mcu_handle h);
int mcu_abi_version(void);

/* ---- model inputs: setinputs!  src/model/initialization.jl:30-40 --------------------------- */
/* Named input arrays of the template (e.g. "x","y" for line; "r","n","x1","x2" for seeds;
 * "y","Xm","rat" for rats; "y","t" for pumps; "X" [N×d row-major],"y" for the GLM).
 * Integer-valued inputs are passed as doubles.  Index inputs ("rat" of rats, "batch" of dyes) are 0-BASED
 * (the Julia shim subtracts 1 from the scripts' 1-based vectors) and range-checked: MCU_ERR_ARG otherwise.
 * Inputs that must agree in length (line x / y, GLM X rows / y) are checked when they are uploaded: MCU_ERR_DIM.  Every template has the reference's dataset as
 * default, so this is optional except for the GLM.                                            */
int mcu_set_data(mcu_handle h, const char* name, int ndim, const int64_t* dims, const double* ptr);

/* ---- sampling scheme: setsamplers!  src/model/initialization.jl:42-48 ---------------------- */
int mcu_set_scheme(mcu_handle h, int n_blocks, const mcu_block_desc* blocks);

/* ---- shapes and names: names(m, true)  src/model/model.jl:231-240 -------------------------- */
/* D = number of unobserved stochastic elements (state record length); n_monitor = monitored columns. */
int mcu_dims(mcu_handle h, int* D, int* n_monitor, int* n_nodes);
/* '\n'-separated names; which = 0 state elements, 1 monitored columns, 2 node names.  Returns needed length if buf too small. */
int mcu_names(mcu_handle h, int which, char* buf, size_t buflen);
/* Number of doubles of sampler tune state per chain for the current scheme (sum over blocks). */
int mcu_tune_size(mcu_handle h, int64_t* n_per_chain);

/* ---- initial values: setinits!  src/model/initialization.jl:3-28 --------------------------- */
/* x is [n_inits × D] (record contiguous, constrained scale); chain c starts from record
 * (chain_offset + c) % n_inits.  jitter_sd > 0 adds Philox N(0, jitter_sd²) noise on the
 * unconstrained scale (stream kind 1; SURVEY.md §8d config 2 "+ jitter").  Resets iter to 0. */
int mcu_set_inits(mcu_handle h, const double* x, int64_t n_inits, double jitter_sd);

/* ---- the engine: mcmc_master!/mcmc_worker!  src/model/mcmc.jl:36-83 ------------------------ */
/* Advances every chain of the handle by `iters` iterations, continuing from the handle's
 * iteration counter (0 after mcu_set_inits; restart semantics of mcmc(mc, iters), mcmc.jl:3-16).
 * Samples with iteration i > burnin and (i - burnin) % thin == 0 are kept (mcmc.jl:76-78).
 * `out` (may be NULL) receives [kept × n_monitor × n_chains], column-major, where
 * kept = number of kept iterations inside this call.                                          */
int mcu_run(mcu_handle h, int64_t iters, int64_t burnin, int64_t thin, double* out, uint32_t flags);
int64_t mcu_kept(int64_t first_iter, int64_t iters, int64_t burnin, int64_t thin);
/* Completes an MCU_RUN_ASYNC run (no-op otherwise): blocks until the handle's stream is idle; device-side failures are reported here. */
int mcu_wait(mcu_handle h);
/* ModelChains.value of the last mcu_run, [kept × n_monitor × n_chains] column-major (src/Mamba.jl:172-185), for callers that passed
 * out = NULL (asynchronous runs; pinned destination buffers: the copy then runs at full PCIe rate without a bounce buffer).        */
int mcu_get_samples(mcu_handle h, double* out);

/* ---- ModelState round trip (src/Mamba.jl:152-155; mcmc.jl:56,82) --------------------------- */
/* values [n_chains × D], tune [n_chains × tune_size] (may be NULL), iter = model.iter.        */
int mcu_get_state(mcu_handle h, double* values, double* tune, int64_t* iter);
int mcu_set_state(mcu_handle h, const double* values, const double* tune, int64_t iter);

/* ---- batched density entry points (parity surface) ----------------------------------------- */
/* logpdf!(m, x, block, transform)  src/model/simulation.jl:77-90 via src/samplers/sampler.jl:102-104.
 * state [B × D]: full model state (constrained) for each evaluation; x [B × k] block vector on
 * the sampler's scale, or NULL to use the block's own values from `state` (unlist, sampler.jl:113-115). */
int mcu_logpdf(mcu_handle h, int block, int64_t B, const double* state, const double* x, double* lp);
/* logpdf(mc::ModelChains, nodekeys)  src/output/modelstats.jl:16-58 (the kernel of dic(mc), modelstats.jl:3-13): the sum of the
 * selected stochastic nodes' log densities (constrained scale, no Jacobian) at B full states [B × D].  Bit f of factor_mask selects
 * factor f: 0 .. n_param_nodes-1 are the unobserved stochastic nodes in mcu_names(h, 2) order, n_param_nodes .. n_factors-1 the
 * observed ones (keys(m, :output)).  mcu_factor_parents gives, as a bitmask over parameter nodes, the nodes factor f reads through
 * Logical nodes (what getsimkeys, modelstats.jl:102-127, finds by walking the graph).                                              */
int mcu_factor_counts(mcu_handle h, int* n_param_nodes, int* n_factors);
int mcu_factor_parents(mcu_handle h, int factor, uint32_t* parent_nodes);
int mcu_logpdf_nodes(mcu_handle h, uint32_t factor_mask, int64_t B, const double* state, double* lp);
/* predict(mc::ModelChains, nodekeys = keys(m, :output))  src/output/modelstats.jl:63-96: one draw from the distribution of every
 * observed node element at each of B full states [B × D] → out [B × n_out] (pass out = NULL to query n_out, the length of the
 * observed node).  Philox stream (seed of the handle, chain = stream_id, iteration = record index, block 0, kind 15): a Normal
 * element takes one normal draw, a Binomial / Poisson / Bernoulli element one uniform (CDF inversion by sequential search).   */
int mcu_predict(mcu_handle h, int64_t B, const double* state, uint32_t stream_id, double* out, int64_t* n_out);
/* logpdfgrad!(block, x, dtype)  src/samplers/sampler.jl:106-111 (+ analytic mode).  g [B × k]. */
int mcu_gradlogpdf(mcu_handle h, int block, int grad_mode, int64_t B, const double* state,
                   const double* x, double* lp, double* g);

/* Likelihood part of the GLM block density for ALL chains of the handle in one pass over X (the kernels behind the
 * GLM/NUTS tick engine): beta [n_chains × d] → lp [n_chains] = Σ_i log Bernoulli(y_i; invlogit(x_i·beta)),
 * grad [n_chains × d] = X'(y − p).  impl 0 = FP64 CUDA-core reference kernel, 1 = fused tcgen05 tensor-core kernel
 * (split-fp16 operands, FP32 accumulation; agrees with impl 0 to ~1e-6 relative).  Device time of the pass is
 * then available from mcu_last_kernel_ms.  (Replaces d+2 interpreted model evaluations per gradient:
 * src/model/simulation.jl:47-51, src/samplers/sampler.jl:106-111.)                                        */
int mcu_glm_gradient(mcu_handle h, int impl, const double* beta, double* lp, double* grad);

/* ---- diagnostics on the way out ------------------------------------------------------------- */
/* All streaming statistics cover the samples kept since the last mcu_set_inits.  They are held
 * per chain on the device (Welford mean/M2 on the raw and log scale, min/max, batch means of
 * size 100) and reduced across chains by reduction kernels; what crosses the ABI is O(p).
 *
 * link(c) for gelmandiag(transform=true): src/output/modelchains.jl:57-76, src/output/chains.jl:237-246.
 * codes[p]: 0 identity, 1 log, 2 logit.  Monitored stochastic columns use their node's link; Logical
 * columns use the reference's data-dependent heuristic (log when every value is > 0, logit when every
 * value is also < 1 — logit moments are kept for Logical columns among the first 64), resolved
 * from minmax[p×2] (pass the all-reduced min/max for multi-GPU, or NULL to use this handle's).  */
int mcu_minmax(mcu_handle h, double* minmax /* [p × 2] */);
int mcu_link_codes(mcu_handle h, int transform, const double* minmax, int* codes);

/* Cross-chain sums for gelmandiag (src/output/gelmandiag.jl:12-29).  With psibar_c / s2_c a chain's
 * mean / variance of column j (on the scale given by codes, NULL = identity) and centres
 * center[j] = (c1, c2) (NULL = 0):  d = psibar_c - c1, e = s2_c - c2,
 *   sums[j][7] = { m, Σd, Σd², Σe, Σe², Σe·d, Σe·d² }   (Σ over this handle's chains).
 * These 7·p doubles are the ONLY thing all-reduced across GPUs.  Two passes (first with
 * center = NULL to get the means, then centred) keep the variances free of cancellation.       */
int mcu_moments(mcu_handle h, const int* codes, const double* center, double* sums, int64_t* n_kept);
/* PSRF and its upper confidence limit from (all-reduced) centred sums: gelmandiag.jl:20-47.
 * psrf [p × 2] row-major, NOT rounded (the reference rounds to 3 dp at gelmandiag.jl:59).        */
int mcu_gelman_from_moments(int64_t n_kept, int p, const double* center, const double* sums,
                            double alpha, double* psrf);
/* Convenience single-handle gelmandiag(c; alpha, transform).                                   */
int mcu_gelman(mcu_handle h, double alpha, int transform, double* psrf);

/* summarystats(c; etype)  src/output/stats.jl:85-94, src/output/mcse.jl:10-33 over the samples stored
 * by the last mcu_run (needs them: no MCU_RUN_NO_STORE): out [p × 5] = mean, SD, naive SE, MCSE, ESS. */
int mcu_summarystats(mcu_handle h, int etype, int batch_size, double* out);
/* Streaming form for chain counts too large to store: per-column sums
 *   sums[j][8] = { C, Σ mean_c, Σ M2_c, Σ (mean_c - c1)², Σ nb_c, Σ nb_c·bmean_c, Σ bM2_c, Σ nb_c (bmean_c - c2)² }
 * (center[j] = (c1, c2) or NULL), all-reducible like mcu_moments; and the single-handle
 * convenience that returns [p × 5] = mean, SD, naive SE, MCSE (batch means of 100, batches never
 * straddle chains), ESS = min((SD/MCSE)², kept per chain) (stats.jl:92).                        */
int mcu_summary_sums(mcu_handle h, const double* center, double* sums);
int mcu_summary_from_sums(int64_t n_kept, int p, const double* center, const double* sums, double* out);
int mcu_summary_streaming(mcu_handle h, double* out);

/* ---- post-processing of a materialised ModelChains.value (host arrays; no handle, no device) ----------------------
 * value: [n iterations × p parameters × m chains], column-major (iteration fastest) — the array mcu_run returns.
 * The Julia shim keeps describe() / gelmandiag(mpsrf = true) / hpd / autocor / changerate on engine output.
 *   mcu_chains_quantile   quantile(c; q)        src/output/stats.jl:74-83     out [p × nq] row-major
 *   mcu_chains_hpd        hpd(c; alpha)         src/output/stats.jl:52-72     out [p × 2]
 *   mcu_chains_autocor    autocor(c; lags, relative)  stats.jl:3-13; lags = index lags on the stored series (× thinning step when
 *                         relative = true, exactly as the reference does); out [p × nlags × m] column-major
 *   mcu_chains_changerate changerate(c)         stats.jl:19-39                out [p + 1] (last = multivariate), NOT rounded
 *   mcu_chains_summarystats  summarystats(c; etype)  src/output/stats.jl:85-94, src/output/mcse.jl:3-46   out [p × 5] row-major: mean, SD, naive SE,
 *                         MCSE, ESS; MCU_ERR_ARG where mcse_bm throws (fewer than 2 batches)
 *   mcu_chains_gelman     gelmandiag(c; alpha, mpsrf, transform)  src/output/gelmandiag.jl:3-60; codes[j] = 1 → log scale, 2 → logit scale (link(c));
 *                         out [(p + mpsrf) × 2] row-major, NOT rounded; multivariate row = (MPSRF, NaN); returns MCU_ERR_ARG for m < 2 */
int mcu_chains_quantile(const double* value, int64_t n, int p, int64_t m, const double* q, int nq, double* out);
int mcu_chains_hpd(const double* value, int64_t n, int p, int64_t m, double alpha, double* out);
int mcu_chains_autocor(const double* value, int64_t n, int p, int64_t m, const int64_t* lags, int nlags, double* out);
int mcu_chains_changerate(const double* value, int64_t n, int p, int64_t m, double* out);
int mcu_chains_gelman(const double* value, int64_t n, int p, int64_t m, double alpha, const int* codes, int mpsrf, double* out);
int mcu_chains_summarystats(const double* value, int64_t n, int p, int64_t m, int etype, int batch_size, double* out);
/* Per-series convergence diagnostics, one row per (parameter, chain); out [p × K × m] column-major, NOT rounded; etype: MCU_ETYPE_*
 * (batch_size only for MCU_ETYPE_BM).  Return MCU_ERR_ARG where the reference throws (window fractions, too few iterations for mcse_bm).
 *   mcu_chains_geweke   gewekediag(c; first, last, etype)        src/output/gewekediag.jl:3-31   K = 2: Z-score, p-value
 *   mcu_chains_heidel   heideldiag(c; alpha, eps, etype)          src/output/heideldiag.jl:3-41   K = 6: burn-in, stationarity, p-value, mean, halfwidth, test
 *                       (start = first(c.range), heideldiag.jl:33)
 *   mcu_chains_raftery  rafterydiag(c; q, r, s, eps)              src/output/rafterydiag.jl:3-61  K = 5: thinning, burn-in, total, nmin, dependence factor
 *                       (range_start, range_step = first(c.range), step(c.range))                                                         */
int mcu_chains_geweke(const double* value, int64_t n, int p, int64_t m, double first, double last, int etype, int batch_size, double* out);
int mcu_chains_heidel(const double* value, int64_t n, int p, int64_t m, double alpha, double eps, int etype, int batch_size, int64_t start, double* out);
int mcu_chains_raftery(const double* value, int64_t n, int p, int64_t m, double q, double r, double s, double eps, int64_t range_start, int64_t range_step, double* out);

/* ---- diagnostics over chains that live on several handles / GPUs -------------------------------------
 * The reference farms chains out to workers and gathers every sample (pmap2(mcmc_worker!, lsts), src/model/mcmc.jl:48-59) before
 * gelmandiag (src/output/gelmandiag.jl:3-60) and summarystats (src/output/stats.jl:85-94) run on the gathered array.  Here every
 * handle reduces its own chains on the device and a packed two-round protocol carries O(p) doubles between handles
 * (mamba.jl_b200/csrc/diagproto.hpp):
 *   round 1 buffer [min p | max p | sum 9p]  — all-reduce the three parts with MIN / MAX / SUM;
 *   round 2 buffer [15p + p(p-1) for p <= 12] — all-reduce with SUM (centred gelman sums + centred summary sums per column, then two sums per
 *                                              column pair for the multivariate PSRF, gelmandiag.jl:49-55).
 * ANY transport can carry the buffers (mcu_diag_round1 → reduce → mcu_diag_round2 → reduce → mcu_diag_finish: Julia worker messaging,
 * MPI, torch.distributed); the built-in transport is NCCL over NVLink: mcu_comm_unique_id on rank 0, the 128-byte id handed to the other
 * ranks by the host language, mcu_comm_init on every rank's handle, then mcu_diag_global does both rounds on the device (reductions,
 * all-reduces and the plan kernel queued back to back on the handle's stream, one synchronisation).  NCCL is bound at run time
 * (dlopen of libnccl.so.2, or $MCU_NCCL_LIB): a process that never asks for a communicator does not need it.
 * Without a communicator mcu_diag_global covers this handle's chains (one call instead of mcu_gelman + mcu_summary_streaming).      */
typedef struct mcu_nccl_id { char internal[128]; } mcu_nccl_id;     /* == ncclUniqueId */
int mcu_diag_sizes(int p, int* n_round1, int* n_round2);
int mcu_monitor_links(mcu_handle h, int* monlink /* [p]: 0 identity, 1 log, -1 Logical column (data-dependent heuristic) */);
int mcu_n_kept(mcu_handle h, int64_t* n_kept);
int mcu_diag_round1(mcu_handle h, double* buf /* [11p] */);
int mcu_diag_round2(mcu_handle h, int transform, const double* reduced1 /* [11p] */, double* buf2 /* [n_round2 of mcu_diag_sizes] */);
/* psrf [p × 2] (not rounded), summary [p × 5] = mean, SD, naive SE, MCSE (batch means of 100), ESS; codes [p] = link code used;
 * mpsrf = the multivariate PSRF of gelmandiag(c; mpsrf = true) from the streamed within-chain covariances (NaN when p > 12, when W is not positive
 * definite, or when a Logical column's link is resolved by the data-dependent heuristic to log / logit: those need the stored draws, mcu_chains_gelman);
 * each output may be NULL */
int mcu_diag_finish(int64_t n_kept, int p, double alpha, const int* monlink, int transform, const double* reduced1,
                    const double* reduced2, double* psrf, double* summary, int* codes, double* mpsrf);
int mcu_comm_unique_id(mcu_nccl_id* id);
int mcu_comm_init(mcu_handle h, int rank, int nranks, const mcu_nccl_id* id);
int mcu_comm_size(mcu_handle h, int* rank, int* nranks);
int mcu_diag_global(mcu_handle h, double alpha, int transform, double* psrf, double* summary, int* codes, double* mpsrf);

/* ---- RNG contract ----------------------------------------------------- */
/* PHILOX: Philox4x32-10, key = seed, counter = (k >> 1, iteration, global chain, block | kind << 16 | stream << 24).
 * Every block update of every chain owns two streams (0 = rand(), 1 = randn()); k counts the draws of a stream in
 * the order the reference consumes them; one Philox block yields two draws (uniforms: words (0,1) / (2,3) as 53-bit
 * fractions; normals: both Box-Muller branches rad·cos / rad·sin).  Full definition: mamba.jl_b200/csrc/rng.cuh.
 * EXTERNAL: the shim stream of north_star — draws are consumed sequentially from
 * u[chain][0..n_per_chain) (uniforms in [0,1)); a normal consumes two (cosine branch).            */
int mcu_set_rng_mode(mcu_handle h, int mode, const double* u, size_t n_per_chain);

/* ---- device info for the harness ------------------------------------------------------------ */
int mcu_device_count(void);
/* Measured FP64 FMA throughput of the handle's device (DFMA microbenchmark, TFLOP/s): the roofline
 * denominator for the CUDA-core small-model kernels, which MEASURED_PEAKS.json does not carry.       */
double mcu_fp64_peak_tflops(mcu_handle h);
/* Number of kernel launches issued by this handle since creation (bench "gpu_launches").      */
int64_t mcu_launch_count(mcu_handle h);
/* Work counters of the gradient-based fused paths since the handle was created, out[4]:
 *   [0] gradient evaluations: leapfrog steps of the warp-per-chain rats kernel / chain-gradients the GLM tick engine actually consumed
 *   [1] GLM ticks (gradient passes)   [2] chain slots those passes carried in total ([0] / [2] = useful fraction of the tensor-core work; a
 *       finished chain idles until its 128-chain group is compacted away)
 *   [3] NUTS iterations that stopped at the tree-depth cap (max_depth doublings; the reference has no cap, src/samplers/nuts.jl:106-124)  */
int mcu_work_count(mcu_handle h, uint64_t* out);
/* Device time in ms of the sampler kernels of the last mcu_run (CUDA events on the launching stream). */
double mcu_last_kernel_ms(mcu_handle h);

#ifdef __cplusplus
}
#endif
#endif /* MAMBACUDA_H */
