// engine.cuh — the generic one-chain-per-thread engine kernels, templated on a model template.
//
// Replaces the reference's per-chain interpreter loop mcmc_worker! → sample!(m) → sampler.eval
// (src/model/mcmc.jl:62-83, src/model/simulation.jl:93-107) by kernels that advance all chains of a
// handle in lockstep.  Memory layout in HBM (structure of arrays, chain fastest ⇒ every warp-wide
// access is one coalesced 256-byte request):
//   state   [D][C]        constrained values of the unobserved stochastic elements
//   tune    [T][C]        sampler tune slots of every block (layout in samplers.cuh)
//   samples [kept][P][C]  thinned monitored values (Chains storage, src/output/chains.jl:5-32)
//   mom     [P*9][C]      streaming per-chain moments for gelmandiag / ESS
//   momn    [3][C]        kept count, in-batch count, number of complete batches
#pragma once
#include "samplers.cuh"

namespace mcu {

constexpr int kMomPerCol = 11;  // mean, M2, lmean, lM2, min, max, bsum, bmean, bM2, logit mean, logit M2 (Logical columns only)
constexpr int kBatch = 100;     // mcse_bm default batch size (src/output/mcse.jl:10)
constexpr int kCoMaxP = 12;     // streaming within-chain covariances (for the multivariate PSRF, gelmandiag.jl:49-55) are kept for up to 12 monitored columns

// logpdf!(block, x) / logpdfgrad!(block, x) for one chain: relist x into the state record
// (invlink when the block samples on the transformed scale), then sum the block's own prior
// factors that are not targets followed by the target factors in topological order, stopping at
// the first non-finite partial sum (src/model/simulation.jl:60-67,77-90).
template <class M>
struct BlockTarget {
  const typename M::Data& d;
  double* s;
  const DevBlock& b;

  // invlink / link of block element i (transformdistribution.jl:6-32): identity, exp / log, or the two-sided (b - a) invlogit(x) + a /
  // logit((v - a) / (b - a)) with invlogit, logit as in src/utils.jl:64-68
  MCU_D double inv(int i, double x) const {
    if (!b.transform) return x;
    const int lk = b.elink[i];
    if (lk == LINK_LOG) return exp(x);
    if (lk == LINK_BOUNDED) { const double lo = b.ebound[2 * i], hi = b.ebound[2 * i + 1]; return (hi - lo) * (1.0 / (exp(-x) + 1.0)) + lo; }
    return x;
  }
  MCU_D double fwd(int i, double v) const {
    if (!b.transform) return v;
    const int lk = b.elink[i];
    if (lk == LINK_LOG) return log(v);
    if (lk == LINK_BOUNDED) { const double lo = b.ebound[2 * i], hi = b.ebound[2 * i + 1]; const double p = (v - lo) / (hi - lo); return log(p / (1.0 - p)); }
    return v;
  }
  MCU_D void relist(const double* x) const { for (int i = 0; i < b.k; ++i) s[b.elem[i]] = inv(i, x[i]); }
  MCU_D void unlist(double* x) const { for (int i = 0; i < b.k; ++i) x[i] = fwd(i, s[b.elem[i]]); }
  MCU_NOINL double eval() const {
    double lp = 0.0;
    const bool tr = b.transform != 0;
    for (int o = 0; o < b.n_own; ++o) {
      const int f = b.own[o];
      if (M::parents(f) & b.mask) continue;        // it is a target: handled below
      lp += M::factor(d, s, f, tr);
      if (!isfinite(lp)) return lp;
    }
    for (int f = 0; f < M::NF; ++f) {
      if (!(M::parents(f) & b.mask)) continue;
      if (!isfinite(lp)) break;
      const bool own = f < M::NN && ((b.mask >> f) & 1u);
      lp += M::factor(d, s, f, tr && own);
    }
    return lp;
  }
  MCU_D double logf(const double* x) const { relist(x); return eval(); }
  MCU_D void put(int i, double x) const { s[b.elem[i]] = inv(i, x); }
  // logpdf!(block, v) when v differs in component i only from the vector evaluated last (whose value is `cur` and which the state record
  // still holds).  Where the template marks the element as local, only the terms that read it are re-evaluated:
  //   logf(v') = logf(v) - terms(v) + terms(v')   — the reference re-evaluates the whole block (amwg.jl:100-106, slice.jl:80-88); every
  // term that reads the element is the element's own prior term or belongs to a target of the block, so nothing else changes.
  MCU_D double logf_comp(const double* v, int i, double cur) const {
    const int e = b.elem[i];
    if (M::elem_local(e) && isfinite(cur)) {
      const bool tr = b.transform != 0;
      const double t_old = M::elem_terms(d, s, e, tr);
      s[e] = inv(i, v[i]);
      const double t_new = M::elem_terms(d, s, e, tr);
      return isnan(t_new) ? neg_inf() : (cur - t_old) + t_new;
    }
    return logf(v);
  }
  // The same for a sequence of candidates of component i around ONE base vector (univariate slice: slice.jl:80-88): terms_at(i) is taken
  // once at the base, every candidate costs one evaluation of the element's own terms.  A candidate outside the support gives -Inf
  // without disturbing the base, so the shrinkage loop never falls back to full block evaluations (it did: a Uniform(0, 1) element
  // whose interval reaches below 0 made every later evaluation of that coordinate a full one).
  MCU_D bool comp_is_local(int i, double base_lp) const { return M::elem_local(b.elem[i]) && isfinite(base_lp); }
  MCU_D double terms_at(int i) const { return M::elem_terms(d, s, b.elem[i], b.transform != 0); }
  MCU_D double logf_comp_base(const double* v, int i, double base_lp, double t_base) const {
    s[b.elem[i]] = inv(i, v[i]);
    const double t_new = M::elem_terms(d, s, b.elem[i], b.transform != 0);
    return isnan(t_new) ? neg_inf() : (base_lp - t_base) + t_new;
  }
  MCU_NOINL void grad_analytic(const double* x, double* g) const {
    relist(x);
    double gj[M::D];
    M::joint_grad(d, s, gj);
    for (int i = 0; i < b.k; ++i) {   // theta = invlink(x): d/dx [lp(theta) + log |dtheta/dx|] = dlp/dtheta * dtheta/dx + d(log-Jacobian)/dx
      const int e = b.elem[i];
      const int lk = b.transform ? b.elink[i] : LINK_IDENT;
      if (lk == LINK_LOG) g[i] = gj[e] * s[e] + 1.0;
      else if (lk == LINK_BOUNDED) {
        const double lo = b.ebound[2 * i], hi = b.ebound[2 * i + 1];
        g[i] = gj[e] * ((s[e] - lo) * (hi - s[e]) / (hi - lo)) + ((hi - s[e]) - (s[e] - lo)) / (hi - lo);
      } else g[i] = gj[e];
    }
  }
  // Calculus.gradient(f, x, :forward / :central): src/model/simulation.jl:47-51
  MCU_NOINL void grad_fd(const double* x0, double* g, int mode) const {
    double x[M::D];
    for (int i = 0; i < b.k; ++i) x[i] = x0[i];
    const double EPS = 2.220446049250313e-16;
    if (mode == 1) {
      const double f0 = logf(x);
      for (int i = 0; i < b.k; ++i) {
        const double h = sqrt(EPS) * fmax(1.0, fabs(x[i]));
        const double old = x[i]; x[i] = old + h;
        g[i] = (logf(x) - f0) / h; x[i] = old;
      }
    } else {
      for (int i = 0; i < b.k; ++i) {
        const double h = cbrt(EPS) * fmax(1.0, fabs(x[i]));
        const double old = x[i];
        x[i] = old + h; const double f1 = logf(x);
        x[i] = old - h; const double f2 = logf(x);
        g[i] = (f1 - f2) / (2.0 * h); x[i] = old;
      }
    }
  }
  MCU_NOINL double logfgrad_mode(const double* x, double* g, int mode) const {   // sampler.jl:106-111
    if (mode == 0) grad_analytic(x, g); else grad_fd(x, g, mode);
    const double lf = logf(x);
    for (int i = 0; i < b.k; ++i) if (!isfinite(g[i])) g[i] = 0.0;
    return lf;
  }
  MCU_D double logfgrad(const double* x, double* g) const { return logfgrad_mode(x, g, b.grad); }
};

struct RunArgs {
  long long n_chains, chain_offset;
  unsigned long long seed;
  long long iter0, iters, burnin, thin;   // this launch advances iterations iter0+1 .. iter0+iters
  long long row0;                         // kept rows that precede this mcu_run call (global row - row0 = row in `samples`)
  int n_blocks, D, P;
  const DevBlock* blocks;
  double* state; double* tune; double* samples; double* mom; double* momn;
  const double* ext_u; unsigned long long ext_n; unsigned long long* ext_pos;
  unsigned long long* work;               // device counter of gradient evaluations (leapfrogs) the kernels add to, or nullptr
  double* comom;                          // [2][P (P - 1) / 2][C] streaming within-chain co-moments (raw scale | node-link scale), or nullptr
  unsigned long long log_mask;            // monitored columns (bit j) whose node link is the log (co-moment set 1 uses log x for them)
  unsigned long long logit_mask;          // monitored columns (bit j) whose link(c) may be the logit: Logical nodes in (0, 1), chains.jl:237-246
};

static MCU_NOINL void moments_update(double* mom, double* momn, size_t C, size_t c, int P, const double* mon, unsigned long long logit_mask = 0ull) {
  const double n = momn[0 * C + c] + 1.0; momn[0 * C + c] = n;
  double bc = momn[1 * C + c] + 1.0;
  const bool bdone = bc >= (double)kBatch;
  double nb = momn[2 * C + c];
  if (bdone) { bc = 0.0; nb += 1.0; momn[2 * C + c] = nb; }
  momn[1 * C + c] = bc;
  for (int j = 0; j < P; ++j) {
    double* q = mom + (size_t)j * kMomPerCol * C + c;
    const double x = mon[j];
    double mean = q[0 * C], M2 = q[1 * C];
    double dl = x - mean; mean += dl / n; M2 += dl * (x - mean);
    q[0 * C] = mean; q[1 * C] = M2;
    const double lx = log(x);
    double lmean = q[2 * C], lM2 = q[3 * C];
    dl = lx - lmean; lmean += dl / n; lM2 += dl * (lx - lmean);
    q[2 * C] = lmean; q[3 * C] = lM2;
    q[4 * C] = n == 1.0 ? x : fmin(q[4 * C], x);
    q[5 * C] = n == 1.0 ? x : fmax(q[5 * C], x);
    if (j < 64 && ((logit_mask >> j) & 1ull)) {
      const double gx = log(x / (1.0 - x));   // logit: src/utils.jl:66
      double gmean = q[9 * C], gM2 = q[10 * C];
      dl = gx - gmean; gmean += dl / n; gM2 += dl * (gx - gmean);
      q[9 * C] = gmean; q[10 * C] = gM2;
    }
    double bsum = q[6 * C] + x;
    if (bdone) {
      const double bm = bsum / (double)kBatch; bsum = 0.0;
      double bmean = q[7 * C], bM2 = q[8 * C];
      const double db = bm - bmean; bmean += db / nb; bM2 += db * (bm - bmean);
      q[7 * C] = bmean; q[8 * C] = bM2;
    }
    q[6 * C] = bsum;
  }
}

// Streaming within-chain co-moments for the multivariate PSRF (MCU_RUN_MPSRF; gelmandiag.jl:49-55): Welford update
// C_ij += (x_i - mean_i^old)(x_j - mean_j^new) on the raw scale (set 0) and on the nodes' own link scale (set 1: log x for log_mask columns).
// Called BEFORE moments_update (it reads the old means) and only when the run asked for it: a separate function so that the default path
// keeps the round-1 code and register allocation (folding it into moments_update cost the fused kernels 3-15 % even when switched off).
static MCU_NOINL void comoments_update(const double* mom, const double* momn, double* comom, size_t C, size_t c, int P, const double* mon,
                                       unsigned long long log_mask) {
  if (P > kCoMaxP || P < 2) return;
  const double n = momn[0 * C + c] + 1.0;
  double d_old[kCoMaxP], r_new[kCoMaxP], dl_old[kCoMaxP], rl_new[kCoMaxP];
  for (int j = 0; j < P; ++j) {
    const double* q = mom + (size_t)j * kMomPerCol * C + c;
    const double x = mon[j];
    const double dl = x - q[0 * C];
    d_old[j] = dl; r_new[j] = x - (q[0 * C] + dl / n);
    if ((log_mask >> j) & 1ull) { const double lx = log(x), dll = lx - q[2 * C]; dl_old[j] = dll; rl_new[j] = lx - (q[2 * C] + dll / n); }
    else { dl_old[j] = d_old[j]; rl_new[j] = r_new[j]; }
  }
  const size_t npair = (size_t)P * (P - 1) / 2;
  size_t k = 0;
  for (int i = 0; i < P; ++i)
    for (int j = i + 1; j < P; ++j, ++k) {
      comom[k * C + c] += d_old[i] * r_new[j];
      comom[(npair + k) * C + c] += dl_old[i] * rl_new[j];
    }
}

// The body of the generic kernel; instantiated behind two __global__ wrappers with different launch bounds (below).
template <class M>
__device__ __forceinline__ void generic_kernel_body(const typename M::Data& data, const RunArgs& a) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_chains) return;
  const size_t C = (size_t)a.n_chains;
  double s[M::D];
  for (int e = 0; e < a.D; ++e) s[e] = a.state[(size_t)e * C + c];
  Draws rng;
  rng.k0 = (uint32_t)a.seed; rng.k1 = (uint32_t)(a.seed >> 32);
  rng.chain = (uint32_t)(a.chain_offset + c);
  rng.ext = a.ext_u ? a.ext_u + (size_t)c * a.ext_n : nullptr;
  rng.ext_n = a.ext_n; rng.ext_pos = a.ext_pos ? a.ext_pos + c : nullptr;
  double v[M::D];
  double mon[M::P];
  for (long long it = 1; it <= a.iters; ++it) {
    const long long iter = a.iter0 + it;                       // m.iter += 1: simulation.jl:94
    for (int bi = 0; bi < a.n_blocks; ++bi) {
      const DevBlock& b = a.blocks[bi];
      BlockTarget<M> tgt{data, s, b};
      TuneRef tn{a.tune + (size_t)b.tune_off * C + c, C};
      rng.seek((uint32_t)iter, (uint32_t)bi, 0);
      tgt.unlist(v);                                           // unlist(block): sampler.jl:113-115
      const bool fresh = iter == 1;                            // sampler.jl:40-45
      const bool isadapt = b.adapt == 1 ? iter <= a.burnin : b.adapt == 0;
      switch (b.kind) {
        case 0: amwg_sample<M::D>(v, b, tn, tgt, rng, fresh, isadapt); break;
        case 1: slice_uni_sample<M::D>(v, b, tgt, rng); break;
        case 2: slice_multi_sample<M::D>(v, b, tgt, rng); break;
        case 3: rwm_sample<M::D>(v, b, tgt, rng); break;
        case 4: if constexpr (M::kGradSamplers) nuts_sample<M::D>(v, b, tn, tgt, rng, fresh, iter <= a.burnin); break;
        case 5: if constexpr (M::kGradSamplers) hmc_sample<M::D>(v, b, tgt, rng); break;
        case 6: if constexpr (M::kGradSamplers) amm_sample<M::D>(v, b, tn, tgt, rng, fresh, isadapt); break;
        case 8: if constexpr (M::kGradSamplers) mala_sample<M::D>(v, b, tgt, rng); break;
        case 7: M::gibbs(data, s, b.own[0], rng, [](double shape, Draws& r) { return rgamma_mt(shape, r); }); tgt.unlist(v); break;   // MCU_GIBBS
      }
      tgt.relist(v);                                           // m[sampler.params] = relist(block, v)
    }
    if (iter > a.burnin && (iter - a.burnin) % a.thin == 0) {  // mcmc.jl:76-78
      M::monitor(data, s, mon);
      if (a.samples) {
        const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;   // iters2inds: src/output/chains.jl:66-87
        for (int j = 0; j < a.P; ++j) a.samples[((size_t)row * a.P + j) * C + c] = mon[j];
      }
      if (a.comom) comoments_update(a.mom, a.momn, a.comom, C, (size_t)c, a.P, mon, a.log_mask);
      moments_update(a.mom, a.momn, C, (size_t)c, a.P, mon, a.logit_mask);
    }
  }
  for (int e = 0; e < a.D; ++e) a.state[(size_t)e * C + c] = s[e];
}

// Two instantiations per template: ptxas' own register choice (128) for chain counts that do not fill the device — the latency of a
// single chain matters there and nothing spills — and the template's tuned MCU_GENERIC_MINB (tpl_*.cu: resident blocks per SM the
// register allocation aims for) for large chain counts (profiles/r1_generic_kernel_occupancy.md).
template <class M>
__global__ void __launch_bounds__(128) run_generic_kernel(typename M::Data data, RunArgs a) { generic_kernel_body<M>(data, a); }
#ifdef MCU_GENERIC_MINB
template <class M>
__global__ void __launch_bounds__(128, MCU_GENERIC_MINB) run_generic_kernel_dense(typename M::Data data, RunArgs a) { generic_kernel_body<M>(data, a); }
#endif

// Batched density entry points (mcu_logpdf / mcu_gradlogpdf): one evaluation per thread.
// state [D][B] chain-fastest; x [k][B] or nullptr (use unlist of state); lp [B]; g [k][B] or nullptr.
template <class M>
__global__ void logpdf_kernel(typename M::Data data, const DevBlock* blocks, int block, long long B, int D,
                              const double* state, const double* x, double* lp, double* g, int grad_mode) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B) return;
  const DevBlock& b = blocks[block];
  double s[M::D], v[M::D], gg[M::D];
  for (int e = 0; e < D; ++e) s[e] = state[(size_t)e * B + c];
  BlockTarget<M> tgt{data, s, b};
  if (x) for (int i = 0; i < b.k; ++i) v[i] = x[(size_t)i * B + c]; else tgt.unlist(v);
  if (g) {
    const double l = tgt.logfgrad_mode(v, gg, grad_mode);
    if (lp) lp[c] = l;
    for (int i = 0; i < b.k; ++i) g[(size_t)i * B + c] = gg[i];
  } else {
    lp[c] = tgt.logf(v);
  }
}


// rand(m[key]) for an observed node element (predict, src/output/modelstats.jl:63-96).  The reference calls Distributions.rand on
// the global RNG; here a Normal takes one draw of the normal stream, the discrete families ONE uniform and invert the CDF by
// sequential search from 0 (pmf recurrences; exact in the same arithmetic as the oracle's, so the integer draws coincide).
static MCU_NOINL double rand_out(int kind, double a, double b, Draws& rng) {
  if (kind == OUT_NORMAL) return a + b * rng.normal();
  const double u = rng.uniform();
  if (kind == OUT_BERNOULLI) return u < a ? 1.0 : 0.0;
  if (kind == OUT_LAPLACE) { const double c = u - 0.5; return a - b * (c < 0 ? -1.0 : 1.0) * log(1.0 - 2.0 * fabs(c)); }   // inverse CDF
  if (kind == OUT_BINOMIAL) {
    const double n = a, p = b, q = 1.0 - p;
    if (!(p > 0.0)) return 0.0;
    if (!(q > 0.0)) return n;
    const double ratio = p / q;
    double pmf = exp(n * log(q)), cdf = pmf, k = 0.0;
    while (u >= cdf && k < n) { pmf *= (n - k) / (k + 1.0) * ratio; k += 1.0; cdf += pmf; }
    return k;
  }
  double pmf = exp(-a), cdf = pmf, k = 0.0;   // Poisson(a)
  while (u >= cdf && k < 100000.0) { k += 1.0; pmf *= a / k; cdf += pmf; }
  return k;
}

// predict(mc, nodekeys): one draw of every observed element at each of B state records; stream = (seed, chain = stream_id,
// iteration = record index, block 0, kind 15), element i takes the next draw(s) of the record's stream.  out [L][B].
template <class M>
__global__ void predict_kernel(typename M::Data data, long long B, int D, const double* state, unsigned long long seed, unsigned stream_id,
                               double* out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B) return;
  double s[M::D];
  for (int e = 0; e < D; ++e) s[e] = state[(size_t)e * B + c];
  Draws rng;
  rng.k0 = (uint32_t)seed; rng.k1 = (uint32_t)(seed >> 32); rng.chain = stream_id; rng.ext = nullptr; rng.ext_pos = nullptr; rng.ext_n = 0;
  rng.seek((uint32_t)c, 0, 15);
  const int L = M::out_len(data);
  for (int i = 0; i < L; ++i) {
    double a, b;
    const int kind = M::out_dist(data, s, i, a, b);
    out[(size_t)i * B + c] = rand_out(kind, a, b, rng);
  }
}

// logpdf(mc, nodekeys) (src/output/modelstats.jl:16-58): sum of the selected factors (node densities on the constrained scale) at B
// states, one per thread; bit f of `mask` = factor f (0 .. NN-1 parameter nodes, NN .. NF-1 observed nodes, i.e. keys(m, :output)).
template <class M>
__global__ void factors_kernel(typename M::Data data, unsigned mask, long long B, int D, const double* state, double* lp) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B) return;
  double s[M::D];
  for (int e = 0; e < D; ++e) s[e] = state[(size_t)e * B + c];
  double sum = 0.0;
  for (int f = 0; f < M::NF; ++f) if ((mask >> f) & 1u) sum += M::factor(data, s, f, false);
  lp[c] = sum;
}

}  // namespace mcu
