// pumps_fast.cu — fused kernel for the reference's own sampling scheme of the `pumps` model (doc/examples/pumps.jl:52-53):
//     [Slice([alpha, beta], 1.0, Univariate), Slice(theta, 1.0, Univariate)]          (constrained scale: slice.jl:47-50)
// One chain per thread, every iteration of an mcu_run call inside ONE launch (as seeds_fast.cu / rats_fast.cu).
//
// What the reference does per shrinkage step: a full block-density evaluation (10 Gamma + 10 Poisson terms, src/model/simulation.jl:77-90).
// What this kernel does:
//   * block (alpha, beta): the ten Gamma(theta_i | alpha, 1/beta) terms enter through SL = sum log theta_i and ST = sum theta_i:
//       logf(alpha, beta) = -alpha + (0.1 - 1) log beta - beta + 10 (alpha log beta - lgamma(alpha)) + (alpha - 1) SL - beta ST + const;
//   * block theta: a move of theta_i changes two terms only,
//       g_i(t) = (alpha - 1 + y_i) log t - t (beta + t_i)      (Gamma prior term + Poisson(y_i | t t_i) term, constants dropped),
//     and the slice test logf(v') >= logf0 + log u is taken on the difference g_i(t') - g_i(t) + (running difference), so an evaluation is
//     one log;
//   * slice.jl:70-71 places ALL interval ends before any coordinate moves (lower = v - width .* rand(n)): with a counter-based stream
//     draw i can be regenerated when component i is reached, so no per-thread arrays are needed.
// Decisions are those of the reference on the same Philox stream up to rounding of the comparisons (tests/test_gpu_parity.py).
#include "fastmath.cuh"

namespace mcu {

namespace {

constexpr int NP = PumpsModel::NPUMP;   // 10 pumps

struct PumpsFastCfg {
  double y[NP], t[NP];
  double w_ab[2], w_th[NP];
};

struct UStreamP {
  const RunArgs& a; uint32_t chain, iter, block, k; Pair cur;
  MCU_D double next() {
    if (!(k & 1u)) cur = draw_uniform_pair(a, chain, iter, block, k >> 1);
    const double u = (k & 1u) ? cur.b : cur.a;
    ++k;
    return u;
  }
  MCU_D double at(uint32_t idx) const {   // random access to draw idx of the block (does not move the cursor)
    const Pair p = draw_uniform_pair(a, chain, iter, block, idx >> 1);
    return (idx & 1u) ? p.b : p.a;
  }
};

#ifndef MCU_PUMPSF_MINB
#define MCU_PUMPSF_MINB 6    // resident blocks per SM (measured 3 / 4 / 5 / 6 / 8 / 10 / 12 / 16: 1.57 / 1.57 / 1.69 / 1.70 / 1.63 / 1.58 / 1.53 / 1.37e9 at 1e6 chains)
#endif
template <int BS>
__global__ void __launch_bounds__(BS, MCU_PUMPSF_MINB) pumps_fast_kernel(const __grid_constant__ PumpsFastCfg cfg, const __grid_constant__ RunArgs a) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const long long c = (long long)blockIdx.x * BS + tid;
  if (c >= a.n_chains) return;
  const size_t C = (size_t)a.n_chains;
  const uint32_t chain = (uint32_t)(a.chain_offset + c);
#define TH(i) smem[(i) * BS + tid]
  double al = a.state[0 * C + c], be = a.state[1 * C + c];
  for (int i = 0; i < NP; ++i) TH(i) = a.state[(size_t)(2 + i) * C + c];

  for (long long it = 1; it <= a.iters; ++it) {
    const long long iter = a.iter0 + it;
    const uint32_t it32 = (uint32_t)iter;
    // ================================================================== block 0: Slice([alpha, beta], 1.0, Univariate)
    {
      double SL = 0.0, ST = 0.0;
      for (int i = 0; i < NP; ++i) { const double th = TH(i); SL += fast_log(th); ST += th; }
      // Exponential(1)(alpha) + Gamma(0.1, 1)(beta) + sum_i Gamma(alpha, 1/beta)(theta_i), -Inf outside the supports
      // (src/model/simulation.jl:60-67: own priors first, early exit on a non-finite partial sum)
      auto logf = [&](double va, double vb) {
        if (!(va >= 0.0) || !(vb >= 0.0)) return -CUDART_INF;
        const double lb = fast_log(vb);
        return -va - 0.9 * lb - vb + (double)NP * (va * lb - lgamma(va)) + (va - 1.0) * SL - vb * ST;
      };
      UStreamP us{a, chain, it32, 0, 0, {0.0, 0.0}};
      double logf0 = logf(al, be);
      double lo0 = al - cfg.w_ab[0] * us.next(), lo1 = be - cfg.w_ab[1] * us.next();
      double up0 = lo0 + cfg.w_ab[0], up1 = lo1 + cfg.w_ab[1];
      {
        const double p0 = logf0 + log_uniform(us.next());
        const double x0 = al;
        double cur = lo0 + (up0 - lo0) * us.next();
        while (true) {
          logf0 = logf(cur, be);
          if (!(logf0 < p0)) break;
          if (cur < x0) lo0 = cur; else up0 = cur;
          cur = lo0 + (up0 - lo0) * us.next();
        }
        al = cur;
      }
      {
        const double p0 = logf0 + log_uniform(us.next());
        const double x0 = be;
        double cur = lo1 + (up1 - lo1) * us.next();
        while (true) {
          logf0 = logf(al, cur);
          if (!(logf0 < p0)) break;
          if (cur < x0) lo1 = cur; else up1 = cur;
          cur = lo1 + (up1 - lo1) * us.next();
        }
        be = cur;
      }
    }
    // ================================================================== block 1: Slice(theta, 1.0, Univariate)
    {
      UStreamP us{a, chain, it32, 1, (uint32_t)NP, {0.0, 0.0}};   // draws 0 .. 9 place the intervals (regenerated below), the cursor starts at 10
      double dcur = 0.0;   // logf(v) - logf(v at the start of the block): the slice levels are differences, the common part cancels
#pragma unroll 1
      for (int i = 0; i < NP; ++i) {
        const double x0 = TH(i);
        double lo = x0 - cfg.w_th[i] * us.at((uint32_t)i), up = lo + cfg.w_th[i];
        const double ca = al - 1.0 + cfg.y[i], cb = be + cfg.t[i];
        const double g0 = ca * fast_log(x0) - x0 * cb;                   // this component's two terms at its current value
        const double base = dcur - g0;                                   // logf(v with theta_i = t) - logf(start) = base + g_i(t)
        const double p0 = dcur + log_uniform(us.next());
        double cur = lo + (up - lo) * us.next();
        double lf;
        while (true) {
          lf = (cur >= 0.0) ? base + (ca * fast_log(cur) - cur * cb) : -CUDART_INF;   // Gamma / Poisson support: theta >= 0
          if (cur == 0.0) lf = -CUDART_INF;
          if (!(lf < p0)) break;
          if (cur < x0) lo = cur; else up = cur;
          cur = lo + (up - lo) * us.next();
        }
        TH(i) = cur; dcur = lf;
      }
    }
    // ================================================================== thinning + streaming moments (mcmc.jl:76-78)
    if (iter > a.burnin && (iter - a.burnin) % a.thin == 0) {
      double mon[PumpsModel::P];
      mon[0] = al; mon[1] = be;
      for (int i = 0; i < NP; ++i) mon[2 + i] = TH(i);
      if (a.samples) {
        const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
        for (int j = 0; j < PumpsModel::P; ++j) a.samples[((size_t)row * PumpsModel::P + j) * C + c] = mon[j];
      }
      if (a.comom) comoments_update(a.mom, a.momn, a.comom, C, (size_t)c, PumpsModel::P, mon, a.log_mask);
      moments_update(a.mom, a.momn, C, (size_t)c, PumpsModel::P, mon);
    }
  }
  a.state[0 * C + c] = al; a.state[1 * C + c] = be;
  for (int i = 0; i < NP; ++i) a.state[(size_t)(2 + i) * C + c] = TH(i);
#undef TH
}

// ---------------------------------------------------------------------------------------------------------------------------------
// BASELINE.json configs[4] / SURVEY.md §8d config 5: [Gibbs(theta), Gibbs(beta), AMWG(alpha)] — the conjugate full conditionals
//   theta_i | . ~ Gamma(alpha + y_i, 1 / (beta + t_i)),  beta | . ~ Gamma(0.1 + 10 alpha, 1 / (1 + sum theta))      (models.cuh, PumpsModel::gibbs)
// and a one-component AMWG update of alpha on x = log alpha (amwg.jl:99-115).  The draws are those of the generic kernel — the same
// Draws cursor per block, the same Marsaglia-Tsang routine — so the Gibbs values are bit-identical; what changes is the AMWG target: the
// ten Gamma(theta_i | alpha, 1/beta) terms enter through SL = sum log theta_i,
//   logf(x') - logf(x) = -(a' - a) + (x' - x) + 10 ((a' - a) log beta - (lgamma(a') - lgamma(a))) + (a' - a) SL,      a = exp(x),
// with lgamma(a) carried from the previous iteration (one lgamma per iteration instead of two block evaluations of ten terms each).
struct PumpsGibbsCfg {
  double y[NP], t[NP];
  double scale, target;
  int adapt, batchsize, tune_off;
};

#ifndef MCU_PUMPSG_PRETEST
#define MCU_PUMPSG_PRETEST 1
#endif
// ---- register-resident block streams for the Gibbs kernel -------------------------------------------------------------------
// rng.cuh's Draws keeps its cursors and the cached second draw behind `this` of two noinline members, i.e. in local memory: in this
// kernel that was 115 LDL + 80 STL per chain-iteration and the second-largest stall (long_scoreboard 3.1 per issue, profiles/
// r2_pumps_gibbs_note.md).  Here the state is a plain struct the compiler keeps in registers; the pair generators stay out of line
// (one copy in the instruction stream: this kernel's top stall is instruction fetch) and take / return everything by value.
// Same draws, same arithmetic as Draws::uniform / Draws::normal (PHILOX mode; the fused kernels never run on an EXTERNAL stream).
static __device__ __noinline__ Pair pg_normal_pair(uint32_t k0, uint32_t k1, uint32_t chain, uint32_t iter, uint32_t blockkind, uint32_t kpair) {
  uint32_t w[4];
  philox4x32_10(kpair, iter, chain, blockkind | (1u << 24), k0, k1, w);
  const double rad = sqrt(-2.0 * fast_log(1.0 - u53(w[0], w[1])));
  const Pair sc = fast_sincos2pi(u53(w[2], w[3]));
  return {rad * sc.b, rad * sc.a};   // (even draw of the stream: cosine branch, odd draw: sine branch)
}
static __device__ __noinline__ Pair pg_uniform_pair(uint32_t k0, uint32_t k1, uint32_t chain, uint32_t iter, uint32_t blockkind, uint32_t kpair) {
  uint32_t w[4];
  philox4x32_10(kpair, iter, chain, blockkind, k0, k1, w);
  return {u53(w[0], w[1]), u53(w[2], w[3])};
}
struct RegDraws {
  uint32_t k0, k1, chain, iter, blockkind, ku, kn, z_blk, u_blk;
  double z_next, u_next;
  MCU_D void seek(uint32_t it, uint32_t block, uint32_t kind) {
    iter = it; blockkind = block | (kind << 16); ku = 0; kn = 0; z_blk = 0xffffffffu; u_blk = 0xffffffffu;
  }
  MCU_D double uniform() {
    if ((ku & 1u) && u_blk == (ku >> 1)) { ++ku; return u_next; }
    const Pair p = pg_uniform_pair(k0, k1, chain, iter, blockkind, ku >> 1);
    u_next = p.b; u_blk = ku >> 1;
    const double u = (ku & 1u) ? p.b : p.a;
    ++ku;
    return u;
  }
  MCU_D double normal() {
    if ((kn & 1u) && z_blk == (kn >> 1)) { ++kn; return z_next; }
    const Pair p = pg_normal_pair(k0, k1, chain, iter, blockkind, kn >> 1);
    z_next = p.b; z_blk = kn >> 1;
    const double z = (kn & 1u) ? p.b : p.a;
    ++kn;
    return z;
  }
};

// Marsaglia-Tsang acceptance: u < 1 - 0.0331 x^4 (squeeze), else log u < x^2 / 2 + d (1 - v + log v).  The second test needs two FP64 logs
// and, although only ~8 % of the lanes reach it, a warp almost always has such a lane: it was 16 % of the kernel's instructions.  It is
// decided on float logs (MUFU.LG2) whenever the two sides differ by more than a band 30x the float error, and in FP64 otherwise — the
// decisions are those of the FP64 test (samplers.cuh rgamma_mt).
MCU_D bool mt_accept(double u, double x2, double d, double v) {
  if (u < 1.0 - 0.0331 * x2 * x2) return true;
#if MCU_PUMPSG_PRETEST
  if (v > 1e-30) {
    const float lu = __log2f((float)u) * 0.693147181f, lv = __log2f((float)v) * 0.693147181f;
    const double rhs = 0.5 * x2 + d * ((1.0 - v) + (double)lv);
    const double band = 1e-5 * (1.0 + fabs((double)lu) + d * (1.0 + fabs((double)lv)));
    const double diff = (double)lu - rhs;
    if (diff < -band) return true;
    if (diff > band) return false;
  }
#endif
  return flog(u) < 0.5 * x2 + d * (1.0 - v + flog(v));
}
// samplers.cuh rgamma_mt on a register-resident stream
template <class R>
MCU_D double rgamma_mt_reg(double a, R& rng) {
  double boost = 1.0;
  if (a < 1.0) { boost = pow(rng.uniform(), 1.0 / a); a += 1.0; }
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x, v;
    do { x = rng.normal(); v = 1.0 + c * x; } while (v <= 0.0);
    v = v * v * v;
    const double u = rng.uniform();
    if (mt_accept(u, x * x, d, v)) return boost * d * v;
  }
}

#ifndef MCU_PUMPSG_REGDRAWS
#define MCU_PUMPSG_REGDRAWS 1
#endif
#ifndef MCU_PUMPSG_MINB
#define MCU_PUMPSG_MINB 10   // resident blocks per SM.  Round 1 (draw state in local memory): 4 / 6 / 8 / 10 / 12 / 14 / 16 gave 2.6 / 2.9 / 3.0 / 3.1 / 3.4 / 3.3 / 3.3e9 chain-iterations/s
                             // at 1e7 chains; round 2 (register-resident draws + float pre-test, 1e6 chains): 6 / 8 / 10 / 12 give 4.41 / 4.40 / 4.64 / 4.58e9
#endif
template <int BS>
__global__ void __launch_bounds__(BS, MCU_PUMPSG_MINB) pumps_gibbs_kernel(const __grid_constant__ PumpsGibbsCfg cfg, const __grid_constant__ RunArgs a) {
  extern __shared__ double smem[];
  __shared__ double sy[NP], st[NP];                 // pump data in shared memory: the lanes of a warp index them with different i
  const int tid = threadIdx.x;
  if (tid < NP) { sy[tid] = cfg.y[tid]; st[tid] = cfg.t[tid]; }
  __syncthreads();
  const long long c = (long long)blockIdx.x * BS + tid;
  if (c >= a.n_chains) return;
  const size_t C = (size_t)a.n_chains;
#define TH(i) smem[(i) * BS + tid]
  double al = a.state[0 * C + c], be = a.state[1 * C + c];
  for (int i = 0; i < NP; ++i) TH(i) = a.state[(size_t)(2 + i) * C + c];
  // AMWG tune record of block 2: m, adapt flag, sigma, accept count (samplers.cuh, amwg_sample)
  double* tn = a.tune + (size_t)cfg.tune_off * C + c;
  double m = tn[0 * C], sigma = tn[2 * C], acc = tn[3 * C];
  bool was = tn[1 * C] != 0.0;
  double lg_al = lgamma(al);
#if MCU_PUMPSG_REGDRAWS
  RegDraws rng;
  rng.k0 = (uint32_t)a.seed; rng.k1 = (uint32_t)(a.seed >> 32); rng.chain = (uint32_t)(a.chain_offset + c);
#else
  Draws rng;
  rng.k0 = (uint32_t)a.seed; rng.k1 = (uint32_t)(a.seed >> 32); rng.chain = (uint32_t)(a.chain_offset + c);
  rng.ext = nullptr; rng.ext_n = 0; rng.ext_pos = nullptr;
#endif

  for (long long it = 1; it <= a.iters; ++it) {
    const long long iter = a.iter0 + it;
    const uint32_t it32 = (uint32_t)iter;
    // ---- block 0: Gibbs(theta).  Marsaglia-Tsang rejects ~4 % of its attempts; with one attempt loop per theta_i a warp would repeat the
    // loop body for the one or two lanes that rejected (ncu: 19 of 32 lanes active on average).  Here every lane walks its OWN component
    // index: a trip is one attempt (one normal, and one uniform when v > 0 — the order rgamma_mt consumes them), a lane that accepts moves
    // on to its next component, a lane that rejects retries, and the warp stays converged until the last lanes finish.
    rng.seek(it32, 0, 0);
    {
      int i = 0;
      double d = 0.0, cc = 0.0, boost = 1.0;
      bool setup = true;
#pragma unroll 1
      while (i < NP) {
        if (setup) {
          double sh = al + sy[i];
          boost = 1.0;
          if (sh < 1.0) { boost = pow(rng.uniform(), 1.0 / sh); sh += 1.0; }
          d = sh - 1.0 / 3.0; cc = fast_rsqrt(9.0 * d);   // (the generic kernel: 1 / sqrt, correctly rounded twice; here <= 2 ulp: the theta draws agree to ~1e-16)
          setup = false;
        }
        const double xn = rng.normal();
        double v = 1.0 + cc * xn;
        if (v > 0.0) {
          v = v * v * v;
          const double u = rng.uniform();
          const double x2 = xn * xn;
          if (mt_accept(u, x2, d, v)) {
            TH(i) = boost * d * v * fast_rcp(be + st[i]);
            ++i; setup = true;
          }
        }
      }
    }
    // ---- block 1: Gibbs(beta)
    rng.seek(it32, 1, 0);
    double sth = 0.0, SL = 0.0;
    for (int i = 0; i < NP; ++i) { const double th = TH(i); sth += th; SL += fast_log(th); }
#if MCU_PUMPSG_REGDRAWS
    be = rgamma_mt_reg(0.1 + (double)NP * al, rng) / (1.0 + sth);
#else
    be = rgamma_mt(0.1 + (double)NP * al, rng) / (1.0 + sth);
#endif
    // ---- block 2: AMWG(alpha) on x = log alpha
    rng.seek(it32, 2, 0);
    const bool adapt = cfg.adapt == 1 ? iter <= a.burnin : cfg.adapt == 0;
    if (iter == 1) { m = 0.0; was = false; sigma = cfg.scale; acc = 0.0; }      // AMWGTune(x, sigma): fresh at the first iteration
    if (adapt && !was) { acc = 0.0; m = 0.0; }                                     // setadapt!: amwg.jl:88-96
    was = adapt;
    if (adapt) m += 1.0;
    const double x = log(al);                                                      // unlist on the link scale ...
    const double a0 = exp(x);                                                      // ... and relist: the generic kernel's round trip, kept bit for bit
    const double z = sigma * rng.normal();
    const double xn = x + z;
    const double an = exp(xn);
    const double lg_an = lgamma(an);
    const double da = an - a0;
    const double delta = -da + (xn - x) + (double)NP * (da * fast_log(be) - (lg_an - lg_al)) + da * SL;
    if (rng.uniform() < exp(delta)) { al = an; lg_al = lg_an; if (adapt) acc += 1.0; }
    else al = a0;
    if (adapt && ((long long)m % cfg.batchsize) == 0) {                            // amwg.jl:74-80
      const double dl = amwg_delta(m, cfg.batchsize);
      sigma *= exp(acc / m < cfg.target ? -dl : dl);
    }
    // ---- thinning + streaming moments (mcmc.jl:76-78)
    if (iter > a.burnin && (iter - a.burnin) % a.thin == 0) {
      double mon[PumpsModel::P];
      mon[0] = al; mon[1] = be;
      for (int i = 0; i < NP; ++i) mon[2 + i] = TH(i);
      if (a.samples) {
        const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
        for (int j = 0; j < PumpsModel::P; ++j) a.samples[((size_t)row * PumpsModel::P + j) * C + c] = mon[j];
      }
      if (a.comom) comoments_update(a.mom, a.momn, a.comom, C, (size_t)c, PumpsModel::P, mon, a.log_mask);
      moments_update(a.mom, a.momn, C, (size_t)c, PumpsModel::P, mon);
    }
  }
  a.state[0 * C + c] = al; a.state[1 * C + c] = be;
  for (int i = 0; i < NP; ++i) a.state[(size_t)(2 + i) * C + c] = TH(i);
  tn[0 * C] = m; tn[1 * C] = was ? 1.0 : 0.0; tn[2 * C] = sigma; tn[3 * C] = acc;
#undef TH
}

}  // namespace

int pumps_gibbs_launch(const double* y, const double* t, int N, const RunArgs& a, const DevBlock& amwg, double scale, cudaStream_t st) {
  if (N != NP) return -2;
  PumpsGibbsCfg cfg;
  for (int i = 0; i < NP; ++i) { cfg.y[i] = y[i]; cfg.t[i] = t[i]; }
  cfg.scale = scale; cfg.target = amwg.target; cfg.adapt = amwg.adapt; cfg.batchsize = amwg.batchsize; cfg.tune_off = amwg.tune_off;
  constexpr int BS = 128;
  const size_t smem = (size_t)BS * NP * sizeof(double);
  pumps_gibbs_kernel<BS><<<(unsigned)((a.n_chains + BS - 1) / BS), BS, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int pumps_fast_launch(const double* y, const double* t, int N, const RunArgs& a, const std::vector<std::vector<double>>& h_scales, cudaStream_t st) {
  if (N != NP) return -2;
  PumpsFastCfg cfg;
  for (int i = 0; i < NP; ++i) { cfg.y[i] = y[i]; cfg.t[i] = t[i]; cfg.w_th[i] = h_scales[1][i]; }
  cfg.w_ab[0] = h_scales[0][0]; cfg.w_ab[1] = h_scales[0][1];
  constexpr int BS = 128;
  const size_t smem = (size_t)BS * NP * sizeof(double);
  pumps_fast_kernel<BS><<<(unsigned)((a.n_chains + BS - 1) / BS), BS, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mcu
