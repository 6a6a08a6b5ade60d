// glm_nuts.cu — NUTS for the Bernoulli-logit GLM template at data sizes where one chain per thread cannot
// evaluate the density (N = 10^6 rows): the "tick" engine of SURVEY.md §7 step 8.
//
// The reference runs, per chain, nuts_sub! → buildtree → leapfrog → logfgrad (src/samplers/nuts.jl:95-180),
// one gradient after the other.  Here a gradient evaluation is a pass over X shared by ALL chains, so the
// sampler of every chain is turned into a resumable state machine:
//
//     advance kernel (one thread per chain)  ── consumes the (logf, grad) it asked for, runs the NUTS
//                                               bookkeeping up to its NEXT gradient request, writes the
//                                               requested position
//     gradient kernels (all chains at once)  ── logf and grad of the likelihood at every requested position
//
// One host-side loop alternates the two ("tick") until every chain has finished its iterations.  A chain
// requests exactly the gradients the reference's recursion would (same leapfrogs, same uniforms in the same
// post-order, same dual averaging, same nutsepsilon), so on the same Philox stream it follows the same
// trajectory as the generic kernel and the oracle; chains sit at different tree depths / iterations in the
// same tick.
//
// Chain-state layout: every vector is [d][C] (chain fastest); the per-level stack is [depth][d][C].
#include "launch.hpp"

namespace mcu {

namespace {

enum Phase { PH_BEGIN = 0, PH_EPS_INIT = 1, PH_EPS_TRIAL = 2, PH_START = 3, PH_LEAF = 4, PH_DONE = 5 };

// scalar slots per chain (doubles)
enum Slot {
  SL_PHASE = 0, SL_ITER, SL_JDRAW /* uniforms drawn */, SL_KN /* normals drawn */, SL_LOGP0, SL_LOGU0, SL_N, SL_TN, SL_TS, SL_ALPHA, SL_NALPHA, SL_J, SL_T, SL_PM,
  SL_EPS_USE, SL_EPS_TRY, SL_EPS_PM, SL_EPS_LOGF0, SL_EPS_D0, SL_EPS_GUARD, SL_SN0,   // SL_SN0 .. SL_SN0 + kMaxDepth - 1: stack n
  SL_COUNT = SL_SN0 + kMaxDepth
};
// vector slots per chain (each d doubles)
enum VSlot { V_CX = 0, V_CR, V_CG, V_XM, V_RM, V_GM, V_XP, V_RP, V_GP, V_TXF, V_TRF, V_TXP, V_SXF0,   // then 3 * kMaxDepth stack vectors
             V_COUNT = V_SXF0 + 3 * kMaxDepth };

struct GlmTickArgs {
  long long C, chain_offset;
  unsigned long long seed;
  long long target_iter, burnin, thin, row0;
  int d, max_depth;
  double target, eps_desc;
  double* state;     // [d][C] current sample v (the handle's state array)
  double* tune;      // [8][C]  NUTS tune slots (samplers.cuh)
  double* sc;        // [SL_COUNT][C]
  double* vec;       // [V_COUNT][d][C]
  double* req;       // [d][C] requested position
  const double* lp;  // [C] likelihood logf at the requested position
  const double* grad;// [d][C] likelihood gradient
  double* samples; double* mom; double* momn;
  int* n_active;     // [2] ring: slot (tick & 1) counts chains still running after this tick
  unsigned long long* work;   // useful chain-gradients: incremented once per chain that still wants the gradient of this tick's pass
  int tick;
};

// One WARP per chain: the lanes stride over the d components of every vector (a thread-per-chain version spends its
// time in d-long chains of dependent global loads/stores), scalars are read by all lanes and written by lane 0, dot
// products are butterfly-reduced, and lane i draws the normal of component i from its own Philox counter.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(128) glm_advance_kernel(GlmTickArgs a) {
  const int lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.n_active[(a.tick + 1) & 1] = 0;   // the other slot was read by the host before this launch
  if (c >= a.C) return;
  const size_t C = (size_t)a.C;
  const int d = a.d;
#define SC(s) a.sc[(size_t)(s) * C + c]
#define SCW(s, v) do { if (lane == 0) a.sc[(size_t)(s) * C + c] = (v); } while (0)
#define VV(v, i) a.vec[((size_t)(v) * d + (i)) * C + c]
#define ST(i) a.state[(size_t)(i) * C + c]
#define TN(s) a.tune[(size_t)(s) * C + c]
#define TNW(s, v) do { if (lane == 0) a.tune[(size_t)(s) * C + c] = (v); } while (0)
#define REQ(i) a.req[(size_t)(i) * C + c]
  int phase = (int)SC(SL_PHASE);
  long long iter = (long long)SC(SL_ITER);
  if (phase == PH_DONE && iter >= a.target_iter) return;
  if (phase == PH_DONE) phase = PH_BEGIN;   // a later mcu_run continues the chain
  const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32), gchain = (uint32_t)(a.chain_offset + c);
  uint32_t jdraw = (uint32_t)SC(SL_JDRAW), kn = (uint32_t)SC(SL_KN);
  // rng.cuh contract: stream 0 = uniforms, stream 1 = normals, two draws per Philox block (block 0, kind 0)
  auto words = [&](uint32_t k, uint32_t stream, uint32_t (&w)[4]) { philox4x32_10(k >> 1, (uint32_t)iter, gchain, stream << 24, k0, k1, w); };
  auto uniform = [&]() {   // same value on every lane
    uint32_t w[4]; words(jdraw, 0u, w);
    const double u = (jdraw & 1u) ? u53(w[2], w[3]) : u53(w[0], w[1]);
    ++jdraw; return u;
  };
  auto normal_at = [&](uint32_t k) {
    uint32_t w[4]; words(k, 1u, w);
    const double ua = u53(w[0], w[1]), ub = u53(w[2], w[3]);
    return (k & 1u) ? box_muller_sin(ua, ub) : box_muller(ua, ub);
  };

  // NUTS tune scalars are kept in registers while the chain advances
  double t_adapt = TN(0), t_alpha = TN(1), t_eps = TN(2), t_epsbar = TN(3), t_Hbar = TN(4), t_m = TN(5), t_mu = TN(6), t_nalpha = TN(7);

  // pending result: full block density = MvNormal(d, sqrt(1000)) prior + likelihood (glm template, models.cuh)
  double lp_full = 0.0;
  auto load_result = [&](int vslot_g) {
    double sq = 0.0; int fin = 1;
    for (int i = lane; i < d; i += 32) {
      const double b = REQ(i);
      double g = a.grad[(size_t)i * C + c] - b / 1000.0;
      sq += b * b; fin &= isfinite(b) ? 1 : 0;
      if (!isfinite(g)) g = 0.0;                       // logpdfgrad!: sampler.jl:110
      VV(vslot_g, i) = g;
    }
    sq = warp_sum(sq);
    fin = __all_sync(0xffffffffu, fin);
    const double prior = fin ? lp_isonormal(sq, (double)d, sqrt(1000.0)) : neg_inf();
    lp_full = prior + a.lp[c];
  };
  auto dotv = [&](int vs) { double s = 0; for (int i = lane; i < d; i += 32) { const double x = VV(vs, i); s += x * x; } return warp_sum(s); };
  auto copyv = [&](int dst, int src) { for (int i = lane; i < d; i += 32) VV(dst, i) = VV(src, i); };
  auto copyv3 = [&](int d0, int s0, int d1, int s1, int d2, int s2) {
    for (int i = lane; i < d; i += 32) { const double x0 = VV(s0, i), x1 = VV(s1, i), x2 = VV(s2, i); VV(d0, i) = x0; VV(d1, i) = x1; VV(d2, i) = x2; }
  };
  auto nouturn = [&](int xminus, int xplus, int rminus, int rplus) {
    double p = 0, q = 0;
    for (int i = lane; i < d; i += 32) { const double df = VV(xplus, i) - VV(xminus, i); p += df * VV(rminus, i); q += df * VV(rplus, i); }
    p = warp_sum(p); q = warp_sum(q);
    return p >= 0 && q >= 0;
  };
  // first half of a leapfrog from (cx, cr, cg) and the gradient request at the new position: nuts.jl:130-131
  auto half_step_and_request = [&](double eps) {
    for (int i = lane; i < d; i += 32) {
      const double rn = VV(V_CR, i) + (0.5 * eps) * VV(V_CG, i); const double xn = VV(V_CX, i) + eps * rn;
      VV(V_CR, i) = rn; VV(V_CX, i) = xn; REQ(i) = xn;
    }
  };
  // start of nuts_sub! (nuts.jl:97-98): r = randn(n), gradient request at the current sample
  auto start_iteration_request = [&]() {
    const bool adapt = iter <= a.burnin;
    if (adapt && t_adapt == 0.0) { t_m = 0.0; t_mu = log(10.0 * t_eps); }    // setadapt!: nuts.jl:84-92
    t_adapt = adapt ? 1.0 : 0.0;
    if (adapt) t_m += 1.0; else if (t_m > 0.0) t_eps = t_epsbar;
    SCW(SL_EPS_USE, t_eps);
    for (int i = lane; i < d; i += 32) { const double x = ST(i); VV(V_CR, i) = normal_at(kn + (uint32_t)i); VV(V_CX, i) = x; VV(V_CG, i) = 0.0; REQ(i) = x; }
    kn += (uint32_t)d;
  };
  double eps_use = SC(SL_EPS_USE);

  bool need_grad = false;
  while (!need_grad) {
    switch (phase) {
      case PH_BEGIN: {
        if (iter >= a.target_iter) { phase = PH_DONE; for (int i = lane; i < d; i += 32) REQ(i) = ST(i); need_grad = true; break; }
        iter += 1; jdraw = 0; kn = 0;
        if (iter == 1) {   // NUTSTune(x, nutsepsilon(x, f)): nuts.jl:17-30
          t_adapt = 0.0; t_alpha = 0.0; t_epsbar = 1.0; t_Hbar = 0.0; t_m = 0.0; t_mu = CUDART_NAN; t_nalpha = 0.0;
          if (a.eps_desc > 0.0) { t_eps = a.eps_desc; }
          else {           // nutsepsilon: nuts.jl:192-205 — r0 = randn(n); leapfrog(x, r0, 0, 0)
            for (int i = lane; i < d; i += 32) { const double x = ST(i); VV(V_RM, i) = normal_at(kn + (uint32_t)i); VV(V_CX, i) = x; REQ(i) = x; }
            kn += (uint32_t)d;
            phase = PH_EPS_INIT; need_grad = true; break;
          }
        }
        start_iteration_request(); eps_use = t_eps;
        phase = PH_START; need_grad = true;
        break;
      }
      case PH_EPS_INIT: {   // (logf0, grad0) at x0 arrived; r0 sits in V_RM, grad0 goes to V_GM
        load_result(V_GM);
        SCW(SL_EPS_LOGF0, lp_full); { const double d0 = dotv(V_RM); SCW(SL_EPS_D0, d0); }
        SCW(SL_EPS_TRY, 1.0); SCW(SL_EPS_PM, 0.0); SCW(SL_EPS_GUARD, 0.0);
        for (int i = lane; i < d; i += 32) {   // trial leapfrog from (x0, r0, grad0) with eps = 1
          const double r = VV(V_RM, i) + 0.5 * VV(V_GM, i); const double x = ST(i) + 1.0 * r;
          VV(V_CR, i) = r; VV(V_CX, i) = x; REQ(i) = x;
        }
        phase = PH_EPS_TRIAL; need_grad = true;
        break;
      }
      case PH_EPS_TRIAL: {
        load_result(V_CG);
        double eps = SC(SL_EPS_TRY);
        double dr = 0.0;
        for (int i = lane; i < d; i += 32) { const double r = VV(V_CR, i) + (0.5 * eps) * VV(V_CG, i); dr += r * r; }
        dr = warp_sum(dr);
        const double prob = exp(lp_full - SC(SL_EPS_LOGF0) - 0.5 * (dr - SC(SL_EPS_D0)));
        double pm = SC(SL_EPS_PM);
        if (pm == 0.0) { pm = prob > 0.5 ? 1.0 : -1.0; }
        const double guard = SC(SL_EPS_GUARD) + 1.0;
        __syncwarp();
        SCW(SL_EPS_PM, pm); SCW(SL_EPS_GUARD, guard);
        if (pow(prob, pm) > pow(0.5, pm) && guard <= 2000.0) {
          eps *= pm > 0 ? 2.0 : 0.5; SCW(SL_EPS_TRY, eps);
          for (int i = lane; i < d; i += 32) {
            const double r = VV(V_RM, i) + (0.5 * eps) * VV(V_GM, i); const double x = ST(i) + eps * r;
            VV(V_CR, i) = r; VV(V_CX, i) = x; REQ(i) = x;
          }
          need_grad = true;
        } else {
          t_eps = eps;
          start_iteration_request(); eps_use = t_eps;
          phase = PH_START; need_grad = true;
        }
        break;
      }
      case PH_START: {   // nuts_sub!: nuts.jl:97-105 after the eps = 0 leapfrog
        load_result(V_CG);
        const double logp0 = lp_full - 0.5 * dotv(V_CR);
        const double logu0 = logp0 + log(uniform());
        SCW(SL_LOGP0, logp0); SCW(SL_LOGU0, logu0);
        copyv3(V_XM, V_CX, V_RM, V_CR, V_GM, V_CG); copyv3(V_XP, V_CX, V_RP, V_CR, V_GP, V_CG);
        SCW(SL_J, 0.0); SCW(SL_N, 1.0);
        const double pm = uniform() > 0.5 ? 1.0 : -1.0;   // first doubling; both edges equal the start point
        SCW(SL_PM, pm); SCW(SL_T, 0.0); SCW(SL_ALPHA, 0.0); SCW(SL_NALPHA, 0.0);
        half_step_and_request(pm * eps_use);
        phase = PH_LEAF; need_grad = true;
        break;
      }
      case PH_LEAF: {    // buildtree leaf (nuts.jl:142-152) + the unrolled merges (samplers.cuh nuts_sub)
        load_result(V_CG);
        const double pm = SC(SL_PM), eps = pm * eps_use;
        double dotr = 0.0;
        for (int i = lane; i < d; i += 32) { const double rn = VV(V_CR, i) + (0.5 * eps) * VV(V_CG, i); VV(V_CR, i) = rn; dotr += rn * rn; }
        dotr = warp_sum(dotr);
        const double logu0 = SC(SL_LOGU0), logp0 = SC(SL_LOGP0);
        const double logpp = lp_full - 0.5 * dotr;
        double Tn = logu0 < logpp ? 1.0 : 0.0;
        bool Ts = logu0 < logpp + 1000.0;
        const double alpha = SC(SL_ALPHA) + fmin(1.0, exp(logpp - logp0));
        const double nalpha = SC(SL_NALPHA) + 1.0;
        const int j = (int)SC(SL_J); const unsigned t = (unsigned)SC(SL_T);
        double n = SC(SL_N);
        __syncwarp();
        SCW(SL_ALPHA, alpha); SCW(SL_NALPHA, nalpha);
        copyv3(V_TXF, V_CX, V_TRF, V_CR, V_TXP, V_CX);
        __syncwarp();
        int l = 0;
        bool parked = false;
        while (l < j) {
          const int sxf = V_SXF0 + 3 * l, srf = sxf + 1, sxp = sxf + 2;
          if ((t >> l) & 1u) {
            const double u = uniform();
            const double nA = SC(SL_SN0 + l);
            if (!(u < Tn / (nA + Tn))) copyv(V_TXP, sxp);
            Tn = nA + Tn;
            const bool ok = pm > 0 ? nouturn(sxf, V_CX, srf, V_CR) : nouturn(V_CX, sxf, V_CR, srf);
            Ts = Ts && ok;
            copyv(V_TXF, sxf); copyv(V_TRF, srf);
            __syncwarp();
            ++l;
          } else if (Ts) {
            copyv3(sxf, V_TXF, srf, V_TRF, sxp, V_TXP); SCW(SL_SN0 + l, Tn);
            parked = true; break;
          } else {
            ++l;
          }
        }
        if (parked) {      // build the sibling: next leaf
          SCW(SL_T, (double)(t + 1));
          half_step_and_request(eps);
          need_grad = true;
          break;
        }
        // tree of depth j complete (or failed): nuts.jl:108-123
        if (pm < 0) copyv3(V_XM, V_CX, V_RM, V_CR, V_GM, V_CG);
        else copyv3(V_XP, V_CX, V_RP, V_CR, V_GP, V_CG);
        __syncwarp();
        if (Ts) { if (uniform() < Tn / n) for (int i = lane; i < d; i += 32) ST(i) = VV(V_TXP, i); }
        const int jn = j + 1;
        n += Tn;
        bool s = Ts && nouturn(V_XM, V_XP, V_RM, V_RP);
        if (s && jn >= a.max_depth) { s = false; if (lane == 0 && a.work) atomicAdd(a.work + 1, 1ull); }   // depth cap bound (reference: no cap, nuts.jl:106-124)
        t_alpha = alpha; t_nalpha = nalpha;
        SCW(SL_N, n); SCW(SL_J, (double)jn);
        if (s) {           // next doubling
          const double pm2 = uniform() > 0.5 ? 1.0 : -1.0;
          SCW(SL_PM, pm2); SCW(SL_T, 0.0); SCW(SL_ALPHA, 0.0); SCW(SL_NALPHA, 0.0);
          if (pm2 < 0) copyv3(V_CX, V_XM, V_CR, V_RM, V_CG, V_GM);
          else copyv3(V_CX, V_XP, V_CR, V_RP, V_CG, V_GP);
          __syncwarp();
          half_step_and_request(pm2 * eps_use);
          need_grad = true;
          break;
        }
        // end of the iteration: dual averaging (nuts.jl:70-75), thinning (mcmc.jl:76-78)
        if (t_adapt != 0.0) {
          const double m = t_m;
          double p = 1.0 / (m + 10.0);
          t_Hbar = (1.0 - p) * t_Hbar + p * (a.target - t_alpha / t_nalpha);
          t_eps = exp(t_mu - sqrt(m) * t_Hbar / 0.05);
          p = pow(m, -0.75);
          t_epsbar = exp(p * log(t_eps) + (1.0 - p) * log(t_epsbar));
        }
        __syncwarp();
        if (iter > a.burnin && (iter - a.burnin) % a.thin == 0) {
          const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
          if (a.samples) for (int i = lane; i < d; i += 32) a.samples[((size_t)row * d + i) * C + c] = ST(i);
          // streaming moments, one monitored column per lane step (same update as engine.cuh moments_update)
          const double nk = a.momn[c] + 1.0;
          double bc = a.momn[C + c] + 1.0; const bool bdone = bc >= (double)kBatch; double nb = a.momn[2 * C + c];
          if (bdone) { bc = 0.0; nb += 1.0; }
          __syncwarp();
          if (lane == 0) { a.momn[c] = nk; a.momn[C + c] = bc; a.momn[2 * C + c] = nb; }
          for (int i = lane; i < d; i += 32) {
            double* q = a.mom + (size_t)i * kMomPerCol * C + c;
            const double x = ST(i);
            double mean = q[0], M2 = q[C]; double dl = x - mean; mean += dl / nk; M2 += dl * (x - mean); q[0] = mean; q[C] = M2;
            const double lx = log(x); double lm = q[2 * C], lM2 = q[3 * C]; dl = lx - lm; lm += dl / nk; lM2 += dl * (lx - lm); q[2 * C] = lm; q[3 * C] = lM2;
            q[4 * C] = nk == 1.0 ? x : fmin(q[4 * C], x); q[5 * C] = nk == 1.0 ? x : fmax(q[5 * C], x);
            double bsum = q[6 * C] + x;
            if (bdone) { const double bm = bsum / (double)kBatch; bsum = 0.0; double bmean = q[7 * C], bM2 = q[8 * C]; const double db = bm - bmean; bmean += db / nb; bM2 += db * (bm - bmean); q[7 * C] = bmean; q[8 * C] = bM2; }
            q[6 * C] = bsum;
          }
        }
        phase = PH_BEGIN;
        break;
      }
      default: need_grad = true; break;
    }
  }
  __syncwarp();
  if (lane == 0) {
    a.sc[(size_t)SL_PHASE * C + c] = (double)phase; a.sc[(size_t)SL_ITER * C + c] = (double)iter; a.sc[(size_t)SL_JDRAW * C + c] = (double)jdraw; a.sc[(size_t)SL_KN * C + c] = (double)kn;
    a.tune[0 * C + c] = t_adapt; a.tune[1 * C + c] = t_alpha; a.tune[2 * C + c] = t_eps; a.tune[3 * C + c] = t_epsbar;
    a.tune[4 * C + c] = t_Hbar; a.tune[5 * C + c] = t_m; a.tune[6 * C + c] = t_mu; a.tune[7 * C + c] = t_nalpha;
    if (phase != PH_DONE) { atomicAdd(&a.n_active[a.tick & 1], 1); if (a.work) atomicAdd(a.work, 1ull); }
  }
#undef SC
#undef SCW
#undef VV
#undef ST
#undef TN
#undef TNW
#undef REQ
}

// ---- reference gradient kernel (FP64, CUDA cores): thread = (chain, row slab) ---------------------------
// lik logf = sum_i [y_i log p_i + (1 - y_i) log(1 - p_i)], grad = X' (y - p), p = invlogit(X beta).
// Deterministic: per-slab partials, then a fold over slabs.  This is the correctness path for the tick
// engine; the tensor-core kernel (glm_tc.cu) replaces it for large N.
template <int DMAX>
__global__ void __launch_bounds__(128) glm_grad_ref_kernel(const double* __restrict__ X, const double* __restrict__ y, int N, int d,
                                                           long long C, const double* __restrict__ req, int rows_per_slab,
                                                           double* __restrict__ part_lp /*[slab][C]*/, double* __restrict__ part_g /*[slab][d][C]*/,
                                                           int family, double sigma) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int slab = blockIdx.y;
  if (c >= C) return;
  double beta[DMAX], g[DMAX];
  for (int j = 0; j < d; ++j) { beta[j] = req[(size_t)j * C + c]; g[j] = 0.0; }
  double lp = 0.0;
  const int i0 = slab * rows_per_slab, i1 = min(N, i0 + rows_per_slab);
  for (int i = i0; i < i1; ++i) {
    const double* xi = X + (size_t)i * d;
    double eta = 0.0;
    for (int j = 0; j < d; ++j) eta += xi[j] * beta[j];
    const double yi = y[i];
    double r;
    if (family == 1) {          // Poisson / log link
      const double lam = exp(eta);
      lp += lp_poisson(yi, lgamma(yi + 1.0), lam);
      r = yi - lam;
    } else if (family == 2) {   // Normal / identity link, known sd
      lp += lp_normal(yi, eta, sigma);
      r = (yi - eta) / (sigma * sigma);
    } else {                    // Bernoulli / logit link
      const double p = 1.0 / (exp(-eta) + 1.0);
      lp += yi == 0.0 ? log(1.0 - p) : log(p);
      r = yi - p;
    }
    for (int j = 0; j < d; ++j) g[j] += r * xi[j];
  }
  part_lp[(size_t)slab * C + c] = lp;
  for (int j = 0; j < d; ++j) part_g[((size_t)slab * d + j) * C + c] = g[j];
}
// out[i] = sum_s part[s][i]
__global__ void glm_fold_kernel(const double* __restrict__ part, int nslab, long long width, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  double s = 0.0;
  for (int k = 0; k < nslab; ++k) s += part[(size_t)k * width + i];
  out[i] = s;
}

}  // namespace

size_t glm_tick_scalar_slots() { return SL_COUNT; }
size_t glm_tick_vector_slots() { return V_COUNT; }

void glm_grad_reference(const double* X, const double* y, int N, int d, long long C, const double* req, int nslab,
                        double* part_lp, double* part_g, double* lp, double* grad, int family, double sigma, cudaStream_t st) {
  const int rows = (N + nslab - 1) / nslab;
  dim3 grid((unsigned)((C + 127) / 128), (unsigned)nslab);
  glm_grad_ref_kernel<kGlmDMax><<<grid, 128, 0, st>>>(X, y, N, d, C, req, rows, part_lp, part_g, family, sigma);
  glm_fold_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(part_lp, nslab, C, lp);
  glm_fold_kernel<<<(unsigned)(((long long)d * C + 255) / 256), 256, 0, st>>>(part_g, nslab, (long long)d * C, grad);
}

// lp[c] = sum_s part_lp[s][c]  and  grad[j][c] = sum_s part_g[s][j][c]  in one launch
__global__ void glm_fold2_kernel(const double* __restrict__ part_lp, const double* __restrict__ part_g, int nslab_lp, int nslab_g,
                                 long long C, long long dC, double* __restrict__ lp, double* __restrict__ grad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dC) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = 0;
    for (; k + 4 <= nslab_g; k += 4) {
      s0 += part_g[(size_t)k * dC + i]; s1 += part_g[(size_t)(k + 1) * dC + i];
      s2 += part_g[(size_t)(k + 2) * dC + i]; s3 += part_g[(size_t)(k + 3) * dC + i];
    }
    for (; k < nslab_g; ++k) s0 += part_g[(size_t)k * dC + i];
    grad[i] = (s0 + s1) + (s2 + s3);
  } else if (i < dC + C) {
    const long long c = i - dC;
    double s = 0.0;
    for (int k = 0; k < nslab_lp; ++k) s += part_lp[(size_t)k * C + c];
    lp[c] = s;
  }
}
// tensor-core path: FP32 gradient partials; lp[c] = lp_const + beta_c . xty + sum_s part_lp[s][c] (xty, lp_const: the family's linear / constant part).
// The partials are indexed by PASS slot k (C slots); with a compaction map the results go to chain map[k] of the Cfull-strided lp / grad / req.
__global__ void glm_fold_tc_kernel(const double* __restrict__ part_lp, const float* __restrict__ part_g, int nslab_lp, int nslab_g,
                                   long long C, int d, const double* __restrict__ req, const double* __restrict__ xty, double lp_const,
                                   double* __restrict__ lp, double* __restrict__ grad, const int* __restrict__ map, long long Cfull,
                                   const double* __restrict__ col_inv) {
  const long long dC = (long long)d * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dC) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = 0;
    for (; k + 4 <= nslab_g; k += 4) {
      s0 += (double)part_g[(size_t)k * dC + i]; s1 += (double)part_g[(size_t)(k + 1) * dC + i];
      s2 += (double)part_g[(size_t)(k + 2) * dC + i]; s3 += (double)part_g[(size_t)(k + 3) * dC + i];
    }
    for (; k < nslab_g; ++k) s0 += (double)part_g[(size_t)k * dC + i];
    const long long j = i / C, slot = i % C;
    const double gsum = (s0 + s1) + (s2 + s3);
    grad[(size_t)j * Cfull + (map ? (long long)map[slot] : slot)] = col_inv ? gsum * col_inv[j] : gsum;   // undo the column factor of the packed X (exact: a power of two)
  } else if (i < dC + C) {
    const long long slot = i - dC;
    const long long c = map ? (long long)map[slot] : slot;
    double s = lp_const;
    for (int k = 0; k < nslab_lp; ++k) s += part_lp[(size_t)k * C + slot];
    for (int j = 0; j < d; ++j) s += req[(size_t)j * Cfull + c] * xty[j];
    lp[c] = s;
  }
}
// Compaction of the tick engine: slot k of the next passes = the k-th chain (in chain order) that is still running.  Chains finish their
// iterations at very different ticks (tree depths differ by up to 2^10), so late passes would otherwise carry mostly idle rows.
// One block; count[0] = number of running chains.
__global__ void __launch_bounds__(1024) glm_compact_kernel(const double* __restrict__ sc, long long C, int* __restrict__ map, int* __restrict__ count) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  const long long per = (C + 1023) / 1024, lo = t * per, hi = lo + per < C ? lo + per : C;
  int n = 0;
  for (long long c = lo; c < hi; ++c) n += (int)sc[(size_t)SL_PHASE * C + c] != PH_DONE;
  part[t] = n;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {           // inclusive scan
    const int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int k = part[t] - n;
  for (long long c = lo; c < hi; ++c) if ((int)sc[(size_t)SL_PHASE * C + c] != PH_DONE) map[k++] = (int)c;
  if (t == 1023) count[0] = part[1023];
}
void glm_compact(const double* sc, long long C, int* map, int* count, cudaStream_t st) { glm_compact_kernel<<<1, 1024, 0, st>>>(sc, C, map, count); }
void glm_fold_tc(const double* part_lp, const float* part_g, int nslab_lp, int nslab_g, int d, long long C, const double* req,
                 const double* xty, double lp_const, double* lp, double* grad, cudaStream_t st, const int* map, long long Cfull, const double* col_inv) {
  const long long dC = (long long)d * C;
  glm_fold_tc_kernel<<<(unsigned)((dC + C + 255) / 256), 256, 0, st>>>(part_lp, part_g, nslab_lp, nslab_g, C, d, req, xty, lp_const, lp, grad, map, map ? Cfull : C, col_inv);
}
void glm_fold(const double* part_lp, const double* part_g, int nslab_lp, int nslab_g, int d, long long C, double* lp, double* grad, cudaStream_t st) {
  const long long dC = (long long)d * C;
  glm_fold2_kernel<<<(unsigned)((dC + C + 255) / 256), 256, 0, st>>>(part_lp, part_g, nslab_lp, nslab_g, C, dC, lp, grad);
}

void glm_advance(const GlmTick& t, cudaStream_t st) {
  GlmTickArgs a;
  a.C = t.C; a.chain_offset = t.chain_offset; a.seed = t.seed; a.target_iter = t.target_iter; a.burnin = t.burnin; a.thin = t.thin;
  a.row0 = t.row0; a.d = t.d; a.max_depth = t.max_depth; a.target = t.target; a.eps_desc = t.eps_desc;
  a.state = t.state; a.tune = t.tune; a.sc = t.sc; a.vec = t.vec; a.req = t.req; a.lp = t.lp; a.grad = t.grad;
  a.samples = t.samples; a.mom = t.mom; a.momn = t.momn; a.n_active = t.n_active; a.work = t.work; a.tick = t.tick;
  glm_advance_kernel<<<(unsigned)((t.C + 3) / 4), 128, 0, st>>>(a);   // one warp per chain, 4 chains per block
}

}  // namespace mcu
