// rats_fast.cu — fused kernel for the reference's own sampling scheme of the `rats` model (doc/examples/rats.jl:112-116):
//     [Slice(s2_c, 10), AMWG(alpha, 100), Slice([mu_alpha, s2_alpha], [100, 10], Univariate), AMWG(beta, 1), Slice([mu_beta, s2_beta], 1, Univariate)]
// One chain per thread, every iteration of an mcu_run call inside ONE launch (as seeds_fast.cu).
//
// What the reference does per iteration (src/samplers/amwg.jl:99-115, slice.jl:66-117 over src/model/simulation.jl:77-90): ~70 full
// block-density evaluations of 150 likelihood + 30 prior terms each.  What this kernel does:
//   * the y-likelihood enters every block only through SEE = sum_ik e_ik^2; an AMWG proposal alpha_i + z (beta_i + z) changes it by
//     -2 z A_i + 5 z^2 (-2 z B_i + z^2 sum_k x_ik^2) with A_i = sum_k e_ik, B_i = sum_k e_ik x_ik of rat i alone (5 residuals), and its
//     prior by a two-term difference: an MH test costs ~25 flops instead of a 180-term evaluation;
//   * the alpha_i (beta_i) are conditionally independent, so two components (the two draws of one Philox block) are updated per trip in
//     branch-free straight-line code, MH tests on the log scale (log u next to the draw);
//   * the Slice blocks see the rats only through sufficient statistics: s2_c through SEE; (mu_alpha, s2_alpha) through the mean c and
//     the centred sum V = sum (alpha_i - c)^2  (sum (alpha_i - mu)^2 = V + 30 (c - mu)^2, no cancellation); likewise beta.
// SEE, c, V are recomputed from the state at the start of the blocks that need them (150 residuals per iteration), so nothing drifts.
// Decisions are those of the reference on the same Philox stream up to rounding of the comparisons (tests/test_gpu_parity.py compares
// trajectories with the oracle and with the generic kernel).
// Layout: alpha, beta in shared memory [element][thread]; scalars in registers; AMWG sigma / accept in the L2-resident tune array,
// prefetched a trip ahead; rat data in the kernel-parameter constant bank, read with warp-uniform indices.
#include "fastmath.cuh"

#ifndef MCU_RATSF_BS
#define MCU_RATSF_BS 128
#endif
#ifndef MCU_RATSF_MINB
#define MCU_RATSF_MINB 3
#endif

namespace mcu {

namespace {

constexpr int NR = RatsModel::NR;   // 30 rats
constexpr int NOBS = 5;             // observations per rat (checked on the host)

struct RatsFastCfg {
  double y[NR][NOBS], x[NR][NOBS];
  double sxx[NR];                   // sum_k x_ik^2
  double w_s2c, w_a[2], w_b[2];     // Slice widths: s2_c; (mu_alpha, s2_alpha); (mu_beta, s2_beta)
  double scale_alpha[NR], scale_beta[NR];
  int adapt[2], batchsize[2], tune_off[2];   // the two AMWG blocks (alpha, beta)
  double target[2];
  double xbar;
};

// sequential uniforms of one block update (stream 0): draw k comes from Philox block k >> 1, half k & 1
struct UStream {
  const RunArgs& a; uint32_t chain, iter, block, k; Pair cur;
  MCU_D double next() {
    if (!(k & 1u)) cur = draw_uniform_pair(a, chain, iter, block, k >> 1);
    const double u = (k & 1u) ? cur.b : cur.a;
    ++k;
    return u;
  }
};


template <int BS>
__global__ void __launch_bounds__(BS, MCU_RATSF_MINB) rats_fast_kernel(const __grid_constant__ RatsFastCfg cfg, const __grid_constant__ RunArgs a) {
  extern __shared__ double smem[];
  double* sal = smem;                 // alpha[i]
  double* sbe = smem + NR * BS;       // beta[i]
  const int tid = threadIdx.x;
  const long long c = (long long)blockIdx.x * BS + tid;
  if (c >= a.n_chains) return;
  const size_t C = (size_t)a.n_chains;
  const uint32_t chain = (uint32_t)(a.chain_offset + c);
#define AL(i) sal[(i) * BS + tid]
#define BE(i) sbe[(i) * BS + tid]
#define TUNE(blk, slot) a.tune[(size_t)(cfg.tune_off[blk] + (slot)) * C + c]
  const double ig_c0 = 0.001 * log(0.001) - lgamma(0.001);
  // ---- chain state: mu_alpha, mu_beta, s2_alpha, s2_beta, s2_c, alpha[30], beta[30]
  double mua = a.state[0 * C + c], mub = a.state[1 * C + c], s2a = a.state[2 * C + c], s2b = a.state[3 * C + c], s2c = a.state[4 * C + c];
  for (int i = 0; i < NR; ++i) { AL(i) = a.state[(size_t)(5 + i) * C + c]; BE(i) = a.state[(size_t)(5 + NR + i) * C + c]; }
  double m1 = TUNE(0, 0), m3 = TUNE(1, 0);
  bool ad1 = TUNE(0, 1) != 0.0, ad3 = TUNE(1, 1) != 0.0;

  for (long long it = 1; it <= a.iters; ++it) {
    const long long iter = a.iter0 + it;
    const uint32_t it32 = (uint32_t)iter;
    if (iter == 1) {   // SamplerVariate(block, sigma): fresh AMWGTune at iter == 1 (sampler.jl:40-45, amwg.jl:14-21)
      m1 = m3 = 0.0; ad1 = ad3 = false;
      for (int i = 0; i < NR; ++i) { TUNE(0, 2 + i) = cfg.scale_alpha[i]; TUNE(0, 2 + NR + i) = 0.0; TUNE(1, 2 + i) = cfg.scale_beta[i]; TUNE(1, 2 + NR + i) = 0.0; }
    }
    // sum of squared residuals at the current state
    double SEE = 0.0;
    for (int i = 0; i < NR; ++i) {
      const double ai = AL(i), bi = BE(i);
#pragma unroll
      for (int k = 0; k < NOBS; ++k) { const double e = cfg.y[i][k] - fma(bi, cfg.x[i][k], ai); SEE = fma(e, e, SEE); }
    }
    // ================================================================== block 0: Slice(s2_c), multivariate form with one element (slice.jl:95-117)
    {
      UStream us{a, chain, it32, 0, 0, {0.0, 0.0}};
      const double n2 = 0.5 * (double)(NR * NOBS);
      auto logf = [&](double v) {   // IG(s2_c) + MvNormal(y | ., sqrt(s2_c)): simulation.jl:77-90 on the cached SEE
        if (!(v >= 0.0)) return -CUDART_INF;
        return ig_c0 - (1.001 + n2) * fast_log(v) - (0.001 + 0.5 * SEE) / v - n2 * kLog2Pi;
      };
      const double p0 = logf(s2c) + log_uniform(us.next());
      double lower = s2c - cfg.w_s2c * us.next();
      double upper = lower + cfg.w_s2c;
      double xv = cfg.w_s2c * us.next() + lower;
      while (logf(xv) < p0) {
        if (xv < s2c) lower = xv; else upper = xv;
        xv = lower + (upper - lower) * us.next();
      }
      s2c = xv;
    }
    // ================================================================== blocks 1 and 3: AMWG(alpha), AMWG(beta)
    // one body for both: WHICH = 0 updates alpha_i (prior Normal(mu_alpha, sqrt(s2_alpha)), dSEE = -2 z A_i + 5 z^2),
    //                    WHICH = 1 updates beta_i  (prior Normal(mu_beta,  sqrt(s2_beta)),  dSEE = -2 z B_i + z^2 sum_k x_ik^2)
    auto amwg_block = [&](auto which, uint32_t block, double& m, bool& ad, double mu, double s2) {
      constexpr int WHICH = decltype(which)::value;
      const bool adapt = cfg.adapt[WHICH] == 1 ? iter <= a.burnin : cfg.adapt[WHICH] == 0;
      if (adapt && !ad) { for (int i = 0; i < NR; ++i) TUNE(WHICH, 2 + NR + i) = 0.0; m = 0.0; }   // setadapt!: amwg.jl:88-96
      ad = adapt;
      if (adapt) m += 1.0;
      const double h_c = 0.5 / s2c, h_p = 0.5 / s2;
      double* tq = &TUNE(WHICH, 2);                 // sigma of component 0 of this chain: components are C doubles apart, the accept counters NR further
      const size_t acoff = (size_t)NR * C;
#pragma unroll 1
      for (int ip = 0; ip < NR / 2; ++ip, tq += 2 * C) {   // two components per trip: the two draws of Philox block ip of each stream
        const int i0 = 2 * ip, i1 = i0 + 1;
        const double sg0 = tq[0], sg1 = tq[C];             // L2 loads through a running pointer, consumed after the draws
        const double ac0 = tq[acoff], ac1 = tq[acoff + C];
        const Pair pz = draw_normal_pair(a, chain, it32, block, ip);
        const Pair pu = draw_uniform_pair(a, chain, it32, block, ip);
        const LogU lu0 = logu_bracket(pu.a), lu1 = logu_bracket(pu.b);   // float bracket of log u; FP64 log only inside the rounding band (same decisions)
        const double z0 = sg0 * pz.a, z1 = sg1 * pz.b;
        const double a0 = AL(i0), b0 = BE(i0), a1 = AL(i1), b1 = BE(i1);
        double S0 = 0.0, S1 = 0.0;                 // A_i (alpha) or B_i (beta) of the two rats
#pragma unroll
        for (int k = 0; k < NOBS; ++k) {
          const double e0 = cfg.y[i0][k] - fma(b0, cfg.x[i0][k], a0), e1 = cfg.y[i1][k] - fma(b1, cfg.x[i1][k], a1);
          if (WHICH == 0) { S0 += e0; S1 += e1; } else { S0 = fma(e0, cfg.x[i0][k], S0); S1 = fma(e1, cfg.x[i1][k], S1); }
        }
        const double q0 = WHICH == 0 ? (double)NOBS : cfg.sxx[i0], q1 = WHICH == 0 ? (double)NOBS : cfg.sxx[i1];
        const double dS0 = z0 * fma(q0, z0, -2.0 * S0), dS1 = z1 * fma(q1, z1, -2.0 * S1);   // change of SEE
        const double d0 = (WHICH == 0 ? a0 : b0) - mu, d1 = (WHICH == 0 ? a1 : b1) - mu;
        const double dl0 = -h_c * dS0 - h_p * z0 * fma(2.0, d0, z0), dl1 = -h_c * dS1 - h_p * z1 * fma(2.0, d1, z1);
        const bool acc0 = logu_less(lu0, dl0), acc1 = logu_less(lu1, dl1);   // rand() < exp(logf' - logf0): amwg.jl:107
        if (acc0) { if (WHICH == 0) AL(i0) = a0 + z0; else BE(i0) = b0 + z0; SEE += dS0; if (adapt) tq[acoff] = ac0 + 1.0; }
        if (acc1) { if (WHICH == 0) AL(i1) = a1 + z1; else BE(i1) = b1 + z1; SEE += dS1; if (adapt) tq[acoff + C] = ac1 + 1.0; }
      }
      if (adapt && ((long long)m % cfg.batchsize[WHICH]) == 0) {   // amwg.jl:74-80
        const double dl = amwg_delta(m, cfg.batchsize[WHICH]);
        const double up = exp(dl), dn = exp(-dl);
        for (int i = 0; i < NR; ++i) TUNE(WHICH, 2 + i) = TUNE(WHICH, 2 + i) * ((TUNE(WHICH, 2 + NR + i) / m < cfg.target[WHICH]) ? dn : up);
      }
    };
    // Slice([mu, s2], [w0, w1], Univariate) on the constrained scale (slice.jl:66-92): Normal(0, 1000) prior of mu, IG prior of s2 and the
    // 30 Normal(mu, sqrt(s2)) terms through their mean cen and centred sum V
    auto slice_mu_s2 = [&](uint32_t block, double& mu, double& s2, double cen, double V, const double* w) {
      UStream us{a, chain, it32, block, 0, {0.0, 0.0}};
      auto logf = [&](double vm, double vs) {
        double lp = -(vm * vm * 1e-6 + kLog2Pi) / 2.0 - 6.907755278982137;     // lp_normal(mu, 0, 1000): log(1000)
        if (!(vs >= 0.0)) return -CUDART_INF;
        const double dc = cen - vm;
        lp += ig_c0 - (1.001 + 0.5 * NR) * fast_log(vs) - (0.001 + 0.5 * fma((double)NR * dc, dc, V)) / vs - 0.5 * NR * kLog2Pi;
        return lp;
      };
      double logf0 = logf(mu, s2);
      double lo0 = mu - w[0] * us.next(), lo1 = s2 - w[1] * us.next();
      double up0 = lo0 + w[0], up1 = lo1 + w[1];
      {   // component 1: mu
        const double p0 = logf0 + log_uniform(us.next());
        const double x0 = mu;
        double cur = lo0 + (up0 - lo0) * us.next();
        while (true) {
          logf0 = logf(cur, s2);
          if (!(logf0 < p0)) break;
          if (cur < x0) lo0 = cur; else up0 = cur;
          cur = lo0 + (up0 - lo0) * us.next();
        }
        mu = cur;
      }
      {   // component 2: s2
        const double p0 = logf0 + log_uniform(us.next());
        const double x0 = s2;
        double cur = lo1 + (up1 - lo1) * us.next();
        while (true) {
          logf0 = logf(mu, cur);
          if (!(logf0 < p0)) break;
          if (cur < x0) lo1 = cur; else up1 = cur;
          cur = lo1 + (up1 - lo1) * us.next();
        }
        s2 = cur;
      }
    };
    auto centred = [&](const double* arr, double& cen, double& V) {
      double s = 0.0;
      for (int i = 0; i < NR; ++i) s += arr[i * BS + tid];
      cen = s / (double)NR;
      double v = 0.0;
      for (int i = 0; i < NR; ++i) { const double d = arr[i * BS + tid] - cen; v = fma(d, d, v); }
      V = v;
    };
    amwg_block(std::integral_constant<int, 0>{}, 1u, m1, ad1, mua, s2a);
    { double cen, V; centred(sal, cen, V); slice_mu_s2(2u, mua, s2a, cen, V, cfg.w_a); }
    amwg_block(std::integral_constant<int, 1>{}, 3u, m3, ad3, mub, s2b);
    { double cen, V; centred(sbe, cen, V); slice_mu_s2(4u, mub, s2b, cen, V, cfg.w_b); }
    // ================================================================== thinning + streaming moments (mcmc.jl:76-78)
    if (iter > a.burnin && (iter - a.burnin) % a.thin == 0) {
      double mon[RatsModel::P];
      mon[0] = mub; mon[1] = mua - cfg.xbar * mub; mon[2] = s2c;   // mu_beta, alpha0 (rats.jl:64-66), s2_c
      if (a.samples) {
        const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
        for (int j = 0; j < RatsModel::P; ++j) a.samples[((size_t)row * RatsModel::P + j) * C + c] = mon[j];
      }
      if (a.comom) comoments_update(a.mom, a.momn, a.comom, C, (size_t)c, RatsModel::P, mon, a.log_mask);
      moments_update(a.mom, a.momn, C, (size_t)c, RatsModel::P, mon);
    }
  }
  // ---- store chain state and tune
  a.state[0 * C + c] = mua; a.state[1 * C + c] = mub; a.state[2 * C + c] = s2a; a.state[3 * C + c] = s2b; a.state[4 * C + c] = s2c;
  for (int i = 0; i < NR; ++i) { a.state[(size_t)(5 + i) * C + c] = AL(i); a.state[(size_t)(5 + NR + i) * C + c] = BE(i); }
  TUNE(0, 0) = m1; TUNE(0, 1) = ad1 ? 1.0 : 0.0; TUNE(1, 0) = m3; TUNE(1, 1) = ad3 ? 1.0 : 0.0;
#undef AL
#undef BE
#undef TUNE
}

}  // namespace

// h_blocks: the five DevBlocks of the scheme; h_scales: their expanded host-side scales.  Returns 0 on success, -2 when the data are not
// 5 observations per rat (the generic kernel takes over).
int rats_fast_launch(const double* y, const double* Xm, const double* rat, int N, double xbar, const RunArgs& a, const DevBlock* h_blocks,
                     const std::vector<std::vector<double>>& h_scales, cudaStream_t st) {
  RatsFastCfg cfg;
  if (N != NR * NOBS) return -2;
  int cnt[NR] = {0};
  for (int k = 0; k < N; ++k) {
    const int i = (int)rat[k];
    if (i < 0 || i >= NR || cnt[i] >= NOBS) return -2;
    cfg.y[i][cnt[i]] = y[k]; cfg.x[i][cnt[i]] = Xm[k]; ++cnt[i];
  }
  for (int i = 0; i < NR; ++i) {
    if (cnt[i] != NOBS) return -2;
    cfg.sxx[i] = 0.0; for (int k = 0; k < NOBS; ++k) cfg.sxx[i] += cfg.x[i][k] * cfg.x[i][k];
    cfg.scale_alpha[i] = h_scales[1][i]; cfg.scale_beta[i] = h_scales[3][i];
  }
  cfg.w_s2c = h_scales[0][0];
  cfg.w_a[0] = h_scales[2][0]; cfg.w_a[1] = h_scales[2][1];
  cfg.w_b[0] = h_scales[4][0]; cfg.w_b[1] = h_scales[4][1];
  const int amwg_idx[2] = {1, 3};
  for (int b = 0; b < 2; ++b) {
    const DevBlock& blk = h_blocks[amwg_idx[b]];
    cfg.adapt[b] = blk.adapt; cfg.batchsize[b] = blk.batchsize; cfg.tune_off[b] = blk.tune_off; cfg.target[b] = blk.target;
  }
  cfg.xbar = xbar;
  constexpr int BS = MCU_RATSF_BS;
  const size_t smem = (size_t)BS * 2 * NR * sizeof(double);
  if (cudaFuncSetAttribute(rats_fast_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  rats_fast_kernel<BS><<<(unsigned)((a.n_chains + BS - 1) / BS), BS, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mcu
