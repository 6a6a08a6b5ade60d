// seeds_fast.cu — fused kernel for the headline configuration: the `seeds` random-effects logistic
// model (doc/examples/seeds.jl:16-56) under the all-AMWG scheme
//     [AMWG(alpha0, alpha1, alpha2, alpha12), AMWG(b), AMWG(s2)]           
// One chain per thread, every iteration of an mcu_run call inside ONE launch, nothing but the thinned
// output and the streaming moments touches HBM.
//
// What the reference does per iteration (src/samplers/amwg.jl:99-115 over src/model/simulation.jl:77-90):
// 29 full block-density evaluations, each re-running every node closure.  What this kernel does:
//   * the per-plate binomial-logit terms ll_i are cached; an AMWG proposal re-evaluates only the plates
//     whose linear predictor changes (alpha0: 21, alpha1: 10, alpha2: 11, alpha12: 5, b_i: 1) and the MH
//     ratio is formed from term differences — 68 term evaluations per iteration instead of 29 x 21;
//   * the (x1, x2) design has four distinct rows, so eta_i = g[group_i] + b_i with four cached group
//     bases, evaluated in the reference's summation order;
//   * the s2 update uses the sufficient statistic sum b_i^2;
//   * e_i = exp(eta_i) and L_i = log(1 + e_i) are cached per plate: an alpha proposal shifts every affected eta_i by
//     the same step z, so e_i' = e_i * exp(z) needs ONE exp per proposal and one log per affected plate
//     (ll_i' - ll_i = r_i z - n_i (L_i' - L_i)); a b_i proposal recomputes e_i = exp(eta_i) afresh, which also stops
//     the rounding drift of the multiplicative updates;
//   * exp / log are own FP64 routines (fdlibm-style reductions, coefficients in __constant__ memory so the DFMAs take
//     them as constant-bank operands instead of two UMOVs each — 20 % of all issued instructions before).
// The accept/reject decisions are those of the reference on the same uniform stream: the same draws
// (Philox counter j = position of the draw inside the block update, rng.cuh), the same proposal, and a
// log-ratio equal to logf(x') - logf(x) up to rounding (~1e-14; tests/test_gpu_parity.py compares
// whole trajectories against the oracle and against the generic kernel).
//
// Layout: alpha, log s2 and their tune state live in registers; b, e, L and the proposed L live in shared memory
// as [element][thread] (conflict-free 8-byte lanes); sigma_b and the b accept counters stay in the (L2-resident)
// tune array and are touched once per plate per iteration; plate
// constants sit in the kernel-parameter constant bank and are read with warp-uniform indices.
// FP64 throughout (the reference is Float64; a decision taken in FP32 would flip ~1e-7 of the time).
#include <type_traits>

#ifndef MCU_SEEDS_ESTRIN
#define MCU_SEEDS_ESTRIN 0
#endif
#define MCU_FASTMATH_ESTRIN MCU_SEEDS_ESTRIN   // before the first include: fastfn.cuh comes in through launch.hpp
#ifndef MCU_LOGU_DOUBLE
#define MCU_LOGU_DOUBLE 1                      // fastmath.cuh: the float bracket of log u is widened to doubles next to the draw (1 % faster here)
#endif
#include "launch.hpp"

#ifndef MCU_SEEDS_BW
#define MCU_SEEDS_BW 2      // plates per trip in the b block (even)
#endif
#ifndef MCU_SEEDS_LOGU
#define MCU_SEEDS_LOGU 2   // MH test on the log scale: 1 = FP64 log u next to the draw (off the critical path); 2 = float bracket of log u + FP64 log inside the
                           // rounding band only (same decisions; 3-7 % slower than 1 with the fdlibm log in round 1, 3 % faster with the table log)
#endif
#ifndef MCU_SEEDS_BS
#define MCU_SEEDS_BS 96
#endif
#ifndef MCU_SEEDS_MINB
#define MCU_SEEDS_MINB 3
#endif
// Switches that were measured and removed again (profiles/r2_seeds_fast_summary.md has the numbers; the code is in the history at commit e00a803 / 12620f6):
// proposal scratch and b in L2 for 14-16 resident warps, draws of the next b trip generated during the current one, all draws of the alpha / s2 blocks up
// front, out-of-line draw functions, table-driven sincos, two threads per chain, producer / consumer warp specialisation.
#ifndef MCU_SEEDS_DEDUP
#define MCU_SEEDS_DEDUP 0   // 1: a single loop over the b trips (one copy of the trip body): measured 12 % SLOWER (every trip then carries the dynamic last-plate guards)
#endif
#ifndef MCU_SEEDS_FSQRT
#define MCU_SEEDS_FSQRT 1   // Box-Muller radius through the branch-free fast_sqrt (fastfn.cuh)
#endif
#ifndef MCU_SEEDS_AW
#define MCU_SEEDS_AW 3      // plates per trip of an alpha proposal (the plate lists are padded with the dummy slot to a multiple of it)
#endif
#ifndef MCU_SEEDS_PF
#define MCU_SEEDS_PF 1   // b block: sigma_b / accept counters of trip t + 1 are fetched (L2) at the top of trip t
#endif
#ifndef MCU_SEEDS_TAB
#define MCU_SEEDS_TAB 1    // table-driven log / exp (fasttab_fn.cuh), tables staged in shared memory; 0 = the fdlibm forms of fastfn.cuh
#endif

#include "fastmath.cuh"
#if MCU_SEEDS_TAB
#include "fasttab.cuh"
#include "fasttab_fn.cuh"
#endif

namespace mcu {

namespace {

constexpr int NPL = SeedsModel::NP;   // 21 plates

constexpr int NSL = NPL + 1;          // shared-memory slots per array: 21 plates + one dummy (e = 0, L = 0, n = 0) that pads the plate lists

struct FastCfg {
  double r[NPL], n[NSL];
  unsigned char alist[4][24];         // plates whose eta depends on alpha_j, padded with the dummy slot to a multiple of MCU_SEEDS_AW
  int atriples[4];                    // trips of MCU_SEEDS_AW plates
  double rsum[4];                     // sum of r_i over the plates that depend on alpha_j
  unsigned char grp[NPL];             // 0:(x1=0,x2=0) 1:(0,1) 2:(1,0) 3:(1,1)
  unsigned amask[4];                  // plates whose eta depends on alpha_j
  unsigned gmask[4];                  // groups whose base depends on alpha_j
  int adapt[3], batchsize[3], tune_off[3];
  double target[3];
  double scale_a[4], scale_b[NPL], scale_s;
  // block 0 = AMM(alpha0..alpha12) (doc/examples/seeds.jl:69): lower Cholesky factor of the initial Sigma (column-major), beta, scale
  double amm_SL[16], amm_beta, amm_scale;
};

#if !MCU_SEEDS_LOGU   // the u < exp(delta) form of the MH test (MCU_SEEDS_LOGU = 0)
MCU_D bool mh_accept(double u, double delta) {   // rand() < exp(logfprime - logf0): amwg.jl:107
  if (delta >= 0.0) return true;                  // u < 1 <= exp(delta)
  if (!(delta > -700.0)) return false;            // exp underflows (or delta is NaN): u < 0 never holds
  return u < fast_exp(delta);
}
// the same decision without branches (the exp is always evaluated), so that two or three independent updates can be
// scheduled into each other's dependency stalls
MCU_D bool mh_accept_nb(double u, double delta) {
  const double ex = fast_exp(fmax(fmin(delta, 0.0), -700.0));
  return delta >= 0.0 ? true : (delta > -700.0 && u < ex);
}
#endif

struct Bases { double g0, g1, g2, g3; };
MCU_D Bases group_bases(double a0, double a1, double a2, double a12) {
  // alpha0 + alpha1*x1 + alpha2*x2 + alpha12*x1*x2 in the reference's order (seeds.jl:22-23)
  Bases g;
  g.g0 = a0;
  g.g1 = a0 + a2;
  g.g2 = a0 + a1;
  g.g3 = ((a0 + a1) + a2) + a12;
  return g;
}
MCU_D double pick(const Bases& g, unsigned grp) {   // warp-uniform select, keeps the bases in registers
  const double lo = (grp & 1u) ? g.g1 : g.g0, hi = (grp & 1u) ? g.g3 : g.g2;
  return (grp & 2u) ? hi : lo;
}


// AMM0: block 0 is AMM(alpha0, alpha1, alpha2, alpha12) (the reference's scheme, doc/examples/seeds.jl:69-71) instead of AMWG
template <int BS, bool AMM0>
__global__ void __launch_bounds__(BS, MCU_SEEDS_MINB) seeds_fast_kernel(const __grid_constant__ FastCfg cfg, const __grid_constant__ RunArgs a) {
  extern __shared__ double smem[];
  double* se = smem;                        // e[i] = exp(eta_i)
  double* sll = smem + NSL * BS;            // L[i] = log(1 + e[i])
  double* sb = smem + 2 * NSL * BS;         // b[i]
  double* sln = smem + 3 * NSL * BS;   // proposed L[i]
  const int tid = threadIdx.x;
  const long long c = (long long)blockIdx.x * BS + tid;
#if MCU_SEEDS_TAB
  double* tlg = smem + 4 * NSL * BS;   // 128 x (invc, logc), 16-byte aligned
  double* tex = tlg + 256;                                              // 128 x 2^(j/128)
  for (int i = tid; i < 256; i += BS) tlg[i] = kLogTabG[i];
  for (int i = tid; i < 128; i += BS) tex[i] = kExpTabG[i];
  __syncthreads();
#define FLOG(x) tab::tlog((x), tlg)
#define FEXP(x) tab::texp((x), tex)
#else
#define FLOG(x) fast_log(x)
#define FEXP(x) fast_exp(x)
#endif
  if (c >= a.n_chains) return;
#if !MCU_SEEDS_LOGU
  auto mh_accept = [&](double u, double delta) { return delta >= 0.0 ? true : (delta > -700.0 ? u < FEXP(delta) : false); };
  auto mh_accept_nb = [&](double u, double delta) { const double ex = FEXP(fmax(fmin(delta, 0.0), -700.0)); return delta >= 0.0 ? true : (delta > -700.0 && u < ex); };
#endif
  // the draws of fastmath.cuh with the kernel's own log (same Philox blocks, same arithmetic otherwise)
  auto log_uniform = [&](double u) { return u > 0.0 ? FLOG(u) : -CUDART_INF; };
  auto draw_logu_pair = [&](uint32_t ch, uint32_t itn, uint32_t blk, uint32_t kpair) -> Pair {   // logs of both uniforms of a Philox block
    const Pair pu = draw_uniform_pair(a, ch, itn, blk, kpair);
    return {log_uniform(pu.a), log_uniform(pu.b)};
  };
  auto draw_normal_pair = [&](const RunArgs& aa, uint32_t ch, uint32_t itn, uint32_t blk, uint32_t kpair) -> Pair {
    uint32_t w[4];
    philox4x32_10(kpair, itn, ch, blk | (1u << 24), (uint32_t)aa.seed, (uint32_t)(aa.seed >> 32), w);
#if MCU_SEEDS_FSQRT
    const double rad = fast_sqrt(-2.0 * FLOG(1.0 - u53(w[0], w[1])));
#else
    const double rad = sqrt(-2.0 * FLOG(1.0 - u53(w[0], w[1])));
#endif
    const Pair sc = fast_sincos2pi(u53(w[2], w[3]));
    return {rad * sc.b, rad * sc.a};
  };
  const size_t C = (size_t)a.n_chains;
  const uint32_t chain = (uint32_t)(a.chain_offset + c);
#define SB(i) sb[(i) * BS + tid]
#define SE(i) se[(i) * BS + tid]
#define SLL(i) sll[(i) * BS + tid]
#define SLN(i) sln[(i) * BS + tid]
#define SSG(i) TUNE(1, 2 + (i))
#define SAC(i) TUNE(1, 2 + NPL + (i))
#define TUNE(blk, slot) a.tune[(size_t)(cfg.tune_off[blk] + (slot)) * C + c]

  // ---- load chain state -------------------------------------------------------------------------
  double al0 = a.state[0 * C + c], al1 = a.state[1 * C + c], al2 = a.state[2 * C + c], al3 = a.state[3 * C + c];
  double s2 = a.state[4 * C + c];
  double x = log(s2);
  for (int i = 0; i < NPL; ++i) SB(i) = a.state[(size_t)(5 + i) * C + c];
  // tune: block 0 [m, adapt, sigma[4], accept[4]]; block 1 [m, adapt, sigma[21], accept[21]]; block 2 [m, adapt, sigma, accept]
  // AMM tune record of block 0 (samplers.cuh): [adapt, m, Mv[4], Mvv[16], SigmaLm[16]] — it stays in the L2-resident tune array
  double m0 = AMM0 ? 0.0 : TUNE(0, 0), m1 = TUNE(1, 0), m2 = TUNE(2, 0);
  bool ad0 = AMM0 ? false : TUNE(0, 1) != 0.0, ad1 = TUNE(1, 1) != 0.0, ad2 = TUNE(2, 1) != 0.0;
  double sg0 = AMM0 ? 0.0 : TUNE(0, 2), sg1 = AMM0 ? 0.0 : TUNE(0, 3), sg2 = AMM0 ? 0.0 : TUNE(0, 4), sg3 = AMM0 ? 0.0 : TUNE(0, 5);
  int ac0 = AMM0 ? 0 : (int)TUNE(0, 6), ac1 = AMM0 ? 0 : (int)TUNE(0, 7), ac2 = AMM0 ? 0 : (int)TUNE(0, 8), ac3 = AMM0 ? 0 : (int)TUNE(0, 9);
  double sgs = TUNE(2, 2); int acs = (int)TUNE(2, 3);

  Bases g = group_bases(al0, al1, al2, al3);
  for (int i = 0; i < NPL; ++i) { const double e = FEXP(pick(g, cfg.grp[i]) + SB(i)); SE(i) = e; SLL(i) = FLOG(1.0 + e); }
  SB(NPL) = 0.0;
  SE(NPL) = 0.0; SLL(NPL) = 0.0; SLN(NPL) = 0.0;   // dummy slot: log(1 + 0 * E) = 0, n = 0

  double mon[SeedsModel::P];
  for (long long it = 1; it <= a.iters; ++it) {
    const long long iter = a.iter0 + it;
    const uint32_t it32 = (uint32_t)iter;
    if (iter == 1) {   // SamplerVariate(block, sigma): fresh AMWGTune at iter == 1 (sampler.jl:40-45, amwg.jl:14-21)
      m0 = m1 = m2 = 0.0; ad0 = ad1 = ad2 = false;
      sg0 = cfg.scale_a[0]; sg1 = cfg.scale_a[1]; sg2 = cfg.scale_a[2]; sg3 = cfg.scale_a[3];
      ac0 = ac1 = ac2 = ac3 = 0;
      for (int i = 0; i < NPL; ++i) { SSG(i) = cfg.scale_b[i]; SAC(i) = 0.0; }
      sgs = cfg.scale_s; acs = 0;
      if (AMM0) for (int i = 0; i < 38; ++i) TUNE(0, i) = 0.0;   // AMMTune(x, Sigma): amm.jl:14-24
    }
    // ================================================================== block 0, AMM form (amm.jl:66-108; device generic form: samplers.cuh amm_sample)
    if (AMM0) {
      const bool adapt = cfg.adapt[0] == 1 ? iter <= a.burnin : cfg.adapt[0] == 0;
      const bool was = TUNE(0, 0) != 0.0;
      double v[4] = {al0, al1, al2, al3};
      if (adapt && !was) {   // setadapt!: amm.jl:97-108
        TUNE(0, 1) = 0.0;
        for (int i = 0; i < 4; ++i) TUNE(0, 2 + i) = v[i];
        for (int i = 0; i < 4; ++i) for (int cc = 0; cc < 4; ++cc) TUNE(0, 6 + i + cc * 4) = v[i] * v[cc];
        for (int i = 0; i < 16; ++i) TUNE(0, 22 + i) = 0.0;
      }
      TUNE(0, 0) = adapt ? 1.0 : 0.0;
      double m = TUNE(0, 1);
      // x = SigmaL randn(4) [mixed with the adapted factor once m > 2n] + v: normals 0..3 (and 4..7) of the block, then one uniform
      const Pair z01 = draw_normal_pair(a, chain, it32, 0, 0), z23 = draw_normal_pair(a, chain, it32, 0, 1);
      const double z[4] = {z01.a, z01.b, z23.a, z23.b};
      double xp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { double acc = 0.0; for (int cc = 0; cc <= i; ++cc) acc += cfg.amm_SL[i + cc * 4] * z[cc]; xp[i] = acc; }
      if (m > 8.0) {
        const Pair y01 = draw_normal_pair(a, chain, it32, 0, 2), y23 = draw_normal_pair(a, chain, it32, 0, 3);
        const double z2[4] = {y01.a, y01.b, y23.a, y23.b};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double acc = 0.0; for (int cc = 0; cc < 4; ++cc) acc += TUNE(0, 22 + i + cc * 4) * z2[cc];
          xp[i] = cfg.amm_beta * xp[i] + (1.0 - cfg.amm_beta) * acc;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) xp[i] += v[i];
      const double lu = log_uniform(draw_uniform_pair(a, chain, it32, 0, 0).a);
      // logf(x) - logf(v): every plate of group q moves by dg[q]; e_i' = e_i exp(dg[grp_i]) (4 exps, 21 logs), priors Normal(0, 1000)
      const Bases gn = group_bases(xp[0], xp[1], xp[2], xp[3]);
      const double dg0 = gn.g0 - g.g0, dg1 = gn.g1 - g.g1, dg2 = gn.g2 - g.g2, dg3 = gn.g3 - g.g3;
      const double E0 = FEXP(dg0), E1 = FEXP(dg1), E2 = FEXP(dg2), E3 = FEXP(dg3);
      double delta = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) delta = fma(-0.5e-6, fma(xp[i], xp[i], -v[i] * v[i]), delta);
      double dLa = 0.0, dLb = 0.0, dLc = 0.0;
#pragma unroll 1
      for (int i = 0; i < NPL; i += 3) {   // 21 plates = 7 trips of three independent logs
        const unsigned q0 = cfg.grp[i], q1 = cfg.grp[i + 1], q2 = cfg.grp[i + 2];
        const double Ea = (q0 & 2u) ? ((q0 & 1u) ? E3 : E2) : ((q0 & 1u) ? E1 : E0), da_ = (q0 & 2u) ? ((q0 & 1u) ? dg3 : dg2) : ((q0 & 1u) ? dg1 : dg0);
        const double Eb = (q1 & 2u) ? ((q1 & 1u) ? E3 : E2) : ((q1 & 1u) ? E1 : E0), db_ = (q1 & 2u) ? ((q1 & 1u) ? dg3 : dg2) : ((q1 & 1u) ? dg1 : dg0);
        const double Ec = (q2 & 2u) ? ((q2 & 1u) ? E3 : E2) : ((q2 & 1u) ? E1 : E0), dc_ = (q2 & 2u) ? ((q2 & 1u) ? dg3 : dg2) : ((q2 & 1u) ? dg1 : dg0);
        const double la = FLOG(fma(SE(i), Ea, 1.0)), lb = FLOG(fma(SE(i + 1), Eb, 1.0)), lc = FLOG(fma(SE(i + 2), Ec, 1.0));
        SLN(i) = la; SLN(i + 1) = lb; SLN(i + 2) = lc;
        dLa += fma(cfg.r[i], da_, -cfg.n[i] * (la - SLL(i)));
        dLb += fma(cfg.r[i + 1], db_, -cfg.n[i + 1] * (lb - SLL(i + 1)));
        dLc += fma(cfg.r[i + 2], dc_, -cfg.n[i + 2] * (lc - SLL(i + 2)));
      }
      delta += (dLa + dLb) + dLc;
      if (lu < delta) {   // rand() < exp(logf(x) - logf(v)): amm.jl:80
        al0 = xp[0]; al1 = xp[1]; al2 = xp[2]; al3 = xp[3];
        v[0] = xp[0]; v[1] = xp[1]; v[2] = xp[2]; v[3] = xp[3];
        g = gn;
        for (int i = 0; i < NPL; ++i) {
          const unsigned q = cfg.grp[i];
          SLL(i) = SLN(i);
          SE(i) = SE(i) * ((q & 2u) ? ((q & 1u) ? E3 : E2) : ((q & 1u) ? E1 : E0));
        }
      }
      if (adapt) {   // running mean / second moment, Sigma = (scale^2 / n / p)(Mvv - Mv Mv'), pivoted Cholesky: amm.jl:83-91
        m += 1.0; TUNE(0, 1) = m;
        const double p = m / (m + 1.0);
        double Sigma[16], PL[16], Mv[4];
        for (int i = 0; i < 4; ++i) { Mv[i] = p * TUNE(0, 2 + i) + (1.0 - p) * v[i]; TUNE(0, 2 + i) = Mv[i]; }
        const double c0 = cfg.amm_scale * cfg.amm_scale / 4.0 / p;
        for (int i = 0; i < 4; ++i) for (int cc = 0; cc < 4; ++cc) {
          const double mvv = p * TUNE(0, 6 + i + cc * 4) + (1.0 - p) * v[i] * v[cc];
          TUNE(0, 6 + i + cc * 4) = mvv;
          Sigma[i + cc * 4] = c0 * (mvv - Mv[i] * Mv[cc]);
        }
        if (pivoted_chol_PL(Sigma, 4, PL) == 4) for (int i = 0; i < 16; ++i) TUNE(0, 22 + i) = PL[i];
      }
    }
    // ================================================================== block 0: AMWG(alpha0..alpha12)
    if (!AMM0) {
      const bool adapt = cfg.adapt[0] == 1 ? iter <= a.burnin : cfg.adapt[0] == 0;
      if (adapt && !ad0) { ac0 = ac1 = ac2 = ac3 = 0; m0 = 0.0; }   // setadapt!: amwg.jl:88-96
      ad0 = adapt;
      if (adapt) m0 += 1.0;
      // components are rotated through slot 0 so the loop stays rolled with everything in registers
      double zc = 0.0, uc = 0.0;   // second draw of the current Philox pair
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        double zn01;
        if ((j & 1) == 0) { const Pair pr = draw_normal_pair(a, chain, it32, 0, j >> 1); zn01 = pr.a; zc = pr.b; } else zn01 = zc;
#if MCU_SEEDS_LOGU == 2
        LogU lu;                                                          // bracket of log(uniform j), off the critical path
        if ((j & 1) == 0) { const Pair pr = draw_uniform_pair(a, chain, it32, 0, j >> 1); lu = logu_bracket(pr.a); uc = pr.b; } else lu = logu_bracket(uc);
#elif MCU_SEEDS_LOGU
        double lu;                                                        // log of uniform j of the block, off the critical path
        if ((j & 1) == 0) { const Pair pr = draw_logu_pair(chain, it32, 0, j >> 1); lu = pr.a; uc = pr.b; } else lu = uc;
#endif
        const double z = sg0 * zn01;                                      // z = sigma .* randn(n): normal j of the block
        const double anew = al0 + z;
        const unsigned pm = cfg.amask[j];
        // proposed group bases, again in the reference's summation order (slot s holds alpha_{(j+s)%4})
        const double q0 = j == 0 ? anew : (j == 1 ? al3 : (j == 2 ? al2 : al1));
        const double q1 = j == 0 ? al1 : (j == 1 ? anew : (j == 2 ? al3 : al2));
        const double q2 = j == 0 ? al2 : (j == 1 ? al1 : (j == 2 ? anew : al3));
        const double q3 = j == 0 ? al3 : (j == 1 ? al2 : (j == 2 ? al1 : anew));
        const Bases gn = group_bases(q0, q1, q2, q3);
        // every affected plate moves by the same step: e_i' = e_i exp(z); ll_i' - ll_i = r_i z - n_i (L_i' - L_i)
        const double E = FEXP(z);
        // AW plates per trip: their logs are independent, so the scheduler fills one chain's DFMA latency with the others
        constexpr int AW = MCU_SEEDS_AW;
        double dL[AW];
#pragma unroll
        for (int w = 0; w < AW; ++w) dL[w] = 0.0;
        const int nt = cfg.atriples[j];
#pragma unroll 1
        for (int k = 0; k < nt; ++k) {
          int ii[AW]; double ln[AW];
#pragma unroll
          for (int w = 0; w < AW; ++w) ii[w] = cfg.alist[j][AW * k + w];
#pragma unroll
          for (int w = 0; w < AW; ++w) ln[w] = FLOG(fma(SE(ii[w]), E, 1.0));
#pragma unroll
          for (int w = 0; w < AW; ++w) { SLN(ii[w]) = ln[w]; dL[w] = fma(cfg.n[ii[w]], ln[w] - SLL(ii[w]), dL[w]); }
        }
        double dLs = dL[0];
#pragma unroll
        for (int w = 1; w < AW; ++w) dLs += dL[w];
        double delta = fma(cfg.rsum[j], z, -dLs);
        {   // Normal(0, 1000) prior of the component: -(z^2 + log 2pi)/2 - log sigma
          delta = fma(-0.5e-6, fma(anew, anew, -al0 * al0), delta);   // (x / 1000)^2 / 2 without the divisions
        }
#if MCU_SEEDS_LOGU == 2
        if (logu_less(lu, delta)) {                                       // rand() < exp(delta): float bracket, FP64 log only inside the rounding band
#elif MCU_SEEDS_LOGU
        if (lu < delta) {                                                 // rand() < exp(delta) on the log scale; lu was formed next to the draw
#else
        double u;                                                         // uniform j of the block
        if ((j & 1) == 0) { const Pair pr = draw_uniform_pair(a, chain, it32, 0, j >> 1); u = pr.a; uc = pr.b; } else u = uc;
        if (mh_accept(u, delta)) {
#endif
          al0 = anew;
          g = gn;   // bases of groups that do not contain alpha_j are recomputed to the same value
          for (int i = 0; i < NPL; ++i) if ((pm >> i) & 1u) { SLL(i) = SLN(i); SE(i) = SE(i) * E; }
          if (adapt) ac0 += 1;
        }
        // rotate (alpha, sigma, accept) so the next component sits in slot 0
        { const double t = al0; al0 = al1; al1 = al2; al2 = al3; al3 = t; }
        { const double t = sg0; sg0 = sg1; sg1 = sg2; sg2 = sg3; sg3 = t; }
        { const int t = ac0; ac0 = ac1; ac1 = ac2; ac2 = ac3; ac3 = t; }
      }
      if (adapt && ((long long)m0 % cfg.batchsize[0]) == 0) {
        const double dl = amwg_delta(m0, cfg.batchsize[0]);
        sg0 *= exp((double)ac0 / m0 < cfg.target[0] ? -dl : dl);
        sg1 *= exp((double)ac1 / m0 < cfg.target[0] ? -dl : dl);
        sg2 *= exp((double)ac2 / m0 < cfg.target[0] ? -dl : dl);
        sg3 *= exp((double)ac3 / m0 < cfg.target[0] ? -dl : dl);
      }
    }
    // ================================================================== block 1: AMWG(b)
    {
      const bool adapt = cfg.adapt[1] == 1 ? iter <= a.burnin : cfg.adapt[1] == 0;
      if (adapt && !ad1) { for (int i = 0; i < NPL; ++i) SAC(i) = 0.0; m1 = 0.0; }
      ad1 = adapt;
      if (adapt) m1 += 1.0;
      const double half_inv_s2 = 0.5 / s2;                                // b ~ Normal(0, sqrt(s2)): -(b/sigma)^2 / 2 = -b^2 / (2 s2)
      // The b_i are conditionally independent given alpha and s2, so W plates (the draws of W / 2 Philox pairs) are updated
      // per trip in straight-line code: W independent exp → log → exp chains for the scheduler to interleave.
#if MCU_SEEDS_PF
      // The tune array is L2-resident (L1 is all shared memory here): ~700 cycles per load.  With the table-driven log / exp a trip is too
      // short to hide that behind its own draws, so the loads run one trip ahead.
      // Addresses: one running pointer per chain (plates are C doubles apart, the accept counters NPL plates after the sigmas) instead of a
      // 64-bit index computation per load.  The fetch for the trip after the last one reads up to three slots past sigma_b / accept of this
      // block: they are the accept slots of plates 0..2 and the four slots of block 2's record (seeds_fast_launch checks the layout), never used.
      static_assert(MCU_SEEDS_BW == 2, "the look-ahead fetch stays inside block 2's four tune slots only for two plates per trip");
      double* tq = &TUNE(1, 2);
      const size_t acoff = (size_t)NPL * C;
      double psg[MCU_SEEDS_BW], pac[MCU_SEEDS_BW];
      auto b_fetch = [&](const double* q) {
#pragma unroll
        for (int w = 0; w < MCU_SEEDS_BW; ++w) { psg[w] = q[(size_t)w * C]; pac[w] = q[acoff + (size_t)w * C]; }
      };
      b_fetch(tq);
#endif
      auto b_trip = [&](auto Wc, int i0) {
        constexpr int W = decltype(Wc)::value;
        int ix[W]; double sg[W], bi[W], zn[W], ac[W]; [[maybe_unused]] double uu[W];
#if MCU_SEEDS_LOGU == 2
        LogU lb[W];
#endif
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const bool real = i0 + w < NPL;
          ix[w] = real ? i0 + w : NPL;                                     // past the last plate: the dummy slot (never accepted)
#if MCU_SEEDS_PF
          sg[w] = psg[w]; ac[w] = pac[w];
#else
          sg[w] = real ? SSG(ix[w]) : 0.0;                                 // global (L2) loads at the top of the trip
          ac[w] = (real && adapt) ? SAC(ix[w]) : 0.0;
#endif
          bi[w] = SB(ix[w]);
        }
#if MCU_SEEDS_PF
        double* const tc = tq;                                             // this trip's plates
        tq += (size_t)W * C;
        b_fetch(tq);
#endif
#pragma unroll
        for (int w = 0; w < W; w += 2) {
          const Pair pz = draw_normal_pair(a, chain, it32, 1, (i0 + w) >> 1);
#if MCU_SEEDS_LOGU == 1
          const Pair pu = draw_logu_pair(chain, it32, 1, (i0 + w) >> 1);
#else
          const Pair pu = draw_uniform_pair(a, chain, it32, 1, (i0 + w) >> 1);
#endif
#if MCU_SEEDS_LOGU == 2
          zn[w] = pz.a; zn[w + 1] = pz.b; lb[w] = logu_bracket(pu.a); lb[w + 1] = logu_bracket(pu.b);
#elif MCU_SEEDS_LOGU
          zn[w] = pz.a; zn[w + 1] = pz.b; uu[w] = pu.a; uu[w + 1] = pu.b;
#else
          zn[w] = pz.a; zn[w + 1] = pz.b; uu[w] = pu.a; uu[w + 1] = pu.b;
#endif
        }
        double bn[W], en[W], ln[W]; bool acc[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const int ir = i0 + w < NPL ? i0 + w : 0;                        // constants of a real plate for the dummy chain
          bn[w] = bi[w] + sg[w] * zn[w];
          en[w] = FEXP(pick(g, cfg.grp[ir]) + bn[w]);                  // fresh e_i: also resets the drift of the alpha updates
          ln[w] = FLOG(1.0 + en[w]);
          const double dl = fma(cfg.r[ir], bn[w] - bi[w], -cfg.n[ix[w]] * (ln[w] - SLL(ix[w]))) - half_inv_s2 * fma(bn[w], bn[w], -bi[w] * bi[w]);
#if MCU_SEEDS_LOGU == 2
          acc[w] = i0 + w < NPL && logu_less(lb[w], dl);
#elif MCU_SEEDS_LOGU
          acc[w] = i0 + w < NPL && uu[w] < dl;
#else
          acc[w] = i0 + w < NPL && mh_accept_nb(uu[w], dl);
#endif
        }
#pragma unroll
        for (int w = 0; w < W; ++w)
#if MCU_SEEDS_PF
          if (acc[w]) { SB(ix[w]) = bn[w]; SE(ix[w]) = en[w]; SLL(ix[w]) = ln[w]; if (adapt) tc[acoff + (size_t)w * C] = ac[w] + 1.0; }
#else
          if (acc[w]) { SB(ix[w]) = bn[w]; SE(ix[w]) = en[w]; SLL(ix[w]) = ln[w]; if (adapt) SAC(ix[w]) = ac[w] + 1.0; }
#endif
      };
      {
        constexpr int W = MCU_SEEDS_BW;
        int i0 = 0;
        if constexpr (W != 2 || !MCU_SEEDS_DEDUP) {   // (W == 2: one loop, one copy of the trip body in the instruction stream)
#pragma unroll 1
          for (; i0 + W <= NPL; i0 += W) b_trip(std::integral_constant<int, W>{}, i0);
        }
#pragma unroll 1
        for (; i0 < NPL; i0 += 2) b_trip(std::integral_constant<int, 2>{}, i0);
      }
      if (adapt && ((long long)m1 % cfg.batchsize[1]) == 0) {
        const double dl = amwg_delta(m1, cfg.batchsize[1]);
        const double up = exp(dl), dn = exp(-dl);
        for (int i = 0; i < NPL; ++i) SSG(i) = SSG(i) * ((SAC(i) / m1 < cfg.target[1]) ? dn : up);
      }
    }
    // ================================================================== block 2: AMWG(s2) on x = log s2
    {
      const bool adapt = cfg.adapt[2] == 1 ? iter <= a.burnin : cfg.adapt[2] == 0;
      if (adapt && !ad2) { acs = 0; m2 = 0.0; }
      ad2 = adapt;
      if (adapt) m2 += 1.0;
      double S = 0.0;
      for (int i = 0; i < NPL; ++i) { const double bi = SB(i); S += bi * bi; }
#if MCU_SEEDS_LOGU == 2
      const LogU lus = logu_bracket(draw_uniform_pair(a, chain, it32, 2, 0).a);
#elif MCU_SEEDS_LOGU
      const double lus = log_uniform(draw_uniform_pair(a, chain, it32, 2, 0).a);
#endif
      const double xn = x + sgs * draw_normal_pair(a, chain, it32, 2, 0).a;
      const double s2n = (xn > -700.0 && xn < 700.0) ? FEXP(xn) : exp(xn);
      const double dx = xn - x;
      const double dinv = 1.0 / s2n - 1.0 / s2;
      // logf(x) = InverseGamma(0.001, 0.001)(s2) + x [log-Jacobian, transformdistribution.jl:75-78]
      //           + sum_i Normal(b_i; 0, sqrt(s2))
      const double delta = -(0.001 + 1.0) * dx - 0.001 * dinv + dx - 0.5 * S * dinv - (double)NPL * 0.5 * dx;
#if MCU_SEEDS_LOGU == 2
      if (logu_less(lus, delta)) { x = xn; s2 = s2n; if (adapt) acs += 1; }
#elif MCU_SEEDS_LOGU
      if (lus < delta) { x = xn; s2 = s2n; if (adapt) acs += 1; }
#else
      const double u = draw_uniform_pair(a, chain, it32, 2, 0).a;
      if (mh_accept(u, delta)) { x = xn; s2 = s2n; if (adapt) acs += 1; }
#endif
      if (adapt && ((long long)m2 % cfg.batchsize[2]) == 0) {
        const double dl = amwg_delta(m2, cfg.batchsize[2]);
        sgs *= exp((double)acs / m2 < cfg.target[2] ? -dl : dl);
      }
    }
    // ================================================================== thinning + streaming moments
    if (iter > a.burnin && (iter - a.burnin) % a.thin == 0) {   // mcmc.jl:76-78
      mon[0] = al0; mon[1] = al1; mon[2] = al2; mon[3] = al3; mon[4] = s2;
      if (a.samples) {
        const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
        for (int j = 0; j < SeedsModel::P; ++j) a.samples[((size_t)row * SeedsModel::P + j) * C + c] = mon[j];
      }
      if (a.comom) comoments_update(a.mom, a.momn, a.comom, C, (size_t)c, SeedsModel::P, mon, a.log_mask);
      moments_update(a.mom, a.momn, C, (size_t)c, SeedsModel::P, mon);
    }
  }
  // ---- store chain state ------------------------------------------------------------------------
  a.state[0 * C + c] = al0; a.state[1 * C + c] = al1; a.state[2 * C + c] = al2; a.state[3 * C + c] = al3;
  a.state[4 * C + c] = s2;
  for (int i = 0; i < NPL; ++i) a.state[(size_t)(5 + i) * C + c] = SB(i);
  if (!AMM0) {
    TUNE(0, 0) = m0; TUNE(0, 1) = ad0 ? 1.0 : 0.0;
    TUNE(0, 2) = sg0; TUNE(0, 3) = sg1; TUNE(0, 4) = sg2; TUNE(0, 5) = sg3;
    TUNE(0, 6) = ac0; TUNE(0, 7) = ac1; TUNE(0, 8) = ac2; TUNE(0, 9) = ac3;
  }
  TUNE(1, 0) = m1; TUNE(1, 1) = ad1 ? 1.0 : 0.0;
  TUNE(2, 0) = m2; TUNE(2, 1) = ad2 ? 1.0 : 0.0; TUNE(2, 2) = sgs; TUNE(2, 3) = acs;
#undef SB
#undef SE
#undef SLL
#undef SLN
#undef SSG
#undef SAC
#undef TUNE
#undef FLOG
#undef FEXP
}

template <int BS, bool AMM0>
int launch_bs(const FastCfg& cfg, const RunArgs& a, cudaStream_t st) {
  const size_t smem = ((size_t)BS * 4 * NSL + (MCU_SEEDS_TAB ? 384 : 0)) * sizeof(double);
  static thread_local int attr_dev = -1;   // the attribute call is slow: once per device
  int dev = 0; cudaGetDevice(&dev);
  if (attr_dev != dev) {
    if (cudaFuncSetAttribute(seeds_fast_kernel<BS, AMM0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    attr_dev = dev;
  }
  const unsigned grid = (unsigned)((a.n_chains + BS - 1) / BS);
  seeds_fast_kernel<BS, AMM0><<<grid, BS, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

// h_blocks: host copies of the three DevBlocks (scale pointers are device pointers; the scales are
// re-read from the host-side scale mirror passed in cfg by the caller).
int seeds_fast_launch(const double* r, const double* n, const double* x1, const double* x2, const RunArgs& a, const DevBlock* h_blocks,
                      const std::vector<std::vector<double>>& h_scales, const double* h_SigmaL, cudaStream_t st) {
  FastCfg cfg;
  // plate constants, scales and (AMM) the Cholesky factor come from the handle's host-side copies of what the generic path uses
  const bool amm0 = h_blocks[0].kind == 6;   // MCU_AMM
  double sa[4] = {0.0, 0.0, 0.0, 0.0}, sb[NPL], ss[1];
  if (!amm0) for (int j = 0; j < 4; ++j) sa[j] = h_scales[0][j];
  for (int i = 0; i < NPL; ++i) sb[i] = h_scales[1][i];
  ss[0] = h_scales[2][0];
  for (int j = 0; j < 4; ++j) { cfg.amask[j] = 0; cfg.gmask[j] = 0; cfg.scale_a[j] = sa[j]; }
  for (int i = 0; i < NPL; ++i) {
    if ((x1[i] != 0.0 && x1[i] != 1.0) || (x2[i] != 0.0 && x2[i] != 1.0)) return -2;   // design must be 0/1 indicators
    cfg.r[i] = r[i]; cfg.n[i] = n[i]; cfg.scale_b[i] = sb[i];
    const int gi = (x1[i] != 0.0 ? 2 : 0) + (x2[i] != 0.0 ? 1 : 0);
    cfg.grp[i] = (unsigned char)gi;
    cfg.amask[0] |= 1u << i;
    if (x1[i] != 0.0) cfg.amask[1] |= 1u << i;
    if (x2[i] != 0.0) cfg.amask[2] |= 1u << i;
    if (x1[i] != 0.0 && x2[i] != 0.0) cfg.amask[3] |= 1u << i;
  }
  for (int j = 0; j < 4; ++j) { cfg.rsum[j] = 0.0; for (int i = 0; i < NPL; ++i) if ((cfg.amask[j] >> i) & 1u) cfg.rsum[j] += r[i]; }
  cfg.n[NPL] = 0.0;
  for (int j = 0; j < 4; ++j) {
    int cnt = 0;
    for (int i = 0; i < NPL; ++i) if ((cfg.amask[j] >> i) & 1u) cfg.alist[j][cnt++] = (unsigned char)i;
    while (cnt % MCU_SEEDS_AW) cfg.alist[j][cnt++] = (unsigned char)NPL;
    cfg.atriples[j] = cnt / MCU_SEEDS_AW;
    for (int k = cnt; k < 24; ++k) cfg.alist[j][k] = (unsigned char)NPL;
  }
  cfg.gmask[0] = 0xFu; cfg.gmask[1] = 0xCu; cfg.gmask[2] = 0xAu; cfg.gmask[3] = 0x8u;
  cfg.scale_s = ss[0];
  if (h_blocks[2].tune_off != h_blocks[1].tune_off + 2 + 2 * NPL) return -2;   // block 2's record must follow block 1's (the b-block prefetch reads into it): else the generic kernel
  for (int b = 0; b < 3; ++b) {
    cfg.adapt[b] = h_blocks[b].adapt; cfg.batchsize[b] = h_blocks[b].batchsize; cfg.tune_off[b] = h_blocks[b].tune_off;
    cfg.target[b] = h_blocks[b].target;
  }
  // 96 threads x 3 blocks/SM = 288 resident chains/SM: 125,000 chains/GPU fit in 3 even rounds
  if (amm0) {
    if (!h_SigmaL) return -1;
    for (int i = 0; i < 16; ++i) cfg.amm_SL[i] = h_SigmaL[i];
    cfg.amm_beta = h_blocks[0].beta; cfg.amm_scale = h_blocks[0].amm_scale;
    return launch_bs<MCU_SEEDS_BS, true>(cfg, a, st);
  }
  return launch_bs<MCU_SEEDS_BS, false>(cfg, a, st);
}

}  // namespace mcu
