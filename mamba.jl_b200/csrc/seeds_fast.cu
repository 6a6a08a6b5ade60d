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

#include "launch.hpp"

#ifndef MCU_SEEDS_BW
#define MCU_SEEDS_BW 2      // plates per trip in the b block (even)
#endif
#ifndef MCU_SEEDS_LOGU
#define MCU_SEEDS_LOGU 1   // MH test on the log scale (log u evaluated off the critical path)
#endif
#ifndef MCU_SEEDS_PIPE
#define MCU_SEEDS_PIPE 0   // b block: draws of trip t + 1 generated during trip t — measured 10 % SLOWER (profiles/r1_seeds_fast_summary.md), kept for the record
#endif
#ifndef MCU_SEEDS_ESTRIN
#define MCU_SEEDS_ESTRIN 0
#endif
#ifndef MCU_SEEDS_BS
#define MCU_SEEDS_BS 96
#endif
#ifndef MCU_SEEDS_MINB
#define MCU_SEEDS_MINB 3
#endif

namespace mcu {

namespace {

constexpr int NPL = SeedsModel::NP;   // 21 plates

constexpr int NSL = NPL + 1;          // shared-memory slots per array: 21 plates + one dummy (e = 0, L = 0, n = 0) that pads the plate lists

struct FastCfg {
  double r[NPL], n[NSL];
  unsigned char alist[4][24];         // plates whose eta depends on alpha_j, padded with the dummy slot to a multiple of 3
  int atriples[4];
  double rsum[4];                     // sum of r_i over the plates that depend on alpha_j
  unsigned char grp[NPL];             // 0:(x1=0,x2=0) 1:(0,1) 2:(1,0) 3:(1,1)
  unsigned amask[4];                  // plates whose eta depends on alpha_j
  unsigned gmask[4];                  // groups whose base depends on alpha_j
  int adapt[3], batchsize[3], tune_off[3];
  double target[3];
  double scale_a[4], scale_b[NPL], scale_s;
};

// rng.cuh contract: stream 0 = uniforms, stream 1 = normals, TWO draws per Philox block: draw k of a stream comes from
// block k >> 1 (first/second half for uniforms, cosine/sine branch of Box-Muller for normals).
struct Pair { double a, b; };
MCU_D double fast_log(double x);
MCU_D Pair fast_sincos2pi(double u);
MCU_D Pair draw_uniform_pair(const RunArgs& a, uint32_t chain, uint32_t iter, uint32_t block, uint32_t kpair) {
  uint32_t w[4];
  philox4x32_10(kpair, iter, chain, block, (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w);
  return {u53(w[0], w[1]), u53(w[2], w[3])};
}
MCU_D Pair draw_normal_pair(const RunArgs& a, uint32_t chain, uint32_t iter, uint32_t block, uint32_t kpair) {
  uint32_t w[4];
  philox4x32_10(kpair, iter, chain, block | (1u << 24), (uint32_t)a.seed, (uint32_t)(a.seed >> 32), w);
  const double rad = sqrt(-2.0 * fast_log(1.0 - u53(w[0], w[1])));
  const Pair sc = fast_sincos2pi(u53(w[2], w[3]));
  return {rad * sc.b, rad * sc.a};   // (rad cos, rad sin)
}

// ---- FP64 exp / log with constant-bank coefficients ------------------------------------------------------
__constant__ double kExpC[12] = {   // 1/n!, n = 13 .. 2 (Horner order); |r| <= ln2/2 ⇒ truncation < 5e-18
  1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.755731922398589e-06,
  2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02,
  1.6666666666666666e-01, 0.5};
__constant__ double kLogC[7] = {   // fdlibm e_log.c Lg7 .. Lg1
  1.479819860511658591e-01, 1.531383769920937332e-01, 1.818357216161805012e-01, 2.222219843214978396e-01,
  2.857142874366239149e-01, 3.999999999940941908e-01, 6.666666666666735130e-01};

// exp(x) for |x| < 700 (callers clamp): x = k ln2 + r, exp(r) by a degree-13 Taylor polynomial, 2^k through the exponent field
MCU_D double fast_exp(double x) {
  const double kf = rint(x * 1.4426950408889634074);
  double r = fma(kf, -6.93147180369123816490e-01, x);
  r = fma(kf, -1.90821492927058770002e-10, r);
#if MCU_SEEDS_ESTRIN
  // Estrin's scheme: dependency depth 6 instead of 12 (kExpC[11 - i] is the coefficient of r^i)
  const double r2 = r * r, r4 = r2 * r2;
  const double b0 = fma(kExpC[10], r, kExpC[11]), b1 = fma(kExpC[8], r, kExpC[9]), b2 = fma(kExpC[6], r, kExpC[7]);
  const double b3 = fma(kExpC[4], r, kExpC[5]), b4 = fma(kExpC[2], r, kExpC[3]), b5 = fma(kExpC[0], r, kExpC[1]);
  const double c0 = fma(b1, r2, b0), c1 = fma(b3, r2, b2), c2 = fma(b5, r2, b4);
  double p = fma(fma(c2, r4, c1), r4, c0);
  p = fma(p, r2, r) + 1.0;                              // 1 + r + r^2 (1/2 + r (1/6 + ...))
#else
  double p = kExpC[0];
#pragma unroll
  for (int i = 1; i < 12; ++i) p = fma(p, r, kExpC[i]);
  p = fma(p * r, r, r) + 1.0;                           // 1 + r + r^2 (1/2 + r (1/6 + ...))
#endif
  const int k = (int)kf;
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// log(x) for normal positive x (fdlibm e_log.c): x = 2^k m, m in [sqrt(1/2), sqrt(2)), f = m - 1, s = f / (2 + f)
MCU_D double fast_log(double x) {
  int hx = __double2hiint(x);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int adj = (hx + 0x95f64) & 0x100000;            // mantissa above sqrt(2): halve m, k += 1
  k += adj >> 20;
  const double m = __hiloint2double(hx | (adj ^ 0x3ff00000), __double2loint(x));
  const double f = m - 1.0;
  const double dnm = 2.0 + f;
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(dnm));
  y = fma(fma(-dnm, y, 1.0), y, y);
  y = fma(fma(-dnm, y, 1.0), y, y);                     // 1 / (2 + f) to full precision
  const double sq = f * y;
  const double z = sq * sq;
#if MCU_SEEDS_ESTRIN
  // fdlibm's own even/odd split (kLogC[7 - i] = Lg_i): two chains of depth 3-4 instead of one of depth 7
  const double w = z * z;
  const double t1 = w * fma(w, fma(w, kLogC[1], kLogC[3]), kLogC[5]);                      // w (Lg2 + w (Lg4 + w Lg6))
  const double t2 = z * fma(w, fma(w, fma(w, kLogC[0], kLogC[2]), kLogC[4]), kLogC[6]);    // z (Lg1 + w (Lg3 + w (Lg5 + w Lg7)))
  const double R = t2 + t1;
#else
  double R = kLogC[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) R = fma(R, z, kLogC[i]);
  R *= z;
#endif
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  // log(1+f) = f - (hfsq - s (hfsq + R));  result = k ln2_hi - ((hfsq - (s (hfsq + R) + k ln2_lo)) - f)
  return fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(sq, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
}

// sin and cos of 2 pi u, u in [0, 1): quadrant reduction is exact (u - q/4), then fdlibm's sin/cos kernels on |x| <= pi/4
__constant__ double kSinC[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};
__constant__ double kCosC[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
                                2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};
MCU_D Pair fast_sincos2pi(double u) {   // returns (sin, cos) of 2 pi u
  const double qf = rint(4.0 * u);
  const int q = (int)qf & 3;
  const double x = 6.283185307179586476925286766559 * fma(qf, -0.25, u);   // |x| <= pi/4
  const double z = x * x;
  double sp = kSinC[0], cp = kCosC[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) { sp = fma(sp, z, kSinC[i]); cp = fma(cp, z, kCosC[i]); }
  const double sn = fma(x * z, sp, x);
  const double cs = fma(z * z, cp, fma(z, -0.5, 1.0));
  // angle x + q pi/2:  cos: q = 0 → cos, 1 → -sin, 2 → -cos, 3 → sin ;  sin: q = 0 → sin, 1 → cos, 2 → -sin, 3 → -cos
  const double cv = (q & 1) ? sn : cs, sv = (q & 1) ? cs : sn;
  return {(q & 2) ? -sv : sv, ((q + 1) & 2) ? -cv : cv};
}

// r log p + (n - r) log(1 - p) with p = invlogit(eta), written as r eta - n softplus(eta)
MCU_D double binlogit_term(double r, double n, double eta) {
  const double e = exp(-fabs(eta));
  return r * eta - n * (fmax(eta, 0.0) + log1p(e));
}

MCU_D bool mh_accept(double u, double delta) {   // rand() < exp(logfprime - logf0): amwg.jl:107
  if (delta >= 0.0) return true;                  // u < 1 <= exp(delta)
  if (!(delta > -700.0)) return false;            // exp underflows (or delta is NaN): u < 0 never holds
  return u < fast_exp(delta);
}
#if MCU_SEEDS_LOGU
// rand() < exp(delta)  <=>  log(rand()) < delta: the log of the uniform does not depend on the proposal, so it is evaluated next to the
// draw, off the exp -> log -> delta critical path of the update (same decision up to rounding of the comparison)
MCU_D double log_uniform(double u) { return u > 0.0 ? fast_log(u) : -CUDART_INF; }
#endif
// the same decision without branches (the exp is always evaluated), so that two or three independent updates can be
// scheduled into each other's dependency stalls
MCU_D bool mh_accept_nb(double u, double delta) {
  const double ex = fast_exp(fmax(fmin(delta, 0.0), -700.0));
  return delta >= 0.0 ? true : (delta > -700.0 && u < ex);
}

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

// AMWG tune update every `batchsize` adaptive iterations: amwg.jl:74-80
MCU_D double amwg_delta(double m, int batchsize) { return fmin(0.01, pow(m / (double)batchsize, -0.5)); }

template <int BS>
__global__ void __launch_bounds__(BS, MCU_SEEDS_MINB) seeds_fast_kernel(const __grid_constant__ FastCfg cfg, const __grid_constant__ RunArgs a) {
  extern __shared__ double smem[];
  double* sb = smem;                        // b[i]
  double* se = smem + NSL * BS;             // e[i] = exp(eta_i)
  double* sll = smem + 2 * NSL * BS;        // L[i] = log(1 + e[i])
  double* sln = smem + 3 * NSL * BS;        // proposed L[i]
  const int tid = threadIdx.x;
  const long long c = (long long)blockIdx.x * BS + tid;
  if (c >= a.n_chains) return;
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
  double m0 = TUNE(0, 0), m1 = TUNE(1, 0), m2 = TUNE(2, 0);
  bool ad0 = TUNE(0, 1) != 0.0, ad1 = TUNE(1, 1) != 0.0, ad2 = TUNE(2, 1) != 0.0;
  double sg0 = TUNE(0, 2), sg1 = TUNE(0, 3), sg2 = TUNE(0, 4), sg3 = TUNE(0, 5);
  int ac0 = (int)TUNE(0, 6), ac1 = (int)TUNE(0, 7), ac2 = (int)TUNE(0, 8), ac3 = (int)TUNE(0, 9);
  double sgs = TUNE(2, 2); int acs = (int)TUNE(2, 3);

  Bases g = group_bases(al0, al1, al2, al3);
  for (int i = 0; i < NPL; ++i) { const double e = fast_exp(pick(g, cfg.grp[i]) + SB(i)); SE(i) = e; SLL(i) = fast_log(1.0 + e); }
  SB(NPL) = 0.0; SE(NPL) = 0.0; SLL(NPL) = 0.0; SLN(NPL) = 0.0;   // dummy slot: log(1 + 0 * E) = 0, n = 0

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
    }
    // ================================================================== block 0: AMWG(alpha0..alpha12)
    {
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
#if MCU_SEEDS_LOGU
        double lu;                                                        // log of uniform j of the block, off the critical path
        if ((j & 1) == 0) { const Pair pr = draw_uniform_pair(a, chain, it32, 0, j >> 1); lu = log_uniform(pr.a); uc = log_uniform(pr.b); } else lu = uc;
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
        const double E = fast_exp(z);
        // three plates per trip: their logs are independent, so the scheduler fills one chain's DFMA latency with the others
        double dLa = 0.0, dLb = 0.0, dLc = 0.0;
        const int nt = cfg.atriples[j];
#pragma unroll 1
        for (int k = 0; k < nt; ++k) {
          const int ia = cfg.alist[j][3 * k], ib = cfg.alist[j][3 * k + 1], ic = cfg.alist[j][3 * k + 2];
          const double la = fast_log(fma(SE(ia), E, 1.0)), lb = fast_log(fma(SE(ib), E, 1.0)), lc = fast_log(fma(SE(ic), E, 1.0));
          SLN(ia) = la; SLN(ib) = lb; SLN(ic) = lc;
          dLa = fma(cfg.n[ia], la - SLL(ia), dLa);
          dLb = fma(cfg.n[ib], lb - SLL(ib), dLb);
          dLc = fma(cfg.n[ic], lc - SLL(ic), dLc);
        }
        double delta = fma(cfg.rsum[j], z, -((dLa + dLb) + dLc));
        {   // Normal(0, 1000) prior of the component: -(z^2 + log 2pi)/2 - log sigma
          delta = fma(-0.5e-6, fma(anew, anew, -al0 * al0), delta);   // (x / 1000)^2 / 2 without the divisions
        }
#if MCU_SEEDS_LOGU
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
      auto b_trip = [&](auto Wc, int i0) {
        constexpr int W = decltype(Wc)::value;
        int ix[W]; double sg[W], bi[W], zn[W], uu[W], ac[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const bool real = i0 + w < NPL;
          ix[w] = real ? i0 + w : NPL;                                     // past the last plate: the dummy slot (never accepted)
          sg[w] = real ? SSG(ix[w]) : 0.0;                                 // global (L2) loads, issued a whole trip ahead of their use
          ac[w] = (real && adapt) ? SAC(ix[w]) : 0.0;
          bi[w] = SB(ix[w]);
        }
#pragma unroll
        for (int w = 0; w < W; w += 2) {
          const Pair pz = draw_normal_pair(a, chain, it32, 1, (i0 + w) >> 1);
          const Pair pu = draw_uniform_pair(a, chain, it32, 1, (i0 + w) >> 1);
#if MCU_SEEDS_LOGU
          zn[w] = pz.a; zn[w + 1] = pz.b; uu[w] = log_uniform(pu.a); uu[w + 1] = log_uniform(pu.b);
#else
          zn[w] = pz.a; zn[w + 1] = pz.b; uu[w] = pu.a; uu[w + 1] = pu.b;
#endif
        }
        double bn[W], en[W], ln[W]; bool acc[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const int ir = i0 + w < NPL ? i0 + w : 0;                        // constants of a real plate for the dummy chain
          bn[w] = bi[w] + sg[w] * zn[w];
          en[w] = fast_exp(pick(g, cfg.grp[ir]) + bn[w]);                  // fresh e_i: also resets the drift of the alpha updates
          ln[w] = fast_log(1.0 + en[w]);
          const double dl = fma(cfg.r[ir], bn[w] - bi[w], -cfg.n[ix[w]] * (ln[w] - SLL(ix[w]))) - half_inv_s2 * fma(bn[w], bn[w], -bi[w] * bi[w]);
#if MCU_SEEDS_LOGU
          acc[w] = i0 + w < NPL && uu[w] < dl;
#else
          acc[w] = i0 + w < NPL && mh_accept_nb(uu[w], dl);
#endif
        }
#pragma unroll
        for (int w = 0; w < W; ++w)
          if (acc[w]) { SB(ix[w]) = bn[w]; SE(ix[w]) = en[w]; SLL(ix[w]) = ln[w]; if (adapt) SAC(ix[w]) = ac[w] + 1.0; }
      };
#if MCU_SEEDS_PIPE && MCU_SEEDS_LOGU
      // Software-pipelined form (two plates per trip): the normals and log-uniforms of trip t + 1 depend on nothing but the counters, so
      // they are generated DURING trip t — four independent dependency chains (two updates, Philox + Box-Muller, Philox + two logs) for
      // the scheduler to interleave, and the exp -> log -> compare chain of a trip no longer waits for its own draws.
      {
        Pair pz = draw_normal_pair(a, chain, it32, 1, 0);
        Pair lu; { const Pair pu = draw_uniform_pair(a, chain, it32, 1, 0); lu.a = log_uniform(pu.a); lu.b = log_uniform(pu.b); }
#pragma unroll 1
        for (int ip = 0; ip < (NPL + 1) / 2; ++ip) {
          const int i0 = 2 * ip;
          const bool two = i0 + 1 < NPL;
          const int i1 = two ? i0 + 1 : NPL, r1 = two ? i0 + 1 : i0;           // odd plate count: the last trip pairs with the dummy slot
          const double sga = SSG(i0), sgb = two ? SSG(i1) : 0.0;               // global (L2) loads, issued a whole trip ahead of their use
          const double aca = adapt ? SAC(i0) : 0.0, acb = (adapt && two) ? SAC(i1) : 0.0;
          const double bia = SB(i0), bib = SB(i1);
          const double za = pz.a, zb = pz.b, lua = lu.a, lub = lu.b;
          // draws of the next trip (one trip past the end is harmless: nothing consumes it)
          pz = draw_normal_pair(a, chain, it32, 1, ip + 1);
          { const Pair pu = draw_uniform_pair(a, chain, it32, 1, ip + 1); lu.a = log_uniform(pu.a); lu.b = log_uniform(pu.b); }
          const double bna = bia + sga * za, bnb = bib + sgb * zb;
          const double ena = fast_exp(pick(g, cfg.grp[i0]) + bna);             // fresh e_i: also resets the drift of the alpha updates
          const double enb = fast_exp(pick(g, cfg.grp[r1]) + bnb);
          const double lna = fast_log(1.0 + ena), lnb = fast_log(1.0 + enb);
          const double da = fma(cfg.r[i0], bna - bia, -cfg.n[i0] * (lna - SLL(i0))) - half_inv_s2 * fma(bna, bna, -bia * bia);
          const double db = fma(cfg.r[r1], bnb - bib, -cfg.n[i1] * (lnb - SLL(i1))) - half_inv_s2 * fma(bnb, bnb, -bib * bib);
          const bool acca = lua < da, accb = two && lub < db;
          if (acca) { SB(i0) = bna; SE(i0) = ena; SLL(i0) = lna; if (adapt) SAC(i0) = aca + 1.0; }
          if (accb) { SB(i1) = bnb; SE(i1) = enb; SLL(i1) = lnb; if (adapt) SAC(i1) = acb + 1.0; }
        }
      }
#else
      {
        constexpr int W = MCU_SEEDS_BW;
        int i0 = 0;
#pragma unroll 1
        for (; i0 + W <= NPL; i0 += W) b_trip(std::integral_constant<int, W>{}, i0);
#pragma unroll 1
        for (; i0 < NPL; i0 += 2) b_trip(std::integral_constant<int, 2>{}, i0);
      }
#endif
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
#if MCU_SEEDS_LOGU
      const double lus = log_uniform(draw_uniform_pair(a, chain, it32, 2, 0).a);
#endif
      const double xn = x + sgs * draw_normal_pair(a, chain, it32, 2, 0).a;
      const double s2n = (xn > -700.0 && xn < 700.0) ? fast_exp(xn) : exp(xn);
      // logf(x) = InverseGamma(0.001, 0.001)(s2) + x [log-Jacobian, transformdistribution.jl:75-78]
      //           + sum_i Normal(b_i; 0, sqrt(s2))
      const double dx = xn - x;
      const double dinv = 1.0 / s2n - 1.0 / s2;
      const double delta = -(0.001 + 1.0) * dx - 0.001 * dinv + dx - 0.5 * S * dinv - (double)NPL * 0.5 * dx;
#if MCU_SEEDS_LOGU
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
      moments_update(a.mom, a.momn, C, (size_t)c, SeedsModel::P, mon);
    }
  }
  // ---- store chain state ------------------------------------------------------------------------
  a.state[0 * C + c] = al0; a.state[1 * C + c] = al1; a.state[2 * C + c] = al2; a.state[3 * C + c] = al3;
  a.state[4 * C + c] = s2;
  for (int i = 0; i < NPL; ++i) a.state[(size_t)(5 + i) * C + c] = SB(i);
  TUNE(0, 0) = m0; TUNE(0, 1) = ad0 ? 1.0 : 0.0;
  TUNE(0, 2) = sg0; TUNE(0, 3) = sg1; TUNE(0, 4) = sg2; TUNE(0, 5) = sg3;
  TUNE(0, 6) = ac0; TUNE(0, 7) = ac1; TUNE(0, 8) = ac2; TUNE(0, 9) = ac3;
  TUNE(1, 0) = m1; TUNE(1, 1) = ad1 ? 1.0 : 0.0;
  TUNE(2, 0) = m2; TUNE(2, 1) = ad2 ? 1.0 : 0.0; TUNE(2, 2) = sgs; TUNE(2, 3) = acs;
#undef SB
#undef SE
#undef SLL
#undef SLN
#undef SSG
#undef SAC
#undef TUNE
}

template <int BS>
int launch_bs(const FastCfg& cfg, const RunArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)BS * 4 * NSL * sizeof(double);
  if (cudaFuncSetAttribute(seeds_fast_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  const unsigned grid = (unsigned)((a.n_chains + BS - 1) / BS);
  seeds_fast_kernel<BS><<<grid, BS, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

// h_blocks: host copies of the three DevBlocks (scale pointers are device pointers; the scales are
// re-read from the host-side scale mirror passed in cfg by the caller).
int seeds_fast_launch(const SeedsModel::Data& d, const RunArgs& a, const DevBlock* h_blocks, cudaStream_t st) {
  (void)d;
  FastCfg cfg;
  // plate constants and scales come down from the device copies the generic path uses, so both
  // paths see identical inputs
  double r[NPL], n[NPL], x1[NPL], x2[NPL], sa[4], sb[NPL], ss[1];
  if (cudaMemcpy(r, d.r, sizeof(r), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(n, d.n, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(x1, d.x1, sizeof(x1), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(x2, d.x2, sizeof(x2), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(sa, h_blocks[0].scale, sizeof(sa), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(sb, h_blocks[1].scale, sizeof(sb), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(ss, h_blocks[2].scale, sizeof(ss), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
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
    while (cnt % 3) cfg.alist[j][cnt++] = (unsigned char)NPL;
    cfg.atriples[j] = cnt / 3;
    for (int k = cnt; k < 24; ++k) cfg.alist[j][k] = (unsigned char)NPL;
  }
  cfg.gmask[0] = 0xFu; cfg.gmask[1] = 0xCu; cfg.gmask[2] = 0xAu; cfg.gmask[3] = 0x8u;
  cfg.scale_s = ss[0];
  for (int b = 0; b < 3; ++b) {
    cfg.adapt[b] = h_blocks[b].adapt; cfg.batchsize[b] = h_blocks[b].batchsize; cfg.tune_off[b] = h_blocks[b].tune_off;
    cfg.target[b] = h_blocks[b].target;
  }
  // 96 threads x 3 blocks/SM = 288 resident chains/SM: 125,000 chains/GPU fit in 3 even rounds
  return launch_bs<MCU_SEEDS_BS>(cfg, a, st);
}

}  // namespace mcu
