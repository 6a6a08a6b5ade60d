// rats_warp.cu — one-chain-per-warp kernel for configs[2]: the `rats` hierarchical normal growth model
// (doc/examples/rats.jl:49-97) under the scheme of SURVEY.md §8d config 3
//     [NUTS(alpha, beta, mu_alpha, mu_beta), Slice(s2_c, s2_alpha, s2_beta; univariate)]
//
// The generic engine runs one chain per thread; for a 62-dimensional NUTS block that means ~20 KB of per-thread
// vectors in local memory (tree edges, the per-level stack of the unrolled buildtree).  Here a chain is a WARP:
//   * lane i < 30 owns rat i: alpha_i, beta_i, its 5 observations and its two components of every NUTS vector;
//     mu_alpha, mu_beta are carried by every lane (warp-uniform);
//   * one leapfrog = 5 fused residuals per lane + butterfly reductions (sum e^2, sum (alpha-mu), sum (alpha-mu)^2, ...)
//     by warp shuffle — the gradient, the log-density and the kinetic energy of a leaf never leave registers;
//   * the U-turn checks (nuts.jl:183-187) are two shuffle-reduced dot products;
//   * tree doubling is the same leaf-by-leaf unrolling of the reference's recursive buildtree (nuts.jl:139-180) as
//     samplers.cuh::nuts_sub — same draws in the same order, same merge rule — with the per-level stack
//     {first leaf x, first leaf r, proposal x, n} in shared memory for levels 0-3 and in an L2-resident scratch
//     for deeper levels; proposals are tracked by reference (which stack level holds it), so nothing is copied at a merge;
//   * dual averaging, nutsepsilon (nuts.jl:63-92,192-205) and the whole Slice block (slice.jl:66-92, on the
//     sufficient statistics sum e^2, sum (alpha-mu_alpha)^2, sum (beta-mu_beta)^2) are warp-uniform scalar code.
// Warps never synchronise with each other: chains sit at different tree depths at the same time.
// The kernel is instruction-fetch bound (ncu: stall_no_instruction dominates once the leaf loop exceeds the ~6 KB L0
// instruction cache: 4 warps per scheduler at different program counters), so the leaf is written for few instructions:
// only logp = logf - r.r/2 is needed per leaf, so its lane parts go through ONE butterfly (3 reductions per leaf with the two
// for the mu gradient), the merge uniforms come from a lane-parallel Philox batch (64 per refill), no FP64 division.
// RNG: the engine's Philox contract (rng.cuh) — normal k of the block update is element k of r = randn(n), so lane l
// evaluates Philox block l (both Box-Muller branches) and the 62 momenta are dealt out by shuffle.
#include <cstdlib>

#ifndef MCU_RATS_MINB
#define MCU_RATS_MINB 3   // resident blocks per SM the register allocation aims for
#endif

#include "launch.hpp"

namespace mcu {

namespace {

constexpr int NR = RatsModel::NR;   // 30 rats
constexpr int NOBS = 5;             // observations per rat (checked on the host)
constexpr int LS = 4;               // stack levels kept in shared memory
constexpr int VEC = 64;             // doubles per stored vector: [0..29] alpha part, [30] mu_alpha, [31] mu_beta, [32..61] beta part
constexpr int kWarpsPerBlock = 4;
constexpr int kSlots = LS * 3 + 7;  // stack levels + 6 tree edges + the iteration's accepted position
constexpr int kWarpDoubles = kSlots * VEC + 16;   // + Sn[10]
constexpr unsigned FULL = 0xffffffffu;

struct WarpCfg {
  double y[NOBS][32], x[NOBS][32];   // [observation][rat]; lanes 30, 31 are padding
  double width[3];                   // Slice widths of s2_c, s2_alpha, s2_beta
  double target, eps_desc, xbar;
  int max_depth, tune_off;
  double* scratch;                   // [warps in the grid][kMaxDepth - LS][3][VEC]
};

struct V4 { double a, b, ma, mb; };  // a, b: this lane's rat (0 on lanes 30, 31); ma, mb: mu_alpha, mu_beta (warp-uniform)

MCU_D double wsum(double v) {        // xor butterfly: every lane ends with the bitwise-identical sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
MCU_D void vst(double* p, const V4& v, int lane) {
  __syncwarp();                      // earlier reads of this slot by other lanes are done
  p[lane] = lane < NR ? v.a : (lane == NR ? v.ma : v.mb);
  p[32 + lane] = v.b;
  __syncwarp();
}
MCU_D V4 vld(const double* p, int lane) {
  V4 v;
  const double t = p[lane];
  v.a = lane < NR ? t : 0.0; v.b = p[32 + lane]; v.ma = p[NR]; v.mb = p[NR + 1];
  return v;
}
MCU_D void vcopy(double* dst, const double* src, int lane) {
  __syncwarp();
  const double t0 = src[lane], t1 = src[32 + lane];
  dst[lane] = t0; dst[32 + lane] = t1;
  __syncwarp();
}

struct WRng {
  uint32_t k0, k1, chain, iter, block, ku, kn, ubase;
  double ua, ub;                       // this lane's two uniforms of the current batch: draws ubase + 2 lane, ubase + 2 lane + 1
  MCU_D void seek(uint32_t it, uint32_t blk) { iter = it; block = blk; ku = 0; kn = 0; ubase = 0xffffff00u; }
  MCU_NOINL void refill() {            // lane l evaluates Philox block (ku >> 1) + l: 64 uniforms per batch
    uint32_t w[4];
    ubase = ku & ~1u;
    philox4x32_10((ubase >> 1) + (uint32_t)(threadIdx.x & 31), iter, chain, block, k0, k1, w);
    ua = u53(w[0], w[1]); ub = u53(w[2], w[3]);
  }
  MCU_D double uniform() {             // draw ku of the block update, warp-uniform
    if (ku - ubase >= 64u) refill();
    const uint32_t idx = ku - ubase;
    const double mine = (idx & 1u) ? ub : ua;
    ++ku;
    return __shfl_sync(FULL, mine, (int)(idx >> 1));
  }
  // r = randn(62): element e is normal kn + e of the block update; lane l evaluates Philox block (kn >> 1) + l
  MCU_NOINL V4 normals62(int lane) {
    uint32_t w[4];
    philox4x32_10((kn >> 1) + (uint32_t)lane, iter, chain, block | (1u << 24), k0, k1, w);
    const double ua_ = u53(w[0], w[1]), ub_ = u53(w[2], w[3]);
    const double zc = box_muller(ua_, ub_), zs = box_muller_sin(ua_, ub_);
    V4 r;
    const int sa = lane >> 1, sb = 15 + (lane >> 1);
    const double ac = __shfl_sync(FULL, zc, sa), as = __shfl_sync(FULL, zs, sa);
    const double bc = __shfl_sync(FULL, zc, sb), bs = __shfl_sync(FULL, zs, sb);
    r.a = lane < NR ? ((lane & 1) ? as : ac) : 0.0;
    r.b = lane < NR ? ((lane & 1) ? bs : bc) : 0.0;
    r.ma = __shfl_sync(FULL, zc, NR); r.mb = __shfl_sync(FULL, zs, NR);
    kn += 62;
    return r;
  }
};

// exp(x) for x <= 0 (acceptance statistic): x = k ln2 + r, degree-11 Taylor polynomial, 2^k through the exponent field
MCU_D double exp_nonpos(double x) {
  x = fmax(x, -700.0);
  const double kf = rint(x * 1.4426950408889634074);
  double r = fma(kf, -6.93147180369123816490e-01, x);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double p = 2.505210838544172e-08;
  p = fma(p, r, 2.755731922398589e-07); p = fma(p, r, 2.755731922398589e-06); p = fma(p, r, 2.48015873015873e-05);
  p = fma(p, r, 1.984126984126984e-04); p = fma(p, r, 1.388888888888889e-03); p = fma(p, r, 8.333333333333333e-03);
  p = fma(p, r, 4.1666666666666664e-02); p = fma(p, r, 1.6666666666666666e-01); p = fma(p, r, 0.5);
  p = fma(p * r, r, r) + 1.0;
  return __hiloint2double(__double2hiint(p) + ((int)kf << 20), __double2loint(p));
}

struct Chain {
  double y[NOBS], xo[NOBS];
  // constants of the NUTS block density while s2_alpha, s2_beta, s2_c are held fixed (set_variances): reciprocals for the
  // gradient, half-reciprocals of sigma^2 for the quadratic forms, and every term of the log-density that does not depend on x
  double is2a, is2b, is2c, qa, qb, qc, lpc;
  double nw;                           // -1 on the 30 rat lanes, 0 on the two padding lanes
  int lane;

  MCU_D void set_variances(double s2a, double s2b, double s2c) {
    is2a = 1.0 / s2a; is2b = 1.0 / s2b; is2c = 1.0 / s2c;
    const double sga = sqrt(s2a), sgb = sqrt(s2b), sgc = sqrt(s2c);
    qa = 0.5 / (sga * sga); qb = 0.5 / (sgb * sgb); qc = 0.5 / (sgc * sgc);
    // lp_normal(mu, 0, 1000) x 2, sum_i lp_normal(alpha_i | mu_alpha, sga), sum_i lp_normal(beta_i | ...), lp_isonormal(y): models.cuh
    lpc = -2.0 * (0.5 * kLog2Pi + log(1000.0)) - (double)NR * (0.5 * kLog2Pi + log(sga)) - (double)NR * (0.5 * kLog2Pi + log(sgb))
          - ((double)(NR * NOBS) * kLog2Pi + (double)(NR * NOBS) * log(sgc * sgc)) / 2.0;
  }
  // One leapfrog step (nuts.jl:129-136, in place) of size eps on the NUTS block density — logpdfgrad!(block, x) is the block density
  // and its analytic gradient (models.cuh: RatsModel::factor / joint_grad; non-finite gradient entries zeroed, sampler.jl:110).
  // Returns logp = logf(x') - r'.r'/2, the only combination buildtree uses (nuts.jl:144-150); its lane parts share one butterfly.
  MCU_D double leapfrog_inl(V4& x, V4& r, V4& g, double eps) const {
    const double h = 0.5 * eps;
    r.a = fma(h, g.a, r.a); r.b = fma(h, g.b, r.b); r.ma = fma(h, g.ma, r.ma); r.mb = fma(h, g.mb, r.mb);
    x.a = fma(eps, r.a, x.a); x.b = fma(eps, r.b, x.b); x.ma = fma(eps, r.ma, x.ma); x.mb = fma(eps, r.mb, x.mb);
    double se = 0.0, sxe = 0.0, see = 0.0;
#pragma unroll
    for (int k = 0; k < NOBS; ++k) {
      const double e = y[k] - fma(x.b, xo[k], x.a);   // padding lanes: y = xo = 0 and x.a = x.b = 0, so e = 0
      se += e; sxe = fma(e, xo[k], sxe); see = fma(e, e, see);
    }
    const double da = fma(nw, x.ma, x.a), db = fma(nw, x.mb, x.b);   // alpha_i - mu_alpha, beta_i - mu_beta; 0 on padding lanes
    double ga = fma(se, is2c, -da * is2a), gb = fma(sxe, is2c, -db * is2b);
    // second half-kick of the lane parts and the lane part of logp first: the three butterflies (sum da, sum db, logp) are
    // independent, so they run interleaved and the leaf pays one shuffle-chain latency instead of two
    if (!isfinite(ga + gb)) { if (!isfinite(ga)) ga = 0.0; if (!isfinite(gb)) gb = 0.0; }   // rare: zero non-finite gradient entries
    g.a = ga; g.b = gb;
    r.a = fma(h, ga, r.a); r.b = fma(h, gb, r.b);
    double hl = -qa * da * da;
    hl = fma(-qb * db, db, hl); hl = fma(-qc, see, hl); hl = fma(-0.5 * r.a, r.a, hl); hl = fma(-0.5 * r.b, r.b, hl);
    double sa = da, sb = db;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double t0 = __shfl_xor_sync(FULL, sa, o), t1 = __shfl_xor_sync(FULL, sb, o), t2 = __shfl_xor_sync(FULL, hl, o);
      sa += t0; sb += t1; hl += t2;
    }
    double gma = fma(sa, is2a, -x.ma * 1e-6), gmb = fma(sb, is2b, -x.mb * 1e-6);
    if (!isfinite(gma + gmb)) { if (!isfinite(gma)) gma = 0.0; if (!isfinite(gmb)) gmb = 0.0; }
    g.ma = gma; g.mb = gmb;
    r.ma = fma(h, gma, r.ma); r.mb = fma(h, gmb, r.mb);
    double H = hl + lpc;
    H = fma(-0.5e-6 * x.ma, x.ma, H); H = fma(-0.5e-6 * x.mb, x.mb, H);
    H = fma(-0.5 * r.ma, r.ma, H); H = fma(-0.5 * r.mb, r.mb, H);
    return H;
  }
};

MCU_NOINL bool nouturn4(const V4& xminus, const V4& xplus, const V4& rminus, const V4& rplus) {   // nuts.jl:183-187
  const double da = xplus.a - xminus.a, db = xplus.b - xminus.b, dma = xplus.ma - xminus.ma, dmb = xplus.mb - xminus.mb;
  const double a = wsum(da * rminus.a + db * rminus.b) + dma * rminus.ma + dmb * rminus.mb;
  const double c = wsum(da * rplus.a + db * rplus.b) + dma * rplus.ma + dmb * rplus.mb;
  return a >= 0 && c >= 0;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32, MCU_RATS_MINB) rats_warp_kernel(const __grid_constant__ WarpCfg cfg, const __grid_constant__ RunArgs a) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = smem + (size_t)warp * kWarpDoubles;
  double* Sn = ws + kSlots * VEC;
  const long long gw = (long long)blockIdx.x * kWarpsPerBlock + warp, GW = (long long)gridDim.x * kWarpsPerBlock;
  double* gscr = cfg.scratch + (size_t)gw * (kMaxDepth - LS) * 3 * VEC;
  auto slot = [&](int level, int which) -> double* {
    return level < LS ? ws + (level * 3 + which) * VEC : gscr + ((level - LS) * 3 + which) * VEC;
  };
  double* e_xm = ws + (LS * 3 + 0) * VEC; double* e_rm = ws + (LS * 3 + 1) * VEC; double* e_gm = ws + (LS * 3 + 2) * VEC;
  double* e_xp = ws + (LS * 3 + 3) * VEC; double* e_rp = ws + (LS * 3 + 4) * VEC; double* e_gp = ws + (LS * 3 + 5) * VEC;
  double* e_v = ws + (LS * 3 + 6) * VEC;
  const size_t C = (size_t)a.n_chains;
  const double ig_c0 = ig001_c0();

  Chain ch;
  ch.lane = lane; ch.nw = lane < NR ? -1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < NOBS; ++k) { ch.y[k] = cfg.y[k][lane]; ch.xo[k] = cfg.x[k][lane]; }

  unsigned long long n_leap = 0;      // leapfrogs of this warp's chains (warp-uniform): the work unit of the roofline (mcu_work_count)
  for (long long c = gw; c < a.n_chains; c += GW) {
    WRng rng;
    rng.k0 = (uint32_t)a.seed; rng.k1 = (uint32_t)(a.seed >> 32); rng.chain = (uint32_t)(a.chain_offset + c);
    // ---- chain state: mu_alpha, mu_beta, s2_alpha, s2_beta, s2_c, alpha[30], beta[30]
    V4 x;
    x.ma = a.state[0 * C + c]; x.mb = a.state[1 * C + c];
    double s2a = a.state[2 * C + c], s2b = a.state[3 * C + c], s2c = a.state[4 * C + c];
    x.a = lane < NR ? a.state[(size_t)(5 + lane) * C + c] : 0.0;
    x.b = lane < NR ? a.state[(size_t)(5 + NR + lane) * C + c] : 0.0;
    double* tn = a.tune + (size_t)cfg.tune_off * C + c;   // slots: 0 adapt, 1 alpha, 2 epsilon, 3 epsilonbar, 4 Hbar, 5 m, 6 mu, 7 nalpha
    double t_adapt = tn[0 * C], t_alpha = tn[1 * C], t_eps = tn[2 * C], t_epsbar = tn[3 * C], t_Hbar = tn[4 * C], t_m = tn[5 * C],
           t_mu = tn[6 * C], t_nalpha = tn[7 * C];

    for (long long it = 1; it <= a.iters; ++it) {
      const long long iter = a.iter0 + it;
      // ================================================================ block 0: NUTS(alpha, beta, mu_alpha, mu_beta)
      rng.seek((uint32_t)iter, 0);
      ch.set_variances(s2a, s2b, s2c);
      const bool adapt = iter <= a.burnin;                       // nuts.jl:52
      if (iter == 1) {                                           // NUTSTune(x, nutsepsilon(x, f)): nuts.jl:17-30 via sampler.jl:40-45
        t_adapt = 0.0; t_alpha = 0.0; t_epsbar = 1.0; t_Hbar = 0.0; t_m = 0.0; t_mu = CUDART_NAN; t_nalpha = 0.0;
        if (cfg.eps_desc > 0.0) {
          t_eps = cfg.eps_desc;
        } else {                                                 // nutsepsilon: nuts.jl:192-205
          V4 r0 = rng.normals62(lane), x0 = x, g0 = {0.0, 0.0, 0.0, 0.0};
          V4 xx = x0;
          const double logp_0 = ch.leapfrog_inl(xx, r0, g0, 0.0);   // logf(x0) - r0.r0/2
          vst(e_xm, x0, lane); vst(e_rm, r0, lane); vst(e_gm, g0, lane);
          double eps = 1.0;
          auto trial = [&](double e) {
            V4 tx = vld(e_xm, lane), tr = vld(e_rm, lane), tg = vld(e_gm, lane);
            return exp(ch.leapfrog_inl(tx, tr, tg, e) - logp_0);
          };
          double prob = trial(eps);
          const int pm = prob > 0.5 ? 1 : -1;
          int guard = 0;
          while (pow(prob, (double)pm) > pow(0.5, (double)pm)) {
            eps *= pm == 1 ? 2.0 : 0.5;
            prob = trial(eps);
            if (++guard > 2000) break;
          }
          t_eps = eps;
        }
      }
      const bool was = t_adapt != 0.0;
      if (adapt && !was) { t_m = 0.0; t_mu = log(10.0 * t_eps); }   // setadapt!: nuts.jl:84-92
      t_adapt = adapt ? 1.0 : 0.0;
      if (adapt) t_m += 1.0; else if (t_m > 0.0) t_eps = t_epsbar;
      {
        // ------------------------------------------------------------ nuts_sub!: nuts.jl:95-126
        const double eps = t_eps;
        V4 cr = rng.normals62(lane), cx = x, cg = {0.0, 0.0, 0.0, 0.0};
        const double logp0 = ch.leapfrog_inl(cx, cr, cg, 0.0);
        const double logu0 = logp0 + log(rng.uniform());
        vst(e_xm, cx, lane); vst(e_xp, cx, lane); vst(e_rm, cr, lane); vst(e_rp, cr, lane); vst(e_gm, cg, lane); vst(e_gp, cg, lane);
        vst(e_v, x, lane);
        int j = 0; double n = 1.0; bool s = true;
        double alpha = 0.0, nalpha = 0.0;
        while (s) {
          const int pm = rng.uniform() > 0.5 ? 1 : -1;
          if (pm == -1) { cx = vld(e_xm, lane); cr = vld(e_rm, lane); cg = vld(e_gm, lane); }
          else { cx = vld(e_xp, lane); cr = vld(e_rp, lane); cg = vld(e_gp, lane); }
          // ---- buildtree(.., pm, j, ..) unrolled leaf by leaf: nuts.jl:139-180
          const unsigned nleaf = 1u << j;
          double Tn = 0.0; bool Ts = true;
          int xp_src = -1;                                       // where the subtree's proposal lives: -1 = the current leaf, else stack level
          alpha = 0.0; nalpha = 0.0;
          for (unsigned t = 0; t < nleaf; ++t) {
            const double logpp = ch.leapfrog_inl(cx, cr, cg, pm * eps);
            Tn = logu0 < logpp ? 1.0 : 0.0;
            Ts = logu0 < logpp + 1000.0;
            alpha += exp_nonpos(fmin(logpp - logp0, 0.0));                   // min(1, exp(logp' - logp0))
            nalpha += 1.0;
            xp_src = -1;
            int l = 0;
            while (l < j) {
              if ((t >> l) & 1u) {   // this subtree is a second half: merge with the pending first half
                const double u = rng.uniform();
                const double nA = Sn[l];
                if (!(u * (nA + Tn) < Tn)) xp_src = l;           // rand() < n''/(n' + n'') fails: keep the first half's proposal
                Tn = nA + Tn;
                // nouturn between the subtree's first leaf (stack) and the current leaf; one code path for both directions:
                // pm = +1: (x - xf).rf >= 0 && (x - xf).r >= 0;  pm = -1: (xf - x).r >= 0 && (xf - x).rf >= 0
                const V4 fx = vld(slot(l, 0), lane), fr = vld(slot(l, 1), lane);
                const double da = cx.a - fx.a, db = cx.b - fx.b, dma = cx.ma - fx.ma, dmb = cx.mb - fx.mb;
                const double p1 = wsum(fma(da, fr.a, db * fr.b)) + fma(dma, fr.ma, dmb * fr.mb);
                const double p2 = wsum(fma(da, cr.a, db * cr.b)) + fma(dma, cr.ma, dmb * cr.mb);
                const double sg = (double)pm;
                const bool ok = sg * p1 >= 0 && sg * p2 >= 0;
                Ts = Ts && ok;
                ++l;
              } else if (Ts) {       // a good first half: park it and build its sibling
                if (l == 0) { vst(slot(0, 0), cx, lane); vst(slot(0, 1), cr, lane); }
                else { vcopy(slot(l, 0), slot(l - 1, 0), lane); vcopy(slot(l, 1), slot(l - 1, 1), lane); }
                if (xp_src < 0) vst(slot(l, 2), cx, lane); else vcopy(slot(l, 2), slot(xp_src, 2), lane);
                if (lane == 0) Sn[l] = Tn;
                __syncwarp();
                break;
              } else {
                ++l;                 // a failed first half: the parent returns it unchanged
              }
            }
            if (l == j) break;
          }
          if (pm == -1) { vst(e_xm, cx, lane); vst(e_rm, cr, lane); vst(e_gm, cg, lane); }
          else { vst(e_xp, cx, lane); vst(e_rp, cr, lane); vst(e_gp, cg, lane); }
          if (Ts) {
            if (rng.uniform() * n < Tn) { if (xp_src < 0) vst(e_v, cx, lane); else vcopy(e_v, slot(xp_src, 2), lane); }
          }
          j += 1;
          n += Tn;
          n_leap += (unsigned long long)nalpha;                  // leaves of this doubling
          if (Ts) {
            const V4 xm = vld(e_xm, lane), xp = vld(e_xp, lane), rm = vld(e_rm, lane), rp = vld(e_rp, lane);
            s = nouturn4(xm, xp, rm, rp);
          } else {
            s = false;
          }
          // the reference would keep doubling (nuts.jl:106-124): counted (mcu_work_count) with an atomic on the spot — a per-warp counter held in
          // registers for the whole kernel cost 15 % (8.65 s instead of 7.5 s per 65,536 x 2,000 launch)
          if (s && j >= cfg.max_depth) { s = false; if (lane == 0 && a.work) atomicAdd(a.work + 1, 1ull); }
        }
        t_alpha = alpha; t_nalpha = nalpha;
        x = vld(e_v, lane);
      }
      if (adapt) {                                               // dual averaging: nuts.jl:66-77
        double p = 1.0 / (t_m + 10.0);                           // t0 = 10
        t_Hbar = (1.0 - p) * t_Hbar + p * (cfg.target - t_alpha / t_nalpha);
        t_eps = exp(t_mu - sqrt(t_m) * t_Hbar / 0.05);           // gamma = 0.05
        p = pow(t_m, -0.75);                                     // kappa = 0.75
        t_epsbar = exp(p * log(t_eps) + (1.0 - p) * log(t_epsbar));
      }
      // ================================================================ block 1: Slice(s2_c, s2_alpha, s2_beta), univariate, constrained scale
      {
        rng.seek((uint32_t)iter, 1);
        const bool act = lane < NR;
        double see = 0.0;
#pragma unroll
        for (int k = 0; k < NOBS; ++k) { const double e = ch.y[k] - (x.a + x.b * ch.xo[k]); see += e * e; }
        const double da = act ? x.a - x.ma : 0.0, db = act ? x.b - x.mb : 0.0;
        const double SEE = wsum(act ? see : 0.0), saa = wsum(da * da), sbb = wsum(db * db);
        // logpdf!(block, v) = IG(s2_c) + IG(s2_alpha) + IG(s2_beta) + sum_i N(alpha_i | mu_alpha, s2_alpha) + sum_i N(beta_i | ..) +
        // N(y | .., s2_c) (simulation.jl:60-67,77-90); with the sufficient statistics fixed it separates into one term per component,
        //   term(x; n, S) = c0 - (1.001 + n/2) log x - (0.001 + S/2) / x - n log(2 pi)/2,   -Inf outside the support,
        // so a univariate update re-evaluates one log and one reciprocal.
        auto term = [&](double xv, double nn, double S) {
          if (!(xv >= 0.0)) return neg_inf();
          return ig_c0 - (1.001 + 0.5 * nn) * log(xv) - (0.001 + 0.5 * S) / xv - 0.5 * nn * kLog2Pi;
        };
        const double cn[3] = {(double)(NR * NOBS), (double)NR, (double)NR}, cS[3] = {SEE, saa, sbb};
        double v[3] = {s2c, s2a, s2b}, lower[3], upper[3], tv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) tv[i] = term(v[i], cn[i], cS[i]);
        double logf0 = (tv[0] + tv[1]) + tv[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) lower[i] = v[i] - cfg.width[i] * rng.uniform();
#pragma unroll
        for (int i = 0; i < 3; ++i) upper[i] = lower[i] + cfg.width[i];
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
          const double p0 = logf0 + log(rng.uniform());
          const double xv = v[i];
          const double others = i == 0 ? tv[1] + tv[2] : (i == 1 ? tv[0] + tv[2] : tv[0] + tv[1]);
          const double nn = i == 0 ? cn[0] : cn[1], S = i == 0 ? cS[0] : (i == 1 ? cS[1] : cS[2]);
          double lo = i == 0 ? lower[0] : (i == 1 ? lower[1] : lower[2]), up = i == 0 ? upper[0] : (i == 1 ? upper[1] : upper[2]);
          double cur = lo + (up - lo) * rng.uniform(), tc;
          while (true) {
            tc = term(cur, nn, S);
            logf0 = others + tc;
            if (!(logf0 < p0)) break;
            if (cur < xv) lo = cur; else up = cur;
            cur = lo + (up - lo) * rng.uniform();
          }
          if (i == 0) { v[0] = cur; tv[0] = tc; } else if (i == 1) { v[1] = cur; tv[1] = tc; } else { v[2] = cur; tv[2] = tc; }
        }
        s2c = v[0]; s2a = v[1]; s2b = v[2];
      }
      // ================================================================ thinning + streaming moments (mcmc.jl:76-78)
      if (iter > a.burnin && (iter - a.burnin) % a.thin == 0 && lane == 0) {
        double mon[RatsModel::P];
        mon[0] = x.mb; mon[1] = x.ma - cfg.xbar * x.mb; mon[2] = s2c;   // mu_beta, alpha0 (rats.jl:64-66), s2_c
        if (a.samples) {
          const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
          for (int jm = 0; jm < RatsModel::P; ++jm) a.samples[((size_t)row * RatsModel::P + jm) * C + c] = mon[jm];
        }
        if (a.comom) comoments_update(a.mom, a.momn, a.comom, C, (size_t)c, RatsModel::P, mon, a.log_mask);
        moments_update(a.mom, a.momn, C, (size_t)c, RatsModel::P, mon);
      }
      __syncwarp();
    }
    // ---- store chain state and tune
    if (lane == 0) {
      a.state[0 * C + c] = x.ma; a.state[1 * C + c] = x.mb; a.state[2 * C + c] = s2a; a.state[3 * C + c] = s2b; a.state[4 * C + c] = s2c;
      tn[0 * C] = t_adapt; tn[1 * C] = t_alpha; tn[2 * C] = t_eps; tn[3 * C] = t_epsbar; tn[4 * C] = t_Hbar; tn[5 * C] = t_m;
      tn[6 * C] = t_mu; tn[7 * C] = t_nalpha;
    }
    if (lane < NR) { a.state[(size_t)(5 + lane) * C + c] = x.a; a.state[(size_t)(5 + NR + lane) * C + c] = x.b; }
    __syncwarp();
  }
  if (lane == 0 && a.work && n_leap) atomicAdd(a.work, n_leap);
}

}  // namespace

// Number of warps the persistent grid will run (the caller sizes the deep-level scratch with it).
int rats_warp_grid(long long n_chains) {
  const size_t smem = (size_t)kWarpsPerBlock * kWarpDoubles * sizeof(double);
  static thread_local int per_sm = 0;
  if (per_sm == 0) {
    if (cudaFuncSetAttribute(rats_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rats_warp_kernel, kWarpsPerBlock * 32, smem) != cudaSuccess || per_sm < 1) return -1;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (n_chains + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long long cap = (long long)sms * per_sm;
  if (const char* e = std::getenv("MCU_RATS_BLOCKS_PER_SM")) { const long long v = std::atoll(e); if (v > 0 && v < per_sm) cap = (long long)sms * v; }   // occupancy experiments
  return (int)(want < cap ? want : cap);
}
size_t rats_warp_scratch_bytes(int grid) { return (size_t)grid * kWarpsPerBlock * (kMaxDepth - LS) * 3 * VEC * sizeof(double); }

// y, Xm, rat: the model inputs on the host (150 entries); returns 0 on success, -2 when the data do not have 5 observations per rat
int rats_warp_launch(const double* y, const double* Xm, const double* rat, int N, double xbar, const RunArgs& a, const DevBlock* h_blocks,
                     const double* h_width, int grid, double* scratch, cudaStream_t st) {
  WarpCfg cfg;
  if (N != NR * NOBS) return -2;
  int cnt[NR] = {0};
  for (int l = 0; l < 32; ++l) for (int k = 0; k < NOBS; ++k) { cfg.y[k][l] = 0.0; cfg.x[k][l] = 0.0; }
  for (int k = 0; k < N; ++k) {
    const int i = (int)rat[k];
    if (i < 0 || i >= NR || cnt[i] >= NOBS) return -2;
    cfg.y[cnt[i]][i] = y[k]; cfg.x[cnt[i]][i] = Xm[k]; ++cnt[i];
  }
  for (int i = 0; i < NR; ++i) if (cnt[i] != NOBS) return -2;
  for (int i = 0; i < 3; ++i) cfg.width[i] = h_width[i];
  const DevBlock& nb = h_blocks[0];
  cfg.target = nb.target; cfg.eps_desc = nb.epsilon; cfg.xbar = xbar;
  cfg.max_depth = nb.max_depth > 0 ? (nb.max_depth < kMaxDepth ? nb.max_depth : kMaxDepth) : kMaxDepth;
  cfg.tune_off = nb.tune_off;
  cfg.scratch = scratch;
  const size_t smem = (size_t)kWarpsPerBlock * kWarpDoubles * sizeof(double);
  rats_warp_kernel<<<grid, kWarpsPerBlock * 32, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mcu
