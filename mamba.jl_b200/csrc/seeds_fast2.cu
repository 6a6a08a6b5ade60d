// seeds_fast2.cu — the fused seeds kernel with TWO THREADS PER CHAIN (round 2; the one-thread-per-chain form is seeds_fast.cu).
//
// Why: seeds_fast_kernel keeps b_i, e_i = exp(eta_i), L_i = log(1 + e_i) and the proposed L_i of its 21 plates in shared memory —
// 704 B per chain — so an SM holds 288 chains = 9 warps, and the kernel is bound by dependent-issue latency (ncu: 2.1 warps per
// scheduler, issue slots 46 % busy, FP64 pipe 33 %: profiles/r1_seeds_fast_summary.md).  More chains do not fit; more THREADS do:
// the 21 plates of a chain are conditionally independent given (alpha, s2), so two adjacent lanes share a chain — lane h owns
// the Philox pairs p = h, h + 2, ... of plates (2p, 2p + 1) — with the same shared-memory footprint per chain and twice the warps
// per SM, each with about half the dependent chain per iteration:
//   * block b (AMWG over the 21 b_i, src/samplers/amwg.jl:99-115): each lane updates its own plates (6 / 5 trips instead of 11);
//   * blocks alpha (AMWG or the reference's AMM, doc/examples/seeds.jl:69-71): the proposal, its exp and the MH test are evaluated by
//     both lanes on identical register copies of (alpha, sigma, accept counts) — same inputs, same instructions, same bits — each
//     lane sums n_i (L_i' - L_i) over its own affected plates and ONE pair shuffle adds the two halves (a + b is commutative, so both
//     lanes hold the identical total and take the identical decision);
//   * block s2: sum b_i^2 the same way.
// Draw order, proposals and decisions are those of the reference on the same stream (Philox counters are position-based, rng.cuh):
// tests/test_gpu_baseline_shapes.py audits 125,000 chains x 2,000 iterations against the oracle.
#include <type_traits>

#include "launch.hpp"
#include "fastmath.cuh"

#ifndef MCU_SEEDS2_BS
#define MCU_SEEDS2_BS 192      // threads per block = 96 chains; 3 blocks per SM = 288 chains (125,000 chains per GPU = 2.93 rounds of that)
#endif
#ifndef MCU_SEEDS2_MAXREG
#define MCU_SEEDS2_MAXREG 112  // 3 blocks x 192 threads per SM; ptxas wants 126 and spills 12 B at 112 (it picks 96 + 176 B of spills when given __launch_bounds__(192, 3))
#endif

namespace mcu {

namespace {

constexpr int NPL = SeedsModel::NP;   // 21 plates
constexpr int HS = 12;                // local slots per lane: up to 11 plates + the dummy slot 11 (e = 0, L = 0, n = 0)
constexpr int DUMMY = HS - 1;

struct Fast2Cfg {
  double rl[2][HS], nl[2][HS];        // r_i, n_i by (lane half, local slot); 0 in unused slots
  double scale_bl[2][HS];             // initial sigma of b by (half, slot)
  unsigned char plate[2][HS];         // local slot -> plate id (NPL: none)
  unsigned char grpl[2][HS];          // design group of the slot's plate: 0:(x1=0,x2=0) 1:(0,1) 2:(1,0) 3:(1,1)
  unsigned char alist[4][2][HS];      // slots whose eta depends on alpha_j, padded with DUMMY to a multiple of 3
  unsigned char atriples[4];          // trips of the alpha_j loop (the longer half's; the shorter one pads with DUMMY)
  unsigned short amask[4][2];         // bit s: slot s depends on alpha_j
  double rsum[4];                     // sum of r_i over the plates that depend on alpha_j (both halves)
  int adapt[3], batchsize[3], tune_off[3];
  double target[3];
  double scale_a[4], scale_s;
  double amm_SL[16], amm_beta, amm_scale;
};

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
MCU_D double pick(const Bases& g, unsigned grp) {
  const double lo = (grp & 1u) ? g.g1 : g.g0, hi = (grp & 1u) ? g.g3 : g.g2;
  return (grp & 2u) ? hi : lo;
}
MCU_D double pick4(double v0, double v1, double v2, double v3, unsigned q) { return (q & 2u) ? ((q & 1u) ? v3 : v2) : ((q & 1u) ? v1 : v0); }

template <int BS, bool AMM0>
__global__ void __maxnreg__(MCU_SEEDS2_MAXREG) seeds_fast2_kernel(const __grid_constant__ Fast2Cfg cfg, const __grid_constant__ RunArgs a) {
  extern __shared__ double smem[];
  double* sb = smem;                        // b[slot]
  double* se = smem + HS * BS;              // e = exp(eta)
  double* sll = smem + 2 * HS * BS;         // L = log(1 + e)
  double* sln = smem + 3 * HS * BS;         // proposed L
  const int tid = threadIdx.x;
  const int half = tid & 1;
  const long long c = ((long long)blockIdx.x * BS + tid) >> 1;
  if (c >= a.n_chains) return;              // both lanes of a chain leave together
  const unsigned pairmask = 3u << (tid & 30);
  const size_t C = (size_t)a.n_chains;
  const uint32_t chain = (uint32_t)(a.chain_offset + c);
#define SB(i) sb[(i) * BS + tid]
#define SE(i) se[(i) * BS + tid]
#define SLL(i) sll[(i) * BS + tid]
#define SLN(i) sln[(i) * BS + tid]
#define TUNE(blk, slot) a.tune[(size_t)(cfg.tune_off[blk] + (slot)) * C + c]
#define SSG(pl) TUNE(1, 2 + (pl))
#define SAC(pl) TUNE(1, 2 + NPL + (pl))
#define PAIRSUM(v) ((v) + __shfl_xor_sync(pairmask, (v), 1))

  // ---- load chain state (scalars: identical register copies in both lanes) ------------------------
  double al0 = a.state[0 * C + c], al1 = a.state[1 * C + c], al2 = a.state[2 * C + c], al3 = a.state[3 * C + c];
  double s2 = a.state[4 * C + c];
  double x = log(s2);
#pragma unroll
  for (int s = 0; s < HS; ++s) { const int pl = cfg.plate[half][s]; SB(s) = pl < NPL ? a.state[(size_t)(5 + pl) * C + c] : 0.0; }
  double m0 = AMM0 ? 0.0 : TUNE(0, 0), m1 = TUNE(1, 0), m2 = TUNE(2, 0);
  bool ad0 = AMM0 ? false : TUNE(0, 1) != 0.0, ad1 = TUNE(1, 1) != 0.0, ad2 = TUNE(2, 1) != 0.0;
  double sg0 = AMM0 ? 0.0 : TUNE(0, 2), sg1 = AMM0 ? 0.0 : TUNE(0, 3), sg2 = AMM0 ? 0.0 : TUNE(0, 4), sg3 = AMM0 ? 0.0 : TUNE(0, 5);
  int ac0 = AMM0 ? 0 : (int)TUNE(0, 6), ac1 = AMM0 ? 0 : (int)TUNE(0, 7), ac2 = AMM0 ? 0 : (int)TUNE(0, 8), ac3 = AMM0 ? 0 : (int)TUNE(0, 9);
  double sgs = TUNE(2, 2); int acs = (int)TUNE(2, 3);

  Bases g = group_bases(al0, al1, al2, al3);
  for (int s = 0; s < HS; ++s) {
    if (cfg.plate[half][s] < NPL) { const double e = fast_exp(pick(g, cfg.grpl[half][s]) + SB(s)); SE(s) = e; SLL(s) = fast_log(1.0 + e); }
    else { SE(s) = 0.0; SLL(s) = 0.0; }
    SLN(s) = 0.0;
  }

  for (long long it = 1; it <= a.iters; ++it) {
    const long long iter = a.iter0 + it;
    const uint32_t it32 = (uint32_t)iter;
    if (iter == 1) {   // SamplerVariate(block, sigma): fresh tune records at iter == 1 (sampler.jl:40-45, amwg.jl:14-21, amm.jl:14-24)
      m0 = m1 = m2 = 0.0; ad0 = ad1 = ad2 = false;
      sg0 = cfg.scale_a[0]; sg1 = cfg.scale_a[1]; sg2 = cfg.scale_a[2]; sg3 = cfg.scale_a[3];
      ac0 = ac1 = ac2 = ac3 = 0;
      for (int s = 0; s < HS; ++s) { const int pl = cfg.plate[half][s]; if (pl < NPL) { SSG(pl) = cfg.scale_bl[half][s]; SAC(pl) = 0.0; } }
      sgs = cfg.scale_s; acs = 0;
      if (AMM0 && half == 0) for (int i = 0; i < 38; ++i) TUNE(0, i) = 0.0;
    }
    // ================================================================== block 0, AMM form (amm.jl:66-108; generic form: samplers.cuh amm_sample)
    if (AMM0) {
      const bool adapt = cfg.adapt[0] == 1 ? iter <= a.burnin : cfg.adapt[0] == 0;
      double v[4] = {al0, al1, al2, al3};
      double xp[4] = {0.0, 0.0, 0.0, 0.0};
      double m = 0.0;
      // the tune record (running mean / second moment / adapted factor) is read-modify-written in the L2-resident tune array: lane 0 of
      // the pair owns it and hands the proposal to its partner
      if (half == 0) {
        const bool was = TUNE(0, 0) != 0.0;
        if (adapt && !was) {   // setadapt!: amm.jl:97-108
          TUNE(0, 1) = 0.0;
          for (int i = 0; i < 4; ++i) TUNE(0, 2 + i) = v[i];
          for (int i = 0; i < 4; ++i) for (int cc = 0; cc < 4; ++cc) TUNE(0, 6 + i + cc * 4) = v[i] * v[cc];
          for (int i = 0; i < 16; ++i) TUNE(0, 22 + i) = 0.0;
        }
        TUNE(0, 0) = adapt ? 1.0 : 0.0;
        m = TUNE(0, 1);
        const Pair z01 = draw_normal_pair(a, chain, it32, 0, 0), z23 = draw_normal_pair(a, chain, it32, 0, 1);
        const double z[4] = {z01.a, z01.b, z23.a, z23.b};
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
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) xp[i] = __shfl_sync(pairmask, xp[i], tid & 30);
      const double lu = log_uniform(draw_uniform_pair(a, chain, it32, 0, 0).a);
      // logf(x) - logf(v): every plate of group q moves by dg[q]; e_i' = e_i exp(dg[grp_i]) (4 exps, 21 logs), priors Normal(0, 1000)
      const Bases gn = group_bases(xp[0], xp[1], xp[2], xp[3]);
      const double dg0 = gn.g0 - g.g0, dg1 = gn.g1 - g.g1, dg2 = gn.g2 - g.g2, dg3 = gn.g3 - g.g3;
      const double E0 = fast_exp(dg0), E1 = fast_exp(dg1), E2 = fast_exp(dg2), E3 = fast_exp(dg3);
      double delta = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) delta = fma(-0.5e-6, fma(xp[i], xp[i], -v[i] * v[i]), delta);
      double dLa = 0.0, dLb = 0.0, dLc = 0.0;
#pragma unroll 1
      for (int s = 0; s < HS; s += 3) {   // own slots, three independent logs per trip (unused slots: e = 0, r = n = 0)
        const unsigned q0 = cfg.grpl[half][s], q1 = cfg.grpl[half][s + 1], q2 = cfg.grpl[half][s + 2];
        const double la = fast_log(fma(SE(s), pick4(E0, E1, E2, E3, q0), 1.0)), lb = fast_log(fma(SE(s + 1), pick4(E0, E1, E2, E3, q1), 1.0)),
                     lc = fast_log(fma(SE(s + 2), pick4(E0, E1, E2, E3, q2), 1.0));
        SLN(s) = la; SLN(s + 1) = lb; SLN(s + 2) = lc;
        dLa += fma(cfg.rl[half][s], pick4(dg0, dg1, dg2, dg3, q0), -cfg.nl[half][s] * (la - SLL(s)));
        dLb += fma(cfg.rl[half][s + 1], pick4(dg0, dg1, dg2, dg3, q1), -cfg.nl[half][s + 1] * (lb - SLL(s + 1)));
        dLc += fma(cfg.rl[half][s + 2], pick4(dg0, dg1, dg2, dg3, q2), -cfg.nl[half][s + 2] * (lc - SLL(s + 2)));
      }
      const double dl_half = (dLa + dLb) + dLc;
      delta += PAIRSUM(dl_half);
      if (lu < delta) {   // rand() < exp(logf(x) - logf(v)): amm.jl:80
        al0 = xp[0]; al1 = xp[1]; al2 = xp[2]; al3 = xp[3];
        v[0] = xp[0]; v[1] = xp[1]; v[2] = xp[2]; v[3] = xp[3];
        g = gn;
        for (int s = 0; s < HS; ++s) if (cfg.plate[half][s] < NPL) { SLL(s) = SLN(s); SE(s) = SE(s) * pick4(E0, E1, E2, E3, cfg.grpl[half][s]); }
      }
      if (adapt && half == 0) {   // running mean / second moment, Sigma = (scale^2 / n / p)(Mvv - Mv Mv'), pivoted Cholesky: amm.jl:83-91
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
      __syncwarp(pairmask);
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
        double lu;                                                        // log of uniform j of the block, off the critical path
        if ((j & 1) == 0) { const Pair pr = draw_uniform_pair(a, chain, it32, 0, j >> 1); lu = log_uniform(pr.a); uc = log_uniform(pr.b); } else lu = uc;
        const double z = sg0 * zn01;                                      // z = sigma .* randn(n): normal j of the block
        const double anew = al0 + z;
        // proposed group bases, again in the reference's summation order (slot s holds alpha_{(j+s)%4})
        const double q0 = j == 0 ? anew : (j == 1 ? al3 : (j == 2 ? al2 : al1));
        const double q1 = j == 0 ? al1 : (j == 1 ? anew : (j == 2 ? al3 : al2));
        const double q2 = j == 0 ? al2 : (j == 1 ? al1 : (j == 2 ? anew : al3));
        const double q3 = j == 0 ? al3 : (j == 1 ? al2 : (j == 2 ? al1 : anew));
        const Bases gn = group_bases(q0, q1, q2, q3);
        // every affected plate moves by the same step: e_i' = e_i exp(z); ll_i' - ll_i = r_i z - n_i (L_i' - L_i)
        const double E = fast_exp(z);
        double dLa = 0.0, dLb = 0.0, dLc = 0.0;
        const int nt = cfg.atriples[j];
#pragma unroll 1
        for (int k = 0; k < nt; ++k) {   // this lane's affected plates, three per trip
          const int ia = cfg.alist[j][half][3 * k], ib = cfg.alist[j][half][3 * k + 1], ic = cfg.alist[j][half][3 * k + 2];
          const double la = fast_log(fma(SE(ia), E, 1.0)), lb = fast_log(fma(SE(ib), E, 1.0)), lc = fast_log(fma(SE(ic), E, 1.0));
          SLN(ia) = la; SLN(ib) = lb; SLN(ic) = lc;
          dLa = fma(cfg.nl[half][ia], la - SLL(ia), dLa);
          dLb = fma(cfg.nl[half][ib], lb - SLL(ib), dLb);
          dLc = fma(cfg.nl[half][ic], lc - SLL(ic), dLc);
        }
        const double dl_half = (dLa + dLb) + dLc;
        double delta = fma(cfg.rsum[j], z, -PAIRSUM(dl_half));
        delta = fma(-0.5e-6, fma(anew, anew, -al0 * al0), delta);         // Normal(0, 1000) prior of the component
        if (lu < delta) {                                                 // rand() < exp(delta) on the log scale
          al0 = anew;
          g = gn;   // bases of groups that do not contain alpha_j are recomputed to the same value
          const unsigned pm = cfg.amask[j][half];
          for (int s = 0; s < HS - 1; ++s) if ((pm >> s) & 1u) { SLL(s) = SLN(s); SE(s) = SE(s) * E; }
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
    // ================================================================== block 1: AMWG(b), this lane's plates
    {
      const bool adapt = cfg.adapt[1] == 1 ? iter <= a.burnin : cfg.adapt[1] == 0;
      if (adapt && !ad1) { for (int s = 0; s < HS; ++s) { const int pl = cfg.plate[half][s]; if (pl < NPL) SAC(pl) = 0.0; } m1 = 0.0; }
      ad1 = adapt;
      if (adapt) m1 += 1.0;
      const double half_inv_s2 = 0.5 / s2;                                // b ~ Normal(0, sqrt(s2)): -(b/sigma)^2 / 2 = -b^2 / (2 s2)
      // trip k = Philox pair 2k + half of the block = plates (2p, 2p + 1) = local slots (2k, 2k + 1): two independent exp -> log -> compare
      // chains in straight-line code.  A slot without a plate (the odd 21st, or the 6th pair of lane 1) works on the dummy and is never accepted.
#pragma unroll 1
      for (int k = 0; k < HS / 2; ++k) {
        const int s0 = 2 * k, s1 = 2 * k + 1;
        const int p0 = cfg.plate[half][s0], p1 = cfg.plate[half][s1];
        const bool r0 = p0 < NPL, r1 = p1 < NPL;
        if (!r0) break;                                                   // (pairs are filled in order: nothing follows an empty pair)
        const int i0 = s0, i1 = r1 ? s1 : DUMMY;
        const double sga = SSG(p0), sgb = r1 ? SSG(p1) : 0.0;            // global (L2) loads, issued ahead of their use
        const double aca = adapt ? SAC(p0) : 0.0, acb = (adapt && r1) ? SAC(p1) : 0.0;
        const double bia = SB(i0), bib = SB(i1);
        const uint32_t pidx = (uint32_t)(2 * k + half);
        const Pair pz = draw_normal_pair(a, chain, it32, 1, pidx);
        const Pair pu = draw_uniform_pair(a, chain, it32, 1, pidx);
        const double lua = log_uniform(pu.a), lub = log_uniform(pu.b);
        const double bna = bia + sga * pz.a, bnb = bib + sgb * pz.b;
        const double ena = fast_exp(pick(g, cfg.grpl[half][s0]) + bna);   // fresh e_i: also resets the drift of the alpha updates
        const double enb = fast_exp(pick(g, cfg.grpl[half][s1]) + bnb);
        const double lna = fast_log(1.0 + ena), lnb = fast_log(1.0 + enb);
        const double da = fma(cfg.rl[half][s0], bna - bia, -cfg.nl[half][s0] * (lna - SLL(i0))) - half_inv_s2 * fma(bna, bna, -bia * bia);
        const double db = fma(cfg.rl[half][s1], bnb - bib, -cfg.nl[half][s1] * (lnb - SLL(i1))) - half_inv_s2 * fma(bnb, bnb, -bib * bib);
        const bool acca = lua < da, accb = r1 && lub < db;
        if (acca) { SB(i0) = bna; SE(i0) = ena; SLL(i0) = lna; if (adapt) SAC(p0) = aca + 1.0; }
        if (accb) { SB(i1) = bnb; SE(i1) = enb; SLL(i1) = lnb; if (adapt) SAC(p1) = acb + 1.0; }
      }
      if (adapt && ((long long)m1 % cfg.batchsize[1]) == 0) {
        const double dl = amwg_delta(m1, cfg.batchsize[1]);
        const double up = exp(dl), dn = exp(-dl);
        for (int s = 0; s < HS; ++s) { const int pl = cfg.plate[half][s]; if (pl < NPL) SSG(pl) = SSG(pl) * ((SAC(pl) / m1 < cfg.target[1]) ? dn : up); }
      }
    }
    // ================================================================== block 2: AMWG(s2) on x = log s2
    {
      const bool adapt = cfg.adapt[2] == 1 ? iter <= a.burnin : cfg.adapt[2] == 0;
      if (adapt && !ad2) { acs = 0; m2 = 0.0; }
      ad2 = adapt;
      if (adapt) m2 += 1.0;
      double Sh = 0.0;
      for (int s = 0; s < HS - 1; ++s) { const double bi = SB(s); Sh += bi * bi; }   // unused slots hold 0
      const double S = PAIRSUM(Sh);
      const double lus = log_uniform(draw_uniform_pair(a, chain, it32, 2, 0).a);
      const double xn = x + sgs * draw_normal_pair(a, chain, it32, 2, 0).a;
      const double s2n = (xn > -700.0 && xn < 700.0) ? fast_exp(xn) : exp(xn);
      // logf(x) = InverseGamma(0.001, 0.001)(s2) + x [log-Jacobian, transformdistribution.jl:75-78] + sum_i Normal(b_i; 0, sqrt(s2))
      const double dx = xn - x;
      const double dinv = 1.0 / s2n - 1.0 / s2;
      const double delta = -(0.001 + 1.0) * dx - 0.001 * dinv + dx - 0.5 * S * dinv - (double)NPL * 0.5 * dx;
      if (lus < delta) { x = xn; s2 = s2n; if (adapt) acs += 1; }
      if (adapt && ((long long)m2 % cfg.batchsize[2]) == 0) {
        const double dl = amwg_delta(m2, cfg.batchsize[2]);
        sgs *= exp((double)acs / m2 < cfg.target[2] ? -dl : dl);
      }
    }
    // ================================================================== thinning + streaming moments (lane 0 of the pair)
    if (half == 0 && iter > a.burnin && (iter - a.burnin) % a.thin == 0) {   // mcmc.jl:76-78
      double mon[SeedsModel::P];
      mon[0] = al0; mon[1] = al1; mon[2] = al2; mon[3] = al3; mon[4] = s2;
      if (a.samples) {
        const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
        for (int j = 0; j < SeedsModel::P; ++j) a.samples[((size_t)row * SeedsModel::P + j) * C + c] = mon[j];
      }
      moments_update(a.mom, a.momn, C, (size_t)c, SeedsModel::P, mon);
    }
  }
  // ---- store chain state ------------------------------------------------------------------------
  for (int s = 0; s < HS; ++s) { const int pl = cfg.plate[half][s]; if (pl < NPL) a.state[(size_t)(5 + pl) * C + c] = SB(s); }
  if (half == 0) {
    a.state[0 * C + c] = al0; a.state[1 * C + c] = al1; a.state[2 * C + c] = al2; a.state[3 * C + c] = al3;
    a.state[4 * C + c] = s2;
    if (!AMM0) {
      TUNE(0, 0) = m0; TUNE(0, 1) = ad0 ? 1.0 : 0.0;
      TUNE(0, 2) = sg0; TUNE(0, 3) = sg1; TUNE(0, 4) = sg2; TUNE(0, 5) = sg3;
      TUNE(0, 6) = ac0; TUNE(0, 7) = ac1; TUNE(0, 8) = ac2; TUNE(0, 9) = ac3;
    }
    TUNE(1, 0) = m1; TUNE(1, 1) = ad1 ? 1.0 : 0.0;
    TUNE(2, 0) = m2; TUNE(2, 1) = ad2 ? 1.0 : 0.0; TUNE(2, 2) = sgs; TUNE(2, 3) = acs;
  }
#undef SB
#undef SE
#undef SLL
#undef SLN
#undef SSG
#undef SAC
#undef TUNE
#undef PAIRSUM
}

template <int BS, bool AMM0>
int launch2(const Fast2Cfg& cfg, const RunArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)BS * 4 * HS * sizeof(double);
  static thread_local int attr_dev = -1;   // the attribute call is slow: once per device
  int dev = 0; cudaGetDevice(&dev);
  if (attr_dev != dev) {
    if (cudaFuncSetAttribute(seeds_fast2_kernel<BS, AMM0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    attr_dev = dev;
  }
  const long long per_block = BS / 2;
  const unsigned grid = (unsigned)((a.n_chains + per_block - 1) / per_block);
  seeds_fast2_kernel<BS, AMM0><<<grid, BS, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

// Same contract as seeds_fast_launch (seeds_fast.cu): 0 on success, -2 when the design is not 0/1 indicators.
int seeds_fast2_launch(const double* r, const double* n, const double* x1, const double* x2, const RunArgs& a, const DevBlock* h_blocks,
                       const std::vector<std::vector<double>>& h_scales, const double* h_SigmaL, cudaStream_t st) {
  Fast2Cfg cfg;
  const bool amm0 = h_blocks[0].kind == 6;   // MCU_AMM
  int grp[NPL]; unsigned amask[4] = {0, 0, 0, 0};
  for (int i = 0; i < NPL; ++i) {
    if ((x1[i] != 0.0 && x1[i] != 1.0) || (x2[i] != 0.0 && x2[i] != 1.0)) return -2;   // design must be 0/1 indicators
    grp[i] = (x1[i] != 0.0 ? 2 : 0) + (x2[i] != 0.0 ? 1 : 0);
    amask[0] |= 1u << i;
    if (x1[i] != 0.0) amask[1] |= 1u << i;
    if (x2[i] != 0.0) amask[2] |= 1u << i;
    if (x1[i] != 0.0 && x2[i] != 0.0) amask[3] |= 1u << i;
  }
  for (int h = 0; h < 2; ++h) {
    for (int s = 0; s < HS; ++s) {
      // lane h owns the Philox pairs p = 2k + h of block b: plates (2p, 2p + 1) in local slots (2k, 2k + 1)
      const int p = 2 * (s >> 1) + h, pl = 2 * p + (s & 1);
      const bool real = pl < NPL && s < HS;
      cfg.plate[h][s] = (unsigned char)(real ? pl : NPL);
      cfg.rl[h][s] = real ? r[pl] : 0.0; cfg.nl[h][s] = real ? n[pl] : 0.0;
      cfg.grpl[h][s] = (unsigned char)(real ? grp[pl] : 0);
      cfg.scale_bl[h][s] = real ? h_scales[1][pl] : 0.0;
    }
    if (cfg.plate[h][DUMMY] != NPL) return -1;   // the dummy slot must stay free (21 plates: lane 0 uses slots 0..10, lane 1 slots 0..9)
  }
  for (int j = 0; j < 4; ++j) {
    cfg.rsum[j] = 0.0; for (int i = 0; i < NPL; ++i) if ((amask[j] >> i) & 1u) cfg.rsum[j] += r[i];
    int trips = 0;
    for (int h = 0; h < 2; ++h) {
      int cnt = 0; cfg.amask[j][h] = 0;
      for (int s = 0; s < HS; ++s) {
        const int pl = cfg.plate[h][s];
        if (pl < NPL && ((amask[j] >> pl) & 1u)) { cfg.alist[j][h][cnt++] = (unsigned char)s; cfg.amask[j][h] |= (unsigned short)(1u << s); }
      }
      for (int k = cnt; k < HS; ++k) cfg.alist[j][h][k] = (unsigned char)DUMMY;
      trips = std::max(trips, (cnt + 2) / 3);
    }
    cfg.atriples[j] = (unsigned char)trips;
    cfg.scale_a[j] = amm0 ? 0.0 : h_scales[0][j];
  }
  cfg.scale_s = h_scales[2][0];
  for (int b = 0; b < 3; ++b) {
    cfg.adapt[b] = h_blocks[b].adapt; cfg.batchsize[b] = h_blocks[b].batchsize; cfg.tune_off[b] = h_blocks[b].tune_off;
    cfg.target[b] = h_blocks[b].target;
  }
  for (int i = 0; i < 16; ++i) cfg.amm_SL[i] = 0.0;
  cfg.amm_beta = 0.0; cfg.amm_scale = 0.0;
  if (amm0) {
    if (!h_SigmaL) return -1;
    for (int i = 0; i < 16; ++i) cfg.amm_SL[i] = h_SigmaL[i];
    cfg.amm_beta = h_blocks[0].beta; cfg.amm_scale = h_blocks[0].amm_scale;
    return launch2<MCU_SEEDS2_BS, true>(cfg, a, st);
  }
  return launch2<MCU_SEEDS2_BS, false>(cfg, a, st);
}

}  // namespace mcu
