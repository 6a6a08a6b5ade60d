// seeds_ws.cuh — warp-specialised form of the fused seeds / AMWG kernel (included by seeds_fast.cu after FastCfg and its helpers).
//
// Why: in seeds_fast_kernel a third of every chain's instruction stream is random-number work — 28 Philox4x32-10 blocks, 14 Box-Muller
// transforms and 27 float brackets of log u per iteration — sitting on the same dependent chain as the updates it feeds, and neither more
// resident warps nor more ILP per warp move that kernel (profiles/r2_seeds_fast_summary.md).  But the draws are COUNTER-BASED: they need
// nothing from the chain's state.  So a block here is CW consumer warps (one chain per thread, exactly the update arithmetic of
// seeds_fast_kernel) plus ONE producer warp that generates the draws of the block's chains a stage ahead into a two-slot ring in shared
// memory; each producer lane serves CW chains (lane, lane + 32, ...), i.e. 2 CW independent Philox chains in flight — the ILP the
// consumer never had.  The consumers' instruction stream shrinks by the draws (and by their registers), the producer's loop is ~700
// instructions that stay in the instruction cache.
//
// A stage = the two normals and the two log-uniform brackets of Philox pair (block, kpair) of the RNG contract (rng.cuh):
//   iteration = stages (0,0) (0,1) | (1,0) ... (1,10) | (2,0)  — 14 stages, consumed in this order.
// Ring slot layout per chain: zn[2] doubles + la[2] floats (24 B); 2 slots.  The exact log u of the rare in-band MH test is recomputed
// from the counter on the spot (cold path), so the uniform itself is not stored.
// Hand-off: named barriers (bar.sync / bar.arrive, ids 1..4; id 0 stays __syncthreads): FULL[slot] producer arrives / consumers sync,
// EMPTY[slot] consumers arrive / producer syncs; every barrier counts all (CW + 1) * 32 threads of the block.
// Decisions, proposals and cached values are those of seeds_fast_kernel bit for bit (same draws, same arithmetic, same order).
#pragma once

namespace mcu {
namespace {

#ifndef MCU_SEEDS_WS_CW
#define MCU_SEEDS_WS_CW 3      // consumer warps per block (96 chains); 3 blocks per SM = 288 chains per SM
#endif
#ifndef MCU_SEEDS_WS_PW
#define MCU_SEEDS_WS_PW 2      // producer warps per block: producer p generates the stages s = p (mod PW) and owns ring slot p (PW = 1 or 2)
#endif

constexpr int kWsStages = 2 + (NPL + 1) / 2 + 1;   // 14

MCU_D void ws_bar_sync(int id, int n) { asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
MCU_D void ws_bar_arrive(int id, int n) { asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// cold path of an MH test: the float bracket could not decide, so log u is evaluated in FP64 from the regenerated uniform
static __device__ __noinline__ bool ws_logu_less_exact(unsigned long long seed, uint32_t chain, uint32_t it32, uint32_t blk, uint32_t kpair, int which,
                                                       double delta) {
  uint32_t w[4];
  philox4x32_10(kpair, it32, chain, blk, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const double u = which ? u53(w[2], w[3]) : u53(w[0], w[1]);
  return log_uniform(u) < delta;
}

template <int CW, int PW>
__global__ void __launch_bounds__((CW + PW) * 32, MCU_SEEDS_MINB) seeds_ws_kernel(const __grid_constant__ FastCfg cfg, const __grid_constant__ RunArgs a) {
  constexpr int BS = CW * 32;           // chains per block
  constexpr int NT = (CW + 1) * 32;     // threads on a ring barrier: the consumers and the producer that owns the slot
  constexpr int NTB = (CW + PW) * 32;   // threads per block
  static_assert(PW == 1 || (PW == 2 && kWsStages % 2 == 0), "a producer owns a ring slot only if the stage parity is the same in every iteration");
  extern __shared__ double smem[];
  double* se = smem;                        // e[i] = exp(eta_i)
  double* sll = smem + NSL * BS;            // L[i] = log(1 + e[i])
  double* sb = smem + 2 * NSL * BS;         // b[i]
  double* sln = smem + 3 * NSL * BS;        // proposed L[i]
  double* rz = smem + 4 * NSL * BS;         // ring: normals [slot][2][BS]
  double* tlg = rz + 2 * 2 * BS;            // 128 x (invc, logc), 16-byte aligned
  double* tex = tlg + 256;                  // 128 x 2^(j/128)
  float* rl = reinterpret_cast<float*>(tex + 128);   // ring: float log u [slot][2][BS]
  const int tid = threadIdx.x;
  for (int i = tid; i < 256; i += NTB) tlg[i] = kLogTabG[i];
  for (int i = tid; i < 128; i += NTB) tex[i] = kExpTabG[i];
  __syncthreads();
#define FLOG(x) tab::tlog((x), tlg)
#define FEXP(x) tab::texp((x), tex)
  const size_t C = (size_t)a.n_chains;
  const long long cbase = (long long)blockIdx.x * BS;

  if (tid >= BS) {
    // ============================================================== producer warp
    const int lane = (tid - BS) & 31, pw = (tid - BS) >> 5;
    uint32_t chq[CW];
#pragma unroll
    for (int q = 0; q < CW; ++q) {
      long long c = cbase + lane + 32 * q; if (c >= a.n_chains) c = a.n_chains - 1;   // ragged last block: duplicates, never consumed
      chq[q] = (uint32_t)(a.chain_offset + c);
    }
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    long long g = pw;                                  // stage counter over the whole call
#pragma unroll 1
    for (long long it = 1; it <= a.iters; ++it) {
      const uint32_t it32 = (uint32_t)(a.iter0 + it);
#pragma unroll 1
      for (int s = pw; s < kWsStages; s += PW, g += PW) {
        const uint32_t blk = s < 2 ? 0u : (s < kWsStages - 1 ? 1u : 2u);
        const uint32_t kp = s < 2 ? (uint32_t)s : (s < kWsStages - 1 ? (uint32_t)(s - 2) : 0u);
        const int slot = (int)(g & 1);
        double za[CW], zb[CW]; float la[CW], lb[CW];
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          uint32_t w[4], v[4];
          philox4x32_10(kp, it32, chq[q], blk | (1u << 24), k0, k1, w);     // stream 1: normals (both Box-Muller branches)
          philox4x32_10(kp, it32, chq[q], blk, k0, k1, v);                  // stream 0: uniforms
          const double rad = fast_sqrt(-2.0 * FLOG(1.0 - u53(w[0], w[1])));
          const Pair sc = fast_sincos2pi(u53(w[2], w[3]));
          za[q] = rad * sc.b; zb[q] = rad * sc.a;
          la[q] = __log2f((float)u53(v[0], v[1])) * 0.693147181f;           // fastmath.cuh logu_bracket
          lb[q] = __log2f((float)u53(v[2], v[3])) * 0.693147181f;
        }
        if (g >= 2) ws_bar_sync(3 + slot, NT);         // EMPTY[slot]: the consumers have taken the previous contents
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          const int col = lane + 32 * q;
          rz[(slot * 2 + 0) * BS + col] = za[q]; rz[(slot * 2 + 1) * BS + col] = zb[q];
          rl[(slot * 2 + 0) * BS + col] = la[q]; rl[(slot * 2 + 1) * BS + col] = lb[q];
        }
        __threadfence_block();
        ws_bar_arrive(1 + slot, NT);                   // FULL[slot]
      }
    }
    return;
  }

  // ================================================================ consumer warps: one chain per thread
  const bool valid = cbase + tid < a.n_chains;          // ragged last block: the surplus threads shadow the last chain and store nothing
  const long long c = valid ? cbase + tid : a.n_chains - 1;
  const uint32_t chain = (uint32_t)(a.chain_offset + c);
  long long g = 0;
  // next stage of the ring: (normal a, normal b, float log u a, float log u b)
  struct Stage { double za, zb; float la, lb; };
  auto take = [&]() -> Stage {
    const int slot = (int)(g & 1); ++g;
    ws_bar_sync(1 + slot, NT);
    Stage st;
    st.za = rz[(slot * 2 + 0) * BS + tid]; st.zb = rz[(slot * 2 + 1) * BS + tid];
    st.la = rl[(slot * 2 + 0) * BS + tid]; st.lb = rl[(slot * 2 + 1) * BS + tid];
    // the loaded values are operands of the arrive: the loads have returned before the slot is handed back
    asm volatile("barrier.arrive %0, %1;" ::"r"(3 + slot), "r"(NT), "d"(st.za), "d"(st.zb), "f"(st.la), "f"(st.lb) : "memory");
    return st;
  };
  // rand() < exp(delta) on the float bracket (fastmath.cuh logu_less, MCU_LOGU_DOUBLE = 0 form); the exact test regenerates the uniform
  auto mh_less = [&](float la, double delta, uint32_t it32, uint32_t blk, uint32_t kp, int which) -> bool {
    const float band = fmaf(fabsf(la), 4e-6f, 4e-6f);
    const float df = (float)delta;
    if (df > la + band) return true;
    if (!(df > la - band)) return false;
    return ws_logu_less_exact(a.seed, chain, it32, blk, kp, which, delta);
  };
#define SB(i) sb[(i) * BS + tid]
#define SE(i) se[(i) * BS + tid]
#define SLL(i) sll[(i) * BS + tid]
#define SLN(i) sln[(i) * BS + tid]
#define SSG(i) TUNE(1, 2 + (i))
#define SAC(i) TUNE(1, 2 + NPL + (i))
#define TUNE(blk, slot) a.tune[(size_t)(cfg.tune_off[blk] + (slot)) * C + c]

  double al0 = a.state[0 * C + c], al1 = a.state[1 * C + c], al2 = a.state[2 * C + c], al3 = a.state[3 * C + c];
  double s2 = a.state[4 * C + c];
  double x = log(s2);
  for (int i = 0; i < NPL; ++i) SB(i) = a.state[(size_t)(5 + i) * C + c];
  double m0 = TUNE(0, 0), m1 = TUNE(1, 0), m2 = TUNE(2, 0);
  bool ad0 = TUNE(0, 1) != 0.0, ad1 = TUNE(1, 1) != 0.0, ad2 = TUNE(2, 1) != 0.0;
  double sg0 = TUNE(0, 2), sg1 = TUNE(0, 3), sg2 = TUNE(0, 4), sg3 = TUNE(0, 5);
  int ac0 = (int)TUNE(0, 6), ac1 = (int)TUNE(0, 7), ac2 = (int)TUNE(0, 8), ac3 = (int)TUNE(0, 9);
  double sgs = TUNE(2, 2); int acs = (int)TUNE(2, 3);

  Bases gb = group_bases(al0, al1, al2, al3);
  for (int i = 0; i < NPL; ++i) { const double e = FEXP(pick(gb, cfg.grp[i]) + SB(i)); SE(i) = e; SLL(i) = FLOG(1.0 + e); }
  SB(NPL) = 0.0; SE(NPL) = 0.0; SLL(NPL) = 0.0; SLN(NPL) = 0.0;   // dummy slot: log(1 + 0 * E) = 0, n = 0

  double mon[SeedsModel::P];
  for (long long it = 1; it <= a.iters; ++it) {
    const long long iter = a.iter0 + it;
    const uint32_t it32 = (uint32_t)iter;
    if (iter == 1) {   // SamplerVariate(block, sigma): fresh AMWGTune at iter == 1 (sampler.jl:40-45, amwg.jl:14-21)
      m0 = m1 = m2 = 0.0; ad0 = ad1 = ad2 = false;
      sg0 = cfg.scale_a[0]; sg1 = cfg.scale_a[1]; sg2 = cfg.scale_a[2]; sg3 = cfg.scale_a[3];
      ac0 = ac1 = ac2 = ac3 = 0;
      if (valid) for (int i = 0; i < NPL; ++i) { SSG(i) = cfg.scale_b[i]; SAC(i) = 0.0; }
      sgs = cfg.scale_s; acs = 0;
    }
    // ================================================================== block 0: AMWG(alpha0..alpha12)  (amwg.jl:99-115)
    {
      const bool adapt = cfg.adapt[0] == 1 ? iter <= a.burnin : cfg.adapt[0] == 0;
      if (adapt && !ad0) { ac0 = ac1 = ac2 = ac3 = 0; m0 = 0.0; }   // setadapt!: amwg.jl:88-96
      ad0 = adapt;
      if (adapt) m0 += 1.0;
      double zc = 0.0; float lc = 0.f;   // second draw of the current stage
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {      // components are rotated through slot 0 so the loop stays rolled with everything in registers
        double zn01; float lu;
        if ((j & 1) == 0) { const Stage st = take(); zn01 = st.za; zc = st.zb; lu = st.la; lc = st.lb; } else { zn01 = zc; lu = lc; }
        const double z = sg0 * zn01;                                      // z = sigma .* randn(n): normal j of the block
        const double anew = al0 + z;
        const unsigned pm = cfg.amask[j];
        const double q0 = j == 0 ? anew : (j == 1 ? al3 : (j == 2 ? al2 : al1));
        const double q1 = j == 0 ? al1 : (j == 1 ? anew : (j == 2 ? al3 : al2));
        const double q2 = j == 0 ? al2 : (j == 1 ? al1 : (j == 2 ? anew : al3));
        const double q3 = j == 0 ? al3 : (j == 1 ? al2 : (j == 2 ? al1 : anew));
        const Bases gn = group_bases(q0, q1, q2, q3);
        const double E = FEXP(z);                                         // every affected plate moves by the same step: e_i' = e_i exp(z)
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
        delta = fma(-0.5e-6, fma(anew, anew, -al0 * al0), delta);        // Normal(0, 1000) prior: (x / 1000)^2 / 2 without the divisions
        if (mh_less(lu, delta, it32, 0u, (uint32_t)(j >> 1), j & 1)) {   // rand() < exp(delta): amwg.jl:107
          al0 = anew;
          gb = gn;
          for (int i = 0; i < NPL; ++i) if ((pm >> i) & 1u) { SLL(i) = SLN(i); SE(i) = SE(i) * E; }
          if (adapt) ac0 += 1;
        }
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
      if (adapt && !ad1) { if (valid) for (int i = 0; i < NPL; ++i) SAC(i) = 0.0; m1 = 0.0; }
      ad1 = adapt;
      if (adapt) m1 += 1.0;
      const double half_inv_s2 = 0.5 / s2;                                // b ~ Normal(0, sqrt(s2)): -b^2 / (2 s2)
      double psg[2], pac[2];                                              // sigma_b / accept counters of the NEXT trip (L2, ~700 cycles)
      auto b_fetch = [&](int i0) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const bool real = i0 + w < NPL;
          psg[w] = real ? SSG(i0 + w) : 0.0;
          pac[w] = (real && adapt) ? SAC(i0 + w) : 0.0;
        }
      };
      b_fetch(0);
      auto b_trip = [&](int i0, bool full) {                             // two conditionally independent plates per trip (one stage)
        int ix[2]; double sg[2], bi[2], ac[2], zn[2], bn[2], en[2], ln[2]; float lu[2]; bool acc[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const bool real = full || i0 + w < NPL;
          ix[w] = real ? i0 + w : NPL;                                     // past the last plate: the dummy slot (never accepted)
          sg[w] = psg[w]; ac[w] = pac[w];
          bi[w] = SB(ix[w]);
        }
        b_fetch(i0 + 2);
        { const Stage st = take(); zn[0] = st.za; zn[1] = st.zb; lu[0] = st.la; lu[1] = st.lb; }
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const bool real = full || i0 + w < NPL;
          const int ir = real ? i0 + w : 0;                                // constants of a real plate for the dummy chain
          bn[w] = bi[w] + sg[w] * zn[w];
          en[w] = FEXP(pick(gb, cfg.grp[ir]) + bn[w]);                     // fresh e_i: also resets the drift of the alpha updates
          ln[w] = FLOG(1.0 + en[w]);
          const double dl = fma(cfg.r[ir], bn[w] - bi[w], -cfg.n[ix[w]] * (ln[w] - SLL(ix[w]))) - half_inv_s2 * fma(bn[w], bn[w], -bi[w] * bi[w]);
          acc[w] = real && mh_less(lu[w], dl, it32, 1u, (uint32_t)(i0 >> 1), w);
        }
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (acc[w]) { SB(ix[w]) = bn[w]; SE(ix[w]) = en[w]; SLL(ix[w]) = ln[w]; if (adapt && valid) SAC(ix[w]) = ac[w] + 1.0; }
      };
      int i0 = 0;
#pragma unroll 1
      for (; i0 + 2 <= NPL; i0 += 2) b_trip(i0, true);
      if (i0 < NPL) b_trip(i0, false);
      if (adapt && ((long long)m1 % cfg.batchsize[1]) == 0) {
        const double dl = amwg_delta(m1, cfg.batchsize[1]);
        const double up = exp(dl), dn = exp(-dl);
        if (valid) for (int i = 0; i < NPL; ++i) SSG(i) = SSG(i) * ((SAC(i) / m1 < cfg.target[1]) ? dn : up);
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
      const Stage st = take();
      const double xn = x + sgs * st.za;
      const double s2n = (xn > -700.0 && xn < 700.0) ? FEXP(xn) : exp(xn);
      // logf(x) = InverseGamma(0.001, 0.001)(s2) + x [log-Jacobian, transformdistribution.jl:75-78] + sum_i Normal(b_i; 0, sqrt(s2))
      const double dx = xn - x;
      const double dinv = 1.0 / s2n - 1.0 / s2;
      const double delta = -(0.001 + 1.0) * dx - 0.001 * dinv + dx - 0.5 * S * dinv - (double)NPL * 0.5 * dx;
      if (mh_less(st.la, delta, it32, 2u, 0u, 0)) { x = xn; s2 = s2n; if (adapt) acs += 1; }
      if (adapt && ((long long)m2 % cfg.batchsize[2]) == 0) {
        const double dl = amwg_delta(m2, cfg.batchsize[2]);
        sgs *= exp((double)acs / m2 < cfg.target[2] ? -dl : dl);
      }
    }
    // ================================================================== thinning + streaming moments (mcmc.jl:76-78)
    if (valid && iter > a.burnin && (iter - a.burnin) % a.thin == 0) {
      mon[0] = al0; mon[1] = al1; mon[2] = al2; mon[3] = al3; mon[4] = s2;
      if (a.samples) {
        const long long row = (iter - a.burnin) / a.thin - 1 - a.row0;
        for (int j = 0; j < SeedsModel::P; ++j) a.samples[((size_t)row * SeedsModel::P + j) * C + c] = mon[j];
      }
      if (a.comom) comoments_update(a.mom, a.momn, a.comom, C, (size_t)c, SeedsModel::P, mon, a.log_mask);
      moments_update(a.mom, a.momn, C, (size_t)c, SeedsModel::P, mon);
    }
  }
  if (valid) {
    a.state[0 * C + c] = al0; a.state[1 * C + c] = al1; a.state[2 * C + c] = al2; a.state[3 * C + c] = al3;
    a.state[4 * C + c] = s2;
    for (int i = 0; i < NPL; ++i) a.state[(size_t)(5 + i) * C + c] = SB(i);
    TUNE(0, 0) = m0; TUNE(0, 1) = ad0 ? 1.0 : 0.0;
    TUNE(0, 2) = sg0; TUNE(0, 3) = sg1; TUNE(0, 4) = sg2; TUNE(0, 5) = sg3;
    TUNE(0, 6) = ac0; TUNE(0, 7) = ac1; TUNE(0, 8) = ac2; TUNE(0, 9) = ac3;
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
#undef FLOG
#undef FEXP
}

template <int CW, int PW>
int launch_ws(const FastCfg& cfg, const RunArgs& a, cudaStream_t st) {
  constexpr int BS = CW * 32;
  const size_t smem = ((size_t)BS * 4 * NSL + 2 * 2 * BS + 384) * sizeof(double) + (size_t)2 * 2 * BS * sizeof(float);
  static thread_local int attr_dev = -1;   // the attribute call is slow: once per device
  int dev = 0; cudaGetDevice(&dev);
  if (attr_dev != dev) {
    if (cudaFuncSetAttribute(seeds_ws_kernel<CW, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    attr_dev = dev;
  }
  const unsigned grid = (unsigned)((a.n_chains + BS - 1) / BS);
  seeds_ws_kernel<CW, PW><<<grid, (CW + PW) * 32, smem, st>>>(cfg, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace
}  // namespace mcu
