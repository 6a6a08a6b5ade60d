// glm_tc.cu — tensor-core likelihood/gradient kernel for the Bernoulli-logit GLM template (sm_100a only).
//
// For every chain c of the handle and the requested position beta_c (req [d][C]):
//     eta = X beta_c,   logf_c = sum_i [y_i eta_i - softplus(eta_i)],   grad_c = X' (y - invlogit(eta))
// i.e. the likelihood part of logpdf!/gradlogpdf! of the GLM block (the reference would run d+2 interpreted
// model evaluations per gradient: src/model/simulation.jl:47-51, src/samplers/sampler.jl:106-111).
//
// One CTA owns 128 chains (UMMA M = 128) and a slab of 128-row tiles of X, and per tile runs the
// FlashAttention-shaped chain   S = Theta X_t'  →  elementwise  →  G += R X_t   entirely on chip:
//   GEMM1  D1[128 chains x 128 rows]  = Theta[128 x d] . X_t'          tcgen05.mma, A and B from shared memory
//   epilogue (4 warps, thread = chain = TMEM lane): tcgen05.ld D1, p = invlogit(eta), logf += y eta - softplus,
//            R = y - p written back to TENSOR MEMORY as fp16 (tcgen05.st)
//   GEMM2  G[128 chains x d]        += R[128 x 128 rows] . X_t         tcgen05.mma, A from TMEM, B = the same
//            shared-memory tile read MN-major; the accumulator G stays in TMEM across the whole slab
// eta never leaves the SM; X is read from HBM once per chain group (bulk-copied tile by tile with
// cp.async.bulk + mbarrier, double buffered).
//
// Precision: operands are split fp16 pairs (x = hi + lo, |lo| <= 2^-11 |hi|) and every product is formed
// as hi*hi + hi*lo + lo*hi with FP32 accumulation (dropped term 2^-22 relative), so eta and the gradient
// carry ~1e-6 relative error — inside north_star's 1e-5 — at 1/3 of the fp16 tensor peak.
// The tensor core truncates when it adds into the FP32 accumulator, so a long running sum drifts linearly
// (measured 1e-4 relative over 27k rows); the gradient accumulator is therefore flushed to FP64 partials every
// FLUSH = 8 tiles (1,024 rows), which bounds the drift at a few 1e-6.
// Partials are written as FP64 and folded deterministically (glm_nuts.cu: glm_fold_kernel).
//
// Shared-memory operand layout: un-swizzled UMMA "interleave" core matrices (8 rows x 16 bytes, 128 bytes
// contiguous), core matrices ordered [row-block][col-block].  X is pre-packed in HBM in exactly this order,
// so a tile is one contiguous bulk copy, and the SAME bytes serve GEMM1 (K-major: SBO = row-block stride,
// LBO = 128) and GEMM2 (MN-major: SBO = 128, LBO = row-block stride).
#include <cuda_fp16.h>

#include "launch.hpp"

namespace mcu {

namespace {

constexpr int TM = 128;   // chains per CTA
constexpr int TR = 128;   // data rows per tile
constexpr int FLUSH = 8;  // tiles between flushes of the TMEM gradient accumulator (see below)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// UMMA shared-memory descriptor, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version for sm_100
  return d;                 // layout_type (bits 61-63) = 0: no swizzle
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float exp2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_half2(__half lo, __half hi) {
  return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
}

// ---- X pre-pack: fp64 [N x d] row-major → per tile [hi | lo | y] ----------------------------------------
// hi/lo blocks: core matrices (8 rows x 8 cols fp16 = 128 B) ordered [row-block][col-block].
__global__ void glm_pack_kernel(const double* __restrict__ X, const double* __restrict__ y, int N, int d, int DP,
                                unsigned char* __restrict__ blob, size_t tile_bytes) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over tiles * TR * (DP/8) 16-byte chunks
  const int CB = DP / 8;
  const long long chunks_per_tile = (long long)TR * CB;
  const long long NT = (N + TR - 1) / TR;
  if (idx >= NT * chunks_per_tile) return;
  const long long t = idx / chunks_per_tile;
  const int rem = (int)(idx % chunks_per_tile);
  const int row = rem / CB, cb = rem % CB;
  const long long grow = t * TR + row;
  __half hi[8], lo[8];
  for (int e = 0; e < 8; ++e) {
    const int col = cb * 8 + e;
    const float x = (grow < N && col < d) ? (float)X[(size_t)grow * d + col] : 0.0f;
    hi[e] = __float2half_rn(x);
    lo[e] = __float2half_rn(x - __half2float(hi[e]));
  }
  unsigned char* tile = blob + (size_t)t * tile_bytes;
  const size_t off = ((size_t)(row / 8) * CB + cb) * 128 + (size_t)(row % 8) * 16;
  uint4 vh, vl;
  vh.x = pack_half2(hi[0], hi[1]); vh.y = pack_half2(hi[2], hi[3]); vh.z = pack_half2(hi[4], hi[5]); vh.w = pack_half2(hi[6], hi[7]);
  vl.x = pack_half2(lo[0], lo[1]); vl.y = pack_half2(lo[2], lo[3]); vl.z = pack_half2(lo[4], lo[5]); vl.w = pack_half2(lo[6], lo[7]);
  *reinterpret_cast<uint4*>(tile + off) = vh;
  *reinterpret_cast<uint4*>(tile + (size_t)TR * DP * 2 + off) = vl;
  if (cb == 0) reinterpret_cast<float*>(tile + (size_t)2 * TR * DP * 2)[row] = grow < N ? (float)y[grow] : 0.5f;   // padding rows: X = 0, so eta = 0 and r x = 0; their softplus(0) is added back below
}

struct TcArgs {
  const unsigned char* blob; size_t tile_bytes;
  int NT, tiles_per_slab, d, DP;
  long long C;
  const double* req;         // [d][C]
  double* part_lp;           // [nslab][C]
  float* part_g;             // [nslab * nsub][d][C]  FP32 flushes of the TMEM accumulator (summed in FP64 by the fold)
  int nsub;
  int n_pad;                 // zero rows appended to the last tile
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int kEpiThreads = 512;   // 16 epilogue warps: warp w owns TMEM lanes 32 (w % 4) .. +31 and the 32-column quarter w / 4
constexpr int kTcThreads = kEpiThreads + 32;   // + 1 issuer warp (bulk copies, tcgen05.mma, commits)

// barrier slots in shared memory
enum { B_FULL0 = 0, B_FULL1, B_D1FULL0, B_D1FULL1, B_RFULL, B_GDONE, B_GREAD, B_COUNT };

// Warp-specialised pipeline.  Per tile t (buffer b = t & 1):
//   issuer  : GEMM2(t)   → G        as soon as epilogue(t) has stored R(t)
//             bulk copy X(t+2) → buffer b once GEMM2(t) has completed, then GEMM1(t+2) → D1[b]
//             (the tensor pipe runs GEMM2(t), GEMM1(t+2) while the 16 epilogue warps work on tile t+1)
//   epilogue: D1[b] → p, logf, R (split fp16) written back INTO D1[b] (R_hi columns 0-63, R_lo 64-127, as
//             FlashAttention keeps P where S was), so R is double buffered for free; every FLUSH tiles G → FP64 partial
//             (16 warps = 4 per scheduler; the 4 warps that share a TMEM lane quarter sync before overwriting D1)
// TMEM columns: D1[0] 0-127, D1[1] 128-255, G 256-383.
__global__ void __launch_bounds__(kTcThreads, 1) glm_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int DP = a.DP, CB = DP / 8;
  const uint32_t op_bytes = (uint32_t)TM * DP * 2;          // one 128 x DP fp16 operand block
  const uint32_t tile_bytes = (uint32_t)a.tile_bytes;       // hi | lo | y
  unsigned char* xbuf[2] = {smem, smem + tile_bytes};
  unsigned char* th_hi = smem + 2 * tile_bytes;
  unsigned char* th_lo = th_hi + op_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(th_lo + op_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
  double* lp_xchg = reinterpret_cast<double*>(bars + B_COUNT + 2);   // [3][128]
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  const int slab = blockIdx.y;
  const int t0 = slab * a.tiles_per_slab, t1 = min(a.NT, t0 + a.tiles_per_slab);
  const int T = max(0, t1 - t0);

  if (warp == kEpiThreads / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar(B_FULL0), 1); mbar_init(bar(B_FULL1), 1);
    mbar_init(bar(B_D1FULL0), 1); mbar_init(bar(B_D1FULL1), 1);
    mbar_init(bar(B_RFULL), kEpiThreads); mbar_init(bar(B_GDONE), 1); mbar_init(bar(B_GREAD), kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Theta tile: chain = row (tid % 128); the two column halves are split over tid / 128; split fp16, blocked layout
  if (tid < kEpiThreads) {
    const int row = tid & 127, grp = tid >> 7;
    const long long c = (long long)blockIdx.x * TM + row;
    for (int cb = grp; cb < CB; cb += kEpiThreads / 128) {
      __half hi[8], lo[8];
      for (int e = 0; e < 8; ++e) {
        const int col = cb * 8 + e;
        const float x = (c < a.C && col < a.d) ? (float)a.req[(size_t)col * a.C + c] : 0.0f;
        hi[e] = __float2half_rn(x);
        lo[e] = __float2half_rn(x - __half2float(hi[e]));
      }
      const size_t off = ((size_t)(row / 8) * CB + cb) * 128 + (size_t)(row % 8) * 16;
      uint4 vh, vl;
      vh.x = pack_half2(hi[0], hi[1]); vh.y = pack_half2(hi[2], hi[3]); vh.z = pack_half2(hi[4], hi[5]); vh.w = pack_half2(hi[6], hi[7]);
      vl.x = pack_half2(lo[0], lo[1]); vl.y = pack_half2(lo[2], lo[3]); vl.z = pack_half2(lo[4], lo[5]); vl.w = pack_half2(lo[6], lo[7]);
      *reinterpret_cast<uint4*>(th_hi + off) = vh;
      *reinterpret_cast<uint4*>(th_lo + off) = vl;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes → visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_d1[2] = {tmem, tmem + 128};
  const uint32_t tm_g = tmem + 256;

  // instruction descriptors (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): F16 x F16 → F32, M = 128
  const uint32_t idesc1 = (1u << 4) | ((uint32_t)(TR >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);                 // N = 128, A/B K-major
  const uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(DP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);     // N = DP, B MN-major
  const uint32_t rb_stride = (uint32_t)CB * 128;   // bytes between 8-row blocks

  if (warp == kEpiThreads / 32) {
    // ======================================================================= issuer (one elected lane)
    if ((tid & 31) == 0 && T > 0) {
      auto load_tile = [&](int t) {
        const int b = t & 1;
        mbar_expect_tx(bar(B_FULL0 + b), tile_bytes);
        bulk_copy_g2s(smem_u32(xbuf[b]), a.blob + (size_t)(t0 + t) * a.tile_bytes, tile_bytes, bar(B_FULL0 + b));
      };
      auto gemm1 = [&](int t) {   // D1[t&1] = Theta . X_t'  (3 split products per 16-wide K step)
        const int b = t & 1;
        const uint32_t xh = smem_u32(xbuf[b]), xl = xh + op_bytes, ah = smem_u32(th_hi), al = smem_u32(th_lo);
        for (int k = 0; k < DP / 16; ++k) {
          const uint32_t ko = (uint32_t)k * 256;   // two col-blocks per K step
          const uint64_t dah = make_desc(ah + ko, 128, rb_stride), dal = make_desc(al + ko, 128, rb_stride);
          const uint64_t dbh = make_desc(xh + ko, 128, rb_stride), dbl = make_desc(xl + ko, 128, rb_stride);
          mma_ss(tm_d1[b], dah, dbh, idesc1, k > 0 ? 1u : 0u);
          mma_ss(tm_d1[b], dah, dbl, idesc1, 1u);
          mma_ss(tm_d1[b], dal, dbh, idesc1, 1u);
        }
        tc_commit(bar(B_D1FULL0 + b));
      };
      load_tile(0);
      if (T > 1) load_tile(1);
      mbar_wait(bar(B_FULL0), 0);
      tc_fence_after();
      gemm1(0);
      if (T > 1) { mbar_wait(bar(B_FULL1), 0); tc_fence_after(); gemm1(1); }
      int nflush = 0;          // completions of B_GREAD consumed so far
      bool prev_flushed = false;
      for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        mbar_wait(bar(B_RFULL), (uint32_t)(t & 1));                 // epilogue(t) stored R(t) into D1[b]
        if (prev_flushed) { mbar_wait(bar(B_GREAD), (uint32_t)(nflush & 1)); ++nflush; }   // G of the previous interval read back
        tc_fence_after();
        {   // GEMM2(t): G += R . X_t  (A = R from tensor memory, B = the X tile read MN-major)
          const uint32_t xh = smem_u32(xbuf[b]), xl = xh + op_bytes;
          const uint32_t rh = tm_d1[b], rl = tm_d1[b] + 64;
          for (int kk = 0; kk < TR / 16; ++kk) {
            const uint32_t ko = (uint32_t)kk * 2 * rb_stride;       // two row-blocks per K step
            const uint64_t dbh = make_desc(xh + ko, rb_stride, 128), dbl = make_desc(xl + ko, rb_stride, 128);
            mma_ts(tm_g, rh + (uint32_t)kk * 8, dbh, idesc2, ((t % FLUSH) != 0 || kk > 0) ? 1u : 0u);
            mma_ts(tm_g, rh + (uint32_t)kk * 8, dbl, idesc2, 1u);
            mma_ts(tm_g, rl + (uint32_t)kk * 8, dbh, idesc2, 1u);
          }
          tc_commit(bar(B_GDONE));
        }
        prev_flushed = (t % FLUSH) == FLUSH - 1 || t == T - 1;
        if (t + 2 < T) {
          mbar_wait(bar(B_GDONE), (uint32_t)(t & 1));               // GEMM2(t) complete → X buffer b and D1[b] are free
          load_tile(t + 2);
          mbar_wait(bar(B_FULL0 + b), (uint32_t)(((t + 2) >> 1) & 1));
          tc_fence_after();
          gemm1(t + 2);
        }
      }
    }
  } else {
    // ======================================================================= epilogue warps
    const int q = warp & 3, cq = warp >> 2;                         // lane quarter, column quarter
    const int lane_row = q * 32 + (tid & 31);                       // TMEM lane = chain row of this CTA
    const long long c = (long long)blockIdx.x * TM + lane_row;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    double lp_acc = 0.0;
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      mbar_wait(bar(B_D1FULL0 + b), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      // y of this warp's 32 rows (shared memory, written by the bulk copy)
      const float4* y4 = reinterpret_cast<const float4*>(smem + (size_t)b * tile_bytes + 2 * (size_t)op_bytes) + cq * 8;
      float sp_sum = 0.0f;
      // this warp's 32 columns of eta → registers; then the four warps of this lane quarter agree that all of D1[b]
      // has been read before any of them overwrites it with R
      uint32_t v0[16], v1[16];
      tmem_ld16(tm_d1[b] + lane_off + (uint32_t)(cq * 32), v0);
      tmem_ld16(tm_d1[b] + lane_off + (uint32_t)(cq * 32 + 16), v1);
      tmem_wait_ld();
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int col0 = cq * 32 + ch * 16;
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 yq = y4[ch * 4 + (e >> 2)];
          const float yy[4] = {yq.x, yq.y, yq.z, yq.w};
          float r4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float eta = __uint_as_float(ch == 0 ? v0[e + u] : v1[e + u]);
            const float ex = exp2f_approx(-1.4426950408889634f * fabsf(eta));   // exp(-|eta|)
            const float w = 1.0f + ex;
            const float s = rcp_approx(w);                                      // invlogit(|eta|)
            const float p = eta >= 0.0f ? s : ex * s;
            sp_sum += fmaf(lg2_approx(w), 0.6931471805599453f, fmaxf(eta, 0.0f));   // softplus(eta)
            r4[u] = yy[u] - p;
          }
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const __half2 h = __floats2half2_rn(r4[u], r4[u + 1]);
            const float2 hb = __half22float2(h);
            const __half2 l = __floats2half2_rn(r4[u] - hb.x, r4[u + 1] - hb.y);
            ph[(e + u) >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            pl[(e + u) >> 1] = *reinterpret_cast<const uint32_t*>(&l);
          }
        }
        tmem_st8(tm_d1[b] + lane_off + (uint32_t)(col0 >> 1), ph);          // R_hi: columns 0-63 of D1[b]
        tmem_st8(tm_d1[b] + lane_off + (uint32_t)(64 + (col0 >> 1)), pl);   // R_lo: columns 64-127
      }
      // logf contribution of the tile: -sum softplus(eta); the sum_i y_i eta_i part is beta . (X'y), added by the fold.
      // Padding rows of the last tile have eta = 0: give their softplus(0) back.
      float lp_tile = -sp_sum;
      if (t0 + t == a.NT - 1 && a.n_pad > 0) {
        const int first_pad = TR - a.n_pad;
        const int lo_c = max(first_pad, cq * 32), hi_c = cq * 32 + 32;
        if (hi_c > lo_c) lp_tile += (float)(hi_c - lo_c) * (lg2_approx(2.0f) * 0.6931471805599453f);
      }
      lp_acc += (double)lp_tile;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar(B_RFULL));
      if ((t % FLUSH) == FLUSH - 1 || t == T - 1) {
        // ---- flush the gradient accumulator to this sub-slab's FP64 partial (16-column blocks dealt over cq)
        mbar_wait(bar(B_GDONE), (uint32_t)(t & 1));
        tc_fence_after();
        const int sub = t / FLUSH;
        for (int j0 = cq * 16; j0 < DP; j0 += 64) {
          uint32_t v[16];
          tmem_ld16(tm_g + lane_off + (uint32_t)j0, v); tmem_wait_ld();
          if (c < a.C)
            for (int e = 0; e < 16; ++e)
              if (j0 + e < a.d) a.part_g[(((size_t)slab * a.nsub + sub) * a.d + (j0 + e)) * a.C + c] = __uint_as_float(v[e]);
        }
        tc_fence_before();
        mbar_arrive(bar(B_GREAD));
      }
    }
    // ---- this slab's logf partial (the four column quarters of a chain are combined through shared memory)
    if (cq > 0) lp_xchg[(cq - 1) * 128 + lane_row] = lp_acc;
    asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory");
    if (cq == 0 && c < a.C) {
      a.part_lp[(size_t)slab * a.C + c] = ((lp_acc + lp_xchg[lane_row]) + lp_xchg[128 + lane_row]) + lp_xchg[256 + lane_row];
      const int used = T > 0 ? (T + FLUSH - 1) / FLUSH : 0;
      for (int sub = used; sub < a.nsub; ++sub)
        for (int j = 0; j < a.d; ++j) a.part_g[(((size_t)slab * a.nsub + sub) * a.d + j) * a.C + c] = 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiThreads / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

size_t glm_tc_tile_bytes(int d) { const int DP = (d + 15) / 16 * 16; return (size_t)2 * TR * DP * 2 + TR * sizeof(float); }
long long glm_tc_num_tiles(long long N) { return (N + TR - 1) / TR; }

void glm_tc_pack(const double* X, const double* y, int N, int d, unsigned char* blob, cudaStream_t st) {
  const int DP = (d + 15) / 16 * 16;
  const long long total = glm_tc_num_tiles(N) * TR * (DP / 8);
  glm_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, y, N, d, DP, blob, glm_tc_tile_bytes(d));
}

int glm_tc_nsub(long long N, int nslab) {
  const long long NT = glm_tc_num_tiles(N), tps = (NT + nslab - 1) / nslab;
  return (int)((tps + FLUSH - 1) / FLUSH);
}

// Returns 0 on success.  part_lp [nslab][C] (FP64, = -sum softplus), part_g [nslab * nsub][d][C] (FP32), to be folded over slabs.
int glm_tc_launch(const unsigned char* blob, int N, int d, long long C, const double* req, int nslab,
                  double* part_lp, float* part_g, cudaStream_t st) {
  TcArgs a;
  a.DP = (d + 15) / 16 * 16;
  a.blob = blob; a.tile_bytes = glm_tc_tile_bytes(d); a.NT = (int)glm_tc_num_tiles(N);
  a.tiles_per_slab = (a.NT + nslab - 1) / nslab; a.nsub = glm_tc_nsub(N, nslab); a.n_pad = (int)(glm_tc_num_tiles(N) * TR - N); a.d = d; a.C = C; a.req = req; a.part_lp = part_lp; a.part_g = part_g;
  const size_t smem = 2 * a.tile_bytes + 2 * (size_t)TM * a.DP * 2 + 8 * (B_COUNT + 2) + 3 * 128 * sizeof(double);
  static thread_local size_t smem_set[64] = {0};   // per device: the attribute call is slow, do it once per size
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || smem_set[dev] != smem) {
    if (cudaFuncSetAttribute(glm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (dev >= 0 && dev < 64) smem_set[dev] = smem;
  }
  dim3 grid((unsigned)((C + TM - 1) / TM), (unsigned)nslab);
  glm_tc_kernel<<<grid, kTcThreads, smem, st>>>(a);
  return 0;   // launch errors surface at the next synchronisation of the handle's stream
}

}  // namespace mcu
