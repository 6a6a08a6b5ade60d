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
// One lane of a converged warp (the warp keeps running converged: descriptors and addresses stay in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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
__global__ void glm_pack_kernel(const double* __restrict__ X, const double* __restrict__ y, int N, int d, int DP, int family,
                                unsigned char* __restrict__ blob, size_t tile_bytes, const double* __restrict__ colscale) {
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
    // colscale[col] is a power of two (1 for columns inside the fp16 range, see glm_tc_pack): the product is exact
    const float x = (grow < N && col < d) ? (float)(X[(size_t)grow * d + col] * (colscale ? colscale[col] : 1.0)) : 0.0f;
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
  if (cb == 0) reinterpret_cast<float*>(tile + (size_t)2 * TR * DP * 2)[row] = grow < N ? (float)(family == 0 ? y[grow] - 0.5 : y[grow]) : 0.0f;   // Bernoulli: y - 1/2, else y; padding rows: X = 0, so eta = 0; their residual meets X = 0 and their logf term is taken back in the kernel
}

struct TcArgs {
  const unsigned char* blob; size_t tile_bytes;
  int NT, tiles_per_slab, d, DP;
  long long C;               // chains of this pass (after compaction: the chains still running, mapped through `map`)
  long long Cfull;           // chains of the handle = stride of req
  const int* map;            // [C] chain index of pass slot k, or nullptr (identity)
  const double* req;         // [d][Cfull]
  double* part_lp;           // [nslab][C]   -ln2 * sum_i [ |s_i| / 2 + log2(1 + 2^-|s_i|) ],  s = eta * log2(e)
  float* part_g;             // [nslab][d][C]  FP32 running sum of the TMEM accumulator flushes (folded over slabs in FP64)
  int n_pad;                 // zero rows appended to the last tile
  double theta_scale;        // log2(e) for the logit / log links (eta arrives as an exponent of 2), 1 for the identity link
  const double* col_inv;     // [d] 1 / colscale (powers of two) or nullptr: X was packed as X diag(colscale), so Theta = beta .* col_inv
  float r_scale;             // Normal: 1 / sigma^2
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int kEpiThreads = 512;   // 16 epilogue warps: warp w owns TMEM lanes 32 (w % 4) .. +31 and the 32-column quarter w / 4
constexpr int kTcThreads = kEpiThreads + 32;   // + 1 issuer warp (bulk copies, tcgen05.mma, commits)
constexpr int kStages = 3;         // X tiles in flight in shared memory

// barrier slots in shared memory
enum { B_FULL0 = 0, B_D1FULL0 = B_FULL0 + kStages, B_RFULL0 = B_D1FULL0 + 2, B_GDONE0 = B_RFULL0 + 2, B_GREAD = B_GDONE0 + 2, B_COUNT };

// Warp-specialised pipeline.  Tile t uses X stage t % 3 and D1 / barrier slot b = t & 1.
//   issuer  : tensor-pipe order  ... G2(t-1), G1(t+1), G2(t), G1(t+2) ...  — GEMM1(t+2) is queued right behind GEMM2(t)
//             (tcgen05.mma of one CTA execute in issue order), so the pipe has GEMM1(t+2) to run while the epilogue warps
//             are busy with tile t+1; X(t+3) is bulk-copied into the stage GEMM2(t) has finished reading.
//   epilogue: D1[b] → sigmoid, softplus sums, R = y - p (split fp16) written back INTO THE SAME 32 COLUMNS the warp read
//             (R_hi in the first 16, R_lo in the last 16: as FlashAttention keeps P where S was), so no warp ever writes
//             columns another warp reads and the epilogue needs no CTA-level synchronisation.
//             Every FLUSH tiles the gradient accumulator G is added into the slab's FP32 partial (L2 resident), one tile
//             late, so that the read never waits for GEMM2.
// Theta (the 128 requested positions, pre-scaled by log2 e, split fp16) lives in TENSOR MEMORY for the whole kernel: both GEMMs
// take A from TMEM, shared memory holds nothing but X tiles.
// TMEM columns: D1[0] 0-127, D1[1] 128-255, G 256-(256+DP), Theta_hi 384-(384+DP/2), Theta_lo 448-(448+DP/2).
// FAM: 0 Bernoulli / logit, 1 Poisson / log, 2 Normal / identity (known sd) — the GLM family of north_star; only the epilogue differs.
template <int FAM>
__global__ void __launch_bounds__(kTcThreads, 1) glm_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int DP = a.DP, CB = DP / 8;
  const uint32_t op_bytes = (uint32_t)TR * DP * 2;          // one 128 x DP fp16 operand block
  const uint32_t tile_bytes = (uint32_t)a.tile_bytes;       // hi | lo | y - 1/2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * tile_bytes);
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
    for (int i = 0; i < kStages; ++i) mbar_init(bar(B_FULL0 + i), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(B_D1FULL0 + i), 1); mbar_init(bar(B_RFULL0 + i), kEpiThreads / 32); mbar_init(bar(B_GDONE0 + i), 1); }
    mbar_init(bar(B_GREAD), kEpiThreads / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_d1[2] = {tmem, tmem + 128};
  const uint32_t tm_g = tmem + 256, tm_th = tmem + 384, tm_tl = tmem + 448;

  // Theta → tensor memory: thread = (chain lane, 16-value chunk); column j of the A operand holds values 2j, 2j+1
  if (tid < kEpiThreads) {
    const int q = warp & 3, cq = warp >> 2;
    const long long c = (long long)blockIdx.x * TM + q * 32 + (tid & 31);
    const long long csrc = (c < a.C && a.map) ? (long long)a.map[c] : c;   // compacted pass: slot c holds chain map[c]
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int ck = cq; ck < DP / 16; ck += 4) {
      uint32_t vh[8], vl[8];
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        float x[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int col = ck * 16 + e + u;
          x[u] = (c < a.C && col < a.d) ? (float)(a.req[(size_t)col * a.Cfull + csrc] * (a.col_inv ? a.col_inv[col] * a.theta_scale : a.theta_scale)) : 0.0f;
        }
        const __half2 h = __floats2half2_rn(x[0], x[1]);
        const float2 hb = __half22float2(h);
        const __half2 l = __floats2half2_rn(x[0] - hb.x, x[1] - hb.y);
        vh[e >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        vl[e >> 1] = *reinterpret_cast<const uint32_t*>(&l);
      }
      tmem_st8(tm_th + lane_off + (uint32_t)(ck * 8), vh);
      tmem_st8(tm_tl + lane_off + (uint32_t)(ck * 8), vl);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // instruction descriptors (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): F16 x F16 → F32, M = 128
  const uint32_t idesc1 = (1u << 4) | ((uint32_t)(TR >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);                 // N = 128, A/B K-major
  const uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(DP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);     // N = DP, B MN-major
  const uint32_t rb_stride = (uint32_t)CB * 128;   // bytes between 8-row blocks

  if (warp == kEpiThreads / 32) {
    // ======================================================================= issuer warp (converged; one elected lane issues)
    if (T > 0) {
      const uint32_t smem_base = smem_u32(smem);
      auto load_tile = [&](int t) {
        const int s = t % kStages;
        if (elect_one()) {
          mbar_expect_tx(bar(B_FULL0 + s), tile_bytes);
          bulk_copy_g2s(smem_base + (uint32_t)s * tile_bytes, a.blob + (size_t)(t0 + t) * a.tile_bytes, tile_bytes, bar(B_FULL0 + s));
        }
        __syncwarp();
      };
      auto gemm1 = [&](int t) {   // D1[t&1] = Theta . X_t'  (3 split products per 16-wide K step; A from tensor memory)
        const int b = t & 1, s = t % kStages;
        mbar_wait(bar(B_FULL0 + s), (uint32_t)((t / kStages) & 1));
        tc_fence_after();
        const uint32_t xh = smem_base + (uint32_t)s * tile_bytes;
        const uint64_t dh0 = make_desc(xh, 128, rb_stride), dl0 = make_desc(xh + op_bytes, 128, rb_stride);
        const uint32_t d1 = tm_d1[b];
        if (elect_one()) {
#pragma unroll 1
          for (int k = 0; k < DP / 16; ++k) {
            const uint64_t dbh = dh0 + (uint64_t)(k * 16), dbl = dl0 + (uint64_t)(k * 16);   // two col-blocks (256 B) per K step
            mma_ts(d1, tm_th + (uint32_t)k * 8, dbh, idesc1, k > 0 ? 1u : 0u);
            mma_ts(d1, tm_th + (uint32_t)k * 8, dbl, idesc1, 1u);
            mma_ts(d1, tm_tl + (uint32_t)k * 8, dbh, idesc1, 1u);
          }
          tc_commit(bar(B_D1FULL0 + b));
        }
        __syncwarp();
      };
      for (int t = 0; t < kStages && t < T; ++t) load_tile(t);
      gemm1(0);
      if (T > 1) gemm1(1);
      int nflush = 0;          // completions of B_GREAD consumed so far
      for (int t = 0; t < T; ++t) {
        const int b = t & 1, s = t % kStages;
        mbar_wait(bar(B_RFULL0 + b), (uint32_t)((t >> 1) & 1));       // epilogue(t) stored R(t) into D1[b]
        if (t > 0 && (t % FLUSH) == 0) { mbar_wait(bar(B_GREAD), (uint32_t)(nflush & 1)); ++nflush; }   // G of the previous interval read back
        tc_fence_after();
        {   // GEMM2(t): G += R . X_t  (A = R from tensor memory, B = the X tile read MN-major)
          const uint32_t xh = smem_base + (uint32_t)s * tile_bytes;
          const uint64_t dh0 = make_desc(xh, rb_stride, 128), dl0 = make_desc(xh + op_bytes, rb_stride, 128);
          const uint32_t step = (2 * rb_stride) >> 4;                  // two row-blocks per K step
          const uint32_t d1 = tm_d1[b];
          const uint32_t acc0 = (t % FLUSH) != 0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll 1
            for (int kk = 0; kk < TR / 16; ++kk) {
              const uint64_t dbh = dh0 + (uint64_t)(kk * step), dbl = dl0 + (uint64_t)(kk * step);
              const uint32_t rh = d1 + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), rl = rh + 16;
              mma_ts(tm_g, rh, dbh, idesc2, kk > 0 ? 1u : acc0);
              mma_ts(tm_g, rh, dbl, idesc2, 1u);
              mma_ts(tm_g, rl, dbh, idesc2, 1u);
            }
            tc_commit(bar(B_GDONE0 + b));
          }
          __syncwarp();
        }
        if (t + 2 < T) {
#if MCU_GLM_TC_CONSERVATIVE
          mbar_wait(bar(B_GDONE0 + b), (uint32_t)((t >> 1) & 1));     // GEMM2(t) complete before D1[b] is overwritten
#endif
          gemm1(t + 2);
        }
        if (t + kStages < T) {
          mbar_wait(bar(B_GDONE0 + b), (uint32_t)((t >> 1) & 1));     // GEMM2(t) complete → X stage s is free
          load_tile(t + kStages);
        }
      }
    }
  } else {
    // ======================================================================= epilogue warps
    const int q = warp & 3, cq = warp >> 2;                         // lane quarter, column quarter
    const int lane = tid & 31;
    const int lane_row = q * 32 + lane;                             // TMEM lane = chain row of this CTA
    const long long c = (long long)blockIdx.x * TM + lane_row;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    double lp_acc = 0.0;
    int pending = -1;                                               // tile whose flush interval has ended and is still to be read
    auto flush = [&](int te) {
      // add the gradient accumulator of the interval ending at tile te into this slab's FP32 partial (16-column blocks dealt over cq)
      mbar_wait(bar(B_GDONE0 + (te & 1)), (uint32_t)((te >> 1) & 1));
      tc_fence_after();
      const bool first = te < FLUSH;
      for (int j0 = cq * 16; j0 < DP; j0 += 64) {
        uint32_t v[16];
        tmem_ld16(tm_g + lane_off + (uint32_t)j0, v); tmem_wait_ld();
        if (c < a.C) {
          float* dst = a.part_g + ((size_t)slab * a.d + j0) * a.C + c;
          float old[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) old[e] = (!first && j0 + e < a.d) ? dst[(size_t)e * a.C] : 0.0f;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (j0 + e < a.d) dst[(size_t)e * a.C] = old[e] + __uint_as_float(v[e]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_GREAD));
    };
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      mbar_wait(bar(B_D1FULL0 + b), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      // y - 1/2 of this warp's 32 rows (shared memory, written by the bulk copy)
      const float4* y4 = reinterpret_cast<const float4*>(smem + (size_t)(t % kStages) * tile_bytes + 2 * (size_t)op_bytes) + cq * 8;
      const uint32_t tcol = tm_d1[b] + lane_off + (uint32_t)(cq * 32);
      uint32_t v[32];
      tmem_ld32(tcol, v);
      tmem_wait_ld();
      float tile_sum;
      if (FAM == 0) {
        float abs_sum = 0.0f, lg_sum = 0.0f;
  #pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t ph[8], pl[8];
          float prod = 1.0f;
  #pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 yq = y4[ch * 4 + (e >> 2)];
            const float ym[4] = {yq.x, yq.y, yq.z, yq.w};
            float r4[4];
  #pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const uint32_t b0 = v[ch * 16 + e + u], b1 = v[ch * 16 + e + u + 1];
              const float s0 = __uint_as_float(b0), s1 = __uint_as_float(b1);   // eta * log2(e)
              const float w0 = 1.0f + exp2f_approx(-fabsf(s0)), w1 = 1.0f + exp2f_approx(-fabsf(s1));
              const float ww = w0 * w1;
              const float rr = rcp_approx(ww);                  // one reciprocal for the pair
              const float h0 = fmaf(rr, w1, -0.5f), h1 = fmaf(rr, w0, -0.5f);   // invlogit(|eta|) - 1/2  in [0, 1/2)
              prod *= ww;
              abs_sum += fabsf(s0); abs_sum += fabsf(s1);
              // r = y - p,  p = 1/2 + copysign(h, eta)
              r4[u] = ym[u] - __uint_as_float(__float_as_uint(h0) | (b0 & 0x80000000u));
              r4[u + 1] = ym[u + 1] - __uint_as_float(__float_as_uint(h1) | (b1 & 0x80000000u));
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
          lg_sum += lg2_approx(prod);                           // sum_16 log2(1 + 2^-|s|) = log2 of the product (<= 2^16)
          tmem_st8(tcol + (uint32_t)(ch * 8), ph);              // R_hi: first 16 of the warp's 32 columns
          tmem_st8(tcol + (uint32_t)(16 + ch * 8), pl);         // R_lo: last 16
        }
        // softplus(eta) = ln2 * (s/2 + |s|/2 + log2(1 + 2^-|s|)); the s/2 part is linear in beta and, like sum y eta, is added by the
        // fold as beta . X'(y - 1/2).  Padding rows of the last tile have s = 0: take their log2(2) = 1 back.
        tile_sum = fmaf(0.5f, abs_sum, lg_sum);
        if (t0 + t == a.NT - 1 && a.n_pad > 0) {
          const int first_pad = TR - a.n_pad;
          const int lo_c = max(first_pad, cq * 32), hi_c = cq * 32 + 32;
          if (hi_c > lo_c) tile_sum -= (float)(hi_c - lo_c);
        }
      } else {
        // Poisson / log:  logf_i = y eta - e^eta (- lgamma(y + 1)),  r = y - e^eta;   Normal / identity:  logf_i = -(y - eta)^2 / (2 sigma^2),
        // r = (y - eta) / sigma^2.  The y eta term of the Poisson family goes through beta . X'y in the fold, like the Bernoulli one.
        float acc = 0.0f;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t ph[8], pl[8];
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 yq = y4[ch * 4 + (e >> 2)];
            const float ym[4] = {yq.x, yq.y, yq.z, yq.w};
            float r4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float s = __uint_as_float(v[ch * 16 + e + u]);
              if (FAM == 1) { const float lam = exp2f_approx(s); acc += lam; r4[u] = ym[u] - lam; }
              else { const float d = ym[u] - s; acc = fmaf(d, d, acc); r4[u] = d * a.r_scale; }
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
          tmem_st8(tcol + (uint32_t)(ch * 8), ph);
          tmem_st8(tcol + (uint32_t)(16 + ch * 8), pl);
        }
        tile_sum = acc;
        if (FAM == 1 && t0 + t == a.NT - 1 && a.n_pad > 0) {   // padding rows: eta = 0, e^0 = 1 each
          const int first_pad = TR - a.n_pad;
          const int lo_c = max(first_pad, cq * 32), hi_c = cq * 32 + 32;
          if (hi_c > lo_c) tile_sum -= (float)(hi_c - lo_c);
        }
      }
      lp_acc += (double)tile_sum;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_RFULL0 + b));
      if (pending >= 0) { flush(pending); pending = -1; }
      if ((t % FLUSH) == FLUSH - 1 || t == T - 1) pending = t;
    }
    if (pending >= 0) flush(pending);
    // ---- this slab's logf partial (the four column quarters of a chain are combined through shared memory)
    if (cq > 0) lp_xchg[(cq - 1) * 128 + lane_row] = lp_acc;
    asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory");
    if (cq == 0 && c < a.C) {
      const double lp_scale = FAM == 0 ? -0.6931471805599453 : (FAM == 1 ? -1.0 : -0.5 * (double)a.r_scale);
      a.part_lp[(size_t)slab * a.C + c] = lp_scale * (((lp_acc + lp_xchg[lane_row]) + lp_xchg[128 + lane_row]) + lp_xchg[256 + lane_row]);
      if (T == 0)
        for (int j = 0; j < a.d; ++j) a.part_g[((size_t)slab * a.d + j) * a.C + c] = 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiThreads / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

size_t glm_tc_tile_bytes(int d) { const int DP = (d + 15) / 16 * 16; return (size_t)2 * TR * DP * 2 + TR * sizeof(float); }
long long glm_tc_num_tiles(long long N) { return (N + TR - 1) / TR; }

// colscale (device, [d]) or nullptr: power-of-two column factors that bring a column whose largest magnitude lies outside [2^-6, 2^14] into
// [1, 2) before the fp16 hi / lo split (fp16 overflows above 65,504 and loses the lo term below ~6e-5); the pass undoes them exactly
// (Theta = beta / colscale on the way in, gradient / colscale in the fold).
void glm_tc_pack(const double* X, const double* y, int N, int d, int family, unsigned char* blob, cudaStream_t st, const double* colscale) {
  const int DP = (d + 15) / 16 * 16;
  const long long total = glm_tc_num_tiles(N) * TR * (DP / 8);
  glm_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, y, N, d, DP, family, blob, glm_tc_tile_bytes(d), colscale);
}

int glm_tc_nsub(long long, int) { return 1; }   // one FP32 gradient partial per slab (the flush intervals are summed in place)

// Returns 0 on success.  part_lp [nslab][C] (FP64), part_g [nslab][d][C] (FP32), to be folded over slabs (glm_fold_tc).
int glm_tc_launch(const unsigned char* blob, int N, int d, long long C, const double* req, int nslab,
                  double* part_lp, float* part_g, int family, double sigma, cudaStream_t st, const int* map, long long Cfull, const double* col_inv) {
  TcArgs a;
  a.col_inv = col_inv;
  a.DP = (d + 15) / 16 * 16;
  if (a.DP > 128) return -2;   // TMEM budget: G (DP columns) + Theta hi/lo (DP/2 each) next to the two D1 buffers
  a.blob = blob; a.tile_bytes = glm_tc_tile_bytes(d); a.NT = (int)glm_tc_num_tiles(N);
  a.tiles_per_slab = (a.NT + nslab - 1) / nslab; a.n_pad = (int)(glm_tc_num_tiles(N) * TR - N); a.d = d; a.C = C; a.Cfull = map ? Cfull : C; a.map = map; a.req = req; a.part_lp = part_lp; a.part_g = part_g;
  a.theta_scale = family == 2 ? 1.0 : 1.4426950408889634; a.r_scale = (float)(1.0 / (sigma * sigma));
  const size_t smem = kStages * a.tile_bytes + 8 * (B_COUNT + 2) + 3 * 128 * sizeof(double);
  static thread_local size_t smem_set[3][64] = {{0}};   // per family and device: the attribute call is slow, do it once per size
  int dev = 0; cudaGetDevice(&dev);
  if (family < 0 || family > 2) return -3;
  const void* fn = family == 0 ? (const void*)glm_tc_kernel<0> : family == 1 ? (const void*)glm_tc_kernel<1> : (const void*)glm_tc_kernel<2>;
  if (dev < 0 || dev >= 64 || smem_set[family][dev] != smem) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    if (dev >= 0 && dev < 64) smem_set[family][dev] = smem;
  }
  dim3 grid((unsigned)((C + TM - 1) / TM), (unsigned)nslab);
  if (family == 0) glm_tc_kernel<0><<<grid, kTcThreads, smem, st>>>(a);
  else if (family == 1) glm_tc_kernel<1><<<grid, kTcThreads, smem, st>>>(a);
  else glm_tc_kernel<2><<<grid, kTcThreads, smem, st>>>(a);
  return 0;   // launch errors surface at the next synchronisation of the handle's stream
}

}  // namespace mcu
