// kern_misc.cu — initialisation, layout transposes and the cross-chain reduction kernels.
#include "launch.hpp"
#include "diagproto.hpp"

namespace mcu {

// setinits! for all chains: state[e][c] = inits[(g % n_inits)][e] (+ jitter on the link scale).
__global__ void init_kernel(long long n_chains, long long chain_offset, unsigned long long seed, int D,
                            const double* inits, long long n_inits, const int* elink, const double* ebound, double jitter_sd, double* state) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  const long long gidx = chain_offset + c;
  const double* rec = inits + (size_t)(gidx % n_inits) * D;
  Draws rng;
  rng.k0 = (uint32_t)seed; rng.k1 = (uint32_t)(seed >> 32); rng.chain = (uint32_t)gidx; rng.ext = nullptr; rng.ext_pos = nullptr; rng.ext_n = 0;
  rng.seek(0, 0, 1);
  for (int e = 0; e < D; ++e) {
    double v = rec[e];
    if (jitter_sd > 0.0) {
      const double z = jitter_sd * rng.normal();
      if (elink[e] == LINK_LOG) v = exp(log(v) + z);
      else if (elink[e] == LINK_BOUNDED) {   // jitter on the two-sided link scale: transformdistribution.jl:9-10, 24-25
        const double lo = ebound[2 * e], hi = ebound[2 * e + 1], p = (v - lo) / (hi - lo);
        v = (hi - lo) * (1.0 / (exp(-(log(p / (1.0 - p)) + z)) + 1.0)) + lo;
      } else v = v + z;
    }
    state[(size_t)e * n_chains + c] = v;
  }
}

// [rows][C] (chain fastest) → out[c][rows] record-contiguous, and the reverse.
__global__ void soa_to_records(const double* soa, double* rec, long long C, int rows) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int e = 0; e < rows; ++e) rec[(size_t)c * rows + e] = soa[(size_t)e * C + c];
}
__global__ void records_to_soa(const double* rec, double* soa, long long C, int rows) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int e = 0; e < rows; ++e) soa[(size_t)e * C + c] = rec[(size_t)c * rows + e];
}
// samples [kept][P][C] → Julia column-major [kept × P × C]
__global__ void samples_to_julia(const double* smp, double* out, long long kept, int P, long long C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over kept*P*C, chain fastest on the read side
  const long long total = kept * P * C;
  if (idx >= total) return;
  const long long c = idx % C; const long long r = idx / C; const long long j = r % P; const long long i = r / P;
  out[i + kept * (j + (long long)P * c)] = smp[idx];
}

// ---- cross-chain reductions ---------------------------------------------------------------------
// Gelman moments (src/output/gelmandiag.jl:12-29): per column j and chain c let psibar = chain mean and
// s2 = chain variance (on the raw or link scale); with centres (c1, c2):
//   d = psibar - c1, e = s2 - c2;  sums = { 1, d, d^2, e, e^2, e d, e d^2 } summed over chains.
// Deterministic two-stage reduction: per-block partials, then one block folds them.
__global__ void gelman_partial_kernel(const double* mom, const double* momn, long long C, int P, const int* use_log,
                                      const double* center, double* partial /*[gridDim.x][P][7]*/) {
  __shared__ double sh[128];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = 0; j < P; ++j) {
    double vals[7] = {0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
      const double n = momn[c];
      const double* q = mom + (size_t)j * kMomPerCol * C + c;
      const int code = use_log[j];   // 0 identity, 1 log, 2 logit
      const double psibar = code == 2 ? q[9 * C] : code == 1 ? q[2 * C] : q[0 * C];
      const double s2 = (code == 2 ? q[10 * C] : code == 1 ? q[3 * C] : q[1 * C]) / (n - 1.0);
      const double d = psibar - (center ? center[j * 2 + 0] : 0.0);
      const double e = s2 - (center ? center[j * 2 + 1] : 0.0);
      vals[0] = 1.0; vals[1] = d; vals[2] = d * d; vals[3] = e; vals[4] = e * e; vals[5] = e * d; vals[6] = e * d * d;
    }
    for (int q7 = 0; q7 < 7; ++q7) {
      sh[threadIdx.x] = vals[q7];
      __syncthreads();
      for (int off = 64; off > 0; off >>= 1) { if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off]; __syncthreads(); }
      if (threadIdx.x == 0) partial[((size_t)blockIdx.x * P + j) * 7 + q7] = sh[0];
      __syncthreads();
    }
  }
}
// generic fold of per-block partial vectors: out[i] = sum_b partial[b][i]
__global__ void fold_kernel(const double* partial, long long nblocks, int width, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  double s = 0.0;
  for (long long b = 0; b < nblocks; ++b) s += partial[(size_t)b * width + i];
  out[i] = s;
}
// min/max of each monitored column over all chains (for the link(c) heuristic, chains.jl:237-246)
__global__ void minmax_partial_kernel(const double* mom, long long C, int P, double* partial /*[grid][P][2]*/) {
  __shared__ double shmin[128], shmax[128];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = 0; j < P; ++j) {
    const double* q = mom + (size_t)j * kMomPerCol * C + c;
    shmin[threadIdx.x] = c < C ? q[4 * C] : CUDART_INF;
    shmax[threadIdx.x] = c < C ? q[5 * C] : -CUDART_INF;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) { shmin[threadIdx.x] = fmin(shmin[threadIdx.x], shmin[threadIdx.x + off]); shmax[threadIdx.x] = fmax(shmax[threadIdx.x], shmax[threadIdx.x + off]); }
      __syncthreads();
    }
    if (threadIdx.x == 0) { partial[((size_t)blockIdx.x * P + j) * 2 + 0] = shmin[0]; partial[((size_t)blockIdx.x * P + j) * 2 + 1] = shmax[0]; }
    __syncthreads();
  }
}
// streaming summary sums per column: { C, sum mean_c, sum M2_c, sum (mean_c - ctr)^2, nb_total, sum bmean_c*nb, sum bM2_c, sum nb (bmean_c - bctr)^2 }
__global__ void summary_partial_kernel(const double* mom, const double* momn, long long C, int P, const double* center /*[P][2] or null*/,
                                       double* partial /*[grid][P][8]*/) {
  __shared__ double sh[128];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = 0; j < P; ++j) {
    double vals[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
      const double* q = mom + (size_t)j * kMomPerCol * C + c;
      const double nb = momn[2 * C + c];
      const double c1 = center ? center[j * 2 + 0] : 0.0, c2 = center ? center[j * 2 + 1] : 0.0;
      vals[0] = 1.0; vals[1] = q[0 * C]; vals[2] = q[1 * C]; vals[3] = (q[0 * C] - c1) * (q[0 * C] - c1);
      vals[4] = nb; vals[5] = nb * q[7 * C]; vals[6] = q[8 * C]; vals[7] = nb * (q[7 * C] - c2) * (q[7 * C] - c2);
    }
    for (int q8 = 0; q8 < 8; ++q8) {
      sh[threadIdx.x] = vals[q8];
      __syncthreads();
      for (int off = 64; off > 0; off >>= 1) { if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off]; __syncthreads(); }
      if (threadIdx.x == 0) partial[((size_t)blockIdx.x * P + j) * 8 + q8] = sh[0];
      __syncthreads();
    }
  }
}

// ---- packed two-round diagnostics protocol (diagproto.hpp) ---------------------------------------------
// block-level reduction of one value per thread (128 threads): op 0 = sum, 1 = min, 2 = max
__device__ __forceinline__ double block_reduce128(double v, int op, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      const double a = sh[threadIdx.x], b = sh[threadIdx.x + off];
      sh[threadIdx.x] = op == 0 ? a + b : op == 1 ? fmin(a, b) : fmax(a, b);
    }
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}
__global__ void diag1_partial_kernel(const double* mom, const double* momn, long long C, int P, unsigned long long logit_mask,
                                     double* partial /*[grid][P][11]*/) {
  __shared__ double sh[128];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = c < C;
  const double n = on ? momn[c] : 2.0, nb = on ? momn[2 * C + c] : 0.0;
  for (int j = 0; j < P; ++j) {
    const double* q = mom + (size_t)j * kMomPerCol * C + c;
    double v[kDiag1];
    v[0] = on ? q[4 * C] : CUDART_INF; v[1] = on ? q[5 * C] : -CUDART_INF;
    const bool lg = j < 64 && ((logit_mask >> j) & 1ull);
    v[2] = on ? 1.0 : 0.0;
    v[3] = on ? q[0 * C] : 0.0; v[4] = on ? q[1 * C] / (n - 1.0) : 0.0;
    v[5] = on ? q[2 * C] : 0.0; v[6] = on ? q[3 * C] / (n - 1.0) : 0.0;
    v[7] = (on && lg) ? q[9 * C] : 0.0; v[8] = (on && lg) ? q[10 * C] / (n - 1.0) : 0.0;
    v[9] = nb; v[10] = on ? nb * q[7 * C] : 0.0;
    for (int k = 0; k < kDiag1; ++k) {
      const double r = block_reduce128(v[k], k == 0 ? 1 : k == 1 ? 2 : 0, sh);
      if (threadIdx.x == 0) partial[((size_t)blockIdx.x * P + j) * kDiag1 + k] = r;
    }
  }
}
// fold of the round-1 partials into the protocol layout [min P | max P | sum 9P]
__global__ void diag1_fold_kernel(const double* partial, long long nblocks, int P, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * kDiag1) return;
  const int j = i / kDiag1, k = i % kDiag1;
  double r = k == 0 ? CUDART_INF : k == 1 ? -CUDART_INF : 0.0;
  for (long long b = 0; b < nblocks; ++b) {
    const double x = partial[(size_t)b * P * kDiag1 + i];
    r = k == 0 ? fmin(r, x) : k == 1 ? fmax(r, x) : r + x;
  }
  out[k == 0 ? j : k == 1 ? P + j : 2 * P + j * kDiagSum1 + (k - 2)] = r;
}
// plan[j] = { code, c1, c2, k1, k2 } from the all-reduced round-1 buffer
__global__ void diag_plan_kernel(const double* r1, const int* monlink, int transform, int P, double* plan) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P) return;
  double ctr[4];
  const int code = diag_plan_column(P, j, monlink[j], transform, r1, ctr);
  plan[j * 5 + 0] = (double)code; plan[j * 5 + 1] = ctr[0]; plan[j * 5 + 2] = ctr[1]; plan[j * 5 + 3] = ctr[2]; plan[j * 5 + 4] = ctr[3];
}
__global__ void diag2_partial_kernel(const double* mom, const double* momn, long long C, int P, const double* plan,
                                     double* partial /*[grid][P][15]*/) {
  __shared__ double sh[128];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = c < C;
  const double n = on ? momn[c] : 2.0, nb = on ? momn[2 * C + c] : 0.0;
  for (int j = 0; j < P; ++j) {
    const double* q = mom + (size_t)j * kMomPerCol * C + c;
    const int code = (int)plan[j * 5];
    const double c1 = plan[j * 5 + 1], c2 = plan[j * 5 + 2], k1 = plan[j * 5 + 3], k2 = plan[j * 5 + 4];
    double v[kDiag2];
    for (int k = 0; k < kDiag2; ++k) v[k] = 0.0;
    if (on) {
      const double psibar = code == 2 ? q[9 * C] : code == 1 ? q[2 * C] : q[0 * C];
      const double s2 = (code == 2 ? q[10 * C] : code == 1 ? q[3 * C] : q[1 * C]) / (n - 1.0);
      const double d = psibar - c1, e = s2 - c2;
      v[0] = 1.0; v[1] = d; v[2] = d * d; v[3] = e; v[4] = e * e; v[5] = e * d; v[6] = e * d * d;
      const double mean = q[0 * C], bmean = q[7 * C];
      v[7] = 1.0; v[8] = mean; v[9] = q[1 * C]; v[10] = (mean - k1) * (mean - k1);
      v[11] = nb; v[12] = nb * bmean; v[13] = q[8 * C]; v[14] = nb * (bmean - k2) * (bmean - k2);
    }
    for (int k = 0; k < kDiag2; ++k) {
      const double r = block_reduce128(v[k], 0, sh);
      if (threadIdx.x == 0) partial[((size_t)blockIdx.x * P + j) * kDiag2 + k] = r;
    }
  }
}
// round 2, pair part: partial[grid][npair][2] = { Σ cov_k(i, j), Σ d_i d_j } over the block's chains (co-moment set `set`: 0 raw, 1 node-link scale)
__global__ void diag_pairs_partial_kernel(const double* mom, const double* momn, const double* comom, long long C, int P, const double* plan, int set,
                                          double* partial) {
  __shared__ double sh[128];
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = c < C;
  const double n = on ? momn[c] : 2.0;
  const int npair = P * (P - 1) / 2;
  double d[kCoMaxP];
  for (int j = 0; j < P; ++j) {
    const double* q = mom + (size_t)j * kMomPerCol * C + c;
    const int code = (int)plan[j * 5];
    d[j] = on ? (code == 2 ? q[9 * C] : code == 1 ? q[2 * C] : q[0 * C]) - plan[j * 5 + 1] : 0.0;
  }
  int k = 0;
  for (int i = 0; i < P; ++i)
    for (int j = i + 1; j < P; ++j, ++k) {
      const double cov = on ? comom[((size_t)set * npair + k) * C + c] / (n - 1.0) : 0.0;
      const double r0 = block_reduce128(cov, 0, sh), r1 = block_reduce128(on ? d[i] * d[j] : 0.0, 0, sh);
      if (threadIdx.x == 0) { partial[((size_t)blockIdx.x * npair + k) * 2] = r0; partial[((size_t)blockIdx.x * npair + k) * 2 + 1] = r1; }
    }
}
void launch_diag_pairs(const double* mom, const double* momn, const double* comom, long long C, int P, const double* plan, int set, double* partial, double* out,
                       cudaStream_t st) {
  const long long nblk = (C + 127) / 128;
  const int npair = P * (P - 1) / 2;
  diag_pairs_partial_kernel<<<(unsigned)nblk, 128, 0, st>>>(mom, momn, comom, C, P, plan, set, partial);
  fold_kernel<<<(unsigned)((npair * 2 + 127) / 128), 128, 0, st>>>(partial, nblk, npair * 2, out);
}
void launch_diag1(const double* mom, const double* momn, long long C, int P, unsigned long long logit_mask, double* partial, double* out, cudaStream_t st) {
  const long long nblk = (C + 127) / 128;
  diag1_partial_kernel<<<(unsigned)nblk, 128, 0, st>>>(mom, momn, C, P, logit_mask, partial);
  diag1_fold_kernel<<<(unsigned)((P * kDiag1 + 127) / 128), 128, 0, st>>>(partial, nblk, P, out);
}
void launch_diag_plan(const double* r1, const int* monlink, int transform, int P, double* plan, cudaStream_t st) {
  diag_plan_kernel<<<(unsigned)((P + 127) / 128), 128, 0, st>>>(r1, monlink, transform, P, plan);
}
void launch_diag2(const double* mom, const double* momn, long long C, int P, const double* plan, double* partial, double* out, cudaStream_t st) {
  const long long nblk = (C + 127) / 128;
  diag2_partial_kernel<<<(unsigned)nblk, 128, 0, st>>>(mom, momn, C, P, plan, partial);
  fold_kernel<<<(unsigned)((P * kDiag2 + 127) / 128), 128, 0, st>>>(partial, nblk, P * kDiag2, out);
}


// FP64 FMA microbenchmark: the roofline denominator for the small-model kernels (not in MEASURED_PEAKS.json).
// 8 independent DFMA chains per thread, 148 x 8 blocks x 256 threads.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
double measure_fp64_peak_tflops(cudaStream_t st) {
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, st);
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;
    if (rep > 0 && ms > 0.f) best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

static inline unsigned gridf(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

void launch_init(long long n_chains, long long chain_offset, unsigned long long seed, int D, const double* inits,
                 long long n_inits, const int* elink, const double* ebound, double jitter_sd, double* state, cudaStream_t st) {
  init_kernel<<<gridf(n_chains, 256), 256, 0, st>>>(n_chains, chain_offset, seed, D, inits, n_inits, elink, ebound, jitter_sd, state);
}
void launch_soa_to_records(const double* soa, double* rec, long long C, int rows, cudaStream_t st) {
  soa_to_records<<<gridf(C, 256), 256, 0, st>>>(soa, rec, C, rows);
}
void launch_records_to_soa(const double* rec, double* soa, long long C, int rows, cudaStream_t st) {
  records_to_soa<<<gridf(C, 256), 256, 0, st>>>(rec, soa, C, rows);
}
void launch_samples_to_julia(const double* smp, double* out, long long kept, int P, long long C, cudaStream_t st) {
  samples_to_julia<<<gridf(kept * P * C, 256), 256, 0, st>>>(smp, out, kept, P, C);
}
void launch_gelman_partial(const double* mom, const double* momn, long long C, int P, const int* use_log,
                           const double* center, double* partial, cudaStream_t st) {
  gelman_partial_kernel<<<gridf(C, 128), 128, 0, st>>>(mom, momn, C, P, use_log, center, partial);
}
void launch_fold(const double* partial, long long nblocks, int width, double* out, cudaStream_t st) {
  fold_kernel<<<gridf(width, 128), 128, 0, st>>>(partial, nblocks, width, out);
}
void launch_minmax_partial(const double* mom, long long C, int P, double* partial, cudaStream_t st) {
  minmax_partial_kernel<<<gridf(C, 128), 128, 0, st>>>(mom, C, P, partial);
}
void launch_summary_partial(const double* mom, const double* momn, long long C, int P, const double* center,
                            double* partial, cudaStream_t st) {
  summary_partial_kernel<<<gridf(C, 128), 128, 0, st>>>(mom, momn, C, P, center, partial);
}

}  // namespace mcu
