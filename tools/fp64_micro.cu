// fp64_micro.cu — what bounds mixed FP64 code on a B200 SM: DFMA latency / throughput against resident warps and per-thread ILP, and how
// integer and conversion instructions share the issue port with it.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_micro fp64_micro.cu
// Output: one line per (kernel, warps per scheduler, ILP): DFMA per clock per SM, issue-slot use.  (profiles/r2_fp64_micro.md reads it.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N_IT = 4096;

// ILP independent DFMA chains per thread
template <int ILP>
__global__ void k_dfma(double* out, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ILP DFMA chains + NI independent integer multiply-add chains (Philox-like IMAD) per DFMA round
template <int ILP, int NI>
__global__ void k_mix_int(double* out, double a, double b, uint32_t m, long long* cyc) {
  double x[ILP]; uint32_t q[4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = threadIdx.x + i;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
      for (int i = 0; i < NI; ++i) q[i & 3] = q[i & 3] * m + (q[(i + 1) & 3] ^ 0x9E3779B9u);
    }
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(q[0] ^ q[1] ^ q[2] ^ q[3]);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ILP DFMA chains + one (rcp64h + int->double conversion) per NF DFMA rounds: the XU instructions of an FP64 log
template <int ILP>
__global__ void k_mix_xu(double* out, double a, double b, long long* cyc) {
  double x[ILP]; double y = 1.5 + threadIdx.x; int k = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
      if ((r & 3) == 0) {   // 2 XU instructions per 4 * ILP DFMA
        double rr; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(y)); y = rr + 1.25;
        k = k * 3 + 1; x[0] += (double)k;
      }
    }
  }
  const long long t1 = clock64();
  double s = y; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
static void run(const char* name, int ilp, double per_thread_dfma, double per_thread_other, F launch) {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* out; long long* cyc; cudaMalloc(&out, sizeof(double) * nsm * 2048); cudaMalloc(&cyc, 8);
  for (int wps : {1, 2, 3, 4, 6, 8}) {   // warps per scheduler: one block of 128 * wps threads per SM
    const int bs = 128 * wps > 1024 ? 1024 : 128 * wps;
    const int nb = nsm * ((128 * wps + bs - 1) / bs);
    if (128 * wps > 1024 && (128 * wps) % 1024) continue;
    launch(nb, bs, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); launch(nb, bs, out, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warps_sm = 4.0 * wps;
    const double dfma_per_clk_sm = per_thread_dfma * warps_sm * 32.0 / (double)c;
    const double issue = (per_thread_dfma + per_thread_other) * wps / (double)c;   // warp instructions per scheduler per clock
    printf("%-10s ilp %d warps/sched %d : %8.1f us  cycles %9lld  DFMA/clk/SM %6.2f (peak 64)  issue/clk/sched %.3f  cycles per dependent DFMA %.2f\n", name, ilp, wps,
           ms * 1e3, c, dfma_per_clk_sm, issue, (double)c / (per_thread_dfma / ilp));
  }
  cudaFree(out); cudaFree(cyc);
}

#define RUN_DFMA(I) run("dfma", I, (double)N_IT * 8 * I, 0.0, [](int nb, int bs, double* o, long long* c) { k_dfma<I><<<nb, bs>>>(o, 1.0000001, 1e-9, c); })
#define RUN_INT(I, NI) run("dfma+int" #NI, I, (double)N_IT * 8 * I, (double)N_IT * 8 * NI * 2, [](int nb, int bs, double* o, long long* c) { k_mix_int<I, NI><<<nb, bs>>>(o, 1.0000001, 1e-9, 0xD2511F53u, c); })
#define RUN_XU(I) run("dfma+xu", I, (double)N_IT * 8 * I, (double)N_IT * 2 * 5, [](int nb, int bs, double* o, long long* c) { k_mix_xu<I><<<nb, bs>>>(o, 1.0000001, 1e-9, c); })

int main() {
  RUN_DFMA(1); RUN_DFMA(2); RUN_DFMA(3); RUN_DFMA(4); RUN_DFMA(8);
  RUN_INT(1, 1); RUN_INT(2, 2); RUN_INT(3, 3); RUN_INT(4, 4); RUN_INT(4, 8);
  RUN_XU(1); RUN_XU(3); RUN_XU(4);
  return 0;
}
