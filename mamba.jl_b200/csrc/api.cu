// api.cu — the C ABI of libmambacuda.so (include/mambacuda.h): handle, data upload, scheme
// descriptors, engine launches, state round trip, batched density entry points and the
// on-device diagnostics.  Host side of what the Julia shim calls instead of
// mcmc_master!/mcmc_worker! (src/model/mcmc.jl:36-83).
//
// There is no CPU fallback anywhere in this file: every compute entry point launches CUDA
// kernels and fails with MCU_ERR_CUDA when no device is usable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "../../include/mambacuda.h"
#include "diagproto.hpp"
#include "hostdiag.hpp"
#include "launch.hpp"

using namespace mcu;

namespace {

thread_local std::string g_create_err;

struct mcu_ctx_impl;

}  // namespace

struct mcu_ctx {
  int tpl = -1;
  long long C = 0, chain_offset = 0;
  int device = 0;
  unsigned long long seed = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::map<std::string, std::vector<double>> inputs;
  std::map<std::string, double*> d_inputs;
  int* d_rat = nullptr;
  bool data_dirty = true;
  int D = 0, P = 0, NN = 0, glm_d = 0;
  std::vector<int> elink_state;   // link code per state element
  int* d_elink_state = nullptr;
  double* d_ebound_state = nullptr;   // (lo, hi) per state element (LINK_BOUNDED elements)
  // scheme
  std::vector<DevBlock> h_blocks;
  DevBlock* d_blocks = nullptr;
  std::vector<void*> scheme_allocs;
  long long tune_size = 0;
  // chain state
  double *d_state = nullptr, *d_tune = nullptr, *d_samples = nullptr, *d_mom = nullptr, *d_momn = nullptr;
  double* d_comom = nullptr;   // [2][P (P - 1) / 2][C] streaming within-chain co-moments (P <= 12), for the multivariate PSRF
  unsigned long long log_mask = 0ull;
  bool comom_ok = false;       // every kept draw since the last inits / state went into the co-moments (MCU_RUN_MPSRF on every run)
  size_t samples_cap = 0; long long samples_kept = 0;
  long long iter = 0;
  bool has_inits = false;
  // rng
  int rng_mode = MCU_RNG_PHILOX; double* d_ext = nullptr; size_t ext_n = 0; unsigned long long* d_ext_pos = nullptr;
  // bookkeeping
  std::string err;
  long long launches = 0;
  double last_ms = 0.0;
  bool seeds_fast_ok = false;
  bool pumps_gibbs_ok = false;                                              // fused pumps Gibbs + AMWG kernel (pumps_fast.cu)
  bool pumps_fast_ok = false;                                               // fused pumps Slice kernel (pumps_fast.cu)
  bool rats_fast_ok = false;                                                // fused rats Slice + AMWG kernel (rats_fast.cu)
  bool rats_warp_ok = false; double* r_scratch = nullptr; int r_grid = 0;   // warp-per-chain rats kernel (rats_warp.cu)
  std::vector<std::vector<double>> h_scales;                                // host mirror of every block's expanded scale
  std::vector<std::vector<double>> h_SigmaL;                                // host mirror of every block's lower Cholesky factor (HMC / MALA / AMM)
  void* d_diag = nullptr; size_t diag_cap = 0;     // persistent scratch of the diagnostics reductions (partials | folded sums | codes | centres)
  void* d_stage = nullptr; size_t stage_cap = 0;   // reusable device staging buffer (no cudaMalloc/cudaFree on the hot API calls)
  // GLM / NUTS tick engine buffers (glm_nuts.cu)
  double *g_sc = nullptr, *g_vec = nullptr, *g_req = nullptr, *g_lp = nullptr, *g_grad = nullptr, *g_part_lp = nullptr, *g_part_g = nullptr;
  int* g_nactive = nullptr; int g_nslab = 0;
  int* g_map = nullptr; long long g_pass_C = 0; int g_pass_nslab = 0;   // tick-engine compaction: pass slot k = chain g_map[k] (g_pass_C = 0: every chain, no map)
  long long compactions = 0; unsigned long long pass_slots = 0;   // chain slots the gradient passes of the tick engine carried (sum over ticks)
  double g_lp_const = 0.0;
  unsigned char* g_blob = nullptr; double* g_xty = nullptr; double* g_colscale = nullptr /* [2d]: column factors of the packed X and their reciprocals, or null */; int g_nslab_tc = 0; int glm_impl = 1; int glm_impl_run = 1;   // 1 = tensor-core kernel, 0 = FP64 reference kernel
  long long ticks = 0;
  unsigned long long* d_work = nullptr;   // device counter of gradient evaluations (rats_warp leapfrogs, GLM useful chain-gradients)
  bool g_reset = false;   // the GLM tick-engine records must be re-initialised before the next run (new inits / state)
  bool pending = false;   // an mcu_run(..., MCU_RUN_ASYNC) has not been waited for yet
  // cross-GPU diagnostics (mcu_comm_init / mcu_diag_global): NCCL communicator of this handle's rank, monitored-column links on the device
  void* comm = nullptr; int comm_rank = 0, comm_nranks = 1;
  int* d_monlink = nullptr; std::vector<int> h_monlink;
  unsigned long long logit_mask = 0ull;
};

namespace {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return MCU_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

int fail(mcu_handle h, int code, const std::string& msg) { h->err = msg; return code; }

// Returns a device buffer of at least `bytes` that lives as long as the handle; contents are scratch.
int stage(mcu_ctx* h, size_t bytes, void** out) {
  if (bytes > h->stage_cap) {
    if (h->d_stage) cudaFree(h->d_stage);
    h->d_stage = nullptr; h->stage_cap = 0;
    cudaError_t e = cudaMalloc(&h->d_stage, bytes);
    if (e != cudaSuccess) { h->err = std::string("cudaMalloc(stage): ") + cudaGetErrorString(e); return MCU_ERR_CUDA; }
    h->stage_cap = bytes;
  }
  *out = h->d_stage;
  return MCU_OK;
}

// device scratch that lives for one API call: freed on every return path (the CK macro returns early on a CUDA error)
struct DevScratch {
  void* p = nullptr;
  DevScratch() = default;
  DevScratch(const DevScratch&) = delete;
  DevScratch& operator=(const DevScratch&) = delete;
  ~DevScratch() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  double* f64() const { return static_cast<double*>(p); }
};

inline unsigned grid_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- template metadata on the host ---------------------------------------------------------------
template <class M> struct Host;

std::vector<double> lchoose_vec(const std::vector<double>& n, const std::vector<double>& r) {
  std::vector<double> lc(n.size());
  for (size_t i = 0; i < n.size(); ++i) lc[i] = std::lgamma(n[i] + 1.0) - std::lgamma(r[i] + 1.0) - std::lgamma(n[i] - r[i] + 1.0);
  return lc;
}

void default_inputs(mcu_ctx* h) {
  auto& in = h->inputs;
  switch (h->tpl) {
    case MCU_TPL_LINE:   // doc/tutorial/line.jl:69-73
      in["x"] = {1, 2, 3, 4, 5}; in["y"] = {1, 3, 3, 3, 5};
      break;
    case MCU_TPL_SEEDS:  // doc/examples/seeds.jl:4-12
      in["r"] = {10, 23, 23, 26, 17, 5, 53, 55, 32, 46, 10, 8, 10, 8, 23, 0, 3, 22, 15, 32, 3};
      in["n"] = {39, 62, 81, 51, 39, 6, 74, 72, 51, 79, 13, 16, 30, 28, 45, 4, 12, 41, 30, 51, 7};
      in["x1"] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
      in["x2"] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1};
      break;
    case MCU_TPL_RATS: {  // doc/examples/rats.jl:4-45
      static const double Y[150] = {
        151, 199, 246, 283, 320, 145, 199, 249, 293, 354, 147, 214, 263, 312, 328, 155, 200, 237, 272, 297,
        135, 188, 230, 280, 323, 159, 210, 252, 298, 331, 141, 189, 231, 275, 305, 159, 201, 248, 297, 338,
        177, 236, 285, 350, 376, 134, 182, 220, 260, 296, 160, 208, 261, 313, 352, 143, 188, 220, 273, 314,
        154, 200, 244, 289, 325, 171, 221, 270, 326, 358, 163, 216, 242, 281, 312, 160, 207, 248, 288, 324,
        142, 187, 234, 280, 316, 156, 203, 243, 283, 317, 157, 212, 259, 307, 336, 152, 203, 246, 286, 321,
        154, 205, 253, 298, 334, 139, 190, 225, 267, 302, 146, 191, 229, 272, 302, 157, 211, 250, 285, 323,
        132, 185, 237, 286, 331, 160, 207, 257, 303, 345, 169, 216, 261, 295, 333, 157, 205, 248, 289, 316,
        137, 180, 219, 258, 291, 153, 200, 244, 286, 324};
      const double xs[5] = {8.0, 15.0, 22.0, 29.0, 36.0};
      std::vector<double> y(Y, Y + 150), rat(150), Xm(150);
      for (int k = 0; k < 150; ++k) { rat[k] = k / 5; Xm[k] = xs[k % 5] - 22.0; }
      in["y"] = y; in["rat"] = rat; in["Xm"] = Xm; in["xbar"] = {22.0};
      break;
    }
    case MCU_TPL_DYES: {  // doc/examples/dyes.jl:4-17
      in["y"] = {1545, 1440, 1440, 1520, 1580, 1540, 1555, 1490, 1560, 1495, 1595, 1550, 1605, 1510, 1560,
                 1445, 1440, 1595, 1465, 1545, 1595, 1630, 1515, 1635, 1625, 1520, 1455, 1450, 1480, 1445};
      std::vector<double> batch(30); for (int k = 0; k < 30; ++k) batch[k] = k / 5;
      in["batch"] = batch;
      break;
    }
    case MCU_TPL_SALM:   // doc/examples/salm.jl:4-11 (y is reshape(..., 3, 6): column-major, plate fastest)
      in["y"] = {15, 21, 29, 16, 18, 21, 16, 26, 33, 27, 41, 60, 33, 38, 41, 20, 27, 42};
      in["x"] = {0, 10, 33, 100, 333, 1000};
      break;
    case MCU_TPL_EQUIV:  // doc/examples/equiv.jl:4-17 (y is a 10 x 2 matrix literal: flattened column-major, subject fastest)
      in["group"] = {1, 1, 2, 2, 2, 1, 1, 1, 2, 2};
      in["y"] = {1.40, 1.64, 1.44, 1.36, 1.65, 1.08, 1.09, 1.25, 1.25, 1.30, 1.65, 1.57, 1.58, 1.68, 1.69, 1.31, 1.43, 1.44, 1.39, 1.52};
      break;
    case MCU_TPL_BLOCKER:  // doc/examples/blocker.jl:4-18
      in["rt"] = {3, 7, 5, 102, 28, 4, 98, 60, 25, 138, 64, 45, 9, 57, 25, 33, 28, 8, 6, 32, 27, 22};
      in["nt"] = {38, 114, 69, 1533, 355, 59, 945, 632, 278, 1916, 873, 263, 291, 858, 154, 207, 251, 151, 174, 209, 391, 680};
      in["rc"] = {3, 14, 11, 127, 27, 6, 152, 48, 37, 188, 52, 47, 16, 45, 31, 38, 12, 6, 3, 40, 43, 39};
      in["nc"] = {39, 116, 93, 1520, 365, 52, 939, 471, 282, 1921, 583, 266, 293, 883, 147, 213, 122, 154, 134, 218, 364, 674};
      break;
    case MCU_TPL_STACKS:  // doc/examples/stacks.jl:4-30 (x: 21 x 3, row-major)
      in["y"] = {42, 37, 37, 28, 18, 18, 19, 20, 15, 14, 14, 13, 11, 12, 8, 7, 8, 8, 9, 15, 15};
      in["x"] = {80, 27, 89, 80, 27, 88, 75, 25, 90, 62, 24, 87, 62, 22, 87, 62, 23, 87, 62, 24, 93, 62, 24, 93, 58, 23, 87, 58, 18, 80, 58, 18, 89,
                 58, 17, 88, 58, 18, 82, 58, 19, 93, 50, 18, 89, 50, 18, 86, 50, 19, 72, 50, 19, 79, 50, 20, 80, 56, 20, 82, 70, 20, 91};
      break;
    case MCU_TPL_MAGNESIUM:  // doc/examples/magnesium.jl:4-9
      in["rt"] = {1, 9, 2, 1, 10, 1, 1, 90};
      in["nt"] = {40, 135, 200, 48, 150, 59, 25, 1159};
      in["rc"] = {2, 23, 7, 1, 8, 9, 3, 118};
      in["nc"] = {36, 135, 200, 46, 148, 56, 23, 1157};
      break;
    case MCU_TPL_OXFORD:  // doc/examples/oxford.jl:4-28
      in["r1"] = {3, 5, 2, 7, 7, 2, 5, 3, 5, 11, 6, 6, 11, 4, 4, 2, 8, 8, 6, 5, 15, 4, 9, 9, 4, 12, 8, 8, 6, 8,
          12, 4, 7, 16, 12, 9, 4, 7, 8, 11, 5, 12, 8, 17, 9, 3, 2, 7, 6, 5, 11, 14, 13, 8, 6, 4, 8, 4, 8, 7,
          15, 15, 9, 9, 5, 6, 3, 9, 12, 14, 16, 17, 8, 8, 9, 5, 9, 11, 6, 14, 21, 16, 6, 9, 8, 9, 8, 4, 11, 11,
          6, 9, 4, 4, 9, 9, 10, 14, 6, 3, 4, 6, 10, 4, 3, 3, 10, 4, 10, 5, 4, 3, 13, 1, 7, 5, 7, 6, 3, 7};
      in["n1"] = {28, 21, 32, 35, 35, 38, 30, 43, 49, 53, 31, 35, 46, 53, 61, 40, 29, 44, 52, 55, 61, 31, 48, 44, 42, 53, 56, 71, 43, 43,
          43, 40, 44, 70, 75, 71, 37, 31, 42, 46, 47, 55, 63, 91, 43, 39, 35, 32, 53, 49, 75, 64, 69, 64, 49, 29, 40, 27, 48, 43,
          61, 77, 55, 60, 46, 28, 33, 32, 46, 57, 56, 78, 58, 52, 31, 28, 46, 42, 45, 63, 71, 69, 43, 50, 31, 34, 54, 46, 58, 62,
          52, 41, 34, 52, 63, 59, 88, 62, 47, 53, 57, 74, 68, 61, 45, 45, 62, 73, 53, 39, 45, 51, 55, 41, 53, 51, 42, 46, 54, 32};
      in["r0"] = {0, 2, 2, 1, 2, 0, 1, 1, 1, 2, 4, 4, 2, 1, 7, 4, 3, 5, 3, 2, 4, 1, 4, 5, 2, 7, 5, 8, 2, 3,
          5, 4, 1, 6, 5, 11, 5, 2, 5, 8, 5, 6, 6, 10, 7, 5, 5, 2, 8, 1, 13, 9, 11, 9, 4, 4, 8, 6, 8, 6,
          8, 14, 6, 5, 5, 2, 4, 2, 9, 5, 6, 7, 5, 10, 3, 2, 1, 7, 9, 13, 9, 11, 4, 8, 2, 3, 7, 4, 7, 5,
          6, 6, 5, 6, 9, 7, 7, 7, 4, 2, 3, 4, 10, 3, 4, 2, 10, 5, 4, 5, 4, 6, 5, 3, 2, 2, 4, 6, 4, 1};
      in["n0"] = {28, 21, 32, 35, 35, 38, 30, 43, 49, 53, 31, 35, 46, 53, 61, 40, 29, 44, 52, 55, 61, 31, 48, 44, 42, 53, 56, 71, 43, 43,
          43, 40, 44, 70, 75, 71, 37, 31, 42, 46, 47, 55, 63, 91, 43, 39, 35, 32, 53, 49, 75, 64, 69, 64, 49, 29, 40, 27, 48, 43,
          61, 77, 55, 60, 46, 28, 33, 32, 46, 57, 56, 78, 58, 52, 31, 28, 46, 42, 45, 63, 71, 69, 43, 50, 31, 34, 54, 46, 58, 62,
          52, 41, 34, 52, 63, 59, 88, 62, 47, 53, 57, 74, 68, 61, 45, 45, 62, 73, 53, 39, 45, 51, 55, 41, 53, 51, 42, 46, 54, 32};
      in["year"] = {-10, -9, -9, -8, -8, -8, -7, -7, -7, -7, -6, -6, -6, -6, -6, -5, -5, -5, -5, -5, -5, -4, -4, -4, -4, -4, -4, -4, -3, -3,
          -3, -3, -3, -3, -3, -3, -2, -2, -2, -2, -2, -2, -2, -2, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0,
          0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3,
          3, 3, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 9, 9, 10};
      break;
    case MCU_TPL_EPIL:  // doc/examples/epil.jl:4-24
      in["y"] = {5, 3, 2, 4, 7, 5, 6, 40, 5, 14, 26, 12, 4, 7, 16, 11, 0, 37, 3, 3, 3, 3, 2, 8, 18, 2, 3, 13, 11, 8,
          0, 3, 2, 4, 22, 5, 2, 3, 4, 2, 0, 5, 11, 10, 19, 1, 6, 2, 102, 4, 8, 1, 18, 6, 3, 1, 2, 0, 1, 3,
          5, 4, 4, 18, 2, 4, 20, 6, 13, 12, 6, 4, 9, 24, 0, 0, 29, 5, 0, 4, 4, 3, 12, 24, 1, 1, 15, 14, 7, 4,
          6, 6, 3, 17, 4, 4, 7, 18, 1, 2, 4, 14, 5, 7, 1, 10, 1, 65, 3, 6, 3, 11, 3, 5, 23, 3, 0, 4, 3, 3,
          0, 1, 9, 8, 0, 21, 6, 6, 6, 8, 6, 12, 10, 0, 3, 28, 2, 6, 3, 3, 3, 2, 76, 2, 4, 13, 9, 9, 3, 1,
          7, 1, 19, 7, 0, 7, 2, 1, 4, 0, 25, 3, 6, 2, 8, 0, 72, 2, 5, 1, 28, 4, 4, 19, 0, 0, 3, 3, 3, 5,
          4, 21, 7, 2, 12, 5, 0, 22, 4, 2, 14, 9, 5, 3, 29, 5, 7, 4, 4, 5, 8, 25, 1, 2, 12, 8, 4, 0, 3, 4,
          3, 16, 4, 4, 7, 5, 0, 0, 3, 15, 8, 7, 3, 8, 0, 63, 4, 7, 5, 13, 0, 3, 8, 1, 0, 2};   // 59 x 4, column-major (patient fastest)
      in["Trt"] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,
          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
      in["Base"] = {11, 11, 6, 8, 66, 27, 12, 52, 23, 10, 52, 33, 18, 42, 87, 50, 18, 111, 18, 20, 12, 9, 17, 28, 55, 9, 10, 47, 76, 38,
          19, 10, 19, 24, 31, 14, 11, 67, 41, 7, 22, 13, 46, 36, 38, 7, 36, 11, 151, 22, 41, 32, 56, 24, 16, 22, 25, 13, 12};
      in["Age"] = {31, 30, 25, 36, 22, 29, 31, 42, 37, 28, 36, 24, 23, 36, 26, 26, 28, 31, 32, 21, 29, 21, 32, 25, 30, 40, 19, 22, 18, 32,
          20, 30, 18, 24, 30, 35, 27, 20, 22, 28, 23, 40, 33, 21, 35, 25, 26, 25, 22, 32, 25, 35, 21, 41, 32, 26, 21, 36, 37};
      in["V4"] = {0, 0, 0, 1};
      break;
    case MCU_TPL_SURGICAL:  // doc/examples/surgical.jl:4-8
      in["r"] = {0, 18, 8, 46, 8, 13, 9, 31, 14, 8, 29, 24};
      in["n"] = {47, 148, 119, 810, 211, 196, 148, 215, 207, 97, 256, 360};
      break;
    case MCU_TPL_PUMPS:  // doc/examples/pumps.jl:4-9
      in["y"] = {5, 1, 5, 14, 3, 19, 1, 1, 4, 22};
      in["t"] = {94.3, 15.7, 62.9, 126, 5.24, 31.4, 1.05, 1.05, 2.1, 10.5};
      break;
    default: break;
  }
}

bool scheme_is_seeds_fast(const mcu_ctx* h);
int upload_inputs(mcu_ctx* h) {
  if (!h->data_dirty) return MCU_OK;
  // inputs that must agree in length (DimensionMismatch in the reference: simulation.jl:20-23)
  if (h->tpl == MCU_TPL_LINE && h->inputs["x"].size() != h->inputs["y"].size()) { h->err = "line: x and y differ in length"; return MCU_ERR_DIM; }
  if (h->tpl == MCU_TPL_GLM_LOGIT && h->glm_d > 0 && h->inputs["X"].size() != h->inputs["y"].size() * (size_t)h->glm_d) { h->err = "GLM: y must have one entry per row of X"; return MCU_ERR_DIM; }
  if (!h->h_blocks.empty()) h->seeds_fast_ok = scheme_is_seeds_fast(h);   // the fused-kernel eligibility depends on the data (0/1 design)
  for (auto& kv : h->d_inputs) cudaFree(kv.second);
  h->d_inputs.clear();
  if (h->d_rat) { cudaFree(h->d_rat); h->d_rat = nullptr; }
  auto in = h->inputs;   // derived arrays
  if (h->tpl == MCU_TPL_SEEDS || h->tpl == MCU_TPL_SURGICAL) in["lc"] = lchoose_vec(in["n"], in["r"]);
  if (h->tpl == MCU_TPL_STACKS) {   // meanx, sdx (sample sd), z = (x - meanx) / sdx: stacks.jl:32-37
    const auto& x = in["x"]; const int N = (int)in["y"].size();
    std::vector<double> mean(3, 0.0), sd(3, 0.0), z((size_t)N * 3);
    for (int j = 0; j < 3; ++j) { for (int i = 0; i < N; ++i) mean[j] += x[i * 3 + j]; mean[j] /= N; }
    for (int j = 0; j < 3; ++j) { double s = 0; for (int i = 0; i < N; ++i) { const double e = x[i * 3 + j] - mean[j]; s += e * e; } sd[j] = std::sqrt(s / (N - 1)); }
    for (int i = 0; i < N; ++i) for (int j = 0; j < 3; ++j) z[i * 3 + j] = (x[i * 3 + j] - mean[j]) / sd[j];
    in["meanx"] = mean; in["sdx"] = sd; in["z"] = z;
  }
  if (h->tpl == MCU_TPL_BLOCKER || h->tpl == MCU_TPL_MAGNESIUM) { in["lcc"] = lchoose_vec(in["nc"], in["rc"]); in["lct"] = lchoose_vec(in["nt"], in["rt"]); }
  if (h->tpl == MCU_TPL_OXFORD) { in["lc1"] = lchoose_vec(in["n1"], in["r1"]); in["lc0"] = lchoose_vec(in["n0"], in["r0"]); }
  if (h->tpl == MCU_TPL_EPIL) {
    // centred covariates logBase4 = log(Base / 4), Trt, BT = logBase4 .* Trt, logAge = log(Age), V4 and their means: epil.jl:25-30
    const auto &Base = in["Base"], &Trt = in["Trt"], &Age = in["Age"], &V4 = in["V4"];
    const int NPT = EpilModel::NPAT, NV = EpilModel::NV;
    std::vector<double> cov((size_t)4 * NPT + NV + 5, 0.0), lb(NPT), bt(NPT), la(NPT);
    double m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
    for (int i = 0; i < NPT; ++i) { lb[i] = std::log(Base[i] / 4.0); bt[i] = lb[i] * Trt[i]; la[i] = std::log(Age[i]); m1 += lb[i]; m2 += Trt[i]; m3 += bt[i]; m4 += la[i]; }
    for (int j = 0; j < NV; ++j) m5 += V4[j];
    m1 /= NPT; m2 /= NPT; m3 /= NPT; m4 /= NPT; m5 /= NV;
    for (int i = 0; i < NPT; ++i) { cov[i] = lb[i] - m1; cov[NPT + i] = Trt[i] - m2; cov[2 * NPT + i] = bt[i] - m3; cov[3 * NPT + i] = la[i] - m4; }
    for (int j = 0; j < NV; ++j) cov[4 * NPT + j] = V4[j] - m5;
    const double bars[5] = {m1, m2, m3, m4, m5};
    for (int k = 0; k < 5; ++k) cov[4 * NPT + NV + k] = bars[k];
    in["cov"] = cov;
  }
  if (h->tpl == MCU_TPL_PUMPS || h->tpl == MCU_TPL_SALM || h->tpl == MCU_TPL_EPIL) { std::vector<double> l; for (double y : in["y"]) l.push_back(std::lgamma(y + 1.0)); in["lgy1"] = l; }
  for (auto& kv : in) {
    if (kv.second.empty()) continue;
    double* p = nullptr;
    CK(cudaMalloc(&p, kv.second.size() * sizeof(double)));
    CK(cudaMemcpyAsync(p, kv.second.data(), kv.second.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->d_inputs[kv.first] = p;
  }
  if (h->tpl == MCU_TPL_DYES) {
    std::vector<int> bt; for (double r : in["batch"]) bt.push_back((int)r);
    CK(cudaMalloc(&h->d_rat, bt.size() * sizeof(int)));
    CK(cudaMemcpyAsync(h->d_rat, bt.data(), bt.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  if (h->tpl == MCU_TPL_RATS) {
    std::vector<int> rat; for (double r : in["rat"]) rat.push_back((int)r);
    CK(cudaMalloc(&h->d_rat, rat.size() * sizeof(int)));
    CK(cudaMemcpyAsync(h->d_rat, rat.data(), rat.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  h->data_dirty = false;
  return MCU_OK;
}

template <> struct Host<LineModel> {
  static LineModel::Data data(mcu_ctx* h) { return {h->d_inputs["x"], h->d_inputs["y"], (int)h->inputs["y"].size()}; }
};
template <> struct Host<SeedsModel> {
  static SeedsModel::Data data(mcu_ctx* h) {
    return {h->d_inputs["r"], h->d_inputs["n"], h->d_inputs["x1"], h->d_inputs["x2"], h->d_inputs["lc"], (int)h->inputs["r"].size()};
  }
};
template <> struct Host<RatsModel> {
  static RatsModel::Data data(mcu_ctx* h) { return {h->d_inputs["y"], h->d_inputs["Xm"], h->d_rat, (int)h->inputs["y"].size(), h->inputs["xbar"][0]}; }
};
template <> struct Host<DyesModel> {
  static DyesModel::Data data(mcu_ctx* h) { return {h->d_inputs["y"], h->d_rat, (int)h->inputs["y"].size()}; }
};
template <> struct Host<SalmModel> {
  static SalmModel::Data data(mcu_ctx* h) { return {h->d_inputs["y"], h->d_inputs["x"], h->d_inputs["lgy1"], (int)h->inputs["y"].size()}; }
};
template <> struct Host<EquivModel> {
  static EquivModel::Data data(mcu_ctx* h) { return {h->d_inputs["y"], h->d_inputs["group"], (int)h->inputs["group"].size()}; }
};
template <> struct Host<BlockerModel> {
  static BlockerModel::Data data(mcu_ctx* h) {
    return {h->d_inputs["rc"], h->d_inputs["nc"], h->d_inputs["rt"], h->d_inputs["nt"], h->d_inputs["lcc"], h->d_inputs["lct"], (int)h->inputs["rc"].size()};
  }
};
template <> struct Host<MagnesiumModel> {
  static MagnesiumModel::Data data(mcu_ctx* h) {
    // s2 = 1 / (rt + 0.5) + 1 / (nt - rt + 0.5) + 1 / (rc + 0.5) + 1 / (nc - rc + 0.5), s2_0 = 1 / mean(1 / s2): magnesium.jl:13-17
    const auto &rt = h->inputs["rt"], &nt = h->inputs["nt"], &rc = h->inputs["rc"], &nc = h->inputs["nc"];
    double sinv = 0.0;
    for (size_t j = 0; j < rt.size(); ++j) sinv += 1.0 / (1.0 / (rt[j] + 0.5) + 1.0 / (nt[j] - rt[j] + 0.5) + 1.0 / (rc[j] + 0.5) + 1.0 / (nc[j] - rc[j] + 0.5));
    const double s2_0 = 1.0 / (sinv / (double)rt.size());
    return {h->d_inputs["rc"], h->d_inputs["nc"], h->d_inputs["rt"], h->d_inputs["nt"], h->d_inputs["lcc"], h->d_inputs["lct"], s2_0, std::sqrt(s2_0 / std::erf(0.75))};
  }
};
template <> struct Host<OxfordModel> {
  static OxfordModel::Data data(mcu_ctx* h) {
    return {h->d_inputs["r1"], h->d_inputs["n1"], h->d_inputs["r0"], h->d_inputs["n0"], h->d_inputs["year"], h->d_inputs["lc1"], h->d_inputs["lc0"]};
  }
};
template <> struct Host<EpilModel> {
  static EpilModel::Data data(mcu_ctx* h) { return {h->d_inputs["y"], h->d_inputs["lgy1"], h->d_inputs["cov"]}; }
};
template <> struct Host<StacksModel> {
  static StacksModel::Data data(mcu_ctx* h) { return {h->d_inputs["y"], h->d_inputs["z"], h->d_inputs["meanx"], h->d_inputs["sdx"], (int)h->inputs["y"].size()}; }
};
template <> struct Host<SurgicalModel> {
  static SurgicalModel::Data data(mcu_ctx* h) { return {h->d_inputs["r"], h->d_inputs["n"], h->d_inputs["lc"], (int)h->inputs["r"].size()}; }
};
template <> struct Host<PumpsModel> {
  static PumpsModel::Data data(mcu_ctx* h) { return {h->d_inputs["y"], h->d_inputs["t"], h->d_inputs["lgy1"], (int)h->inputs["y"].size()}; }
};
inline int glm_family(mcu_ctx* h) { auto it = h->inputs.find("family"); return it == h->inputs.end() || it->second.empty() ? 0 : (int)it->second[0]; }
inline double glm_sigma(mcu_ctx* h) { auto it = h->inputs.find("sigma"); return it == h->inputs.end() || it->second.empty() ? 1.0 : it->second[0]; }
template <> struct Host<GlmM> {
  static GlmM::Data data(mcu_ctx* h) { return {h->d_inputs["X"], h->d_inputs["y"], (int)h->inputs["y"].size(), h->glm_d, glm_family(h), glm_sigma(h)}; }
};

#define MCU_DISPATCH(h, BODY)                                                      \
  switch ((h)->tpl) {                                                              \
    case MCU_TPL_LINE: { typedef LineModel M; BODY; break; }                       \
    case MCU_TPL_SEEDS: { typedef SeedsModel M; BODY; break; }                     \
    case MCU_TPL_RATS: { typedef RatsModel M; BODY; break; }                       \
    case MCU_TPL_PUMPS: { typedef PumpsModel M; BODY; break; }                     \
    case MCU_TPL_GLM_LOGIT: { typedef GlmM M; BODY; break; }                       \
    case MCU_TPL_SURGICAL: { typedef SurgicalModel M; BODY; break; }               \
    case MCU_TPL_DYES: { typedef DyesModel M; BODY; break; }                       \
    case MCU_TPL_SALM: { typedef SalmModel M; BODY; break; }                       \
    case MCU_TPL_BLOCKER: { typedef BlockerModel M; BODY; break; }                 \
    case MCU_TPL_STACKS: { typedef StacksModel M; BODY; break; }                   \
    case MCU_TPL_EQUIV: { typedef EquivModel M; BODY; break; }                     \
    case MCU_TPL_MAGNESIUM: { typedef MagnesiumModel M; BODY; break; }             \
    case MCU_TPL_OXFORD: { typedef OxfordModel M; BODY; break; }                   \
    case MCU_TPL_EPIL: { typedef EpilModel M; BODY; break; }                       \
    default: return fail(h, MCU_ERR_ARG, "unknown template");                      \
  }

struct TplInfo {
  int D, P, NN; std::vector<int> off, len, link, monlink; std::vector<std::string> node_names;
  std::vector<int> elink;        // link code of every state element (the node's, unless the node is an array of different distributions)
  std::vector<double> ebound;    // (lo, hi) of every state element; read for LINK_BOUNDED elements only
};

template <class M> TplInfo tpl_info_fixed() {
  TplInfo t; t.D = M::D; t.P = M::P; t.NN = M::NN;
  for (int n = 0; n < M::NN; ++n) { t.off.push_back(M::node_off(n)); t.len.push_back(M::node_len(n)); t.link.push_back(M::node_link(n)); t.node_names.push_back(M::node_name(n)); }
  for (int j = 0; j < M::P; ++j) t.monlink.push_back(M::mon_link(j));
  t.elink.assign(M::D, LINK_IDENT); t.ebound.assign(2 * (size_t)M::D, 0.0);
  for (int n = 0; n < M::NN; ++n) for (int e = M::node_off(n); e < M::node_off(n) + M::node_len(n); ++e) {
    t.elink[e] = M::elem_link(e, M::node_link(n));
    M::elem_bounds(e, t.ebound[2 * e], t.ebound[2 * e + 1]);
  }
  return t;
}
TplInfo tpl_info(const mcu_ctx* h) {
  switch (h->tpl) {
    case MCU_TPL_LINE: return tpl_info_fixed<LineModel>();
    case MCU_TPL_SEEDS: return tpl_info_fixed<SeedsModel>();
    case MCU_TPL_RATS: return tpl_info_fixed<RatsModel>();
    case MCU_TPL_PUMPS: return tpl_info_fixed<PumpsModel>();
    case MCU_TPL_SURGICAL: return tpl_info_fixed<SurgicalModel>();
    case MCU_TPL_DYES: return tpl_info_fixed<DyesModel>();
    case MCU_TPL_SALM: return tpl_info_fixed<SalmModel>();
    case MCU_TPL_BLOCKER: return tpl_info_fixed<BlockerModel>();
    case MCU_TPL_STACKS: return tpl_info_fixed<StacksModel>();
    case MCU_TPL_EQUIV: return tpl_info_fixed<EquivModel>();
    case MCU_TPL_MAGNESIUM: return tpl_info_fixed<MagnesiumModel>();
    case MCU_TPL_OXFORD: return tpl_info_fixed<OxfordModel>();
    case MCU_TPL_EPIL: return tpl_info_fixed<EpilModel>();
    default: {
      TplInfo t; t.D = h->glm_d; t.P = h->glm_d; t.NN = 1;
      t.off = {0}; t.len = {h->glm_d}; t.link = {LINK_IDENT}; t.node_names = {"beta"};
      t.monlink.assign(h->glm_d, LINK_IDENT);
      t.elink.assign(h->glm_d, LINK_IDENT); t.ebound.assign(2 * (size_t)h->glm_d, 0.0);
      return t;
    }
  }
}

std::string names_of(const mcu_ctx* h, int which) {
  TplInfo t = tpl_info(h);
  std::string s;
  auto add = [&](const std::string& x) { if (!s.empty()) s += "\n"; s += x; };
  if (which == 2) { for (auto& n : t.node_names) add(n); return s; }
  if (which == 1) {
    switch (h->tpl) {
      case MCU_TPL_LINE: return LineModel::monitor_names();
      case MCU_TPL_SEEDS: return SeedsModel::monitor_names();
      case MCU_TPL_RATS: return RatsModel::monitor_names();
      case MCU_TPL_PUMPS: return PumpsModel::monitor_names();
      case MCU_TPL_SURGICAL: return SurgicalModel::monitor_names();
      case MCU_TPL_DYES: return DyesModel::monitor_names();
      case MCU_TPL_SALM: return SalmModel::monitor_names();
      case MCU_TPL_BLOCKER: return BlockerModel::monitor_names();
      case MCU_TPL_STACKS: return StacksModel::monitor_names();
      case MCU_TPL_EQUIV: return EquivModel::monitor_names();
      case MCU_TPL_MAGNESIUM: return MagnesiumModel::monitor_names();
      case MCU_TPL_OXFORD: return OxfordModel::monitor_names();
      case MCU_TPL_EPIL: return EpilModel::monitor_names();
      default: break;
    }
  }
  for (int n = 0; n < t.NN; ++n) {
    if (t.len[n] == 1 && h->tpl != MCU_TPL_GLM_LOGIT) add(t.node_names[n]);
    else for (int e = 0; e < t.len[n]; ++e) add(t.node_names[n] + "[" + std::to_string(e + 1) + "]");
  }
  return s;
}

void free_scheme(mcu_ctx* h) {
  for (void* p : h->scheme_allocs) cudaFree(p);
  h->scheme_allocs.clear();
  if (h->d_blocks) { cudaFree(h->d_blocks); h->d_blocks = nullptr; }
  h->h_blocks.clear();
}
// packed X / y tiles, X'y and the family constants: depend on the data only (they survive mcu_set_inits / mcu_set_state)
void free_glm_data(mcu_ctx* h) {
  cudaFree(h->g_blob); h->g_blob = nullptr; cudaFree(h->g_xty); h->g_xty = nullptr; cudaFree(h->g_colscale); h->g_colscale = nullptr;
}
// tick-engine state of the chains
void free_glm_buffers(mcu_ctx* h) {
  cudaFree(h->g_sc); cudaFree(h->g_vec); cudaFree(h->g_req); cudaFree(h->g_lp); cudaFree(h->g_grad); cudaFree(h->g_part_lp); cudaFree(h->g_part_g);
  cudaFree(h->g_nactive); cudaFree(h->g_map); h->g_map = nullptr; h->g_pass_C = 0;
  h->g_sc = h->g_vec = h->g_req = h->g_lp = h->g_grad = h->g_part_lp = h->g_part_g = nullptr; h->g_nactive = nullptr;
}
void free_chain_buffers(mcu_ctx* h) {
  free_glm_buffers(h);
  cudaFree(h->d_state); cudaFree(h->d_tune); cudaFree(h->d_samples); cudaFree(h->d_mom); cudaFree(h->d_momn); cudaFree(h->d_comom);
  h->d_state = h->d_tune = h->d_samples = h->d_mom = h->d_momn = h->d_comom = nullptr;
  h->samples_cap = 0; h->samples_kept = 0;
}

bool chol_lower_host(const std::vector<double>& A, int n, std::vector<double>& L) {
  L.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    double d = A[j + (size_t)j * n];
    for (int c = 0; c < j; ++c) d -= L[j + (size_t)c * n] * L[j + (size_t)c * n];
    if (!(d > 0.0)) return false;
    const double dj = std::sqrt(d); L[j + (size_t)j * n] = dj;
    for (int i = j + 1; i < n; ++i) {
      double a = A[i + (size_t)j * n];
      for (int c = 0; c < j; ++c) a -= L[i + (size_t)c * n] * L[j + (size_t)c * n];
      L[i + (size_t)j * n] = a / dj;
    }
  }
  return true;
}

int ensure_chain_buffers(mcu_ctx* h) {
  if (h->d_state) return MCU_OK;
  const size_t C = (size_t)h->C;
  CK(cudaMalloc(&h->d_state, sizeof(double) * C * std::max(1, h->D)));
  CK(cudaMalloc(&h->d_tune, sizeof(double) * C * (size_t)std::max(1LL, h->tune_size)));
  CK(cudaMalloc(&h->d_mom, sizeof(double) * C * (size_t)h->P * kMomPerCol));
  CK(cudaMalloc(&h->d_momn, sizeof(double) * C * 3));
  CK(cudaMemsetAsync(h->d_tune, 0, sizeof(double) * C * (size_t)std::max(1LL, h->tune_size), h->stream));
  CK(cudaMemsetAsync(h->d_mom, 0, sizeof(double) * C * (size_t)h->P * kMomPerCol, h->stream));
  CK(cudaMemsetAsync(h->d_momn, 0, sizeof(double) * C * 3, h->stream));
  if (diag_npair(h->P) > 0) {
    CK(cudaMalloc(&h->d_comom, sizeof(double) * C * 2 * (size_t)diag_npair(h->P)));
    CK(cudaMemsetAsync(h->d_comom, 0, sizeof(double) * C * 2 * (size_t)diag_npair(h->P), h->stream));
  }
  return MCU_OK;
}

bool scheme_is_seeds_fast(const mcu_ctx* h) {
  // AMWG(alpha0..alpha12) , AMWG(b) , AMWG(s2) — SURVEY.md §8d config 2 scheme A — or the reference's scheme B with AMM(alpha0..alpha12)
  // as the first block (doc/examples/seeds.jl:69-71)
  if (h->tpl != MCU_TPL_SEEDS || h->h_blocks.size() != 3) return false;
  const DevBlock& a = h->h_blocks[0]; const DevBlock& b = h->h_blocks[1]; const DevBlock& c = h->h_blocks[2];
  if ((a.kind != MCU_AMWG && a.kind != MCU_AMM) || b.kind != MCU_AMWG || c.kind != MCU_AMWG) return false;
  if (a.n_own != 4 || a.own[0] != 0 || a.own[1] != 1 || a.own[2] != 2 || a.own[3] != 3) return false;
  if (b.n_own != 1 || b.own[0] != 5) return false;
  if (c.n_own != 1 || c.own[0] != 4) return false;
  if ((int)h->inputs.at("r").size() != SeedsModel::NP) return false;
  for (const char* nm : {"x1", "x2"}) for (double v : h->inputs.at(nm)) if (v != 0.0 && v != 1.0) return false;   // 0/1 design → 4 group bases
  return true;
}

bool scheme_is_pumps_fast(const mcu_ctx* h) {
  // Slice([alpha, beta], 1.0, Univariate), Slice(theta, 1.0, Univariate) on the constrained scale: doc/examples/pumps.jl:52-53
  if (h->tpl != MCU_TPL_PUMPS || h->h_blocks.size() != 2) return false;
  const DevBlock& a = h->h_blocks[0]; const DevBlock& b = h->h_blocks[1];
  return a.kind == MCU_SLICE_UNI && a.transform == 0 && a.n_own == 2 && a.own[0] == 0 && a.own[1] == 1 &&
         b.kind == MCU_SLICE_UNI && b.transform == 0 && b.n_own == 1 && b.own[0] == 2;
}

bool scheme_is_pumps_gibbs(const mcu_ctx* h) {
  // Gibbs(theta), Gibbs(beta), AMWG(alpha): BASELINE.json configs[4] / SURVEY.md §8d config 5
  if (h->tpl != MCU_TPL_PUMPS || h->h_blocks.size() != 3) return false;
  const DevBlock* b = h->h_blocks.data();
  return b[0].kind == MCU_GIBBS && b[0].n_own == 1 && b[0].own[0] == 2 && b[1].kind == MCU_GIBBS && b[1].n_own == 1 && b[1].own[0] == 1 &&
         b[2].kind == MCU_AMWG && b[2].n_own == 1 && b[2].own[0] == 0;
}

bool scheme_is_rats_fast(const mcu_ctx* h) {
  // Slice(s2_c), AMWG(alpha), Slice([mu_alpha, s2_alpha]; univariate), AMWG(beta), Slice([mu_beta, s2_beta]; univariate): doc/examples/rats.jl:112-116
  if (h->tpl != MCU_TPL_RATS || h->h_blocks.size() != 5) return false;
  const DevBlock* b = h->h_blocks.data();
  auto own1 = [&](int k, int n0) { return b[k].n_own == 1 && b[k].own[0] == n0; };
  auto own2 = [&](int k, int n0, int n1) { return b[k].n_own == 2 && b[k].own[0] == n0 && b[k].own[1] == n1; };
  return b[0].kind == MCU_SLICE_MULTI && b[0].transform == 0 && own1(0, 4) && b[1].kind == MCU_AMWG && own1(1, 5) &&
         b[2].kind == MCU_SLICE_UNI && b[2].transform == 0 && own2(2, 0, 2) && b[3].kind == MCU_AMWG && own1(3, 6) &&
         b[4].kind == MCU_SLICE_UNI && b[4].transform == 0 && own2(4, 1, 3);
}

bool scheme_is_rats_warp(const mcu_ctx* h) {
  // NUTS(alpha, beta, mu_alpha, mu_beta) , Slice(s2_c, s2_alpha, s2_beta; univariate, constrained) — SURVEY.md §8d config 3
  if (h->tpl != MCU_TPL_RATS || h->h_blocks.size() != 2) return false;
  const DevBlock& a = h->h_blocks[0]; const DevBlock& b = h->h_blocks[1];
  if (a.kind != MCU_NUTS || a.grad != MCU_GRAD_ANALYTIC || b.kind != MCU_SLICE_UNI || b.transform != 0) return false;
  if (a.n_own != 4 || a.own[0] != 5 || a.own[1] != 6 || a.own[2] != 0 || a.own[3] != 1) return false;
  if (b.n_own != 3 || b.own[0] != 4 || b.own[1] != 2 || b.own[2] != 3) return false;
  return true;
}

bool scheme_is_glm_tick(const mcu_ctx* h) {
  if (h->tpl != MCU_TPL_GLM_LOGIT || h->h_blocks.size() != 1) return false;
  const DevBlock& b = h->h_blocks[0];
  return b.kind == MCU_NUTS && b.grad == MCU_GRAD_ANALYTIC && b.n_own == 1;
}

// one CTA per SM (tensor memory and shared memory are both fully used): pick the number of row slabs that minimises
// waves x tiles-per-slab, i.e. the idle SMs of the last wave (4,096 chains = 32 groups: 9 slabs = 288 CTAs = 1.95 waves
// instead of 4 slabs = 128 CTAs on 148 SMs)
int glm_tc_choose_nslab(long long N, long long C, int sms) {
  const long long groups = (C + 127) / 128;
  const long long NT = glm_tc_num_tiles(N);
  long long ns = 1; double best = 1e300;
  for (long long cand = 1; cand <= NT && cand * groups <= 8LL * sms; ++cand) {
    const long long waves = (cand * groups + sms - 1) / sms, tps = (NT + cand - 1) / cand;
    const double cost = (double)waves * ((double)tps + 6.0);   // + ~6 tiles of prologue / epilogue per CTA
    if (cost < best - 1e-9) { best = cost; ns = cand; }
  }
  return (int)ns;
}

int reset_glm_records(mcu_ctx* h) {
  const size_t C = (size_t)h->C, d = (size_t)h->D;
  const size_t nsc = glm_tick_scalar_slots(), nv = glm_tick_vector_slots();
  CK(cudaMemsetAsync(h->g_sc, 0, sizeof(double) * nsc * C, h->stream));
  CK(cudaMemsetAsync(h->g_vec, 0, sizeof(double) * nv * d * C, h->stream));
  CK(cudaMemsetAsync(h->g_req, 0, sizeof(double) * d * C, h->stream));
  CK(cudaMemsetAsync(h->g_lp, 0, sizeof(double) * C, h->stream));
  CK(cudaMemsetAsync(h->g_grad, 0, sizeof(double) * d * C, h->stream));
  CK(cudaMemsetAsync(h->g_nactive, 0, 4 * sizeof(int), h->stream));
  // resume point: every chain's own iteration counter starts at the handle's
  if (h->iter > 0) {
    std::vector<double> it(C, (double)h->iter);
    CK(cudaMemcpyAsync(h->g_sc + 1 * C, it.data(), sizeof(double) * C, cudaMemcpyHostToDevice, h->stream));   // slot 1 = SL_ITER
    CK(cudaStreamSynchronize(h->stream));
  }
  h->g_reset = false;
  return MCU_OK;
}

int ensure_glm_buffers(mcu_ctx* h) {
  if (h->g_sc) return h->g_reset ? reset_glm_records(h) : MCU_OK;
  const size_t C = (size_t)h->C, d = (size_t)h->D;
  const size_t nsc = glm_tick_scalar_slots(), nv = glm_tick_vector_slots();
  int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
  long long nslab = ((long long)n_sm * 512 + h->C - 1) / h->C;
  const long long N = (long long)h->inputs["y"].size();
  if (nslab > (N + 63) / 64) nslab = (N + 63) / 64;
  if (nslab < 1) nslab = 1;
  const long long nslab_ref = nslab; h->g_nslab = (int)nslab_ref;
  h->g_nslab_tc = glm_tc_choose_nslab(N, h->C, n_sm);
  {
    // FP64 / FP32 partials: one per (slab, chain slot); a compacted pass has fewer 128-chain groups and more slabs, slabs x groups <= 8 SMs
    const long long npart = (long long)h->g_nslab_tc * glm_tc_nsub(N, h->g_nslab_tc);
    if (npart > nslab) nslab = npart;
    const long long cap = (8LL * n_sm * 128 + h->C - 1) / h->C;   // slabs x 128-chain groups of any compacted pass, in units of C slots
    if (cap > nslab) nslab = cap;
  }
  if (const char* e = std::getenv("MCU_GLM_IMPL")) h->glm_impl = std::atoi(e);
  CK(cudaMalloc(&h->g_sc, sizeof(double) * nsc * C));
  CK(cudaMalloc(&h->g_vec, sizeof(double) * nv * d * C));
  CK(cudaMalloc(&h->g_req, sizeof(double) * d * C));
  CK(cudaMalloc(&h->g_lp, sizeof(double) * C));
  CK(cudaMalloc(&h->g_grad, sizeof(double) * d * C));
  CK(cudaMalloc(&h->g_part_lp, sizeof(double) * nslab * C));
  CK(cudaMalloc(&h->g_part_g, sizeof(double) * nslab * d * C));
  CK(cudaMalloc(&h->g_nactive, 4 * sizeof(int)));
  CK(cudaMalloc(&h->g_map, sizeof(int) * C));
  h->g_pass_C = 0;
  CK(cudaMemsetAsync(h->g_nactive, 0, 4 * sizeof(int), h->stream));
  if (!h->g_blob) {
    const size_t blob_bytes = (size_t)glm_tc_num_tiles(N) * glm_tc_tile_bytes(h->D);
    CK(cudaMalloc(&h->g_blob, blob_bytes));
    const int fam = glm_family(h);
    {
      // fp16 hi / lo operands: a column whose largest magnitude is above 2^14 (fp16 overflows at 65,504) or below 2^-6 (the lo term goes
      // subnormal) is packed times a power of two that brings it into [1, 2); the pass divides Theta and the gradient by the same factor
      const std::vector<double>& Xs = h->inputs["X"];
      std::vector<double> cmax(h->D, 0.0), cs(2 * (size_t)h->D, 1.0);
      for (long long i = 0; i < N; ++i) for (int j = 0; j < h->D; ++j) { const double v = std::fabs(Xs[(size_t)i * h->D + j]); if (v > cmax[j]) cmax[j] = v; }
      bool any = false;
      for (int j = 0; j < h->D; ++j) {
        if (!std::isfinite(cmax[j])) return fail(h, MCU_ERR_ARG, "GLM design matrix holds a non-finite value");
        if (cmax[j] > 0.0 && (cmax[j] > 16384.0 || cmax[j] < 0.015625)) { int e = 0; std::frexp(cmax[j], &e); cs[j] = std::ldexp(1.0, 1 - e); cs[h->D + j] = std::ldexp(1.0, e - 1); any = true; }
      }
      if (any) {
        CK(cudaMalloc(&h->g_colscale, sizeof(double) * 2 * h->D));
        CK(cudaMemcpyAsync(h->g_colscale, cs.data(), sizeof(double) * 2 * h->D, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
      }
    }
    glm_tc_pack(h->d_inputs["X"], h->d_inputs["y"], (int)N, h->D, fam, h->g_blob, h->stream, h->g_colscale); h->launches++;
    // X'(y - 1/2) (FP64, once): sum_i (y_i - 1/2) eta_i = beta . X'(y - 1/2) is the part of the log-likelihood that is linear in
    // beta (y eta from the Bernoulli term, -eta/2 from softplus(eta) = eta/2 + |eta|/2 + log(1 + e^-|eta|)); the fold adds it
    std::vector<double> xty(h->D, 0.0);
    const std::vector<double>& Xh = h->inputs["X"]; const std::vector<double>& yh = h->inputs["y"];
    // Bernoulli: X'(y - 1/2) (see above); Poisson: X'y, constant -sum lgamma(y + 1); Normal: the kernel forms the whole quadratic,
    // constant -N (log sigma + log(2 pi) / 2)
    h->g_lp_const = 0.0;
    if (fam != 2) for (long long i = 0; i < N; ++i) { const double yi = fam == 0 ? yh[i] - 0.5 : yh[i]; for (int j = 0; j < h->D; ++j) xty[j] += yi * Xh[(size_t)i * h->D + j]; }
    if (fam == 1) for (long long i = 0; i < N; ++i) h->g_lp_const -= std::lgamma(yh[i] + 1.0);
    if (fam == 2) h->g_lp_const = -(double)N * (std::log(glm_sigma(h)) + 0.5 * 1.8378770664093454835606594728112);
    CK(cudaMalloc(&h->g_xty, sizeof(double) * h->D));
    CK(cudaMemcpyAsync(h->g_xty, xty.data(), sizeof(double) * h->D, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  { const int rc = reset_glm_records(h); if (rc) return rc; }
  return MCU_OK;
}

int glm_gradient_dispatch(mcu_ctx* h, int N) {
  if (h->glm_impl == 1) {
    const bool cmp = h->g_pass_C > 0;                                   // compacted pass: only the chains that are still running
    const long long Cp = cmp ? h->g_pass_C : h->C; const int ns = cmp ? h->g_pass_nslab : h->g_nslab_tc;
    const int* map = cmp ? h->g_map : nullptr;
    if (glm_tc_launch(h->g_blob, N, h->D, Cp, h->g_req, ns, h->g_part_lp, reinterpret_cast<float*>(h->g_part_g), glm_family(h), glm_sigma(h), h->stream, map, h->C, h->g_colscale ? h->g_colscale + h->D : nullptr) != 0)
      return fail(h, MCU_ERR_CUDA, "glm_tc_kernel launch failed");
    glm_fold_tc(h->g_part_lp, reinterpret_cast<const float*>(h->g_part_g), ns, ns * glm_tc_nsub(N, ns), h->D, Cp,
                h->g_req, h->g_xty, h->g_lp_const, h->g_lp, h->g_grad, h->stream, map, h->C, h->g_colscale ? h->g_colscale + h->D : nullptr);
  } else {
    glm_grad_reference(h->d_inputs["X"], h->d_inputs["y"], N, h->D, h->C, h->g_req, h->g_nslab, h->g_part_lp, h->g_part_g,
                       h->g_lp, h->g_grad, glm_family(h), glm_sigma(h), h->stream);
    h->launches += 1;
  }
  h->launches += 2;
  return MCU_OK;
}

int run_glm_tick(mcu_ctx* h, long long iters, long long burnin, long long thin, const RunArgs& a) {
  int rc = ensure_glm_buffers(h); if (rc) return rc;
  const DevBlock& b = h->h_blocks[0];
  GlmTick t;
  t.C = h->C; t.chain_offset = h->chain_offset; t.seed = h->seed; t.target_iter = h->iter + iters; t.burnin = burnin; t.thin = thin;
  t.row0 = a.row0; t.d = h->D;
  t.max_depth = b.max_depth > 0 ? (b.max_depth < kMaxDepth ? b.max_depth : kMaxDepth) : kMaxDepth;
  t.target = b.target; t.eps_desc = b.epsilon;
  t.state = h->d_state; t.tune = h->d_tune + (size_t)b.tune_off * h->C; t.sc = h->g_sc; t.vec = h->g_vec; t.req = h->g_req;
  t.lp = h->g_lp; t.grad = h->g_grad; t.samples = a.samples; t.mom = h->d_mom; t.momn = h->d_momn; t.n_active = h->g_nactive; t.work = h->d_work;
  const int N = (int)h->inputs["y"].size();
  // A tick = advance every chain to its next gradient request, then one gradient pass.  The host only looks at
  // the running-chain counter every kCheck ticks (finished chains idle; at most kCheck - 1 passes are wasted at the end).
  int kCheck = 8;
  if (const char* e = std::getenv("MCU_GLM_CHECK")) { const int v = std::atoi(e); if (v > 0) kCheck = v; }
  int tick = 0;
  h->g_pass_C = 0;                                   // every chain runs again
  bool compact = h->glm_impl_run == 1 && h->glm_impl == 1;
  if (const char* e = std::getenv("MCU_GLM_COMPACT")) compact = compact && std::atoi(e) != 0;
  int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
  while (true) {
    int slot = 0;
    for (int k = 0; k < kCheck; ++k, ++tick) {
      t.tick = tick; slot = tick & 1;
      glm_advance(t, h->stream); h->launches++;
      const int saved = h->glm_impl; if (h->glm_impl_run == 0) h->glm_impl = 0;
      rc = glm_gradient_dispatch(h, N); h->glm_impl = saved;
      if (rc) return rc;
      h->ticks++; h->pass_slots += (unsigned long long)(h->g_pass_C > 0 ? h->g_pass_C : h->C);
    }
    int active = 0;
    CK(cudaMemcpyAsync(&active, h->g_nactive + slot, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (active == 0) break;
    // Chains finish at very different ticks (NUTS trees differ in depth by up to 2^10 leaves): once a whole 128-chain group of the
    // current pass has gone idle, the running chains are compacted into the leading pass slots and the tensor-core pass shrinks with them
    const long long cur = h->g_pass_C > 0 ? h->g_pass_C : h->C;
    if (compact && (active + 127) / 128 < (cur + 127) / 128) {
      glm_compact(h->g_sc, h->C, h->g_map, h->g_nactive + 2, h->stream); h->launches++;
      int cnt = 0;
      CK(cudaMemcpyAsync(&cnt, h->g_nactive + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (cnt > 0) { h->g_pass_C = cnt; h->g_pass_nslab = glm_tc_choose_nslab(N, cnt, n_sm); h->compactions++; }
    }
  }
  h->g_pass_C = 0;
  return MCU_OK;
}


// ---- NCCL, bound at run time ------------------------------------------------------------------------------------------------
// libmambacuda.so has no link-time dependency on NCCL: the library is looked up when a communicator is first asked for
// (dlopen finds the copy the host process already loaded, e.g. torch's, else the system one), so single-GPU users need no NCCL at all.
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, mcu_nccl_id, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable ABI)

NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return &api;
  const char* names[] = {std::getenv("MCU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) { api.err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return &api; }
  auto sym = [&](const char* n) { void* f = dlsym(api.lib, n); if (!f && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + n; return f; };
  api.GetUniqueId = reinterpret_cast<int (*)(void*)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(void**, int, mcu_nccl_id, int)>(sym("ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclAllReduce"));
  api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
  api.CommDestroy = reinterpret_cast<int (*)(void*)>(sym("ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  if (!api.err.empty()) { dlclose(api.lib); api.lib = nullptr; }
  return &api;
}
#define NK(call)                                                                                   \
  do {                                                                                             \
    const int r_ = (call);                                                                         \
    if (r_ != 0) {                                                                                 \
      h->err = std::string(#call) + ": " + (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "NCCL error"); \
      return MCU_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

// device scratch of the two-round protocol: [partials | round-1 buffer 11P | plan 5P | round-2 buffer 15P]
struct DiagBufs { double* partial; double* r1; double* plan; double* r2; };
int diag_bufs(mcu_ctx* h, DiagBufs* b) {
  const long long nblk = grid_for(h->C, 128);
  const size_t P = (size_t)h->P;
  const size_t npair = (size_t)diag_npair(h->P);
  const size_t need = sizeof(double) * ((size_t)nblk * std::max(P * kDiag2, 2 * npair) + P * (kDiag1 + 5 + kDiag2) + 2 * npair) + 64;
  if (need > h->diag_cap) {
    if (h->d_diag) cudaFree(h->d_diag);
    h->d_diag = nullptr; h->diag_cap = 0;
    if (cudaMalloc(&h->d_diag, need) != cudaSuccess) return fail(h, MCU_ERR_CUDA, "cudaMalloc(diagnostics scratch) failed");
    h->diag_cap = need;
  }
  b->partial = static_cast<double*>(h->d_diag);
  b->r1 = b->partial + (size_t)nblk * std::max(P * kDiag2, 2 * npair);
  b->plan = b->r1 + P * kDiag1;
  b->r2 = b->plan + P * 5;
  return MCU_OK;
}
int diag_finish_host(int64_t n_kept, int p, double alpha, const int* monlink, int transform, const double* r1, const double* r2,
                     double* psrf, double* summary, int* codes_out, double* mpsrf) {
  std::vector<int> codes(p); std::vector<double> ctrs((size_t)p * 4);
  for (int j = 0; j < p; ++j) {
    double* ctr = ctrs.data() + (size_t)j * 4;
    const int code = diag_plan_column(p, j, monlink[j], transform, r1, ctr);
    if (code == 2 && j >= 64) return MCU_ERR_UNSUPPORTED;   // logit moments are streamed for the first 64 columns only
    codes[j] = code;
    if (codes_out) codes_out[j] = code;
    const double* s = r2 + (size_t)j * kDiag2;
    if (psrf) {
      if (s[0] < 2.0) return MCU_ERR_ARG;                    // "less than 2 chains supplied to gelman diagnostic": gelmandiag.jl:6-7
      hostdiag::gelman_column((double)n_kept, ctr[0], ctr[1], s, alpha, psrf + j * 2);
    }
    if (summary) hostdiag::summary_column((double)n_kept, ctr[2], ctr[3], s + 7, summary + j * 5);
  }
  if (mpsrf) {   // gelmandiag.jl:49-55 from the streamed within-chain co-moments: W = mean_k S_k, B = n cov_k(chain means)
    *mpsrf = NAN;
    const int npair = diag_npair(p);
    bool ok = npair > 0;
    for (int j = 0; j < p && ok; ++j) ok = codes[j] == (transform && monlink[j] == kPlanLinkLog ? 1 : 0);   // co-moments exist on the raw and the node-link scale only
    if (ok) {
      const double m = r2[0], n = (double)n_kept;
      std::vector<double> W((size_t)p * p), B((size_t)p * p);
      for (int j = 0; j < p; ++j) {
        const double* s = r2 + (size_t)j * kDiag2;
        W[j * p + j] = ctrs[(size_t)j * 4 + 1] + s[3] / m;
        B[j * p + j] = n * (s[2] - s[1] * s[1] / m) / (m - 1.0);
      }
      const double* pr = r2 + (size_t)p * kDiag2;
      int k = 0;
      for (int i = 0; i < p; ++i)
        for (int j = i + 1; j < p; ++j, ++k) {
          const double di = r2[(size_t)i * kDiag2 + 1], dj = r2[(size_t)j * kDiag2 + 1];
          W[i * p + j] = W[j * p + i] = pr[2 * k] / m;
          B[i * p + j] = B[j * p + i] = n * (pr[2 * k + 1] - di * dj / m) / (m - 1.0);
        }
      *mpsrf = hostdiag::mpsrf_from_WB(W, B, (long long)n_kept, (long long)m, p);
    }
  }
  return MCU_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int mcu_abi_version(void) { return MCU_ABI_VERSION; }

int mcu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* mcu_last_error(mcu_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int mcu_create(int template_id, int64_t n_chains, int64_t chain_offset, int device, uint64_t seed, mcu_handle* out) {
  if (!out) { g_create_err = "out is NULL"; return MCU_ERR_ARG; }
  *out = nullptr;
  if (template_id < 0 || template_id >= MCU_N_TEMPLATES) { g_create_err = "unknown template id"; return MCU_ERR_ARG; }
  if (n_chains < 1) { g_create_err = "n_chains must be positive"; return MCU_ERR_ARG; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); libmambacuda has no CPU fallback";
    return MCU_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return MCU_ERR_ARG; }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return MCU_ERR_CUDA; }
  mcu_ctx* h = new mcu_ctx();
  h->tpl = template_id; h->C = n_chains; h->chain_offset = chain_offset; h->device = device; h->seed = seed;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
    g_create_err = "cannot create stream/events"; delete h; return MCU_ERR_CUDA;
  }
  default_inputs(h);
  TplInfo t = tpl_info(h);
  h->D = t.D; h->P = t.P; h->NN = t.NN;
  *out = h;
  return MCU_OK;
}

int mcu_destroy(mcu_handle h) {
  if (!h) return MCU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_scheme(h); free_chain_buffers(h);
  if (h->comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy(h->comm);
  cudaFree(h->d_monlink); cudaFree(h->d_work);
  for (auto& kv : h->d_inputs) cudaFree(kv.second);
  cudaFree(h->d_rat); cudaFree(h->d_elink_state); cudaFree(h->d_ebound_state); cudaFree(h->d_ext); cudaFree(h->d_ext_pos); cudaFree(h->d_stage); cudaFree(h->d_diag); cudaFree(h->r_scratch); free_glm_data(h);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaStreamDestroy(h->stream);
  delete h;
  return MCU_OK;
}

int mcu_set_data(mcu_handle h, const char* name, int ndim, const int64_t* dims, const double* ptr) {
  if (!h || !name || !dims || !ptr || ndim < 1) return h ? fail(h, MCU_ERR_ARG, "bad argument to mcu_set_data") : MCU_ERR_ARG;
  size_t n = 1; for (int i = 0; i < ndim; ++i) n *= (size_t)dims[i];
  const std::string nm(name);
  if (h->tpl == MCU_TPL_GLM_LOGIT && nm == "X") {
    if (ndim != 2) return fail(h, MCU_ERR_DIM, "X must be [N × d]");
    if (dims[1] < 1 || dims[1] > kGlmDMax) return fail(h, MCU_ERR_UNSUPPORTED, "GLM template supports 1 <= d <= 128");
    if (h->glm_d != (int)dims[1]) { free_chain_buffers(h); free_scheme(h); h->has_inits = false; }
    h->glm_d = (int)dims[1]; h->D = h->glm_d; h->P = h->glm_d;
  } else if (h->tpl != MCU_TPL_GLM_LOGIT) {
    if (h->inputs.find(nm) == h->inputs.end()) return fail(h, MCU_ERR_ARG, "template has no input named " + nm);
    if (h->tpl == MCU_TPL_SEEDS && n != (size_t)SeedsModel::NP) return fail(h, MCU_ERR_DIM, "seeds inputs have 21 entries");
    if (h->tpl == MCU_TPL_PUMPS && n != (size_t)PumpsModel::NPUMP) return fail(h, MCU_ERR_DIM, "pumps inputs have 10 entries");
    if (h->tpl == MCU_TPL_SURGICAL && n != (size_t)SurgicalModel::NH) return fail(h, MCU_ERR_DIM, "surgical inputs have 12 entries");
    if (h->tpl == MCU_TPL_DYES && n != 30) return fail(h, MCU_ERR_DIM, "dyes inputs have 30 entries");
    if (h->tpl == MCU_TPL_STACKS && n != (nm == "x" ? 63u : 21u)) return fail(h, MCU_ERR_DIM, "stacks inputs: y has 21 entries, x 21 x 3");
    if (h->tpl == MCU_TPL_BLOCKER && n != (size_t)BlockerModel::NT) return fail(h, MCU_ERR_DIM, "blocker inputs have 22 entries");
    if (h->tpl == MCU_TPL_MAGNESIUM && n != (size_t)MagnesiumModel::NTR) return fail(h, MCU_ERR_DIM, "magnesium inputs have 8 entries");
    if (h->tpl == MCU_TPL_OXFORD && n != (size_t)OxfordModel::K) return fail(h, MCU_ERR_DIM, "oxford inputs have 120 entries");
    if (h->tpl == MCU_TPL_EPIL && n != (nm == "y" ? 236u : nm == "V4" ? 4u : 59u)) return fail(h, MCU_ERR_DIM, "epil inputs: y has 236 entries (59 x 4), V4 has 4, Trt / Base / Age have 59");
    if (h->tpl == MCU_TPL_SALM && n != (nm == "x" ? 6u : 18u)) return fail(h, MCU_ERR_DIM, "salm inputs: y has 18 entries (3 x 6), x has 6");
    if (h->tpl == MCU_TPL_EQUIV && n != (nm == "group" ? 10u : 20u)) return fail(h, MCU_ERR_DIM, "equiv inputs: y has 20 entries (10 x 2), group has 10");
    if (h->tpl == MCU_TPL_RATS && nm != "xbar" && n != 150) return fail(h, MCU_ERR_DIM, "rats inputs have 150 entries");
    // index inputs are 0-based offsets into the state record (the Julia shim converts the scripts' 1-based rat / batch: rats.jl:42, dyes.jl:16)
    if (h->tpl == MCU_TPL_RATS && nm == "rat") for (size_t i = 0; i < n; ++i) if (!(ptr[i] >= 0.0 && ptr[i] < 30.0 && ptr[i] == std::floor(ptr[i]))) return fail(h, MCU_ERR_ARG, "rat must hold 0-based integer indices in [0, 30)");
    if (h->tpl == MCU_TPL_DYES && nm == "batch") for (size_t i = 0; i < n; ++i) if (!(ptr[i] >= 0.0 && ptr[i] < 6.0 && ptr[i] == std::floor(ptr[i]))) return fail(h, MCU_ERR_ARG, "batch must hold 0-based integer indices in [0, 6)");
  }
  if (h->tpl == MCU_TPL_GLM_LOGIT && nm == "family" && (n != 1 || !(ptr[0] == 0.0 || ptr[0] == 1.0 || ptr[0] == 2.0)))
    return fail(h, MCU_ERR_ARG, "family must be 0 (Bernoulli / logit), 1 (Poisson / log) or 2 (Normal / identity)");
  if (h->tpl == MCU_TPL_GLM_LOGIT && nm == "sigma" && (n != 1 || !(ptr[0] > 0.0))) return fail(h, MCU_ERR_ARG, "sigma must be positive");
  if (h->tpl == MCU_TPL_GLM_LOGIT) { free_glm_buffers(h); free_glm_data(h); }   // packed X / y, X'y and the constants depend on the data
  h->inputs[nm].assign(ptr, ptr + n);
  h->data_dirty = true;
  return MCU_OK;
}

int mcu_dims(mcu_handle h, int* D, int* n_monitor, int* n_nodes) {
  if (!h) return MCU_ERR_ARG;
  if (D) *D = h->D;
  if (n_monitor) *n_monitor = h->P;
  if (n_nodes) *n_nodes = h->NN;
  return MCU_OK;
}

int mcu_names(mcu_handle h, int which, char* buf, size_t buflen) {
  if (!h) return MCU_ERR_ARG;
  const std::string s = names_of(h, which);
  if (!buf || s.size() + 1 > buflen) return (int)s.size() + 1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return MCU_OK;
}

int mcu_tune_size(mcu_handle h, int64_t* n) {
  if (!h || !n) return MCU_ERR_ARG;
  *n = h->tune_size;
  return MCU_OK;
}

int mcu_set_scheme(mcu_handle h, int n_blocks, const mcu_block_desc* blocks) {
  if (!h || n_blocks < 1 || !blocks) return h ? fail(h, MCU_ERR_ARG, "bad argument to mcu_set_scheme") : MCU_ERR_ARG;
  CK(cudaSetDevice(h->device));
  if (h->tpl == MCU_TPL_GLM_LOGIT && h->glm_d == 0) return fail(h, MCU_ERR_STATE, "inputs must be set before the scheme (GLM needs X)");
  TplInfo t = tpl_info(h);
  free_scheme(h);
  free_chain_buffers(h);
  h->has_inits = false;
  long long toff = 0;
  std::vector<DevBlock> hb;
  std::vector<std::vector<double>> h_scales, h_SigmaL;
  for (int bi = 0; bi < n_blocks; ++bi) {
    const mcu_block_desc& d = blocks[bi];
    if (d.kind < MCU_AMWG || d.kind > MCU_MALA) return fail(h, MCU_ERR_ARG, "unknown sampler kind");
    if (d.kind == MCU_GIBBS) {
      bool ok = false;
      if (d.n_nodes == 1) { MCU_DISPATCH(h, ok = M::has_gibbs(d.nodes[0])); }
      if (!ok) return fail(h, MCU_ERR_UNSUPPORTED, "no conjugate full conditional for this node on the device (user-defined samplers have no device equivalent)");
    }
    if (d.n_nodes < 1 || d.n_nodes > MCU_MAX_BLOCK_NODES) return fail(h, MCU_ERR_ARG, "block must name 1..8 nodes");
    if (d.kind == MCU_NUTS || d.kind == MCU_HMC || d.kind == MCU_MALA || d.kind == MCU_AMM) {
      bool grad_ok = true;
      MCU_DISPATCH(h, grad_ok = M::kGradSamplers);
      if (!grad_ok) return fail(h, MCU_ERR_UNSUPPORTED, "NUTS / HMC / MALA / AMM are not compiled for this template (240+ state elements per chain): use AMWG / Slice / RWM blocks");
    }
    DevBlock b; std::memset(&b, 0, sizeof(b));
    b.kind = d.kind;
    b.transform = (d.kind == MCU_SLICE_UNI || d.kind == MCU_SLICE_MULTI) ? (d.transform != 0) : (d.kind == MCU_GIBBS ? 0 : 1);   // slice.jl:47-50; others sampler files :53
    b.adapt = d.adapt;
    if (d.adapt < MCU_ADAPT_ALL || d.adapt > MCU_ADAPT_NONE) return fail(h, MCU_ERR_ARG, "adapt must be one of :all, :burnin, or :none");   // amwg.jl:49-50
    b.batchsize = d.batchsize > 0 ? d.batchsize : 50;
    b.proposal = d.proposal; b.L = d.L; b.grad = d.grad; b.max_depth = d.max_depth;
    b.target = d.target > 0 ? d.target : (d.kind == MCU_NUTS ? 0.6 : 0.44);
    b.epsilon = d.epsilon; b.beta = d.beta > 0 ? d.beta : 0.05; b.amm_scale = d.amm_scale > 0 ? d.amm_scale : 2.38;
    std::vector<int> elem, elink; std::vector<double> ebound; bool any_bounded = false;
    for (int i = 0; i < d.n_nodes; ++i) {
      const int n = d.nodes[i];
      if (n < 0 || n >= t.NN) return fail(h, MCU_ERR_ARG, "node id out of range");
      if (b.mask & (1u << n)) return fail(h, MCU_ERR_ARG, "node listed twice in a block");
      b.mask |= 1u << n; b.own[i] = n;
      for (int e = 0; e < t.len[n]; ++e) {
        const int se = t.off[n] + e;
        elem.push_back(se); elink.push_back(t.elink[se]); ebound.push_back(t.ebound[2 * se]); ebound.push_back(t.ebound[2 * se + 1]);
        any_bounded = any_bounded || t.elink[se] == LINK_BOUNDED;
      }
    }
    b.n_own = d.n_nodes; b.k = (int)elem.size();
    const int k = b.k;
    // scale: sigma / width / scale  (validate(): amwg.jl:38-43, slice.jl:34-40, rwm.jl:40-46)
    std::vector<double> sc;
    const bool needs_scale = d.kind == MCU_AMWG || d.kind == MCU_SLICE_UNI || d.kind == MCU_SLICE_MULTI || d.kind == MCU_RWM;
    if (needs_scale) {
      if (!d.scale || (d.n_scale != 1 && d.n_scale != k))
        return fail(h, MCU_ERR_ARG, "length(scale) differs from variate length " + std::to_string(k));
      for (int i = 0; i < k; ++i) sc.push_back(d.n_scale == 1 ? d.scale[0] : d.scale[i]);
    }
    std::vector<double> SL;
    if (d.kind == MCU_AMM || ((d.kind == MCU_HMC || d.kind == MCU_MALA) && d.scale)) {
      if (!d.scale || d.n_scale != k * k) return fail(h, MCU_ERR_ARG, "Sigma dimension differs from variate length " + std::to_string(k));   // amm.jl:36-41, hmc.jl:36-41
      if (d.kind == MCU_AMM && k > kAmmMaxK) return fail(h, MCU_ERR_UNSUPPORTED, "AMM blocks are limited to 8 elements on the device");
      std::vector<double> S(d.scale, d.scale + (size_t)k * k);
      if (!chol_lower_host(S, k, SL)) return fail(h, MCU_ERR_ARG, "Sigma is not positive definite");
    }
    if (d.kind == MCU_HMC && (d.L < 1 || !(d.epsilon > 0))) return fail(h, MCU_ERR_ARG, "HMC needs epsilon > 0 and L >= 1");
    if (d.kind == MCU_MALA && !(d.epsilon > 0)) return fail(h, MCU_ERR_ARG, "MALA needs epsilon > 0");
    if (d.kind == MCU_RWM && (d.proposal < 0 || d.proposal > MCU_PROP_TRIWEIGHT)) return fail(h, MCU_ERR_UNSUPPORTED, "RWM proposal not available on the device");
    if (d.grad < 0 || d.grad > 2) return fail(h, MCU_ERR_ARG, "unknown gradient mode");
    int *de = nullptr, *dl = nullptr; double *ds = nullptr, *dS = nullptr;
    CK(cudaMalloc(&de, sizeof(int) * k)); h->scheme_allocs.push_back(de);
    CK(cudaMalloc(&dl, sizeof(int) * k)); h->scheme_allocs.push_back(dl);
    CK(cudaMemcpy(de, elem.data(), sizeof(int) * k, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dl, elink.data(), sizeof(int) * k, cudaMemcpyHostToDevice));
    double* db = nullptr;
    if (any_bounded) { CK(cudaMalloc(&db, sizeof(double) * 2 * k)); h->scheme_allocs.push_back(db); CK(cudaMemcpy(db, ebound.data(), sizeof(double) * 2 * k, cudaMemcpyHostToDevice)); }
    if (!sc.empty()) { CK(cudaMalloc(&ds, sizeof(double) * k)); h->scheme_allocs.push_back(ds); CK(cudaMemcpy(ds, sc.data(), sizeof(double) * k, cudaMemcpyHostToDevice)); }
    if (!SL.empty()) { CK(cudaMalloc(&dS, sizeof(double) * k * k)); h->scheme_allocs.push_back(dS); CK(cudaMemcpy(dS, SL.data(), sizeof(double) * k * k, cudaMemcpyHostToDevice)); }
    b.elem = de; b.elink = dl; b.ebound = db; b.scale = ds; b.SigmaL = dS;
    h_scales.push_back(sc);
    h_SigmaL.push_back(SL);
    b.tune_off = (int)toff;
    switch (d.kind) {   // tune record layout: see samplers.cuh
      case MCU_AMWG: toff += 2 + 2 * k; break;
      case MCU_NUTS: toff += 8; break;
      case MCU_AMM: toff += 2 + k + 2 * k * k; break;
      default: break;
    }
    hb.push_back(b);
  }
  h->h_blocks = hb;
  h->tune_size = toff;
  CK(cudaMalloc(&h->d_blocks, sizeof(DevBlock) * hb.size()));
  CK(cudaMemcpy(h->d_blocks, hb.data(), sizeof(DevBlock) * hb.size(), cudaMemcpyHostToDevice));
  // link code of every state element (for init jitter)
  h->elink_state = t.elink;
  cudaFree(h->d_elink_state); h->d_elink_state = nullptr; cudaFree(h->d_ebound_state); h->d_ebound_state = nullptr;
  CK(cudaMalloc(&h->d_elink_state, sizeof(int) * h->D));
  CK(cudaMemcpy(h->d_elink_state, h->elink_state.data(), sizeof(int) * h->D, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_ebound_state, sizeof(double) * 2 * h->D));
  CK(cudaMemcpy(h->d_ebound_state, t.ebound.data(), sizeof(double) * 2 * h->D, cudaMemcpyHostToDevice));
  h->h_scales = h_scales; h->h_SigmaL = h_SigmaL;
  h->h_monlink = t.monlink; h->logit_mask = 0ull; h->log_mask = 0ull;
  for (int j = 0; j < h->P && j < 64; ++j) { if (t.monlink[j] == LINK_HEUR) h->logit_mask |= 1ull << j; if (t.monlink[j] == LINK_LOG) h->log_mask |= 1ull << j; }
  cudaFree(h->d_monlink); h->d_monlink = nullptr;
  CK(cudaMalloc(&h->d_monlink, sizeof(int) * std::max(1, h->P)));
  CK(cudaMemcpy(h->d_monlink, t.monlink.data(), sizeof(int) * h->P, cudaMemcpyHostToDevice));
  h->seeds_fast_ok = scheme_is_seeds_fast(h);
  h->rats_warp_ok = scheme_is_rats_warp(h);
  h->rats_fast_ok = scheme_is_rats_fast(h);
  h->pumps_fast_ok = scheme_is_pumps_fast(h);
  h->pumps_gibbs_ok = scheme_is_pumps_gibbs(h);
  return MCU_OK;
}

int mcu_set_inits(mcu_handle h, const double* x, int64_t n_inits, double jitter_sd) {
  if (!h || !x) return h ? fail(h, MCU_ERR_ARG, "missing initial values") : MCU_ERR_ARG;
  if (n_inits < 1) return fail(h, MCU_ERR_ARG, "fewer initial values than chains");   // mcmc.jl:24-25
  if (h->h_blocks.empty()) return fail(h, MCU_ERR_STATE, "set the sampling scheme before the initial values");
  CK(cudaSetDevice(h->device));
  int rc = upload_inputs(h); if (rc) return rc;
  rc = ensure_chain_buffers(h); if (rc) return rc;
  double* d_in = nullptr;
  rc = stage(h, sizeof(double) * (size_t)n_inits * h->D, (void**)&d_in); if (rc) return rc;
  CK(cudaMemcpyAsync(d_in, x, sizeof(double) * (size_t)n_inits * h->D, cudaMemcpyHostToDevice, h->stream));
  launch_init(h->C, h->chain_offset, h->seed, h->D, d_in, n_inits, h->d_elink_state, h->d_ebound_state, jitter_sd, h->d_state, h->stream);
  h->launches++;
  const size_t C = (size_t)h->C;
  CK(cudaMemsetAsync(h->d_tune, 0, sizeof(double) * C * (size_t)std::max(1LL, h->tune_size), h->stream));
  CK(cudaMemsetAsync(h->d_mom, 0, sizeof(double) * C * (size_t)h->P * kMomPerCol, h->stream));
  CK(cudaMemsetAsync(h->d_momn, 0, sizeof(double) * C * 3, h->stream));
  if (h->d_comom) CK(cudaMemsetAsync(h->d_comom, 0, sizeof(double) * C * 2 * (size_t)diag_npair(h->P), h->stream));
  if (h->d_ext_pos) CK(cudaMemsetAsync(h->d_ext_pos, 0, sizeof(unsigned long long) * C, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->iter = 0; h->has_inits = true; h->samples_kept = 0; h->comom_ok = h->d_comom != nullptr;
  h->g_reset = true;   // the tick engine's chain records start over (buffers are kept: no cudaMalloc / cudaFree per mcmc() call)
  return MCU_OK;
}

int64_t mcu_kept(int64_t first_iter, int64_t iters, int64_t burnin, int64_t thin) {
  // number of i in (first_iter, first_iter + iters] with i > burnin and (i - burnin) % thin == 0
  if (thin < 1) return 0;
  auto upto = [&](int64_t i) { return i > burnin ? (i - burnin) / thin : 0; };
  return upto(first_iter + iters) - upto(first_iter);
}

int mcu_set_rng_mode(mcu_handle h, int mode, const double* u, size_t n_per_chain) {
  if (!h) return MCU_ERR_ARG;
  CK(cudaSetDevice(h->device));
  cudaFree(h->d_ext); cudaFree(h->d_ext_pos); h->d_ext = nullptr; h->d_ext_pos = nullptr; h->ext_n = 0;
  h->rng_mode = MCU_RNG_PHILOX;
  if (mode == MCU_RNG_PHILOX) return MCU_OK;
  if (mode != MCU_RNG_EXTERNAL || !u || n_per_chain == 0) return fail(h, MCU_ERR_ARG, "bad RNG mode / stream");
  CK(cudaMalloc(&h->d_ext, sizeof(double) * (size_t)h->C * n_per_chain));
  CK(cudaMemcpy(h->d_ext, u, sizeof(double) * (size_t)h->C * n_per_chain, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_ext_pos, sizeof(unsigned long long) * (size_t)h->C));
  CK(cudaMemset(h->d_ext_pos, 0, sizeof(unsigned long long) * (size_t)h->C));
  h->ext_n = n_per_chain; h->rng_mode = MCU_RNG_EXTERNAL;
  return MCU_OK;
}

int mcu_run(mcu_handle h, int64_t iters, int64_t burnin, int64_t thin, double* out, uint32_t flags) {
  if (!h) return MCU_ERR_ARG;
  if (iters < 1) return fail(h, MCU_ERR_ARG, "iters must be positive");
  if (thin < 1) return fail(h, MCU_ERR_ARG, "thin must be positive");
  if (h->iter == 0 && iters <= burnin && !(flags & MCU_RUN_PARTIAL)) return fail(h, MCU_ERR_ARG, "burnin is greater than or equal to iters");   // mcmc.jl:22-23
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "initial values must be set before mcu_run");
  CK(cudaSetDevice(h->device));
  int rc = upload_inputs(h); if (rc) return rc;
  const long long kept = mcu_kept(h->iter, iters, burnin, thin);
  const bool store = out != nullptr || !(flags & MCU_RUN_NO_STORE);
  const size_t C = (size_t)h->C;
  if (store && kept > 0) {
    const size_t need = (size_t)kept * h->P * C;
    if (need > h->samples_cap) {
      cudaFree(h->d_samples); h->d_samples = nullptr; h->samples_cap = 0;
      CK(cudaMalloc(&h->d_samples, sizeof(double) * need));
      h->samples_cap = need;
    }
  }
  h->samples_kept = store ? kept : 0;
  RunArgs a;
  a.n_chains = h->C; a.chain_offset = h->chain_offset; a.seed = h->seed;
  a.burnin = burnin; a.thin = thin;
  a.row0 = h->iter > burnin ? (h->iter - burnin) / thin : 0;
  a.n_blocks = (int)h->h_blocks.size(); a.D = h->D; a.P = h->P; a.blocks = h->d_blocks;
  a.state = h->d_state; a.tune = h->d_tune; a.samples = (store && kept > 0) ? h->d_samples : nullptr;
  a.mom = h->d_mom; a.momn = h->d_momn;
  a.logit_mask = h->logit_mask; a.log_mask = h->log_mask;
  a.comom = (flags & MCU_RUN_MPSRF) ? h->d_comom : nullptr;
  if (kept > 0 && !a.comom) h->comom_ok = false;
  if (!h->d_work) { CK(cudaMalloc(&h->d_work, 2 * sizeof(unsigned long long))); CK(cudaMemset(h->d_work, 0, 2 * sizeof(unsigned long long))); }
  a.work = h->d_work;
  a.ext_u = h->rng_mode == MCU_RNG_EXTERNAL ? h->d_ext : nullptr; a.ext_n = h->ext_n; a.ext_pos = h->d_ext_pos;
  bool fast = h->seeds_fast_ok && !(flags & MCU_RUN_FORCE_GENERIC) && h->rng_mode == MCU_RNG_PHILOX;
  const bool glm_tick = scheme_is_glm_tick(h) && !(flags & MCU_RUN_FORCE_GENERIC) && h->rng_mode == MCU_RNG_PHILOX;
  long long chunk = 256;
  if (const char* e = std::getenv("MCU_CHUNK_ITERS")) { long long v = std::atoll(e); if (v > 0) chunk = v; }
  bool rats_warp = h->rats_warp_ok && !(flags & MCU_RUN_FORCE_GENERIC) && h->rng_mode == MCU_RNG_PHILOX;
  if (rats_warp && h->r_grid == 0) {
    const int g = rats_warp_grid(h->C);
    if (g < 1) return fail(h, MCU_ERR_CUDA, "rats_warp_kernel: cannot size the grid");
    CK(cudaMalloc(&h->r_scratch, rats_warp_scratch_bytes(g)));
    h->r_grid = g;
  }
  bool rats_fast = h->rats_fast_ok && !(flags & MCU_RUN_FORCE_GENERIC) && h->rng_mode == MCU_RNG_PHILOX;
  const bool pumps_fast = h->pumps_fast_ok && !(flags & MCU_RUN_FORCE_GENERIC) && h->rng_mode == MCU_RNG_PHILOX;
  const bool pumps_gibbs = h->pumps_gibbs_ok && !(flags & MCU_RUN_FORCE_GENERIC) && h->rng_mode == MCU_RNG_PHILOX;
  if (fast || rats_warp || rats_fast || pumps_fast || pumps_gibbs) chunk = iters;   // the fused kernels keep everything on chip for the whole call
  CK(cudaEventRecord(h->ev0, h->stream));
  long long done = 0;
  if (glm_tick) {
    if (flags & MCU_RUN_GLM_REFERENCE) h->glm_impl_run = 0; else h->glm_impl_run = 1;
    rc = run_glm_tick(h, iters, burnin, thin, a);
    if (rc) return rc;
    done = iters;
  }
  while (done < iters) {
    const long long n = std::min(chunk, iters - done);
    a.iter0 = h->iter + done; a.iters = n;
    if (fast) {
      rc = seeds_fast_launch(h->inputs["r"].data(), h->inputs["n"].data(), h->inputs["x1"].data(), h->inputs["x2"].data(), a, h->h_blocks.data(),
                             h->h_scales, h->h_SigmaL.empty() || h->h_SigmaL[0].empty() ? nullptr : h->h_SigmaL[0].data(), h->stream);
      if (rc == -2) { fast = false; chunk = 256; continue; }   // design is not 0/1 indicators: the generic kernel takes over
      if (rc) return fail(h, MCU_ERR_CUDA, "seeds_fast launch failed");
    } else if (pumps_gibbs) {
      rc = pumps_gibbs_launch(h->inputs["y"].data(), h->inputs["t"].data(), (int)h->inputs["y"].size(), a, h->h_blocks[2], h->h_scales[2][0], h->stream);
      if (rc) return fail(h, MCU_ERR_CUDA, "pumps_gibbs launch failed");
    } else if (pumps_fast) {
      rc = pumps_fast_launch(h->inputs["y"].data(), h->inputs["t"].data(), (int)h->inputs["y"].size(), a, h->h_scales, h->stream);
      if (rc) return fail(h, MCU_ERR_CUDA, "pumps_fast launch failed");
    } else if (rats_fast) {
      rc = rats_fast_launch(h->inputs["y"].data(), h->inputs["Xm"].data(), h->inputs["rat"].data(), (int)h->inputs["y"].size(),
                            h->inputs["xbar"][0], a, h->h_blocks.data(), h->h_scales, h->stream);
      if (rc == -2) { rats_fast = false; chunk = 256; continue; }   // data are not 5 observations per rat: the generic kernel takes over
      if (rc) return fail(h, MCU_ERR_CUDA, "rats_fast launch failed");
    } else if (rats_warp) {
      rc = rats_warp_launch(h->inputs["y"].data(), h->inputs["Xm"].data(), h->inputs["rat"].data(), (int)h->inputs["y"].size(),
                            h->inputs["xbar"][0], a, h->h_blocks.data(), h->h_scales[1].data(), h->r_grid, h->r_scratch, h->stream);
      if (rc == -2) { rats_warp = false; chunk = 256; continue; }   // data are not 5 observations per rat: the generic kernel takes over
      if (rc) return fail(h, MCU_ERR_CUDA, "rats_warp launch failed");
    } else {
      MCU_DISPATCH(h, launch_run(Host<M>::data(h), a, h->stream));
    }
    h->launches++;
    done += n;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->iter += iters;
  h->pending = true;
  if (flags & MCU_RUN_ASYNC) {
    if (out) return fail(h, MCU_ERR_ARG, "MCU_RUN_ASYNC returns before the samples exist: pass out = NULL and fetch them with mcu_get_samples");
    return MCU_OK;
  }
  rc = mcu_wait(h); if (rc) return rc;
  if (out && kept > 0) return mcu_get_samples(h, out);
  return MCU_OK;
}

// Completes an mcu_run(..., MCU_RUN_ASYNC): blocks until the handle's stream is idle and reports what the run left behind.
int mcu_wait(mcu_handle h) {
  if (!h) return MCU_ERR_ARG;
  if (!h->pending) return MCU_OK;
  CK(cudaSetDevice(h->device));
  h->pending = false;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->last_ms = ms;
  if (h->rng_mode == MCU_RNG_EXTERNAL) {   // a shim stream that ran dry would silently feed constant draws: report it
    const size_t C = (size_t)h->C;
    std::vector<unsigned long long> pos(C);
    CK(cudaMemcpy(pos.data(), h->d_ext_pos, sizeof(unsigned long long) * C, cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < C; ++c) if (pos[c] == ~0ull) return fail(h, MCU_ERR_STATE, "external uniform stream exhausted (chain " + std::to_string(c) + "): supply more draws per chain");
  }
  return MCU_OK;
}

// ModelChains.value of the last mcu_run ([kept x n_monitor x n_chains], column-major): device transpose into the handle's staging
// buffer (no allocation on the hot call) and one D2H copy — give it pinned memory for an asynchronous, full-rate transfer.
int mcu_get_samples(mcu_handle h, double* out) {
  if (!h || !out) return MCU_ERR_ARG;
  int rc = mcu_wait(h); if (rc) return rc;
  if (h->samples_kept < 1 || !h->d_samples) return fail(h, MCU_ERR_STATE, "no stored samples (run without MCU_RUN_NO_STORE)");
  CK(cudaSetDevice(h->device));
  double* d_out = nullptr;
  const size_t total = (size_t)h->samples_kept * h->P * (size_t)h->C;
  rc = stage(h, sizeof(double) * total, (void**)&d_out); if (rc) return rc;
  launch_samples_to_julia(h->d_samples, d_out, h->samples_kept, h->P, h->C, h->stream);
  h->launches++;
  CK(cudaMemcpyAsync(out, d_out, sizeof(double) * total, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return MCU_OK;
}

int mcu_get_state(mcu_handle h, double* values, double* tune, int64_t* iter) {
  if (!h) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no chain state yet");
  CK(cudaSetDevice(h->device));
  const size_t C = (size_t)h->C;
  if (values) {
    double* tmp = nullptr; int rc = stage(h, sizeof(double) * C * h->D, (void**)&tmp); if (rc) return rc;
    launch_soa_to_records(h->d_state, tmp, h->C, h->D, h->stream); h->launches++;
    CK(cudaMemcpyAsync(values, tmp, sizeof(double) * C * h->D, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (tune && h->tune_size > 0) {
    double* tmp = nullptr; int rc = stage(h, sizeof(double) * C * h->tune_size, (void**)&tmp); if (rc) return rc;
    launch_soa_to_records(h->d_tune, tmp, h->C, (int)h->tune_size, h->stream); h->launches++;
    CK(cudaMemcpyAsync(tune, tmp, sizeof(double) * C * h->tune_size, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (iter) *iter = h->iter;
  CK(cudaGetLastError());
  return MCU_OK;
}

int mcu_set_state(mcu_handle h, const double* values, const double* tune, int64_t iter) {
  if (!h || !values) return h ? fail(h, MCU_ERR_ARG, "values is NULL") : MCU_ERR_ARG;
  if (h->h_blocks.empty()) return fail(h, MCU_ERR_STATE, "set the sampling scheme before the state");
  // the tune records are (re)created only at iteration 1 (sampler.jl:40-45): a chain resumed later needs the ones it had
  if (!tune && h->tune_size > 0 && iter > 0) return fail(h, MCU_ERR_ARG, "a state at iteration > 0 needs its sampler tune records");
  CK(cudaSetDevice(h->device));
  int rc = upload_inputs(h); if (rc) return rc;
  rc = ensure_chain_buffers(h); if (rc) return rc;
  const size_t C = (size_t)h->C;
  {
    double* tmp = nullptr; rc = stage(h, sizeof(double) * C * h->D, (void**)&tmp); if (rc) return rc;
    CK(cudaMemcpyAsync(tmp, values, sizeof(double) * C * h->D, cudaMemcpyHostToDevice, h->stream));
    launch_records_to_soa(tmp, h->d_state, h->C, h->D, h->stream); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
  }
  if (tune && h->tune_size > 0) {
    double* tmp = nullptr; rc = stage(h, sizeof(double) * C * h->tune_size, (void**)&tmp); if (rc) return rc;
    CK(cudaMemcpyAsync(tmp, tune, sizeof(double) * C * h->tune_size, cudaMemcpyHostToDevice, h->stream));
    launch_records_to_soa(tmp, h->d_tune, h->C, (int)h->tune_size, h->stream); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
  } else if (h->tune_size > 0) {
    CK(cudaMemsetAsync(h->d_tune, 0, sizeof(double) * C * (size_t)h->tune_size, h->stream));
  }
  // a state set from outside starts a new history: the streaming moments and the stored samples of the previous one are dropped
  CK(cudaMemsetAsync(h->d_mom, 0, sizeof(double) * C * (size_t)h->P * kMomPerCol, h->stream));
  CK(cudaMemsetAsync(h->d_momn, 0, sizeof(double) * C * 3, h->stream));
  if (h->d_comom) CK(cudaMemsetAsync(h->d_comom, 0, sizeof(double) * C * 2 * (size_t)diag_npair(h->P), h->stream));
  if (h->d_ext_pos) CK(cudaMemsetAsync(h->d_ext_pos, 0, sizeof(unsigned long long) * C, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->iter = iter; h->has_inits = true; h->samples_kept = 0; h->comom_ok = h->d_comom != nullptr;
  h->g_reset = true;
  return MCU_OK;
}

static int density_call(mcu_handle h, int block, int grad_mode, int64_t B, const double* state, const double* x, double* lp, double* g) {
  if (!h || !state || B < 1) return h ? fail(h, MCU_ERR_ARG, "bad argument") : MCU_ERR_ARG;
  if (block < 0 || block >= (int)h->h_blocks.size()) return fail(h, MCU_ERR_ARG, "block index out of range");
  CK(cudaSetDevice(h->device));
  int rc = upload_inputs(h); if (rc) return rc;
  const int k = h->h_blocks[block].k, D = h->D;
  DevScratch b_rec, b_state, b_x, b_lp, b_g, b_tmp;
  double *d_x = nullptr, *d_g = nullptr, *d_tmp = nullptr;
  CK(b_rec.alloc(sizeof(double) * B * std::max(D, k)));
  CK(b_state.alloc(sizeof(double) * B * D));
  double *d_rec = b_rec.f64(), *d_state = b_state.f64();
  CK(cudaMemcpyAsync(d_rec, state, sizeof(double) * B * D, cudaMemcpyHostToDevice, h->stream));
  launch_records_to_soa(d_rec, d_state, B, D, h->stream); h->launches++;
  if (x) {
    CK(b_x.alloc(sizeof(double) * B * k)); d_x = b_x.f64();
    CK(cudaMemcpyAsync(d_rec, x, sizeof(double) * B * k, cudaMemcpyHostToDevice, h->stream));
    launch_records_to_soa(d_rec, d_x, B, k, h->stream); h->launches++;
  }
  CK(b_lp.alloc(sizeof(double) * B));
  double* d_lp = b_lp.f64();
  if (g) { CK(b_g.alloc(sizeof(double) * B * k)); CK(b_tmp.alloc(sizeof(double) * B * k)); d_g = b_g.f64(); d_tmp = b_tmp.f64(); }
  MCU_DISPATCH(h, launch_logpdf(Host<M>::data(h), h->d_blocks, block, B, D, d_state, d_x, d_lp, d_g, grad_mode, h->stream));
  h->launches++;
  if (lp) CK(cudaMemcpyAsync(lp, d_lp, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
  if (g) {
    launch_soa_to_records(d_g, d_tmp, B, k, h->stream); h->launches++;
    CK(cudaMemcpyAsync(g, d_tmp, sizeof(double) * B * k, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return MCU_OK;
}

int mcu_factor_counts(mcu_handle h, int* n_param_nodes, int* n_factors) {
  if (!h) return MCU_ERR_ARG;
  int nf = 0;
  MCU_DISPATCH(h, nf = M::NF);
  if (n_param_nodes) *n_param_nodes = tpl_info(h).NN;
  if (n_factors) *n_factors = nf;
  return MCU_OK;
}

int mcu_factor_parents(mcu_handle h, int factor, uint32_t* parent_nodes) {
  if (!h || !parent_nodes) return MCU_ERR_ARG;
  int nf = 0; uint32_t m = 0;
  MCU_DISPATCH(h, nf = M::NF; if (factor >= 0 && factor < nf) m = M::parents(factor));
  if (factor < 0 || factor >= nf) return fail(h, MCU_ERR_ARG, "factor index out of range");
  *parent_nodes = m;
  return MCU_OK;
}

int mcu_logpdf_nodes(mcu_handle h, uint32_t factor_mask, int64_t B, const double* state, double* lp) {
  if (!h || !state || !lp || B < 1) return h ? fail(h, MCU_ERR_ARG, "bad argument") : MCU_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc = upload_inputs(h); if (rc) return rc;
  const int D = h->D;
  DevScratch rec, soa, out;
  CK(rec.alloc(sizeof(double) * B * D));
  CK(soa.alloc(sizeof(double) * B * D));
  CK(out.alloc(sizeof(double) * B));
  double *d_rec = rec.f64(), *d_state = soa.f64(), *d_lp = out.f64();
  CK(cudaMemcpyAsync(d_rec, state, sizeof(double) * B * D, cudaMemcpyHostToDevice, h->stream));
  launch_records_to_soa(d_rec, d_state, B, D, h->stream); h->launches++;
  MCU_DISPATCH(h, launch_factors(Host<M>::data(h), factor_mask, B, D, d_state, d_lp, h->stream));
  h->launches++;
  CK(cudaMemcpyAsync(lp, d_lp, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return MCU_OK;
}

int mcu_predict(mcu_handle h, int64_t B, const double* state, uint32_t stream_id, double* out, int64_t* n_out) {
  if (!h) return MCU_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc = upload_inputs(h); if (rc) return rc;
  int L = 0;
  MCU_DISPATCH(h, L = launch_predict(Host<M>::data(h), 0, h->D, nullptr, 0ull, 0u, nullptr, h->stream));
  if (n_out) *n_out = L;
  if (!out) return MCU_OK;   // size query
  if (!state || B < 1) return fail(h, MCU_ERR_ARG, "bad argument");
  const int D = h->D;
  DevScratch rec, soa, draws, tmp;
  CK(rec.alloc(sizeof(double) * B * D));
  CK(soa.alloc(sizeof(double) * B * D));
  CK(draws.alloc(sizeof(double) * B * L));
  CK(tmp.alloc(sizeof(double) * B * L));
  double *d_rec = rec.f64(), *d_state = soa.f64(), *d_out = draws.f64(), *d_tmp = tmp.f64();
  CK(cudaMemcpyAsync(d_rec, state, sizeof(double) * B * D, cudaMemcpyHostToDevice, h->stream));
  launch_records_to_soa(d_rec, d_state, B, D, h->stream); h->launches++;
  MCU_DISPATCH(h, launch_predict(Host<M>::data(h), B, D, d_state, h->seed, stream_id, d_out, h->stream));
  h->launches++;
  launch_soa_to_records(d_out, d_tmp, B, L, h->stream); h->launches++;
  CK(cudaMemcpyAsync(out, d_tmp, sizeof(double) * B * L, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return MCU_OK;
}

int mcu_logpdf(mcu_handle h, int block, int64_t B, const double* state, const double* x, double* lp) {
  if (!lp) return h ? fail(h, MCU_ERR_ARG, "lp is NULL") : MCU_ERR_ARG;
  return density_call(h, block, 0, B, state, x, lp, nullptr);
}
int mcu_gradlogpdf(mcu_handle h, int block, int grad_mode, int64_t B, const double* state, const double* x, double* lp, double* g) {
  if (!g) return h ? fail(h, MCU_ERR_ARG, "g is NULL") : MCU_ERR_ARG;
  if (grad_mode < 0 || grad_mode > 2) return fail(h, MCU_ERR_ARG, "unknown gradient mode");
  return density_call(h, block, grad_mode, B, state, x, lp, g);
}

// ---- diagnostics ----------------------------------------------------------------------------------
// Diagnostics scratch that lives as long as the handle (no cudaMalloc / cudaFree — each a device-wide sync — on the per-step calls):
// layout [partials: nblk * width doubles | folded sums: width doubles | centres: 2 P doubles | codes: P ints]
struct DiagScratch { double* partial; double* out; double* center; int* codes; };
static int diag_scratch(mcu_ctx* h, long long nblk, int width, DiagScratch* s) {
  const size_t need = sizeof(double) * ((size_t)nblk * width + width + 2 * (size_t)h->P) + sizeof(int) * (size_t)h->P + 64;
  if (need > h->diag_cap) {
    if (h->d_diag) cudaFree(h->d_diag);
    h->d_diag = nullptr; h->diag_cap = 0;
    if (cudaMalloc(&h->d_diag, need) != cudaSuccess) return fail(h, MCU_ERR_CUDA, "cudaMalloc(diagnostics scratch) failed");
    h->diag_cap = need;
  }
  s->partial = static_cast<double*>(h->d_diag);
  s->out = s->partial + (size_t)nblk * width;
  s->center = s->out + width;
  s->codes = reinterpret_cast<int*>(s->center + 2 * (size_t)h->P);
  return MCU_OK;
}
static int reduce_partials(mcu_handle h, const DiagScratch& s, long long nblk, int width, double* host_out) {
  launch_fold(s.partial, nblk, width, s.out, h->stream); h->launches++;
  CK(cudaMemcpyAsync(host_out, s.out, sizeof(double) * width, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return MCU_OK;
}

int mcu_minmax(mcu_handle h, double* minmax) {
  if (!h || !minmax) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no samples yet");
  CK(cudaSetDevice(h->device));
  const long long nblk = grid_for(h->C, 128);
  DiagScratch ds; int rc0 = diag_scratch(h, nblk, h->P * 2, &ds); if (rc0) return rc0;
  launch_minmax_partial(h->d_mom, h->C, h->P, ds.partial, h->stream); h->launches++;
  std::vector<double> part((size_t)nblk * h->P * 2);
  CK(cudaMemcpyAsync(part.data(), ds.partial, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  for (int j = 0; j < h->P; ++j) {
    double mn = INFINITY, mx = -INFINITY;
    for (long long b = 0; b < nblk; ++b) { mn = std::fmin(mn, part[((size_t)b * h->P + j) * 2]); mx = std::fmax(mx, part[((size_t)b * h->P + j) * 2 + 1]); }
    minmax[j * 2] = mn; minmax[j * 2 + 1] = mx;
  }
  return MCU_OK;
}

int mcu_link_codes(mcu_handle h, int transform, const double* minmax, int* codes) {
  if (!h || !codes) return MCU_ERR_ARG;
  TplInfo t = tpl_info(h);
  std::vector<double> mm;
  for (int j = 0; j < h->P; ++j) {
    int c = 0;
    if (transform) {
      if (t.monlink[j] == LINK_LOG) c = 1;
      else if (t.monlink[j] == LINK_HEUR) {
        if (!minmax && mm.empty()) { mm.resize((size_t)h->P * 2); int rc = mcu_minmax(h, mm.data()); if (rc) return rc; }
        const double* q = minmax ? minmax : mm.data();
        if (q[j * 2] > 0.0) {
          c = 1;
          if (q[j * 2 + 1] < 1.0) {   // all values in (0, 1): logit (chains.jl:241-243); the streaming record keeps logit moments for the first 64 columns
            if (j >= 64) return fail(h, MCU_ERR_UNSUPPORTED, "logit link beyond monitored column 64 needs stored samples");
            c = 2;
          }
        }
      }
    }
    codes[j] = c;
  }
  return MCU_OK;
}

int mcu_moments(mcu_handle h, const int* codes, const double* center, double* sums, int64_t* n_kept) {
  if (!h || !sums) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no samples yet");
  CK(cudaSetDevice(h->device));
  const int P = h->P;
  const long long nblk = grid_for(h->C, 128);
  DiagScratch ds; int rc = diag_scratch(h, nblk, P * 7, &ds); if (rc) return rc;
  std::vector<int> zc(P, 0);
  CK(cudaMemcpyAsync(ds.codes, codes ? codes : zc.data(), sizeof(int) * P, cudaMemcpyHostToDevice, h->stream));
  if (center) CK(cudaMemcpyAsync(ds.center, center, sizeof(double) * P * 2, cudaMemcpyHostToDevice, h->stream));
  launch_gelman_partial(h->d_mom, h->d_momn, h->C, P, ds.codes, center ? ds.center : nullptr, ds.partial, h->stream); h->launches++;
  double n0 = 0;
  if (n_kept) CK(cudaMemcpyAsync(&n0, h->d_momn, sizeof(double), cudaMemcpyDeviceToHost, h->stream));   // completes with the sync below
  rc = reduce_partials(h, ds, nblk, P * 7, sums);
  if (rc) return rc;
  if (n_kept) *n_kept = (int64_t)n0;
  return MCU_OK;
}

int mcu_gelman_from_moments(int64_t n_kept, int p, const double* center, const double* sums, double alpha, double* psrf) {
  if (!sums || !psrf || p < 1) return MCU_ERR_ARG;
  for (int j = 0; j < p; ++j) {
    if (sums[j * 7] < 2.0) return MCU_ERR_ARG;   // "less than 2 chains supplied to gelman diagnostic": gelmandiag.jl:6-7
    hostdiag::gelman_column((double)n_kept, center ? center[j * 2] : 0.0, center ? center[j * 2 + 1] : 0.0, sums + j * 7, alpha, psrf + j * 2);
  }
  return MCU_OK;
}

int mcu_gelman(mcu_handle h, double alpha, int transform, double* psrf) {
  if (!h || !psrf) return MCU_ERR_ARG;
  const int P = h->P;
  if (h->C < 2) return fail(h, MCU_ERR_ARG, "less than 2 chains supplied to gelman diagnostic");
  std::vector<int> codes(P); std::vector<double> s0((size_t)P * 7), s1((size_t)P * 7), center((size_t)P * 2);
  int rc = mcu_link_codes(h, transform, nullptr, codes.data()); if (rc) return rc;
  int64_t n = 0;
  rc = mcu_moments(h, codes.data(), nullptr, s0.data(), &n); if (rc) return rc;
  for (int j = 0; j < P; ++j) { center[j * 2] = s0[j * 7 + 1] / s0[j * 7]; center[j * 2 + 1] = s0[j * 7 + 3] / s0[j * 7]; }
  rc = mcu_moments(h, codes.data(), center.data(), s1.data(), &n); if (rc) return rc;
  if (n < 2) return fail(h, MCU_ERR_STATE, "fewer than 2 kept samples per chain");
  return mcu_gelman_from_moments(n, P, center.data(), s1.data(), alpha, psrf);
}

int mcu_summary_sums(mcu_handle h, const double* center, double* sums) {
  if (!h || !sums) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no samples yet");
  CK(cudaSetDevice(h->device));
  const int P = h->P;
  const long long nblk = grid_for(h->C, 128);
  DiagScratch ds; int rc = diag_scratch(h, nblk, P * 8, &ds); if (rc) return rc;
  if (center) CK(cudaMemcpyAsync(ds.center, center, sizeof(double) * P * 2, cudaMemcpyHostToDevice, h->stream));
  launch_summary_partial(h->d_mom, h->d_momn, h->C, P, center ? ds.center : nullptr, ds.partial, h->stream); h->launches++;
  return reduce_partials(h, ds, nblk, P * 8, sums);
}

int mcu_summary_from_sums(int64_t n_kept, int p, const double* center, const double* sums, double* out) {
  if (!sums || !out || !center) return MCU_ERR_ARG;
  for (int j = 0; j < p; ++j) hostdiag::summary_column((double)n_kept, center[j * 2], center[j * 2 + 1], sums + j * 8, out + j * 5);
  return MCU_OK;
}

int mcu_summary_streaming(mcu_handle h, double* out) {
  if (!h || !out) return MCU_ERR_ARG;
  const int P = h->P;
  std::vector<double> s0((size_t)P * 8), s1((size_t)P * 8), center((size_t)P * 2);
  int rc = mcu_summary_sums(h, nullptr, s0.data()); if (rc) return rc;
  for (int j = 0; j < P; ++j) { center[j * 2] = s0[j * 8 + 1] / s0[j * 8]; center[j * 2 + 1] = s0[j * 8 + 4] > 0 ? s0[j * 8 + 5] / s0[j * 8 + 4] : 0.0; }
  rc = mcu_summary_sums(h, center.data(), s1.data()); if (rc) return rc;
  double n0 = 0; CK(cudaMemcpy(&n0, h->d_momn, sizeof(double), cudaMemcpyDeviceToHost));
  return mcu_summary_from_sums((int64_t)n0, P, center.data(), s1.data(), out);
}

int mcu_summarystats(mcu_handle h, int etype, int batch_size, double* out) {
  if (!h || !out) return MCU_ERR_ARG;
  if (h->samples_kept < 1 || !h->d_samples) return fail(h, MCU_ERR_STATE, "no stored samples (run without MCU_RUN_NO_STORE)");
  if (etype != MCU_ETYPE_BM && etype != MCU_ETYPE_IMSE) return fail(h, MCU_ERR_ARG, "unsupported mcse method");   // mcse.jl:3-8
  CK(cudaSetDevice(h->device));
  const size_t total = (size_t)h->samples_kept * h->P * (size_t)h->C;
  std::vector<double> smp(total);
  CK(cudaMemcpy(smp.data(), h->d_samples, sizeof(double) * total, cudaMemcpyDeviceToHost));
  if (batch_size < 1) batch_size = 100;
  const int rc = hostdiag::summarystats_soa(smp.data(), h->samples_kept, h->P, h->C, etype, batch_size, out);
  if (rc) return fail(h, MCU_ERR_ARG, "iterations are < 2 * batch size");   // mcse.jl:13-16
  return MCU_OK;
}

int mcu_chains_quantile(const double* value, int64_t n, int p, int64_t m, const double* q, int nq, double* out) {
  if (!value || !q || !out || n < 1 || p < 1 || m < 1 || nq < 1) return MCU_ERR_ARG;
  hostdiag::chains_quantile(value, n, p, m, q, nq, out);
  return MCU_OK;
}
int mcu_chains_hpd(const double* value, int64_t n, int p, int64_t m, double alpha, double* out) {
  if (!value || !out || n < 1 || p < 1 || m < 1 || !(alpha > 0.0 && alpha < 1.0)) return MCU_ERR_ARG;
  hostdiag::chains_hpd(value, n, p, m, alpha, out);
  return MCU_OK;
}
int mcu_chains_autocor(const double* value, int64_t n, int p, int64_t m, const int64_t* lags, int nlags, double* out) {
  if (!value || !lags || !out || n < 2 || p < 1 || m < 1 || nlags < 1) return MCU_ERR_ARG;
  std::vector<long long> lg(lags, lags + nlags);
  hostdiag::chains_autocor(value, n, p, m, lg.data(), nlags, out);
  return MCU_OK;
}
int mcu_chains_changerate(const double* value, int64_t n, int p, int64_t m, double* out) {
  if (!value || !out || n < 2 || p < 1 || m < 1) return MCU_ERR_ARG;
  hostdiag::chains_changerate(value, n, p, m, out);
  return MCU_OK;
}
int mcu_chains_gelman(const double* value, int64_t n, int p, int64_t m, double alpha, const int* codes, int mpsrf, double* out) {
  if (!value || !out || n < 2 || p < 1) return MCU_ERR_ARG;
  return hostdiag::chains_gelman(value, n, p, m, alpha, codes, mpsrf != 0, out) ? MCU_ERR_ARG : MCU_OK;   // "less than 2 chains": gelmandiag.jl:6-7
}

int mcu_chains_geweke(const double* value, int64_t n, int p, int64_t m, double first, double last, int etype, int batch_size, double* out) {
  if (!value || !out || n < 4 || p < 1 || m < 1 || etype < 0 || etype > 2) return MCU_ERR_ARG;
  if (!(first > 0.0 && first < 1.0) || !(last > 0.0 && last < 1.0) || first + last > 1.0) return MCU_ERR_ARG;   // gewekediag.jl:5-11
  if (batch_size < 1) batch_size = 100;
  return hostdiag::chains_series(value, n, p, m, 2, out, [&](const double* x, double* r) { return hostdiag::geweke_vec(x, n, first, last, etype, batch_size, r); }) ? MCU_ERR_ARG : MCU_OK;
}
int mcu_chains_heidel(const double* value, int64_t n, int p, int64_t m, double alpha, double eps, int etype, int batch_size, int64_t start, double* out) {
  if (!value || !out || n < 4 || p < 1 || m < 1 || etype < 0 || etype > 2 || !(alpha > 0.0 && alpha < 1.0)) return MCU_ERR_ARG;
  if (batch_size < 1) batch_size = 100;
  return hostdiag::chains_series(value, n, p, m, 6, out, [&](const double* x, double* r) { return hostdiag::heidel_vec(x, n, alpha, eps, etype, batch_size, start, r); }) ? MCU_ERR_ARG : MCU_OK;
}
int mcu_chains_raftery(const double* value, int64_t n, int p, int64_t m, double q, double r, double s, double eps, int64_t range_start, int64_t range_step, double* out) {
  if (!value || !out || n < 3 || p < 1 || m < 1 || !(q > 0.0 && q < 1.0) || !(r > 0.0) || !(s > 0.0 && s < 1.0) || range_step < 1) return MCU_ERR_ARG;
  return hostdiag::chains_series(value, n, p, m, 5, out, [&](const double* x, double* rr) { hostdiag::raftery_vec(x, n, q, r, s, eps, range_start, range_step, rr); return 0; }) ? MCU_ERR_ARG : MCU_OK;
}

int mcu_chains_summarystats(const double* value, int64_t n, int p, int64_t m, int etype, int batch_size, double* out) {
  if (!value || !out || n < 1 || p < 1 || m < 1 || etype < MCU_ETYPE_BM || etype > MCU_ETYPE_IPSE) return MCU_ERR_ARG;
  if (batch_size < 1) batch_size = 100;
  return hostdiag::chains_summarystats(value, n, p, m, etype, batch_size, out) ? MCU_ERR_ARG : MCU_OK;
}

// ---- cross-handle diagnostics: the packed two-round protocol (diagproto.hpp) ---------------------------------------------------------
int mcu_diag_sizes(int p, int* n_round1, int* n_round2) {
  if (p < 1) return MCU_ERR_ARG;
  if (n_round1) *n_round1 = p * kDiag1;
  if (n_round2) *n_round2 = p * kDiag2 + 2 * diag_npair(p);
  return MCU_OK;
}
int mcu_monitor_links(mcu_handle h, int* monlink) {
  if (!h || !monlink) return MCU_ERR_ARG;
  if (h->h_monlink.empty()) return fail(h, MCU_ERR_STATE, "set the sampling scheme first");
  for (int j = 0; j < h->P; ++j) monlink[j] = h->h_monlink[j];
  return MCU_OK;
}
int mcu_diag_round1(mcu_handle h, double* buf) {
  if (!h || !buf) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no samples yet");
  CK(cudaSetDevice(h->device));
  DiagBufs b; int rc = diag_bufs(h, &b); if (rc) return rc;
  launch_diag1(h->d_mom, h->d_momn, h->C, h->P, h->logit_mask, b.partial, b.r1, h->stream); h->launches += 2;
  CK(cudaMemcpyAsync(buf, b.r1, sizeof(double) * h->P * kDiag1, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return MCU_OK;
}
int mcu_diag_round2(mcu_handle h, int transform, const double* reduced1, double* buf2) {
  if (!h || !reduced1 || !buf2) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no samples yet");
  CK(cudaSetDevice(h->device));
  DiagBufs b; int rc = diag_bufs(h, &b); if (rc) return rc;
  CK(cudaMemcpyAsync(b.r1, reduced1, sizeof(double) * h->P * kDiag1, cudaMemcpyHostToDevice, h->stream));
  launch_diag_plan(b.r1, h->d_monlink, transform, h->P, b.plan, h->stream);
  launch_diag2(h->d_mom, h->d_momn, h->C, h->P, b.plan, b.partial, b.r2, h->stream); h->launches += 3;
  const int npair = diag_npair(h->P);   // the protocol buffer always carries the pair slots (p <= 12); NaN when the co-moments were not streamed
  if (npair > 0 && h->d_comom && h->comom_ok) { launch_diag_pairs(h->d_mom, h->d_momn, h->d_comom, h->C, h->P, b.plan, transform ? 1 : 0, b.partial, b.r2 + (size_t)h->P * kDiag2, h->stream); h->launches += 2; }
  else if (npair > 0) { std::vector<double> nanv(2 * (size_t)npair, NAN); CK(cudaMemcpyAsync(b.r2 + (size_t)h->P * kDiag2, nanv.data(), sizeof(double) * nanv.size(), cudaMemcpyHostToDevice, h->stream)); CK(cudaStreamSynchronize(h->stream)); }
  CK(cudaMemcpyAsync(buf2, b.r2, sizeof(double) * ((size_t)h->P * kDiag2 + 2 * (size_t)npair), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return MCU_OK;
}
int mcu_diag_finish(int64_t n_kept, int p, double alpha, const int* monlink, int transform, const double* reduced1, const double* reduced2,
                    double* psrf, double* summary, int* codes, double* mpsrf) {
  if (p < 1 || !monlink || !reduced1 || !reduced2) return MCU_ERR_ARG;
  return diag_finish_host(n_kept, p, alpha, monlink, transform, reduced1, reduced2, psrf, summary, codes, mpsrf);
}
int mcu_n_kept(mcu_handle h, int64_t* n_kept) {
  if (!h || !n_kept) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no samples yet");
  CK(cudaSetDevice(h->device));
  double n0 = 0; CK(cudaMemcpy(&n0, h->d_momn, sizeof(double), cudaMemcpyDeviceToHost));
  *n_kept = (int64_t)n0;
  return MCU_OK;
}

// ---- built-in transport: NCCL over NVLink (one rank per handle; replaces the gather of pmap2(mcmc_worker!, ...), mcmc.jl:48-59) ------
int mcu_comm_unique_id(mcu_nccl_id* id) {
  if (!id) return MCU_ERR_ARG;
  NcclApi* n = nccl_api();
  if (!n->lib) { g_create_err = n->err; return MCU_ERR_UNSUPPORTED; }
  return n->GetUniqueId(id) == 0 ? MCU_OK : MCU_ERR_CUDA;
}
int mcu_comm_init(mcu_handle h, int rank, int nranks, const mcu_nccl_id* id) {
  if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return h ? fail(h, MCU_ERR_ARG, "bad rank / nranks / id") : MCU_ERR_ARG;
  NcclApi* n = nccl_api();
  if (!n->lib) return fail(h, MCU_ERR_UNSUPPORTED, n->err);
  CK(cudaSetDevice(h->device));
  if (h->comm) { n->CommDestroy(h->comm); h->comm = nullptr; }
  NK(n->CommInitRank(&h->comm, nranks, *id, rank));
  h->comm_rank = rank; h->comm_nranks = nranks;
  return MCU_OK;
}
int mcu_comm_size(mcu_handle h, int* rank, int* nranks) {
  if (!h) return MCU_ERR_ARG;
  if (rank) *rank = h->comm_rank;
  if (nranks) *nranks = h->comm ? h->comm_nranks : 1;
  return MCU_OK;
}
// gelmandiag + summarystats over the chains of EVERY rank of the communicator (of this handle alone without one): both protocol rounds
// stay on the device — reductions, NCCL all-reduces (round 1: MIN / MAX / SUM in one group; round 2: SUM) and the plan kernel are queued
// on the handle's stream back to back; one synchronisation, then O(p) host arithmetic (F quantile etc., hostdiag.hpp).
int mcu_diag_global(mcu_handle h, double alpha, int transform, double* psrf, double* summary, int* codes, double* mpsrf) {
  if (!h || (!psrf && !summary && !codes && !mpsrf)) return MCU_ERR_ARG;
  if (!h->has_inits) return fail(h, MCU_ERR_STATE, "no samples yet");
  CK(cudaSetDevice(h->device));
  const int P = h->P;
  DiagBufs b; int rc = diag_bufs(h, &b); if (rc) return rc;
  launch_diag1(h->d_mom, h->d_momn, h->C, P, h->logit_mask, b.partial, b.r1, h->stream); h->launches += 2;
  NcclApi* n = h->comm ? nccl_api() : nullptr;
  if (n) {
    NK(n->GroupStart());
    NK(n->AllReduce(b.r1, b.r1, (size_t)P, kNcclFloat64, kNcclMin, h->comm, h->stream));
    NK(n->AllReduce(b.r1 + P, b.r1 + P, (size_t)P, kNcclFloat64, kNcclMax, h->comm, h->stream));
    NK(n->AllReduce(b.r1 + 2 * P, b.r1 + 2 * P, (size_t)P * kDiagSum1, kNcclFloat64, kNcclSum, h->comm, h->stream));
    NK(n->GroupEnd());
  }
  launch_diag_plan(b.r1, h->d_monlink, transform, P, b.plan, h->stream);
  launch_diag2(h->d_mom, h->d_momn, h->C, P, b.plan, b.partial, b.r2, h->stream); h->launches += 3;
  const int npair = (h->d_comom && h->comom_ok && mpsrf) ? diag_npair(P) : 0;   // every rank of a communicator must agree on this (same flags on every rank)
  if (npair > 0) { launch_diag_pairs(h->d_mom, h->d_momn, h->d_comom, h->C, P, b.plan, transform ? 1 : 0, b.partial, b.r2 + (size_t)P * kDiag2, h->stream); h->launches += 2; }
  const size_t n2 = (size_t)P * kDiag2 + 2 * (size_t)npair;
  if (n) NK(n->AllReduce(b.r2, b.r2, n2, kNcclFloat64, kNcclSum, h->comm, h->stream));
  std::vector<double> host((size_t)P * kDiag1 + n2 + 1);
  CK(cudaMemcpyAsync(host.data(), b.r1, sizeof(double) * P * kDiag1, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(host.data() + (size_t)P * kDiag1, b.r2, sizeof(double) * n2, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(host.data() + (size_t)P * kDiag1 + n2, h->d_momn, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  const int64_t n_kept = (int64_t)host[(size_t)P * kDiag1 + n2];
  if (psrf && n_kept < 2) return fail(h, MCU_ERR_STATE, "fewer than 2 kept samples per chain");
  rc = diag_finish_host(n_kept, P, alpha, h->h_monlink.data(), transform, host.data(), host.data() + (size_t)P * kDiag1, psrf, summary, codes, npair > 0 ? mpsrf : nullptr);
  if (mpsrf && npair == 0) *mpsrf = NAN;
  if (rc == MCU_ERR_UNSUPPORTED) return fail(h, rc, "logit link beyond monitored column 64 needs stored samples");
  if (rc == MCU_ERR_ARG) return fail(h, rc, "less than 2 chains supplied to gelman diagnostic");
  return rc;
}

double mcu_fp64_peak_tflops(mcu_handle h) {
  if (!h) return -1.0;
  if (cudaSetDevice(h->device) != cudaSuccess) return -1.0;
  h->launches += 6;
  return measure_fp64_peak_tflops(h->stream);
}
int mcu_glm_gradient(mcu_handle h, int impl, const double* beta, double* lp, double* grad) {
  if (!h || !beta || !lp || !grad) return h ? fail(h, MCU_ERR_ARG, "NULL argument") : MCU_ERR_ARG;
  if (h->tpl != MCU_TPL_GLM_LOGIT || h->glm_d == 0) return fail(h, MCU_ERR_STATE, "GLM template with inputs X, y required");
  if (impl != 0 && impl != 1) return fail(h, MCU_ERR_ARG, "impl must be 0 (reference) or 1 (tensor core)");
  CK(cudaSetDevice(h->device));
  int rc = upload_inputs(h); if (rc) return rc;
  rc = ensure_glm_buffers(h); if (rc) return rc;
  const size_t C = (size_t)h->C; const int d = h->D; const int N = (int)h->inputs["y"].size();
  double* tmp = nullptr; CK(cudaMalloc(&tmp, sizeof(double) * C * d));
  CK(cudaMemcpyAsync(tmp, beta, sizeof(double) * C * d, cudaMemcpyHostToDevice, h->stream));
  launch_records_to_soa(tmp, h->g_req, h->C, d, h->stream); h->launches++;
  const int saved = h->glm_impl; h->glm_impl = impl; h->g_pass_C = 0;
  CK(cudaEventRecord(h->ev0, h->stream));
  rc = glm_gradient_dispatch(h, N);
  CK(cudaEventRecord(h->ev1, h->stream));
  h->glm_impl = saved;
  if (rc) { cudaFree(tmp); return rc; }
  launch_soa_to_records(h->g_grad, tmp, h->C, d, h->stream); h->launches++;
  CK(cudaMemcpyAsync(grad, tmp, sizeof(double) * C * d, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(lp, h->g_lp, sizeof(double) * C, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(tmp);
  CK(cudaGetLastError());
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1)); h->last_ms = ms;
  return MCU_OK;
}

int64_t mcu_launch_count(mcu_handle h) { return h ? h->launches : 0; }
int mcu_work_count(mcu_handle h, uint64_t* out) {
  if (!h || !out) return MCU_ERR_ARG;
  unsigned long long w[2] = {0, 0};
  if (h->d_work) { CK(cudaSetDevice(h->device)); CK(cudaMemcpy(w, h->d_work, sizeof(w), cudaMemcpyDeviceToHost)); }
  out[0] = w[0]; out[1] = (uint64_t)h->ticks; out[2] = h->pass_slots; out[3] = w[1];
  return MCU_OK;
}
double mcu_last_kernel_ms(mcu_handle h) { return h ? h->last_ms : 0.0; }

}  // extern "C"
