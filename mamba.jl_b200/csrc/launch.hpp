// launch.hpp — host-callable launch wrappers; each model template's kernels live in their own
// translation unit (tpl_*.cu) so the library builds in parallel.
#pragma once
#include <vector>

#include "engine.cuh"

namespace mcu {

constexpr int kGlmDMax = 128;
typedef GlmModel<kGlmDMax> GlmM;

#define MCU_DECLARE_TPL(M)                                                                                   \
  void launch_run(const M::Data& d, const RunArgs& a, cudaStream_t st);                                      \
  void launch_logpdf(const M::Data& d, const DevBlock* blocks, int block, long long B, int D,                \
                     const double* state, const double* x, double* lp, double* g, int grad_mode, cudaStream_t st);   \
  void launch_factors(const M::Data& d, unsigned mask, long long B, int D, const double* state, double* lp, cudaStream_t st); \
  int launch_predict(const M::Data& d, long long B, int D, const double* state, unsigned long long seed, unsigned stream_id,   \
                     double* out, cudaStream_t st);
MCU_DECLARE_TPL(LineModel)
MCU_DECLARE_TPL(SeedsModel)
MCU_DECLARE_TPL(RatsModel)
MCU_DECLARE_TPL(PumpsModel)
MCU_DECLARE_TPL(GlmM)
MCU_DECLARE_TPL(SurgicalModel)
MCU_DECLARE_TPL(DyesModel)
MCU_DECLARE_TPL(SalmModel)
MCU_DECLARE_TPL(EquivModel)
MCU_DECLARE_TPL(BlockerModel)
MCU_DECLARE_TPL(StacksModel)
MCU_DECLARE_TPL(MagnesiumModel)
MCU_DECLARE_TPL(OxfordModel)
MCU_DECLARE_TPL(EpilModel)

// fewer chains than one wave at the default occupancy (148 SMs x 4 blocks x 128 threads): the low-latency instantiation
#ifdef MCU_GENERIC_MINB
#define MCU_LAUNCH_GENERIC(M)                                                                                \
    const unsigned grid_ = (unsigned)((a.n_chains + 127) / 128);                                             \
    if (a.n_chains >= 148LL * 4 * 128) run_generic_kernel_dense<M><<<grid_, 128, 0, st>>>(d, a);             \
    else run_generic_kernel<M><<<grid_, 128, 0, st>>>(d, a);
#else
#define MCU_LAUNCH_GENERIC(M) run_generic_kernel<M><<<(unsigned)((a.n_chains + 127) / 128), 128, 0, st>>>(d, a);
#endif

#define MCU_DEFINE_TPL(M)                                                                                    \
  void launch_run(const M::Data& d, const RunArgs& a, cudaStream_t st) {                                     \
    MCU_LAUNCH_GENERIC(M)                                                                                    \
  }                                                                                                          \
  void launch_logpdf(const M::Data& d, const DevBlock* blocks, int block, long long B, int D,                \
                     const double* state, const double* x, double* lp, double* g, int grad_mode, cudaStream_t st) { \
    logpdf_kernel<M><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(d, blocks, block, B, D, state, x, lp, g, grad_mode); \
  }                                                                                                          \
  void launch_factors(const M::Data& d, unsigned mask, long long B, int D, const double* state, double* lp, cudaStream_t st) { \
    factors_kernel<M><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(d, mask, B, D, state, lp);               \
  }                                                                                                          \
  int launch_predict(const M::Data& d, long long B, int D, const double* state, unsigned long long seed, unsigned stream_id,   \
                     double* out, cudaStream_t st) {                                                         \
    if (out) predict_kernel<M><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(d, B, D, state, seed, stream_id, out); \
    return M::out_len(d);                                                                                    \
  }

// misc kernels (kern_misc.cu)
void launch_init(long long n_chains, long long chain_offset, unsigned long long seed, int D, const double* inits,
                 long long n_inits, const int* elink, const double* ebound, double jitter_sd, double* state, cudaStream_t st);
void launch_soa_to_records(const double* soa, double* rec, long long C, int rows, cudaStream_t st);
void launch_records_to_soa(const double* rec, double* soa, long long C, int rows, cudaStream_t st);
void launch_samples_to_julia(const double* smp, double* out, long long kept, int P, long long C, cudaStream_t st);
void launch_gelman_partial(const double* mom, const double* momn, long long C, int P, const int* use_log,
                           const double* center, double* partial, cudaStream_t st);
void launch_fold(const double* partial, long long nblocks, int width, double* out, cudaStream_t st);
void launch_minmax_partial(const double* mom, long long C, int P, double* partial, cudaStream_t st);
void launch_summary_partial(const double* mom, const double* momn, long long C, int P, const double* center,
                            double* partial, cudaStream_t st);

// packed two-round diagnostics protocol (diagproto.hpp): round-1 reductions into [min P | max P | sum 9P], plan, round-2 reductions (15 P)
void launch_diag1(const double* mom, const double* momn, long long C, int P, unsigned long long logit_mask, double* partial, double* out, cudaStream_t st);
void launch_diag_plan(const double* r1, const int* monlink, int transform, int P, double* plan, cudaStream_t st);
void launch_diag2(const double* mom, const double* momn, long long C, int P, const double* plan, double* partial, double* out, cudaStream_t st);
void launch_diag_pairs(const double* mom, const double* momn, const double* comom, long long C, int P, const double* plan, int set, double* partial, double* out,
                       cudaStream_t st);

double measure_fp64_peak_tflops(cudaStream_t st);

// fused seeds/AMWG kernel (seeds_fast.cu); returns 0 on success
int seeds_fast_launch(const double* r, const double* n, const double* x1, const double* x2, const RunArgs& a, const DevBlock* h_blocks,
                      const std::vector<std::vector<double>>& h_scales, const double* h_SigmaL, cudaStream_t st);


// fused pumps kernel for the reference's Slice scheme (pumps_fast.cu); returns 0 on success
int pumps_fast_launch(const double* y, const double* t, int N, const RunArgs& a, const std::vector<std::vector<double>>& h_scales, cudaStream_t st);
// fused pumps kernel for [Gibbs(theta), Gibbs(beta), AMWG(alpha)] (BASELINE.json configs[4]); `amwg` is the host copy of block 2
int pumps_gibbs_launch(const double* y, const double* t, int N, const RunArgs& a, const DevBlock& amwg, double scale, cudaStream_t st);

// fused rats kernel for the reference's Slice + AMWG scheme (rats_fast.cu); returns 0 on success, -2 if the data layout is not 30 x 5
int rats_fast_launch(const double* y, const double* Xm, const double* rat, int N, double xbar, const RunArgs& a, const DevBlock* h_blocks,
                     const std::vector<std::vector<double>>& h_scales, cudaStream_t st);

// warp-per-chain rats kernel, NUTS + Slice (rats_warp.cu); launch returns 0 on success
int rats_warp_grid(long long n_chains);
size_t rats_warp_scratch_bytes(int grid);
int rats_warp_launch(const double* y, const double* Xm, const double* rat, int N, double xbar, const RunArgs& a, const DevBlock* h_blocks,
                     const double* h_width, int grid, double* scratch, cudaStream_t st);

// ---- GLM / NUTS tick engine (glm_nuts.cu) ----------------------------------------------------------------
struct GlmTick {
  long long C, chain_offset;
  unsigned long long seed;
  long long target_iter, burnin, thin, row0;
  int d, max_depth;
  double target, eps_desc;
  double* state; double* tune; double* sc; double* vec; double* req; const double* lp; const double* grad;
  double* samples; double* mom; double* momn;
  int* n_active;   // [2]
  unsigned long long* work;
  int tick;
};
size_t glm_tick_scalar_slots();
size_t glm_tick_vector_slots();
void glm_advance(const GlmTick& t, cudaStream_t st);
void glm_grad_reference(const double* X, const double* y, int N, int d, long long C, const double* req, int nslab,
                        double* part_lp, double* part_g, double* lp, double* grad, int family, double sigma, cudaStream_t st);

// ---- tensor-core GLM likelihood/gradient kernel (glm_tc.cu) ---------------------------------------------
size_t glm_tc_tile_bytes(int d);
long long glm_tc_num_tiles(long long N);
int glm_tc_nsub(long long N, int nslab);
void glm_tc_pack(const double* X, const double* y, int N, int d, int family, unsigned char* blob, cudaStream_t st, const double* colscale = nullptr);
// C = chains of the pass; with a compaction map (slot k = chain map[k]) req / lp / grad keep the handle's stride Cfull
int glm_tc_launch(const unsigned char* blob, int N, int d, long long C, const double* req, int nslab,
                  double* part_lp, float* part_g, int family, double sigma, cudaStream_t st, const int* map = nullptr, long long Cfull = 0,
                  const double* col_inv = nullptr);
void glm_fold_tc(const double* part_lp, const float* part_g, int nslab_lp, int nslab_g, int d, long long C, const double* req,
                 const double* xty, double lp_const, double* lp, double* grad, cudaStream_t st, const int* map = nullptr, long long Cfull = 0,
                 const double* col_inv = nullptr);
void glm_compact(const double* sc, long long C, int* map, int* count, cudaStream_t st);
void glm_fold(const double* part_lp, const double* part_g, int nslab_lp, int nslab_g, int d, long long C, double* lp, double* grad, cudaStream_t st);

}  // namespace mcu
