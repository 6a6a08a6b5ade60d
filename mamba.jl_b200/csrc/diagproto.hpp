// diagproto.hpp — the packed two-round protocol behind cross-chain diagnostics (gelmandiag + summarystats) over chains that
// live on several handles / GPUs.  The reference farms chains out to workers and gathers the whole sample array
// (src/model/mcmc.jl:48-59) before src/output/gelmandiag.jl:3-60 and src/output/stats.jl:85-94 run on it; here each handle reduces its
// own chains' streaming moments on the device and only O(p) doubles per round cross handle boundaries:
//
//   round 1  [ min (p) | max (p) | sum1 (9 p) ]   all-reduced with MIN / MAX / SUM.  Per column j:
//            sum1[j] = { m, Σ ψ̄, Σ s², Σ ψ̄_log, Σ s²_log, Σ ψ̄_logit, Σ s²_logit, Σ nb, Σ nb·bmean }
//            (ψ̄ / s² = a chain's mean / variance of the column on the identity, log and logit scale; nb / bmean = its number of
//            complete batches of 100 and their mean).  From the reduced buffer every handle derives the SAME plan: the link code
//            of link(c) (src/output/modelchains.jl:57-76, chains.jl:237-246 — the heuristic needs the global min / max) and the
//            centres of round 2.
//   round 2  sum2 (15 p), all-reduced with SUM.  Per column j:
//            { m, Σd, Σd², Σe, Σe², Σed, Σed² }  with d = ψ̄ - c1, e = s² - c2 on the planned scale   (gelmandiag.jl:12-29)
//            { C, Σ mean, Σ M2, Σ (mean - k1)², Σ nb, Σ nb·bmean, Σ bM2, Σ nb (bmean - k2)² }          (stats.jl:85-94, mcse.jl:10-19)
//            followed, when p <= 12, by two sums per column pair i < j (p (p - 1) / 2 pairs, in the order (0,1), (0,2), ..., (1,2), ...):
//            { Σ cov_k(i, j), Σ d_i d_j }  — the off-diagonals of W = mean of the within-chain covariances and of B / n = cov of the chain
//            means, for the multivariate PSRF (gelmandiag.jl:49-55) without the draws.  The within-chain co-moments are streamed on the raw
//            scale and on the nodes' own link scale; a column whose link(c) is resolved by the data-dependent heuristic to log / logit has
//            no streamed co-moments on that scale, and the MPSRF is then reported as NaN (it needs the stored draws).
//   The centring keeps the variances free of cancellation; any transport may carry the buffers (NCCL inside libmambacuda,
//   torch.distributed / gloo in the CPU tests, Julia's own worker messaging).
#pragma once

#if defined(__CUDACC__)
#define MCU_PROTO_HD __host__ __device__ inline
#else
#define MCU_PROTO_HD inline
#endif

namespace mcu {

constexpr int kDiag1 = 11;       // values per column produced by a handle in round 1: min, max, 9 sums
constexpr int kDiagSum1 = 9;
constexpr int kDiag2 = 15;       // values per column in round 2
constexpr int kDiagCoMaxP = 12;  // == kCoMaxP of engine.cuh
MCU_PROTO_HD int diag_npair(int p) { return (p > 1 && p <= kDiagCoMaxP) ? p * (p - 1) / 2 : 0; }
constexpr int kPlanLinkLog = 1, kPlanLinkHeur = -1;   // == LINK_LOG / LINK_HEUR of models.cuh

// link code (0 identity, 1 log, 2 logit) and round-2 centres { c1, c2, k1, k2 } of column j from the all-reduced round-1 buffer
MCU_PROTO_HD int diag_plan_column(int p, int j, int monlink, int transform, const double* r1, double* ctr) {
  const double mn = r1[j], mx = r1[p + j];
  const double* s = r1 + 2 * p + (long long)j * kDiagSum1;
  int code = 0;
  if (transform) {
    if (monlink == kPlanLinkLog) code = 1;
    else if (monlink == kPlanLinkHeur && mn > 0.0) code = mx < 1.0 ? 2 : 1;    // chains.jl:239-243
  }
  const double m = s[0];
  ctr[0] = s[1 + 2 * code] / m;
  ctr[1] = s[2 + 2 * code] / m;
  ctr[2] = s[1] / m;
  ctr[3] = s[7] > 0.0 ? s[8] / s[7] : 0.0;
  return code;
}

}  // namespace mcu
