// models.cuh — the fixed library of model templates, compiled to device functions.
//
// The reference interprets its density: every logpdf! re-runs node closures that allocate
// Distribution objects (src/model/dependent.jl:176-179, src/model/simulation.jl:77-90).  Here each
// template is a struct of static device functions over a flat state record `s[D]` holding the
// unobserved stochastic elements on the CONSTRAINED scale:
//
//   factor(f)      log density of stochastic node f given its parents (logpdf_sub semantics:
//                  -Inf outside the support, src/distributions/distributionstruct.jl:138-158; plus the
//                  log-Jacobian of the node's link when `transform`,
//                  src/distributions/transformdistribution.jl:66-78)
//   parents(f)     bitmask of the parameter nodes factor f depends on through Logical nodes — the
//                  transpose of the reference's `targets` (src/model/model.jl:17-25,
//                  src/model/graph.jl:93-103)
//   joint_grad     analytic gradient of the sum of all factors w.r.t. every state element
//                  (replaces Calculus.gradient finite differences, src/model/simulation.jl:47-51)
//   monitor        unlist(m, true): the monitored columns (src/model/simulation.jl:114-121)
//
// Factors 0..NN-1 are the parameter nodes' own densities (factor f belongs to node f); factors
// NN.. are observed (data) nodes.  Factor order is topological and fixes summation order.
// Templates: line (doc/tutorial/line.jl:5-25), seeds (doc/examples/seeds.jl:16-56),
// rats (doc/examples/rats.jl:49-97), pumps (doc/examples/pumps.jl:12-39), glm (synthetic).
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fastfn.cuh"

namespace mcu {

#define MCU_HD __host__ __device__ __forceinline__
#define MCU_D __device__ __forceinline__
#define MCU_NOINL __device__ __noinline__

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
// link codes of a state element (src/distributions/transformdistribution.jl): identity (:53-61), log (lower bound 0: :66-78),
// two-sided logit((x - a) / (b - a)) with log-Jacobian log((x - a)(b - x) / (b - a)) (:6-48; the unit interval of :83-93 is (a, b) = (0, 1));
// LINK_HEUR marks a monitored Logical column whose link(c) is data-dependent (src/output/chains.jl:237-246)
constexpr int LINK_IDENT = 0, LINK_LOG = 1, LINK_BOUNDED = 2, LINK_HEUR = -1;
constexpr int OUT_NORMAL = 0, OUT_BINOMIAL = 1, OUT_POISSON = 2, OUT_BERNOULLI = 3, OUT_LAPLACE = 4;

MCU_D double neg_inf() { return -CUDART_INF; }

// log / exp of the density formulas: the constant-bank fdlibm kernels of fastfn.cuh (< 1 ulp, no 64-bit immediates to materialise) on
// their domain, libm outside it (zero, subnormal, infinite or NaN arguments; |x| >= 700 for exp)
// Inlined by default (loop-invariant logs hoist out of the term loops: stacks 2x); a translation unit may define MCU_DENSITY_MATH_NOINLINE
// to call them instead, which is faster where the inlined bodies push the kernel out of its register / instruction-cache budget
// (measured per template, profiles/r1_templates_bench.json: surgical, blocker, equiv).
#ifdef MCU_DENSITY_MATH_NOINLINE
#define MCU_DMATH static __device__ __noinline__
#else
#define MCU_DMATH MCU_D
#endif
MCU_DMATH double flog(double x) { return (x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308) ? fast_log(x) : log(x); }
MCU_DMATH double fexp(double x) { return (x > -700.0 && x < 700.0) ? fast_exp(x) : exp(x); }

// ---- univariate log densities (Distributions.jl formulas) -------------------
MCU_D double lp_normal(double x, double mu, double sigma) {
  if (isnan(x)) return neg_inf();
  const double z = (x - mu) / sigma;
  return -(z * z + kLog2Pi) / 2.0 - flog(sigma);
}
// InverseGamma(shape a, scale th), with lgamma(a) and a*flog(th) folded into c0 = a*flog(th) - lgamma(a)
MCU_D double lp_invgamma(double x, double a, double th, double c0, bool transform) {
  if (!(x >= 0.0)) return neg_inf();
  const double lx = flog(x);
  double lp = c0 - (a + 1.0) * lx - th / x;
  if (transform) lp += lx;
  return lp;
}
MCU_D double lp_gamma(double x, double a, double th, bool transform) {   // Gamma(shape a, scale th)
  if (!(x >= 0.0)) return neg_inf();
  const double lx = flog(x);
  double lp = -lgamma(a) - a * flog(th) + (a - 1.0) * lx - x / th;
  if (transform) lp += lx;
  return lp;
}
MCU_D double lp_exponential(double x, double th, bool transform) {       // Exponential(scale th)
  if (!(x >= 0.0)) return neg_inf();
  const double lambda = 1.0 / th;
  double lp = flog(lambda) - lambda * x;
  if (transform) lp += flog(x);
  return lp;
}
// Binomial(n, p = invlogit(eta)) at integer r, lc = lchoose(n, r) precomputed on the host
MCU_D double lp_binomial_logit(double r, double n, double lc, double eta) {
  const double p = 1.0 / (fexp(-eta) + 1.0);      // invlogit: src/utils.jl:64
  const double q = 1.0 - p;
  if (p == 0.0) return r == 0.0 ? 0.0 : neg_inf();
  if (q == 0.0) return r == n ? 0.0 : neg_inf();
  double lp = lc;
  if (r > 0.0) lp += r * flog(p);
  if (n - r > 0.0) lp += (n - r) * flog(q);
  return lp;
}
MCU_D double lp_poisson(double y, double lgy1 /* lgamma(y+1) */, double lam) {
  if (lam == 0.0) return y == 0.0 ? 0.0 : neg_inf();
  return y * flog(lam) - lam - lgy1;
}
// Laplace(location mu, scale theta): -(|x - mu| / theta + flog(2 theta))
MCU_D double lp_laplace(double x, double mu, double theta) { return -(fabs(x - mu) / theta + flog(2.0 * theta)); }
MCU_D double lp_bernoulli_logit(double y, double eta) {
  const double p = 1.0 / (fexp(-eta) + 1.0);
  return y == 0.0 ? flog(1.0 - p) : flog(p);
}
// MvNormal(mu, sigma) isotropic: -(d log 2pi + d log sigma^2)/2 - (|x-mu|^2 / sigma^2)/2
MCU_D double lp_isonormal(double sq, double d, double sigma) {
  const double v = sigma * sigma;
  return -(d * kLog2Pi + d * flog(v)) / 2.0 - (sq / v) / 2.0;
}
MCU_D double d_invgamma(double x, double a, double th) { return -(a + 1.0) / x + th / (x * x); }

MCU_D double digamma_d(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
                   f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + flog(x) - 0.5 / x + t;
}

// c0 of InverseGamma(0.001, 0.001): 0.001*flog(0.001) - lgamma(0.001)
MCU_D double ig001_c0() { return 0.001 * flog(0.001) - lgamma(0.001); }

// Defaults every template inherits: an element takes its node's link (arrays of different distributions override elem_link), bounded
// links take their interval from elem_bounds.
struct TplBase {
  static constexpr bool kGradSamplers = true;   // false: NUTS / HMC / MALA / AMM are compiled out of the generic kernel (templates whose state is too
                                                // large for ~40 block-sized per-thread vectors); mcu_set_scheme answers MCU_ERR_UNSUPPORTED for them
  MCU_HD static int elem_link(int /*e*/, int node_link) { return node_link; }
  MCU_HD static void elem_bounds(int /*e*/, double& lo, double& hi) { lo = 0.0; hi = 1.0; }
};
// Uniform(a, b) (Distributions.jl: -log(b - a) on [a, b]) with the two-sided link's log-Jacobian when the block samples on the link scale
MCU_D double lp_uniform(double x, double a, double b, bool transform) {
  if (!(x >= a && x <= b)) return neg_inf();
  double lp = -flog(b - a);
  if (transform) lp += flog((x - a) * (b - x) / (b - a));      // transformdistribution.jl:39-40
  return lp;
}
// Binomial(n, p) at integer r, lc = lchoose(n, r) precomputed on the host (closed form of dbinom, as lp_binomial_logit)
MCU_D double lp_binomial_p(double r, double n, double lc, double p) {
  const double q = 1.0 - p;
  if (p == 0.0) return r == 0.0 ? 0.0 : neg_inf();
  if (q == 0.0) return r == n ? 0.0 : neg_inf();
  double lp = lc;
  if (r > 0.0) lp += r * flog(p);
  if (n - r > 0.0) lp += (n - r) * flog(q);
  return lp;
}

// =============================================================================== line
struct LineModel : TplBase {
  static constexpr int D = 3, NN = 2, NF = 3, P = 3, MAXN = 64;
  struct Data { const double* x; const double* y; int N; };
  MCU_HD static int node_off(int n) { return n == 0 ? 0 : 2; }
  MCU_HD static int node_len(int n) { return n == 0 ? 2 : 1; }
  MCU_HD static int node_link(int n) { return n == 1 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 2 ? 0x3u : 0u; }
  MCU_HD static int mon_link(int j) { return j == 2 ? LINK_LOG : LINK_IDENT; }
  static const char* node_name(int n) { static const char* nm[] = {"beta", "s2"}; return nm[n]; }
  static const char* state_names() { return "beta[1]\nbeta[2]\ns2"; }
  static const char* monitor_names() { return "beta[1]\nbeta[2]\ns2"; }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    switch (f) {
      case 0: {  // beta ~ MvNormal(2, sqrt(1000))
        if (!isfinite(s[0]) || !isfinite(s[1])) return neg_inf();
        return lp_isonormal(s[0] * s[0] + s[1] * s[1], 2.0, sqrt(1000.0));
      }
      case 1: return lp_invgamma(s[2], 0.001, 0.001, ig001_c0(), transform);
      default: {  // y ~ MvNormal(xmat * beta, sqrt(s2))
        double sq = 0.0;
        for (int i = 0; i < d.N; ++i) { const double mu = 1.0 * s[0] + d.x[i] * s[1]; const double e = d.y[i] - mu; sq += e * e; }
        return lp_isonormal(sq, (double)d.N, sqrt(s[2]));
      }
    }
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    double sr = 0, sxr = 0, srr = 0;
    for (int i = 0; i < d.N; ++i) { const double r = d.y[i] - s[0] - s[1] * d.x[i]; sr += r; sxr += d.x[i] * r; srr += r * r; }
    g[0] = sr / s[2] - s[0] / 1000.0;
    g[1] = sxr / s[2] - s[1] / 1000.0;
    g[2] = -0.5 * (double)d.N / s[2] + 0.5 * srr / (s[2] * s[2]) + d_invgamma(s[2], 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int) { return false; }   // no element whose move touches only a few terms
  MCU_D static double elem_terms(const Data&, const double*, int, bool) { return 0.0; }
  // The tutorial's user-defined Gibbs samplers (doc/tutorial/line.jl:27-45, scheme3):
  //   beta | . ~ MvNormal(mu, Sigma),  Sigma = inv(X'X / s2 + I / 1000),  mu = Sigma X'y / s2        (Gibbs_beta; prior mean 0, invcov I / 1000)
  //   s2 | .   ~ InverseGamma(N / 2 + 0.001, sum (y - mu)^2 / 2 + 0.001)                               (Gibbs_s2)
  // Draw order: two normals (z1, z2; beta = mu + L z with L the lower Cholesky factor of Sigma); one Gamma(shape) variate, s2 = scale / G.
  MCU_HD static bool has_gibbs(int node) { return node == 0 || node == 1; }
  template <class R, class G> MCU_D static void gibbs(const Data& d, double* s, int node, R& rng, G rgamma) {
    if (node == 0) {
      double sx = 0, sxx = 0, sy = 0, sxy = 0;
      for (int i = 0; i < d.N; ++i) { sx += d.x[i]; sxx += d.x[i] * d.x[i]; sy += d.y[i]; sxy += d.x[i] * d.y[i]; }
      const double s2 = s[2];
      const double a11 = (double)d.N / s2 + 1.0 / 1000.0, a12 = sx / s2, a22 = sxx / s2 + 1.0 / 1000.0;
      const double det = a11 * a22 - a12 * a12;
      const double S11 = a22 / det, S12 = -a12 / det, S22 = a11 / det;
      const double r1 = sy / s2, r2 = sxy / s2;
      const double m1 = S11 * r1 + S12 * r2, m2 = S12 * r1 + S22 * r2;
      const double l11 = sqrt(S11), l21 = S12 / l11, l22 = sqrt(S22 - l21 * l21);
      const double z1 = rng.normal(), z2 = rng.normal();
      s[0] = m1 + l11 * z1;
      s[1] = m2 + l21 * z1 + l22 * z2;
    } else {
      double ss = 0.0;
      for (int i = 0; i < d.N; ++i) { const double r = d.y[i] - (1.0 * s[0] + d.x[i] * s[1]); ss += r * r; }
      const double a = (double)d.N / 2.0 + 0.001, b = ss / 2.0 + 0.001;
      s[2] = b / rgamma(a, rng);
    }
  }
  MCU_D static void monitor(const Data&, const double* s, double* out) { out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; }
  // distribution of observed element i given the state (predict, src/output/modelstats.jl:63-96):
  // returns OUT_NORMAL (a = mean, b = sd), OUT_BINOMIAL (a = n, b = p), OUT_POISSON (a = rate) or OUT_BERNOULLI (a = p)
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int i, double& a, double& b) { a = 1.0 * s[0] + d.x[i] * s[1]; b = sqrt(s[2]); return OUT_NORMAL; }
};

// =============================================================================== seeds
struct SeedsModel : TplBase {
  static constexpr int D = 26, NN = 6, NF = 7, P = 5, NP = 21;
  struct Data { const double* r; const double* n; const double* x1; const double* x2; const double* lc; int N; };
  MCU_HD static int node_off(int n) { return n <= 4 ? n : 5; }
  MCU_HD static int node_len(int n) { return n == 5 ? NP : 1; }
  MCU_HD static int node_link(int n) { return n == 4 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 5 ? 0x10u : f == 6 ? 0x2Fu : 0u; }
  MCU_HD static int mon_link(int j) { return j == 4 ? LINK_LOG : LINK_IDENT; }
  static const char* node_name(int n) { static const char* nm[] = {"alpha0", "alpha1", "alpha2", "alpha12", "s2", "b"}; return nm[n]; }
  static const char* state_names() {
    return "alpha0\nalpha1\nalpha2\nalpha12\ns2\nb[1]\nb[2]\nb[3]\nb[4]\nb[5]\nb[6]\nb[7]\nb[8]\nb[9]\nb[10]\nb[11]\nb[12]\nb[13]\nb[14]\nb[15]\nb[16]\nb[17]\nb[18]\nb[19]\nb[20]\nb[21]";
  }
  static const char* monitor_names() { return "alpha0\nalpha1\nalpha2\nalpha12\ns2"; }
  MCU_D static double eta(const Data& d, const double* s, int i) {
    return s[0] + s[1] * d.x1[i] + s[2] * d.x2[i] + s[3] * d.x1[i] * d.x2[i] + s[5 + i];
  }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f < 4) return lp_normal(s[f], 0.0, 1000.0);
    if (f == 4) return lp_invgamma(s[4], 0.001, 0.001, ig001_c0(), transform);
    if (f == 5) {  // b ~ Normal(0, sqrt(s2)), one distribution for the 21-vector
      const double sigma = sqrt(s[4]);
      double lp = 0.0;
      for (int i = 0; i < d.N; ++i) lp += lp_normal(s[5 + i], 0.0, sigma);
      return lp;
    }
    double lp = 0.0;  // r[i] ~ Binomial(n[i], invlogit(eta_i))
    for (int i = 0; i < d.N; ++i) lp += lp_binomial_logit(d.r[i], d.n[i], d.lc[i], eta(d, s, i));
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    double g0 = 0, g1 = 0, g2 = 0, g12 = 0, sbb = 0;
    const double s2 = s[4];
    for (int i = 0; i < d.N; ++i) {
      const double p = 1.0 / (fexp(-eta(d, s, i)) + 1.0);
      const double de = d.r[i] - d.n[i] * p;
      g0 += de; g1 += d.x1[i] * de; g2 += d.x2[i] * de; g12 += d.x1[i] * d.x2[i] * de;
      g[5 + i] = de - s[5 + i] / s2;
      sbb += s[5 + i] * s[5 + i];
    }
    g[0] = g0 - s[0] / 1e6; g[1] = g1 - s[1] / 1e6; g[2] = g2 - s[2] / 1e6; g[3] = g12 - s[3] / 1e6;
    g[4] = -0.5 * (double)d.N / s2 + 0.5 * sbb / (s2 * s2) + d_invgamma(s2, 0.001, 0.001);
  }
  // elem_local(e): a move of state element e changes only a few terms of the joint density; elem_terms = their sum at the current state
  // (own prior term, with the log-Jacobian when the block samples on the transformed scale, + the likelihood terms that read e).
  // The generic AMWG / univariate-slice samplers then form logf(v') = logf(v) - terms(v) + terms(v') instead of a full block evaluation.
  MCU_HD static bool elem_local(int e) { return e >= 5; }
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    const int i = e - 5;
    return lp_normal(s[e], 0.0, sqrt(s[4])) + lp_binomial_logit(d.r[i], d.n[i], d.lc[i], eta(d, s, i));
  }
  MCU_HD static bool has_gibbs(int) { return false; }   // no conjugate full conditional registered for this template (MCU_GIBBS)
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data&, const double* s, double* out) {
    for (int j = 0; j < 5; ++j) out[j] = s[j];
  }
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int i, double& a, double& b) { a = d.n[i]; b = 1.0 / (fexp(-eta(d, s, i)) + 1.0); return OUT_BINOMIAL; }
};

// =============================================================================== rats
struct RatsModel : TplBase {
  // state: mu_alpha, mu_beta, s2_alpha, s2_beta, s2_c, alpha[30], beta[30]
  static constexpr int D = 65, NN = 7, NF = 8, P = 3, NR = 30;
  struct Data { const double* y; const double* Xm; const int* rat; int N; double xbar; };
  MCU_HD static int node_off(int n) { return n < 5 ? n : (n == 5 ? 5 : 35); }
  MCU_HD static int node_len(int n) { return n < 5 ? 1 : NR; }
  MCU_HD static int node_link(int n) { return (n >= 2 && n <= 4) ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) {
    return f == 5 ? 0x05u /* mu_alpha, s2_alpha */ : f == 6 ? 0x0Au /* mu_beta, s2_beta */ : f == 7 ? 0x70u /* s2_c, alpha, beta */ : 0u;
  }
  // monitored: mu_beta (stochastic, identity), alpha0 (Logical → heuristic link), s2_c (log)
  MCU_HD static int mon_link(int j) { return j == 0 ? LINK_IDENT : j == 1 ? LINK_HEUR : LINK_LOG; }
  static const char* node_name(int n) { static const char* nm[] = {"mu_alpha", "mu_beta", "s2_alpha", "s2_beta", "s2_c", "alpha", "beta"}; return nm[n]; }
  static const char* state_names() { return nullptr; }   // generated: scalar nodes + alpha[1..30] + beta[1..30]
  static const char* monitor_names() { return "mu_beta\nalpha0\ns2_c"; }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f < 2) return lp_normal(s[f], 0.0, 1000.0);
    if (f < 5) return lp_invgamma(s[f], 0.001, 0.001, ig001_c0(), transform);
    if (f == 5) { const double sg = sqrt(s[2]); double lp = 0; for (int i = 0; i < NR; ++i) lp += lp_normal(s[5 + i], s[0], sg); return lp; }
    if (f == 6) { const double sg = sqrt(s[3]); double lp = 0; for (int i = 0; i < NR; ++i) lp += lp_normal(s[35 + i], s[1], sg); return lp; }
    double sq = 0.0;  // y ~ MvNormal(alpha[rat] + beta[rat] .* Xm, sqrt(s2_c))
    for (int k = 0; k < d.N; ++k) {
      const int i = d.rat[k];
      if (!isfinite(s[5 + i]) || !isfinite(s[35 + i])) { /* mu non-finite: logpdf is NaN/-Inf either way */ }
      const double mu = s[5 + i] + s[35 + i] * d.Xm[k];
      const double e = d.y[k] - mu; sq += e * e;
    }
    return lp_isonormal(sq, (double)d.N, sqrt(s[4]));
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double mua = s[0], mub = s[1], s2a = s[2], s2b = s[3], s2c = s[4];
    for (int i = 0; i < 2 * NR; ++i) g[5 + i] = 0.0;
    double see = 0;
    for (int k = 0; k < d.N; ++k) {
      const int i = d.rat[k];
      const double e = d.y[k] - (s[5 + i] + s[35 + i] * d.Xm[k]);
      g[5 + i] += e / s2c; g[35 + i] += e * d.Xm[k] / s2c; see += e * e;
    }
    double sa = 0, saa = 0, sb = 0, sbb = 0;
    for (int i = 0; i < NR; ++i) {
      const double da = s[5 + i] - mua, db = s[35 + i] - mub;
      g[5 + i] -= da / s2a; g[35 + i] -= db / s2b;
      sa += da; saa += da * da; sb += db; sbb += db * db;
    }
    g[0] = sa / s2a - mua / 1e6;
    g[1] = sb / s2b - mub / 1e6;
    g[2] = -0.5 * NR / s2a + 0.5 * saa / (s2a * s2a) + d_invgamma(s2a, 0.001, 0.001);
    g[3] = -0.5 * NR / s2b + 0.5 * sbb / (s2b * s2b) + d_invgamma(s2b, 0.001, 0.001);
    g[4] = -0.5 * (double)d.N / s2c + 0.5 * see / (s2c * s2c) + d_invgamma(s2c, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 5; }   // alpha_i, beta_i: own Normal term + rat i's five observations
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    const int i = e < 35 ? e - 5 : e - 35;
    const double own = e < 35 ? lp_normal(s[e], s[0], sqrt(s[2])) : lp_normal(s[e], s[1], sqrt(s[3]));
    double sq = 0.0;
    for (int k = 0; k < d.N; ++k) if (d.rat[k] == i) { const double r = d.y[k] - (s[5 + i] + s[35 + i] * d.Xm[k]); sq += r * r; }
    return own - (sq / s[4]) / 2.0;
  }
  MCU_HD static bool has_gibbs(int) { return false; }   // no conjugate full conditional registered for this template (MCU_GIBBS)
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data& d, const double* s, double* out) {
    out[0] = s[1]; out[1] = s[0] - d.xbar * s[1]; out[2] = s[4];   // alpha0 = mu_alpha - xbar * mu_beta (rats.jl:64-66)
  }
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int k, double& a, double& b) { const int i = d.rat[k]; a = s[5 + i] + s[35 + i] * d.Xm[k]; b = sqrt(s[4]); return OUT_NORMAL; }
};

// =============================================================================== pumps
struct PumpsModel : TplBase {
  static constexpr int D = 12, NN = 3, NF = 4, P = 12, NPUMP = 10;
  struct Data { const double* y; const double* t; const double* lgy1; int N; };
  MCU_HD static int node_off(int n) { return n; }
  MCU_HD static int node_len(int n) { return n == 2 ? NPUMP : 1; }
  MCU_HD static int node_link(int) { return LINK_LOG; }
  MCU_HD static uint32_t parents(int f) { return f == 2 ? 0x3u : f == 3 ? 0x4u : 0u; }
  MCU_HD static int mon_link(int) { return LINK_LOG; }
  static const char* node_name(int n) { static const char* nm[] = {"alpha", "beta", "theta"}; return nm[n]; }
  static const char* state_names() { return "alpha\nbeta\ntheta[1]\ntheta[2]\ntheta[3]\ntheta[4]\ntheta[5]\ntheta[6]\ntheta[7]\ntheta[8]\ntheta[9]\ntheta[10]"; }
  static const char* monitor_names() { return state_names(); }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f == 0) return lp_exponential(s[0], 1.0, transform);
    if (f == 1) return lp_gamma(s[1], 0.1, 1.0, transform);
    if (f == 2) {  // theta ~ Gamma(alpha, 1 / beta), one distribution for the 10-vector
      const double a = s[0], th = 1.0 / s[1];
      const double c = -lgamma(a) - a * flog(th);
      double lp = 0.0;
      for (int i = 0; i < d.N; ++i) {
        const double x = s[2 + i];
        if (!(x >= 0.0)) { lp += neg_inf(); continue; }
        const double lx = flog(x);
        double t = c + (a - 1.0) * lx - x / th;
        if (transform) t += lx;
        lp += t;
      }
      return lp;
    }
    double lp = 0.0;  // y[i] ~ Poisson(theta[i] * t[i])
    for (int i = 0; i < d.N; ++i) lp += lp_poisson(d.y[i], d.lgy1[i], s[2 + i] * d.t[i]);
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double al = s[0], be = s[1];
    double slog = 0, sth = 0; const double N = (double)d.N;
    for (int i = 0; i < d.N; ++i) {
      const double th = s[2 + i];
      g[2 + i] = d.y[i] / th - d.t[i] + (al - 1.0) / th - be;
      slog += flog(th); sth += th;
    }
    g[0] = N * flog(be) + slog - N * digamma_d(al) - 1.0;
    g[1] = N * al / be - sth + (0.1 - 1.0) / be - 1.0;
  }
  MCU_D static void monitor(const Data&, const double* s, double* out) { for (int j = 0; j < 12; ++j) out[j] = s[j]; }
  // conjugate full conditionals (MCU_GIBBS): theta_i | . ~ Gamma(alpha + y_i, 1/(beta + t_i)); beta | . ~ Gamma(0.1 + N alpha, 1/(1 + sum theta))
  MCU_HD static bool elem_local(int e) { return e >= 2; }   // theta_i: its Gamma term + its Poisson term
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool transform) {
    const int i = e - 2;
    const double own = lp_gamma(s[e], s[0], 1.0 / s[1], transform);
    if (!isfinite(own)) return own;                       // outside the support: the reference stops at the first non-finite partial sum
    return own + lp_poisson(d.y[i], d.lgy1[i], s[e] * d.t[i]);
  }
  MCU_HD static bool has_gibbs(int node) { return node == 1 || node == 2; }
  template <class R, class G>
  MCU_D static void gibbs(const Data& d, double* s, int node, R& rng, G rgamma) {
    if (node == 2) {
      for (int i = 0; i < d.N; ++i) s[2 + i] = rgamma(s[0] + d.y[i], rng) / (s[1] + d.t[i]);
    } else {
      double sth = 0.0;
      for (int i = 0; i < d.N; ++i) sth += s[2 + i];
      s[1] = rgamma(0.1 + (double)d.N * s[0], rng) / (1.0 + sth);
    }
  }
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int i, double& a, double& b) { a = s[2 + i] * d.t[i]; b = 0.0; return OUT_POISSON; }
};

// =============================================================================== surgical
// doc/examples/surgical.jl:11-43: r_i ~ Binomial(n_i, invlogit(b_i)), b_i ~ Normal(mu, sqrt(s2)), mu ~ Normal(0, 1000),
// s2 ~ InverseGamma(0.001, 0.001); Logical p = invlogit(b), pop_mean = invlogit(mu).  State: mu, s2, b[12].
struct SurgicalModel : TplBase {
  static constexpr int D = 14, NN = 3, NF = 4, P = 15, NH = 12;
  struct Data { const double* r; const double* n; const double* lc; int N; };
  MCU_HD static int node_off(int n) { return n; }
  MCU_HD static int node_len(int n) { return n == 2 ? NH : 1; }
  MCU_HD static int node_link(int n) { return n == 1 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 2 ? 0x3u /* mu, s2 */ : f == 3 ? 0x4u /* b */ : 0u; }
  // monitored: mu (stochastic, identity), pop_mean (Logical), s2 (log), p[12] (Logical)
  MCU_HD static int mon_link(int j) { return j == 0 ? LINK_IDENT : j == 2 ? LINK_LOG : LINK_HEUR; }
  static const char* node_name(int n) { static const char* nm[] = {"mu", "s2", "b"}; return nm[n]; }
  static const char* state_names() { return "mu\ns2\nb[1]\nb[2]\nb[3]\nb[4]\nb[5]\nb[6]\nb[7]\nb[8]\nb[9]\nb[10]\nb[11]\nb[12]"; }
  static const char* monitor_names() { return "mu\npop_mean\ns2\np[1]\np[2]\np[3]\np[4]\np[5]\np[6]\np[7]\np[8]\np[9]\np[10]\np[11]\np[12]"; }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f == 0) return lp_normal(s[0], 0.0, 1000.0);
    if (f == 1) return lp_invgamma(s[1], 0.001, 0.001, ig001_c0(), transform);
    if (f == 2) {   // b ~ Normal(mu, sqrt(s2)), one distribution for the 12-vector
      const double sigma = sqrt(s[1]);
      double lp = 0.0;
      for (int i = 0; i < d.N; ++i) lp += lp_normal(s[2 + i], s[0], sigma);
      return lp;
    }
    double lp = 0.0;   // r[i] ~ Binomial(n[i], invlogit(b[i]))
    for (int i = 0; i < d.N; ++i) lp += lp_binomial_logit(d.r[i], d.n[i], d.lc[i], s[2 + i]);
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double mu = s[0], s2 = s[1];
    double sd = 0.0, sdd = 0.0;
    for (int i = 0; i < d.N; ++i) {
      const double p = 1.0 / (fexp(-s[2 + i]) + 1.0), db = s[2 + i] - mu;
      g[2 + i] = (d.r[i] - d.n[i] * p) - db / s2;
      sd += db; sdd += db * db;
    }
    g[0] = sd / s2 - mu / 1e6;
    g[1] = -0.5 * (double)d.N / s2 + 0.5 * sdd / (s2 * s2) + d_invgamma(s2, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 2; }   // b_i: its Normal term + its Binomial term
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    const int i = e - 2;
    return lp_normal(s[e], s[0], sqrt(s[1])) + lp_binomial_logit(d.r[i], d.n[i], d.lc[i], s[e]);
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data& d, const double* s, double* out) {
    out[0] = s[0]; out[1] = 1.0 / (fexp(-s[0]) + 1.0); out[2] = s[1];   // pop_mean = invlogit(mu): surgical.jl:34-36
    for (int i = 0; i < d.N; ++i) out[3 + i] = 1.0 / (fexp(-s[2 + i]) + 1.0);   // p = invlogit(b): surgical.jl:19-21
  }
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int i, double& a, double& b) { a = d.n[i]; b = 1.0 / (fexp(-s[2 + i]) + 1.0); return OUT_BINOMIAL; }
};

// =============================================================================== dyes
// doc/examples/dyes.jl:22-47: y_k ~ MvNormal(mu[batch_k], sqrt(s2_within)) (30-dim iso), mu_i ~ Normal(theta, sqrt(s2_between)) (6 batches),
// theta ~ Normal(0, 1000), s2_within, s2_between ~ InverseGamma(0.001, 0.001).  State / monitors (order of doc/examples/dyes.rst):
// s2_between, theta, s2_within, mu[6].
struct DyesModel : TplBase {
  static constexpr int D = 9, NN = 4, NF = 5, P = 9, NB = 6;
  struct Data { const double* y; const int* batch; int N; };
  MCU_HD static int node_off(int n) { return n; }
  MCU_HD static int node_len(int n) { return n == 3 ? NB : 1; }
  MCU_HD static int node_link(int n) { return (n == 0 || n == 2) ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 3 ? 0x3u /* s2_between, theta */ : f == 4 ? 0xCu /* s2_within, mu */ : 0u; }
  MCU_HD static int mon_link(int j) { return (j == 0 || j == 2) ? LINK_LOG : LINK_IDENT; }
  static const char* node_name(int n) { static const char* nm[] = {"s2_between", "theta", "s2_within", "mu"}; return nm[n]; }
  static const char* state_names() { return "s2_between\ntheta\ns2_within\nmu[1]\nmu[2]\nmu[3]\nmu[4]\nmu[5]\nmu[6]"; }
  static const char* monitor_names() { return state_names(); }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f == 0) return lp_invgamma(s[0], 0.001, 0.001, ig001_c0(), transform);
    if (f == 1) return lp_normal(s[1], 0.0, 1000.0);
    if (f == 2) return lp_invgamma(s[2], 0.001, 0.001, ig001_c0(), transform);
    if (f == 3) { const double sg = sqrt(s[0]); double lp = 0.0; for (int i = 0; i < NB; ++i) lp += lp_normal(s[3 + i], s[1], sg); return lp; }
    double sq = 0.0;
    for (int k = 0; k < d.N; ++k) { const double e = d.y[k] - s[3 + d.batch[k]]; sq += e * e; }
    return lp_isonormal(sq, (double)d.N, sqrt(s[2]));
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double s2b = s[0], th = s[1], s2w = s[2];
    for (int i = 0; i < NB; ++i) g[3 + i] = 0.0;
    double see = 0.0;
    for (int k = 0; k < d.N; ++k) { const int i = d.batch[k]; const double e = d.y[k] - s[3 + i]; g[3 + i] += e / s2w; see += e * e; }
    double sd = 0.0, sdd = 0.0;
    for (int i = 0; i < NB; ++i) { const double dm = s[3 + i] - th; g[3 + i] -= dm / s2b; sd += dm; sdd += dm * dm; }
    g[1] = sd / s2b - th / 1e6;
    g[0] = -0.5 * NB / s2b + 0.5 * sdd / (s2b * s2b) + d_invgamma(s2b, 0.001, 0.001);
    g[2] = -0.5 * (double)d.N / s2w + 0.5 * see / (s2w * s2w) + d_invgamma(s2w, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 3; }   // mu_i: its Normal term + the five samples of batch i
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    const int i = e - 3;
    double sq = 0.0;
    for (int k = 0; k < d.N; ++k) if (d.batch[k] == i) { const double r = d.y[k] - s[e]; sq += r * r; }
    return lp_normal(s[e], s[1], sqrt(s[0])) - (sq / s[2]) / 2.0;
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data&, const double* s, double* out) { for (int j = 0; j < 9; ++j) out[j] = s[j]; }
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int k, double& a, double& b) { a = s[3 + d.batch[k]]; b = sqrt(s[2]); return OUT_NORMAL; }
};

// =============================================================================== salm
// doc/examples/salm.jl:16-53 (data :4-11): 3 plates x 6 doses of a mutagenicity assay, Poisson counts with a log-linear dose
// response and an extra-Poisson random effect per plate/dose.  Matrices are flattened column-major (e = plate + 3 dose), as unlist
// does (src/model/dependent.jl:192-195).  State: s2, gamma, beta, alpha, lambda[18] (the first four are the monitored columns, in the
// order of doc/examples/salm.rst).
struct SalmModel : TplBase {
  static constexpr int D = 22, NN = 5, NF = 6, P = 4, NY = 18, NPLATE = 3;
  struct Data { const double* y; const double* x; const double* lgy1; int N; };
  MCU_HD static int node_off(int n) { return n; }
  MCU_HD static int node_len(int n) { return n == 4 ? NY : 1; }
  MCU_HD static int node_link(int n) { return n == 0 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 4 ? 0x1u /* s2 */ : f == 5 ? 0x1Eu /* gamma, beta, alpha, lambda */ : 0u; }
  MCU_HD static int mon_link(int j) { return j == 0 ? LINK_LOG : LINK_IDENT; }   // lambda and y are declared with monitor = false
  static const char* node_name(int n) { static const char* nm[] = {"s2", "gamma", "beta", "alpha", "lambda"}; return nm[n]; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return "s2\ngamma\nbeta\nalpha"; }
  MCU_D static double mu(const Data& d, const double* s, int e) {   // fexp(alpha + beta flog(x_j + 10) + gamma x_j + lambda_ij): salm.jl:22
    const double x = d.x[e / NPLATE];
    return fexp(s[3] + s[2] * flog(x + 10.0) + s[1] * x + s[4 + e]);
  }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f == 0) return lp_invgamma(s[0], 0.001, 0.001, ig001_c0(), transform);
    if (f < 4) return lp_normal(s[f], 0.0, 1000.0);
    if (f == 4) {   // lambda ~ Normal(0, sqrt(s2)), one distribution for the 3 x 6 matrix
      const double sigma = sqrt(s[0]);
      double lp = 0.0;
      for (int e = 0; e < NY; ++e) lp += lp_normal(s[4 + e], 0.0, sigma);
      return lp;
    }
    double lp = 0.0;   // y[i, j] ~ Poisson(mu_ij)
    for (int e = 0; e < NY; ++e) lp += lp_poisson(d.y[e], d.lgy1[e], mu(d, s, e));
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    double ga = 0, gb = 0, gg = 0, sll = 0;
    const double s2 = s[0];
    for (int e = 0; e < NY; ++e) {
      const double x = d.x[e / NPLATE];
      const double r = d.y[e] - mu(d, s, e);
      ga += r; gb += r * flog(x + 10.0); gg += r * x;
      g[4 + e] = r - s[4 + e] / s2;
      sll += s[4 + e] * s[4 + e];
    }
    g[3] = ga - s[3] / 1e6; g[2] = gb - s[2] / 1e6; g[1] = gg - s[1] / 1e6;
    g[0] = -0.5 * (double)NY / s2 + 0.5 * sll / (s2 * s2) + d_invgamma(s2, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 4; }   // lambda_ij: its Normal term + its Poisson term
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    const int i = e - 4;
    return lp_normal(s[e], 0.0, sqrt(s[0])) + lp_poisson(d.y[i], d.lgy1[i], mu(d, s, i));
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data&, const double* s, double* out) { for (int j = 0; j < 4; ++j) out[j] = s[j]; }
  MCU_HD static int out_len(const Data&) { return NY; }
  MCU_D static int out_dist(const Data& d, const double* s, int e, double& a, double& b) { a = mu(d, s, e); b = 0.0; return OUT_POISSON; }
};

// =============================================================================== equiv
// doc/examples/equiv.jl:25-75 (data :4-22): two-period crossover bioequivalence trial, 10 subjects x 2 periods, Normal responses with
// treatment (phi), period (pi) and subject-by-period (delta) effects; theta = fexp(phi) and equiv = 1{0.8 < theta < 1.2} are Logical.
// Matrices column-major (e = subject + 10 period).  State: s2_2, s2_1, pi, phi, mu, delta[20]; monitored s2_2, s2_1, pi, phi, theta,
// equiv, mu (the order of doc/examples/equiv.rst).
struct EquivModel : TplBase {
  static constexpr int D = 25, NN = 6, NF = 7, P = 7, NS = 10, NY = 20;
  struct Data { const double* y; const double* group; int N; };
  MCU_HD static int node_off(int n) { return n; }
  MCU_HD static int node_len(int n) { return n == 5 ? NY : 1; }
  MCU_HD static int node_link(int n) { return n < 2 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 5 ? 0x1u /* s2_2 */ : f == 6 ? 0x3Eu /* s2_1, pi, phi, mu, delta */ : 0u; }
  MCU_HD static int mon_link(int j) { return j < 2 ? LINK_LOG : (j == 4 || j == 5) ? LINK_HEUR : LINK_IDENT; }
  static const char* node_name(int n) { static const char* nm[] = {"s2_2", "s2_1", "pi", "phi", "mu", "delta"}; return nm[n]; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return "s2_2\ns2_1\npi\nphi\ntheta\nequiv\nmu"; }
  // m_ij = mu + (-1)^(T[i,j] - 1) phi / 2 + (-1)^(j - 1) pi / 2 + delta[i,j],  T = [group  3 - group]  (equiv.jl:22, 33-34)
  MCU_D static double sphi(const Data& d, int e) { const int i = e % NS, j = e / NS; const double T = j == 0 ? d.group[i] : 3.0 - d.group[i]; return T == 1.0 ? 1.0 : -1.0; }
  MCU_D static double mean(const Data& d, const double* s, int e) {
    const double spi = e < NS ? 1.0 : -1.0;
    return s[4] + sphi(d, e) * s[3] / 2.0 + spi * s[2] / 2.0 + s[5 + e];
  }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f < 2) return lp_invgamma(s[f], 0.001, 0.001, ig001_c0(), transform);
    if (f < 5) return lp_normal(s[f], 0.0, 1000.0);
    if (f == 5) { const double sg = sqrt(s[0]); double lp = 0.0; for (int e = 0; e < NY; ++e) lp += lp_normal(s[5 + e], 0.0, sg); return lp; }
    const double sg = sqrt(s[1]);
    double lp = 0.0;
    for (int e = 0; e < NY; ++e) lp += lp_normal(d.y[e], mean(d, s, e), sg);
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double s22 = s[0], s21 = s[1];
    double gm = 0, gp = 0, gq = 0, see = 0, sdd = 0;
    for (int e = 0; e < NY; ++e) {
      const double res = d.y[e] - mean(d, s, e);
      const double r = res / s21;
      gm += r; gp += r * sphi(d, e) / 2.0; gq += r * (e < NS ? 0.5 : -0.5);
      g[5 + e] = r - s[5 + e] / s22;
      see += res * res; sdd += s[5 + e] * s[5 + e];
    }
    g[4] = gm - s[4] / 1e6; g[3] = gp - s[3] / 1e6; g[2] = gq - s[2] / 1e6;
    g[1] = -0.5 * (double)NY / s21 + 0.5 * see / (s21 * s21) + d_invgamma(s21, 0.001, 0.001);
    g[0] = -0.5 * (double)NY / s22 + 0.5 * sdd / (s22 * s22) + d_invgamma(s22, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 5; }   // delta_ij: its Normal term + its observation
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    const int i = e - 5;
    return lp_normal(s[e], 0.0, sqrt(s[0])) + lp_normal(d.y[i], mean(d, s, i), sqrt(s[1]));
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data&, const double* s, double* out) {
    const double theta = fexp(s[3]);
    out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; out[3] = s[3]; out[4] = theta; out[5] = (0.8 < theta && theta < 1.2) ? 1.0 : 0.0; out[6] = s[4];
  }
  MCU_HD static int out_len(const Data&) { return NY; }
  MCU_D static int out_dist(const Data& d, const double* s, int e, double& a, double& b) { a = mean(d, s, e); b = sqrt(s[1]); return OUT_NORMAL; }
};

// =============================================================================== blocker
// doc/examples/blocker.jl:22-69 (data :4-18): meta-analysis of 22 beta-blocker trials, two observed Binomial nodes (control and treated
// arms), trial baselines mu[22], trial effects delta[22] ~ Normal(d, sqrt(s2)) and a predictive effect delta_new.
// State: s2, d, delta_new, mu[22], delta[22]; monitored s2, d, delta_new (the order of doc/examples/blocker.rst).
struct BlockerModel : TplBase {
  static constexpr int D = 47, NN = 5, NF = 7, P = 3, NT = 22;
  struct Data { const double* rc; const double* nc; const double* rt; const double* nt; const double* lcc; const double* lct; int N; };
  MCU_HD static int node_off(int n) { return n < 3 ? n : (n == 3 ? 3 : 3 + NT); }
  MCU_HD static int node_len(int n) { return n < 3 ? 1 : NT; }
  MCU_HD static int node_link(int n) { return n == 0 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) {
    return (f == 2 || f == 4) ? 0x3u /* s2, d */ : f == 5 ? 0x8u /* mu */ : f == 6 ? 0x18u /* mu, delta */ : 0u;
  }
  MCU_HD static int mon_link(int j) { return j == 0 ? LINK_LOG : LINK_IDENT; }
  static const char* node_name(int n) { static const char* nm[] = {"s2", "d", "delta_new", "mu", "delta"}; return nm[n]; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return "s2\nd\ndelta_new"; }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f == 0) return lp_invgamma(s[0], 0.001, 0.001, ig001_c0(), transform);
    if (f == 1) return lp_normal(s[1], 0.0, 1000.0);
    if (f == 2) return lp_normal(s[2], s[1], sqrt(s[0]));
    if (f == 3) { double lp = 0.0; for (int i = 0; i < NT; ++i) lp += lp_normal(s[3 + i], 0.0, 1000.0); return lp; }
    if (f == 4) { const double sg = sqrt(s[0]); double lp = 0.0; for (int i = 0; i < NT; ++i) lp += lp_normal(s[3 + NT + i], s[1], sg); return lp; }
    double lp = 0.0;
    if (f == 5) { for (int i = 0; i < NT; ++i) lp += lp_binomial_logit(d.rc[i], d.nc[i], d.lcc[i], s[3 + i]); return lp; }          // rc ~ Binomial(nc, invlogit(mu))
    for (int i = 0; i < NT; ++i) lp += lp_binomial_logit(d.rt[i], d.nt[i], d.lct[i], s[3 + i] + s[3 + NT + i]);                      // rt ~ Binomial(nt, invlogit(mu + delta))
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double s2 = s[0], dd = s[1];
    double sd = s[2] - dd, sdd = sd * sd;
    g[2] = -(s[2] - dd) / s2;
    for (int i = 0; i < NT; ++i) {
      const double mu = s[3 + i], dl = s[3 + NT + i];
      const double pc = 1.0 / (fexp(-mu) + 1.0), pt = 1.0 / (fexp(-(mu + dl)) + 1.0);
      const double rt = d.rt[i] - d.nt[i] * pt;
      g[3 + i] = (d.rc[i] - d.nc[i] * pc) + rt - mu / 1e6;
      g[3 + NT + i] = rt - (dl - dd) / s2;
      sd += dl - dd; sdd += (dl - dd) * (dl - dd);
    }
    g[1] = sd / s2 - dd / 1e6;
    g[0] = -0.5 * (double)(NT + 1) / s2 + 0.5 * sdd / (s2 * s2) + d_invgamma(s2, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 2; }   // delta_new: one Normal; mu_i: prior + both arms of trial i; delta_i: prior + treated arm
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    if (e == 2) return lp_normal(s[2], s[1], sqrt(s[0]));
    if (e < 3 + NT) {
      const int i = e - 3;
      return lp_normal(s[e], 0.0, 1000.0) + lp_binomial_logit(d.rc[i], d.nc[i], d.lcc[i], s[e]) + lp_binomial_logit(d.rt[i], d.nt[i], d.lct[i], s[e] + s[e + NT]);
    }
    const int i = e - 3 - NT;
    return lp_normal(s[e], s[1], sqrt(s[0])) + lp_binomial_logit(d.rt[i], d.nt[i], d.lct[i], s[3 + i] + s[e]);
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data&, const double* s, double* out) { out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; }
  MCU_HD static int out_len(const Data&) { return 2 * NT; }   // rc[22] then rt[22]
  MCU_D static int out_dist(const Data& d, const double* s, int e, double& a, double& b) {
    if (e < NT) { a = d.nc[e]; b = 1.0 / (fexp(-s[3 + e]) + 1.0); }
    else { const int i = e - NT; a = d.nt[i]; b = 1.0 / (fexp(-(s[3 + i] + s[3 + NT + i])) + 1.0); }
    return OUT_BINOMIAL;
  }
};

// =============================================================================== stacks
// doc/examples/stacks.jl:41-94 (data :4-38): stack-loss regression on standardised covariates z with a Laplace likelihood,
// y[i] ~ Laplace(beta0 + z[i,:] . beta, s2).  Every monitored quantity is a Logical node: b = beta ./ sdx, b0 = beta0 - b . meanx,
// sigma = sqrt(2) s2, outlier[i] = |y[i] - mu[i]| / sigma > 2.5 for i in (1, 3, 4, 21).  State: beta0, beta[3], s2.
struct StacksModel : TplBase {
  static constexpr int D = 5, NN = 3, NF = 4, P = 9, NOBS = 21;
  struct Data { const double* y; const double* z; const double* meanx; const double* sdx; int N; };
  MCU_HD static int node_off(int n) { return n == 0 ? 0 : (n == 1 ? 1 : 4); }
  MCU_HD static int node_len(int n) { return n == 1 ? 3 : 1; }
  MCU_HD static int node_link(int n) { return n == 2 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 3 ? 0x7u : 0u; }
  MCU_HD static int mon_link(int) { return LINK_HEUR; }   // no monitored stochastic node
  static const char* node_name(int n) { static const char* nm[] = {"beta0", "beta", "s2"}; return nm[n]; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return "b[1]\nb[2]\nb[3]\nb0\nsigma\noutlier[1]\noutlier[3]\noutlier[4]\noutlier[21]"; }
  MCU_D static double mu(const Data& d, const double* s, int i) { return s[0] + (d.z[i * 3] * s[1] + d.z[i * 3 + 1] * s[2] + d.z[i * 3 + 2] * s[3]); }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f == 0) return lp_normal(s[0], 0.0, 1000.0);
    if (f == 1) { double lp = 0.0; for (int j = 0; j < 3; ++j) lp += lp_normal(s[1 + j], 0.0, 1000.0); return lp; }
    if (f == 2) return lp_invgamma(s[4], 0.001, 0.001, ig001_c0(), transform);
    double lp = 0.0;
    for (int i = 0; i < d.N; ++i) lp += lp_laplace(d.y[i], mu(d, s, i), s[4]);
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double th = s[4];
    double g0 = 0, g1 = 0, g2 = 0, g3 = 0, sabs = 0;
    for (int i = 0; i < d.N; ++i) {
      const double e = d.y[i] - mu(d, s, i);
      const double sgn = e > 0 ? 1.0 : (e < 0 ? -1.0 : 0.0);
      g0 += sgn; g1 += sgn * d.z[i * 3]; g2 += sgn * d.z[i * 3 + 1]; g3 += sgn * d.z[i * 3 + 2];
      sabs += fabs(e);
    }
    g[0] = g0 / th - s[0] / 1e6; g[1] = g1 / th - s[1] / 1e6; g[2] = g2 / th - s[2] / 1e6; g[3] = g3 / th - s[3] / 1e6;
    g[4] = -(double)d.N / th + sabs / (th * th) + d_invgamma(th, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int) { return false; }   // no element whose move touches only a few terms
  MCU_D static double elem_terms(const Data&, const double*, int, bool) { return 0.0; }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data& d, const double* s, double* out) {
    double dot = 0.0;
    for (int j = 0; j < 3; ++j) { out[j] = s[1 + j] / d.sdx[j]; dot += out[j] * d.meanx[j]; }
    out[3] = s[0] - dot;
    const double sigma = sqrt(2.0) * s[4];
    out[4] = sigma;
    const int idx[4] = {0, 2, 3, 20};
    for (int q = 0; q < 4; ++q) out[5 + q] = fabs((d.y[idx[q]] - mu(d, s, idx[q])) / sigma) > 2.5 ? 1.0 : 0.0;
  }
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int i, double& a, double& b) { a = mu(d, s, i); b = s[4]; return OUT_LAPLACE; }
};

// =============================================================================== magnesium
// doc/examples/magnesium.jl:21-82 (data :4-17): meta-analysis of 8 trials under six priors for the between-trial sd tau.  The example
// whose parameter nodes carry bounded distributions: pc ~ Uniform(0, 1), mu ~ Uniform(-10, 10) — sampled by AMWG on the two-sided link
// (transformdistribution.jl:6-48) — and priors = [InverseGamma(.001, .001), Uniform(0, 50) x 2, Uniform(0, 1) x 2,
// Truncated(Normal(0, sqrt(s2_0 / erf(0.75))), 0, Inf)].  6 x 8 matrices are column-major (prior index i fastest), as in Julia.
// State: priors[6], mu[6], theta[48], pc[48]; monitored (both Logical): tau[6], OR[6] = exp(mu).
struct MagnesiumModel : TplBase {
  static constexpr int D = 108, NN = 4, NF = 6, P = 12, NPR = 6, NTR = 8, NE = 48;
  struct Data { const double* rc; const double* nc; const double* rt; const double* nt; const double* lcc; const double* lct; double s2_0; double sd6; };
  MCU_HD static int node_off(int n) { return n == 0 ? 0 : n == 1 ? 6 : n == 2 ? 12 : 60; }
  MCU_HD static int node_len(int n) { return n < 2 ? NPR : NE; }
  MCU_HD static int node_link(int n) { return n == 2 ? LINK_IDENT : LINK_BOUNDED; }
  MCU_HD static int elem_link(int e, int node_link) { return (e == 0 || e == 5) ? LINK_LOG : node_link; }   // InverseGamma / Normal truncated to [0, Inf)
  MCU_HD static void elem_bounds(int e, double& lo, double& hi) {
    if (e == 1 || e == 2) { lo = 0.0; hi = 50.0; }
    else if (e >= 6 && e < 12) { lo = -10.0; hi = 10.0; }
    else { lo = 0.0; hi = 1.0; }
  }
  MCU_HD static uint32_t parents(int f) { return f == 2 ? 0x3u /* priors (through tau), mu */ : f == 4 ? 0x8u /* pc */ : f == 5 ? 0xCu /* theta, pc */ : 0u; }
  MCU_HD static int mon_link(int) { return LINK_HEUR; }
  static const char* node_name(int n) { static const char* nm[] = {"priors", "mu", "theta", "pc"}; return nm[n]; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return "tau[1]\ntau[2]\ntau[3]\ntau[4]\ntau[5]\ntau[6]\nOR[1]\nOR[2]\nOR[3]\nOR[4]\nOR[5]\nOR[6]"; }
  // tau = Logical(priors, s2_0): magnesium.jl:60-70
  MCU_D static double tau(const Data& d, const double* s, int i) {
    switch (i) {
      case 0: return sqrt(s[0]);
      case 1: return sqrt(s[1]);
      case 2: return s[2];
      case 3: return sqrt(d.s2_0 * (1.0 / s[3] - 1.0));
      case 4: return sqrt(d.s2_0) * (1.0 / s[4] - 1.0);
      default: return sqrt(s[5]);
    }
  }
  MCU_D static double prior_term(const Data& d, const double* s, int k, bool transform) {
    const double x = s[k];
    switch (k) {
      case 0: return lp_invgamma(x, 0.001, 0.001, ig001_c0(), transform);
      case 1: case 2: return lp_uniform(x, 0.0, 50.0, transform);
      case 3: case 4: return lp_uniform(x, 0.0, 1.0, transform);
      default: {   // Truncated(Normal(0, sd6), 0, Inf): logpdf(Normal) - log(1/2) on [0, Inf) (Distributions/truncate.jl)
        if (!(x >= 0.0)) return neg_inf();
        double lp = lp_normal(x, 0.0, d.sd6) - flog(0.5);
        if (transform) lp += flog(x);
        return lp;
      }
    }
  }
  MCU_D static double rt_term(const Data& d, const double* s, int e) {   // rtx[i, j] ~ Binomial(nt[j], invlogit(theta + logit(pc))): magnesium.jl:37-49
    const int j = e / NPR;
    const double pc = s[60 + e], phi = flog(pc / (1.0 - pc));
    const double pt = 1.0 / (fexp(-(s[12 + e] + phi)) + 1.0);
    return lp_binomial_p(d.rt[j], d.nt[j], d.lct[j], pt);
  }
  MCU_D static double rc_term(const Data& d, const double* s, int e) { const int j = e / NPR; return lp_binomial_p(d.rc[j], d.nc[j], d.lcc[j], s[60 + e]); }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    double lp = 0.0;
    if (f == 0) { for (int k = 0; k < NPR; ++k) lp += prior_term(d, s, k, transform); return lp; }
    if (f == 1) { for (int i = 0; i < NPR; ++i) lp += lp_uniform(s[6 + i], -10.0, 10.0, transform); return lp; }
    if (f == 2) {   // theta[i, j] ~ Normal(mu[i], tau[i])
      double t[NPR]; for (int i = 0; i < NPR; ++i) t[i] = tau(d, s, i);
      for (int e = 0; e < NE; ++e) lp += lp_normal(s[12 + e], s[6 + e % NPR], t[e % NPR]);
      return lp;
    }
    if (f == 3) { for (int e = 0; e < NE; ++e) lp += lp_uniform(s[60 + e], 0.0, 1.0, transform); return lp; }
    if (f == 4) { for (int e = 0; e < NE; ++e) lp += rc_term(d, s, e); return lp; }
    for (int e = 0; e < NE; ++e) lp += rt_term(d, s, e);
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    double dmu[NPR] = {0, 0, 0, 0, 0, 0}, dtau[NPR] = {0, 0, 0, 0, 0, 0}, t[NPR];
    for (int i = 0; i < NPR; ++i) t[i] = tau(d, s, i);
    for (int e = 0; e < NE; ++e) {
      const int i = e % NPR, j = e / NPR;
      const double pc = s[60 + e], th = s[12 + e];
      const double pt = 1.0 / (fexp(-(th + flog(pc / (1.0 - pc)))) + 1.0);
      const double r = d.rt[j] - d.nt[j] * pt, dev = th - s[6 + i], t2 = t[i] * t[i];
      g[12 + e] = -dev / t2 + r;
      g[60 + e] = d.rc[j] / pc - (d.nc[j] - d.rc[j]) / (1.0 - pc) + r / (pc * (1.0 - pc));
      dmu[i] += dev / t2;
      dtau[i] += -1.0 / t[i] + dev * dev / (t2 * t[i]);
    }
    for (int i = 0; i < NPR; ++i) g[6 + i] = dmu[i];
    g[0] = dtau[0] / (2.0 * t[0]) + d_invgamma(s[0], 0.001, 0.001);
    g[1] = dtau[1] / (2.0 * t[1]);
    g[2] = dtau[2];
    g[3] = dtau[3] * (-d.s2_0 / (s[3] * s[3])) / (2.0 * t[3]);
    g[4] = dtau[4] * (-sqrt(d.s2_0) / (s[4] * s[4]));
    g[5] = dtau[5] / (2.0 * t[5]) - s[5] / (d.sd6 * d.sd6);
  }
  // every element is local: priors[k] / mu[i]: own prior + the 8 Normal terms of row k / i; theta[e]: its Normal term + its treated arm;
  // pc[e]: its Uniform term + both arms of cell e
  MCU_HD static bool elem_local(int) { return true; }
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool transform) {
    if (e < 12) {
      const int i = e < 6 ? e : e - 6;
      const double own = e < 6 ? prior_term(d, s, e, transform) : lp_uniform(s[e], -10.0, 10.0, transform);
      if (!(own > neg_inf())) return neg_inf();
      const double ti = tau(d, s, i);
      double lp = own;
      for (int j = 0; j < NTR; ++j) lp += lp_normal(s[12 + i + NPR * j], s[6 + i], ti);
      return lp;
    }
    if (e < 60) { const int q = e - 12; return lp_normal(s[e], s[6 + q % NPR], tau(d, s, q % NPR)) + rt_term(d, s, q); }
    const int q = e - 60;
    const double own = lp_uniform(s[e], 0.0, 1.0, transform);
    if (!(own > neg_inf())) return neg_inf();
    return own + rc_term(d, s, q) + rt_term(d, s, q);
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data& d, const double* s, double* out) {
    for (int i = 0; i < NPR; ++i) { out[i] = tau(d, s, i); out[NPR + i] = fexp(s[6 + i]); }   // OR = exp(mu): magnesium.jl:56-58
  }
  MCU_HD static int out_len(const Data&) { return 2 * NE; }   // rcx[48] then rtx[48]
  MCU_D static int out_dist(const Data& d, const double* s, int o, double& a, double& b) {
    if (o < NE) { a = d.nc[o / NPR]; b = s[60 + o]; }
    else { const int e = o - NE; const double pc = s[60 + e]; a = d.nt[e / NPR]; b = 1.0 / (fexp(-(s[12 + e] + flog(pc / (1.0 - pc)))) + 1.0); }
    return OUT_BINOMIAL;
  }
};

// =============================================================================== oxford
// doc/examples/oxford.jl:31-82 (data :4-28): 120 strata of a case-control study — 244 unobserved elements per chain, the largest state of
// the example corpus next to epil.  alpha, beta1, beta2 ~ Normal(0, 1000), s2 ~ InverseGamma(.001, .001), b[120] ~ Normal(0, sqrt(s2)),
// mu[120] ~ Normal(0, 1000); r0[i] ~ Binomial(n0[i], invlogit(mu[i])), r1[i] ~ Binomial(n1[i], invlogit(mu[i] + alpha + beta1 year[i] +
// beta2 (year[i]^2 - 22) + b[i])).  State: alpha, beta1, beta2, s2, b[120], mu[120]; monitored alpha, beta1, beta2, s2.
// kGradSamplers = false: the gradient-based samplers keep ~40 block-sized vectors per thread (NUTS tree stack), which at this state size
// needs the warp-per-chain layout of rats_warp.cu; the script's own scheme (AMWG + three multivariate Slice blocks) runs on the generic kernel.
struct OxfordModel : TplBase {
  static constexpr int D = 244, NN = 6, NF = 8, P = 4, K = 120;
  static constexpr bool kGradSamplers = false;
  struct Data { const double* r1; const double* n1; const double* r0; const double* n0; const double* year; const double* lc1; const double* lc0; };
  MCU_HD static int node_off(int n) { return n < 4 ? n : (n == 4 ? 4 : 4 + K); }
  MCU_HD static int node_len(int n) { return n < 4 ? 1 : K; }
  MCU_HD static int node_link(int n) { return n == 3 ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 4 ? 0x08u /* s2 */ : f == 6 ? 0x20u /* mu */ : f == 7 ? 0x37u /* alpha, beta1, beta2, b, mu */ : 0u; }
  MCU_HD static int mon_link(int j) { return j == 3 ? LINK_LOG : LINK_IDENT; }
  static const char* node_name(int n) { static const char* nm[] = {"alpha", "beta1", "beta2", "s2", "b", "mu"}; return nm[n]; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return "alpha\nbeta1\nbeta2\ns2"; }
  MCU_D static double eta1(const Data& d, const double* s, int i) {
    const double yr = d.year[i];
    return s[4 + K + i] + s[0] + s[1] * yr + s[2] * (yr * yr - 22.0) + s[4 + i];
  }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f < 3) return lp_normal(s[f], 0.0, 1000.0);
    if (f == 3) return lp_invgamma(s[3], 0.001, 0.001, ig001_c0(), transform);
    double lp = 0.0;
    if (f == 4) { const double sg = sqrt(s[3]); for (int i = 0; i < K; ++i) lp += lp_normal(s[4 + i], 0.0, sg); return lp; }
    if (f == 5) { for (int i = 0; i < K; ++i) lp += lp_normal(s[4 + K + i], 0.0, 1000.0); return lp; }
    if (f == 6) { for (int i = 0; i < K; ++i) lp += lp_binomial_logit(d.r0[i], d.n0[i], d.lc0[i], s[4 + K + i]); return lp; }
    for (int i = 0; i < K; ++i) lp += lp_binomial_logit(d.r1[i], d.n1[i], d.lc1[i], eta1(d, s, i));
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double s2 = s[3];
    double ga = 0, g1 = 0, g2 = 0, sbb = 0;
    for (int i = 0; i < K; ++i) {
      const double yr = d.year[i], q = yr * yr - 22.0;
      const double res1 = d.r1[i] - d.n1[i] * (1.0 / (fexp(-eta1(d, s, i)) + 1.0));
      const double res0 = d.r0[i] - d.n0[i] * (1.0 / (fexp(-s[4 + K + i]) + 1.0));
      ga += res1; g1 += res1 * yr; g2 += res1 * q;
      g[4 + i] = res1 - s[4 + i] / s2;
      g[4 + K + i] = res0 + res1 - s[4 + K + i] / 1e6;
      sbb += s[4 + i] * s[4 + i];
    }
    g[0] = ga - s[0] / 1e6; g[1] = g1 - s[1] / 1e6; g[2] = g2 - s[2] / 1e6;
    g[3] = -0.5 * (double)K / s2 + 0.5 * sbb / (s2 * s2) + d_invgamma(s2, 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 4; }   // b_i: its Normal term + case term i; mu_i: its Normal term + both arms of stratum i
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    if (e < 4 + K) { const int i = e - 4; return lp_normal(s[e], 0.0, sqrt(s[3])) + lp_binomial_logit(d.r1[i], d.n1[i], d.lc1[i], eta1(d, s, i)); }
    const int i = e - 4 - K;
    return lp_normal(s[e], 0.0, 1000.0) + lp_binomial_logit(d.r0[i], d.n0[i], d.lc0[i], s[e]) + lp_binomial_logit(d.r1[i], d.n1[i], d.lc1[i], eta1(d, s, i));
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data&, const double* s, double* out) { for (int j = 0; j < 4; ++j) out[j] = s[j]; }
  MCU_HD static int out_len(const Data&) { return 2 * K; }   // r0[120] then r1[120]
  MCU_D static int out_dist(const Data& d, const double* s, int o, double& a, double& b) {
    if (o < K) { a = d.n0[o]; b = 1.0 / (fexp(-s[4 + K + o]) + 1.0); }
    else { const int i = o - K; a = d.n1[i]; b = 1.0 / (fexp(-eta1(d, s, i)) + 1.0); }
    return OUT_BINOMIAL;
  }
};

// =============================================================================== epil
// doc/examples/epil.jl:33-111 (data :4-30): Poisson GLMM of seizure counts, 59 patients x 4 visits — 303 unobserved elements per chain.
// a0 and the five coefficients ~ Normal(0, 100), s2_b1, s2_b ~ InverseGamma(.001, .001), b1[59] ~ Normal(0, sqrt(s2_b1)),
// b[59 x 4] ~ Normal(0, sqrt(s2_b)) (column-major, patient fastest), y[i, j] ~ Poisson(exp(eta_ij)) with centred covariates (epil.jl:25-30).
// State: a0, alpha_Base, alpha_Trt, alpha_BT, alpha_Age, alpha_V4, s2_b1, s2_b, b1[59], b[236]; monitored: the five coefficients,
// alpha0 (Logical, epil.jl:85-91), s2_b1, s2_b.  cov = [x1[59] | x2[59] | x3[59] | x4[59] | x5[4] | the five covariate means] (host-derived).
struct EpilModel : TplBase {
  static constexpr int D = 303, NN = 10, NF = 11, P = 8, NPAT = 59, NV = 4, NOBS = 236;
  static constexpr bool kGradSamplers = false;
  struct Data { const double* y; const double* lgy1; const double* cov; };
  MCU_HD static int node_off(int n) { return n < 8 ? n : (n == 8 ? 8 : 8 + NPAT); }
  MCU_HD static int node_len(int n) { return n < 8 ? 1 : (n == 8 ? NPAT : NOBS); }
  MCU_HD static int node_link(int n) { return (n == 6 || n == 7) ? LINK_LOG : LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 8 ? 0x040u /* s2_b1 */ : f == 9 ? 0x080u /* s2_b */ : f == 10 ? 0x33Fu /* a0, coefficients, b1, b */ : 0u; }
  MCU_HD static int mon_link(int j) { return j < 5 ? LINK_IDENT : (j == 5 ? LINK_HEUR : LINK_LOG); }
  static const char* node_name(int n) { static const char* nm[] = {"a0", "alpha_Base", "alpha_Trt", "alpha_BT", "alpha_Age", "alpha_V4", "s2_b1", "s2_b", "b1", "b"}; return nm[n]; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return "alpha_Base\nalpha_Trt\nalpha_BT\nalpha_Age\nalpha_V4\nalpha0\ns2_b1\ns2_b"; }
  MCU_D static double eta(const Data& d, const double* s, int i, int j) {
    const double* c = d.cov;
    return s[0] + s[1] * c[i] + s[2] * c[NPAT + i] + s[3] * c[2 * NPAT + i] + s[4] * c[3 * NPAT + i] + s[5] * c[4 * NPAT + j] + s[8 + i] + s[8 + NPAT + i + NPAT * j];
  }
  MCU_D static double y_term(const Data& d, const double* s, int i, int j) { const int o = i + NPAT * j; return lp_poisson(d.y[o], d.lgy1[o], fexp(eta(d, s, i, j))); }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool transform) {
    if (f < 6) return lp_normal(s[f], 0.0, 100.0);
    if (f < 8) return lp_invgamma(s[f], 0.001, 0.001, ig001_c0(), transform);
    double lp = 0.0;
    if (f == 8) { const double sg = sqrt(s[6]); for (int i = 0; i < NPAT; ++i) lp += lp_normal(s[8 + i], 0.0, sg); return lp; }
    if (f == 9) { const double sg = sqrt(s[7]); for (int o = 0; o < NOBS; ++o) lp += lp_normal(s[8 + NPAT + o], 0.0, sg); return lp; }
    for (int j = 0; j < NV; ++j) for (int i = 0; i < NPAT; ++i) lp += y_term(d, s, i, j);
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    const double* c = d.cov;
    double ga[6] = {0, 0, 0, 0, 0, 0}, sb1 = 0, sb = 0;
    for (int i = 0; i < NPAT; ++i) g[8 + i] = 0.0;
    for (int j = 0; j < NV; ++j) for (int i = 0; i < NPAT; ++i) {
      const int o = i + NPAT * j;
      const double res = d.y[o] - fexp(eta(d, s, i, j));
      ga[0] += res; ga[1] += res * c[i]; ga[2] += res * c[NPAT + i]; ga[3] += res * c[2 * NPAT + i]; ga[4] += res * c[3 * NPAT + i]; ga[5] += res * c[4 * NPAT + j];
      g[8 + i] += res;
      g[8 + NPAT + o] = res - s[8 + NPAT + o] / s[7];
      sb += s[8 + NPAT + o] * s[8 + NPAT + o];
    }
    for (int i = 0; i < NPAT; ++i) { g[8 + i] -= s[8 + i] / s[6]; sb1 += s[8 + i] * s[8 + i]; }
    for (int k = 0; k < 6; ++k) g[k] = ga[k] - s[k] / 1e4;
    g[6] = -0.5 * (double)NPAT / s[6] + 0.5 * sb1 / (s[6] * s[6]) + d_invgamma(s[6], 0.001, 0.001);
    g[7] = -0.5 * (double)NOBS / s[7] + 0.5 * sb / (s[7] * s[7]) + d_invgamma(s[7], 0.001, 0.001);
  }
  MCU_HD static bool elem_local(int e) { return e >= 8; }   // b1_i: its Normal term + patient i's four visits; b_ij: its Normal term + one visit
  MCU_D static double elem_terms(const Data& d, const double* s, int e, bool) {
    if (e < 8 + NPAT) {
      const int i = e - 8;
      double lp = lp_normal(s[e], 0.0, sqrt(s[6]));
      for (int j = 0; j < NV; ++j) lp += y_term(d, s, i, j);
      return lp;
    }
    const int o = e - 8 - NPAT;
    return lp_normal(s[e], 0.0, sqrt(s[7])) + y_term(d, s, o % NPAT, o / NPAT);
  }
  MCU_HD static bool has_gibbs(int) { return false; }
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data& d, const double* s, double* out) {
    const double* bar = d.cov + 4 * NPAT + NV;
    for (int k = 0; k < 5; ++k) out[k] = s[1 + k];
    out[5] = s[0] - s[1] * bar[0] - s[2] * bar[1] - s[3] * bar[2] - s[4] * bar[3] - s[5] * bar[4];   // alpha0: epil.jl:85-91
    out[6] = s[6]; out[7] = s[7];
  }
  MCU_HD static int out_len(const Data&) { return NOBS; }
  MCU_D static int out_dist(const Data& d, const double* s, int o, double& a, double& b) { a = fexp(eta(d, s, o % NPAT, o / NPAT)); b = 0.0; return OUT_POISSON; }
};

// =============================================================================== glm (CUDA-core form)
// y_i ~ Bernoulli(invlogit(X[i,:] . beta)), beta ~ MvNormal(d, sqrt(1000)).  This per-chain form is the
// small-N path used by the generic kernel; the large-N path is the fused tensor-core kernel.
template <int DMAX>
struct GlmModel : TplBase {
  static constexpr int D = DMAX, NN = 1, NF = 2, P = DMAX;
  // family: 0 Bernoulli / logit, 1 Poisson / log, 2 Normal / identity with known sd sigma
  struct Data { const double* X; const double* y; int N; int d; int family; double sigma; };
  MCU_HD static int node_off(int) { return 0; }
  MCU_HD static int node_len(int) { return DMAX; }
  MCU_HD static int node_link(int) { return LINK_IDENT; }
  MCU_HD static uint32_t parents(int f) { return f == 1 ? 0x1u : 0u; }
  MCU_HD static int mon_link(int) { return LINK_IDENT; }
  static const char* node_name(int) { return "beta"; }
  static const char* state_names() { return nullptr; }
  static const char* monitor_names() { return nullptr; }
  MCU_NOINL static double factor(const Data& d, const double* s, int f, bool) {
    if (f == 0) {
      double sq = 0; for (int j = 0; j < d.d; ++j) { if (!isfinite(s[j])) return neg_inf(); sq += s[j] * s[j]; }
      return lp_isonormal(sq, (double)d.d, sqrt(1000.0));
    }
    double lp = 0.0;
    for (int i = 0; i < d.N; ++i) {
      double eta = 0; for (int j = 0; j < d.d; ++j) eta += d.X[(size_t)i * d.d + j] * s[j];
      lp += d.family == 1 ? lp_poisson(d.y[i], lgamma(d.y[i] + 1.0), fexp(eta)) : d.family == 2 ? lp_normal(d.y[i], eta, d.sigma) : lp_bernoulli_logit(d.y[i], eta);
    }
    return lp;
  }
  MCU_NOINL static void joint_grad(const Data& d, const double* s, double* g) {
    for (int j = 0; j < d.d; ++j) g[j] = -s[j] / 1000.0;
    for (int i = 0; i < d.N; ++i) {
      double eta = 0; for (int j = 0; j < d.d; ++j) eta += d.X[(size_t)i * d.d + j] * s[j];
      const double r = d.family == 1 ? d.y[i] - fexp(eta) : d.family == 2 ? (d.y[i] - eta) / (d.sigma * d.sigma) : d.y[i] - 1.0 / (fexp(-eta) + 1.0);
      for (int j = 0; j < d.d; ++j) g[j] += r * d.X[(size_t)i * d.d + j];
    }
  }
  MCU_HD static bool elem_local(int) { return false; }   // no element whose move touches only a few terms
  MCU_D static double elem_terms(const Data&, const double*, int, bool) { return 0.0; }
  MCU_HD static bool has_gibbs(int) { return false; }   // no conjugate full conditional registered for this template (MCU_GIBBS)
  template <class R, class G> MCU_D static void gibbs(const Data&, double*, int, R&, G) {}
  MCU_D static void monitor(const Data& d, const double* s, double* out) { for (int j = 0; j < d.d; ++j) out[j] = s[j]; }
  MCU_HD static int out_len(const Data& d) { return d.N; }
  MCU_D static int out_dist(const Data& d, const double* s, int i, double& a, double& b) {
    double eta = 0; for (int j = 0; j < d.d; ++j) eta += d.X[(size_t)i * d.d + j] * s[j];
    if (d.family == 1) { a = fexp(eta); b = 0.0; return OUT_POISSON; }
    if (d.family == 2) { a = eta; b = d.sigma; return OUT_NORMAL; }
    a = 1.0 / (fexp(-eta) + 1.0); b = 0.0; return OUT_BERNOULLI;
  }
};

}  // namespace mcu
