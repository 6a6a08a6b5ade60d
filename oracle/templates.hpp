// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
//
// templates.hpp — the model scripts on the hot path, restated node by node.
//   line  : doc/tutorial/line.jl:5-25, data :69-73
//   seeds : doc/examples/seeds.jl:16-56, data :4-12
//   rats  : doc/examples/rats.jl:49-97, data :4-45
//   pumps : doc/examples/pumps.jl:12-39, data :4-9
//   glm   : synthetic Bernoulli-logit regression (no reference file; structure of seeds.jl:18-28
//           with the prior of line.jl:18)
// Node order inside each template is a valid topological order and stands in for
// keys(m, :dependent) (model.jl:112-120), whose exact order in the reference depends on Dict
// iteration order; it only fixes column order and floating-point summation order.
// The analytic joint gradients are hand-derived (the reference has none: it differentiates
// numerically, simulation.jl:47-51) and are validated against the FD gradient in tests.
#pragma once
#include "model.hpp"

namespace orc {

enum { TPL_LINE = 0, TPL_SEEDS = 1, TPL_RATS = 2, TPL_PUMPS = 3, TPL_GLM = 4, TPL_SURGICAL = 5, TPL_DYES = 6, TPL_SALM = 7, TPL_EQUIV = 8, TPL_BLOCKER = 9, TPL_STACKS = 10, TPL_MAGNESIUM = 11, TPL_OXFORD = 12, TPL_EPIL = 13 };

inline Node make_node(const std::string& name, bool stochastic, int len, bool scalar, bool monitored,
                      bool observed = false) {
  Node n; n.name = name; n.stochastic = stochastic; n.len = len; n.scalar = scalar; n.observed = observed;
  n.value.assign(len, stochastic ? 0.0 : NAN);
  if (monitored) for (int i = 0; i < len; ++i) n.monitor.push_back(i);  // setmonitor!: dependent.jl:33-51
  return n;
}

inline double ig_dlogpdf(double a, double th, double x) { return -(a + 1.0) / x + th / (x * x); }

// ------------------------------------------------------------------------------------------
inline Model make_line() {
  Model m; m.template_id = TPL_LINE;
  m.inputs["x"] = {1, 2, 3, 4, 5};
  m.inputs["y"] = {1, 3, 3, 3, 5};
  // beta = Stochastic(1, () -> MvNormal(2, sqrt(1000)))
  { Node n = make_node("beta", true, 2, false, true);
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::MVNORMAL_ISO; s.distr.mu.assign(2, 0.0); s.distr.sigma = std::sqrt(1000.0); };
    m.nodes.push_back(n); }
  // s2 = Stochastic(() -> InverseGamma(0.001, 0.001))
  { Node n = make_node("s2", true, 1, true, true);
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
    m.nodes.push_back(n); }
  // mu = Logical(1, (xmat, beta) -> xmat * beta, false) ; xmat = [ones(5) x]  (line.jl:73)
  { Node n = make_node("mu", false, 5, false, false);
    n.sources = {0};
    n.eval = [](const Model& mm, Node& l) {
      const auto& x = mm.in("x"); const auto& b = mm.val(0);
      l.value.resize(x.size());
      for (size_t i = 0; i < x.size(); ++i) l.value[i] = 1.0 * b[0] + x[i] * b[1];
    };
    m.nodes.push_back(n); }
  // y = Stochastic(1, (mu, s2) -> MvNormal(mu, sqrt(s2)), false)
  { Node n = make_node("y", true, 5, false, false, true);
    n.sources = {2, 1};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::MVNORMAL_ISO; s.distr.mu = mm.val(2); s.distr.sigma = std::sqrt(mm.val(1)[0]); };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {
    const auto& x = mm.in("x"); const auto& y = mm.in("y");
    const auto& b = mm.val(0); double s2 = mm.val(1)[0];
    double sr = 0, sxr = 0, srr = 0;
    for (size_t i = 0; i < x.size(); ++i) { double r = y[i] - b[0] - b[1] * x[i]; sr += r; sxr += x[i] * r; srr += r * r; }
    g[0] = sr / s2 - b[0] / 1000.0;
    g[1] = sxr / s2 - b[1] / 1000.0;
    g[2] = -0.5 * (double)x.size() / s2 + 0.5 * srr / (s2 * s2) + ig_dlogpdf(0.001, 0.001, s2);
  };
  // the tutorial's Gibbs_beta / Gibbs_s2 (doc/tutorial/line.jl:27-45): conjugate full conditionals; same draw order as the device template
  m.gibbs = [](Model& mm, int node, Rng& rng) {
    const auto& x = mm.in("x"); const auto& y = mm.in("y");
    const size_t N = y.size();
    if (node == 0) {
      double sx = 0, sxx = 0, sy = 0, sxy = 0;
      for (size_t i = 0; i < N; ++i) { sx += x[i]; sxx += x[i] * x[i]; sy += y[i]; sxy += x[i] * y[i]; }
      const double s2 = mm.val(1)[0];
      const double a11 = (double)N / s2 + 1.0 / 1000.0, a12 = sx / s2, a22 = sxx / s2 + 1.0 / 1000.0;
      const double det = a11 * a22 - a12 * a12;
      const double S11 = a22 / det, S12 = -a12 / det, S22 = a11 / det;
      const double r1 = sy / s2, r2 = sxy / s2;
      const double m1 = S11 * r1 + S12 * r2, m2 = S12 * r1 + S22 * r2;
      const double l11 = std::sqrt(S11), l21 = S12 / l11, l22 = std::sqrt(S22 - l21 * l21);
      const double z1 = rng.normal(), z2 = rng.normal();
      mm.nodes[0].value[0] = m1 + l11 * z1;
      mm.nodes[0].value[1] = m2 + l21 * z1 + l22 * z2;
      return true;
    }
    if (node == 1) {
      const auto& be = mm.val(0);
      double ss = 0.0;
      for (size_t i = 0; i < N; ++i) { const double r = y[i] - (1.0 * be[0] + x[i] * be[1]); ss += r * r; }
      const double a = (double)N / 2.0 + 0.001, b = ss / 2.0 + 0.001;
      mm.nodes[1].value[0] = b / rgamma_mt(a, rng);
      return true;
    }
    return false;
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
inline Model make_seeds() {
  Model m; m.template_id = TPL_SEEDS;
  m.inputs["r"] = {10, 23, 23, 26, 17, 5, 53, 55, 32, 46, 10, 8, 10, 8, 23, 0, 3, 22, 15, 32, 3};
  m.inputs["n"] = {39, 62, 81, 51, 39, 6, 74, 72, 51, 79, 13, 16, 30, 28, 45, 4, 12, 41, 30, 51, 7};
  m.inputs["x1"] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  m.inputs["x2"] = {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1};
  const char* an[4] = {"alpha0", "alpha1", "alpha2", "alpha12"};
  for (int k = 0; k < 4; ++k) {   // alpha* = Stochastic(() -> Normal(0, 1000))
    Node n = make_node(an[k], true, 1, true, true);
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
    m.nodes.push_back(n);
  }
  { Node n = make_node("s2", true, 1, true, true);   // InverseGamma(0.001, 0.001)
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
    m.nodes.push_back(n); }
  { Node n = make_node("b", true, 21, false, false);  // b = Stochastic(1, s2 -> Normal(0, sqrt(s2)), false)
    n.sources = {4};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, std::sqrt(mm.val(4)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("r", true, 21, false, false, true);  // seeds.jl:18-28
    n.sources = {0, 1, 2, 3, 5};
    n.eval = [](const Model& mm, Node& s) {
      const auto& x1 = mm.in("x1"); const auto& x2 = mm.in("x2"); const auto& nn = mm.in("n");
      double a0 = mm.val(0)[0], a1 = mm.val(1)[0], a2 = mm.val(2)[0], a12 = mm.val(3)[0];
      const auto& b = mm.val(5);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(nn.size());
      for (size_t i = 0; i < nn.size(); ++i) {
        double p = invlogit(a0 + a1 * x1[i] + a2 * x2[i] + a12 * x1[i] * x2[i] + b[i]);
        s.distr.arr[i] = {D_BINOMIAL, nn[i], p};
      }
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {
    const auto& x1 = mm.in("x1"); const auto& x2 = mm.in("x2"); const auto& nn = mm.in("n"); const auto& r = mm.in("r");
    double a0 = mm.val(0)[0], a1 = mm.val(1)[0], a2 = mm.val(2)[0], a12 = mm.val(3)[0], s2 = mm.val(4)[0];
    const auto& b = mm.val(5);
    double g0 = 0, g1 = 0, g2 = 0, g12 = 0, sbb = 0;
    for (size_t i = 0; i < nn.size(); ++i) {
      double p = invlogit(a0 + a1 * x1[i] + a2 * x2[i] + a12 * x1[i] * x2[i] + b[i]);
      double de = r[i] - nn[i] * p;
      g0 += de; g1 += x1[i] * de; g2 += x2[i] * de; g12 += x1[i] * x2[i] * de;
      g[5 + i] = de - b[i] / s2;
      sbb += b[i] * b[i];
    }
    g[0] = g0 - a0 / 1e6; g[1] = g1 - a1 / 1e6; g[2] = g2 - a2 / 1e6; g[3] = g12 - a12 / 1e6;
    g[4] = -0.5 * (double)nn.size() / s2 + 0.5 * sbb / (s2 * s2) + ig_dlogpdf(0.001, 0.001, s2);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
inline Model make_rats() {
  Model m; m.template_id = TPL_RATS;
  static const double Y[150] = {
    151, 199, 246, 283, 320, 145, 199, 249, 293, 354, 147, 214, 263, 312, 328, 155, 200, 237, 272, 297,
    135, 188, 230, 280, 323, 159, 210, 252, 298, 331, 141, 189, 231, 275, 305, 159, 201, 248, 297, 338,
    177, 236, 285, 350, 376, 134, 182, 220, 260, 296, 160, 208, 261, 313, 352, 143, 188, 220, 273, 314,
    154, 200, 244, 289, 325, 171, 221, 270, 326, 358, 163, 216, 242, 281, 312, 160, 207, 248, 288, 324,
    142, 187, 234, 280, 316, 156, 203, 243, 283, 317, 157, 212, 259, 307, 336, 152, 203, 246, 286, 321,
    154, 205, 253, 298, 334, 139, 190, 225, 267, 302, 146, 191, 229, 272, 302, 157, 211, 250, 285, 323,
    132, 185, 237, 286, 331, 160, 207, 257, 303, 345, 169, 216, 261, 295, 333, 157, 205, 248, 289, 316,
    137, 180, 219, 258, 291, 153, 200, 244, 286, 324};
  // rats.jl:38-45: the script's :y is a 30x5 Julia matrix literal written row by row and the
  // node is a 150-vector; y[k] pairs with rat[k] = div(k-1,5)+1, week[k] = (k-1)%5+1.
  // (Julia's vec() of that matrix would be column-major; upstream Mamba passes y as the
  // row-wise flat vector — SURVEY.md §8d config 3 "y flat 150-vector" — which is what is used.)
  std::vector<double> y(Y, Y + 150), rat(150), Xm(150);
  const double xs[5] = {8.0, 15.0, 22.0, 29.0, 36.0};
  double xbar = 22.0;
  for (int k = 0; k < 150; ++k) { rat[k] = k / 5; Xm[k] = xs[k % 5] - xbar; }
  m.inputs["y"] = y; m.inputs["rat"] = rat; m.inputs["Xm"] = Xm; m.inputs["xbar"] = {xbar};
  auto prior_norm = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
  auto prior_ig = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
  { Node n = make_node("mu_alpha", true, 1, true, false); n.eval = prior_norm; m.nodes.push_back(n); }   // 0
  { Node n = make_node("mu_beta", true, 1, true, true); n.eval = prior_norm; m.nodes.push_back(n); }     // 1
  { Node n = make_node("alpha0", false, 1, true, true);                                                   // 2
    n.sources = {0, 1};
    n.eval = [](const Model& mm, Node& l) { l.value.assign(1, mm.val(0)[0] - mm.in("xbar")[0] * mm.val(1)[0]); };
    m.nodes.push_back(n); }
  { Node n = make_node("s2_alpha", true, 1, true, false); n.eval = prior_ig; m.nodes.push_back(n); }     // 3
  { Node n = make_node("s2_beta", true, 1, true, false); n.eval = prior_ig; m.nodes.push_back(n); }      // 4
  { Node n = make_node("s2_c", true, 1, true, true); n.eval = prior_ig; m.nodes.push_back(n); }          // 5
  { Node n = make_node("alpha", true, 30, false, false);                                                 // 6
    n.sources = {0, 3};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, mm.val(0)[0], std::sqrt(mm.val(3)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("beta", true, 30, false, false);                                                  // 7
    n.sources = {1, 4};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, mm.val(1)[0], std::sqrt(mm.val(4)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 150, false, false, true);                                              // 8
    n.sources = {6, 7, 5};
    n.eval = [](const Model& mm, Node& s) {
      const auto& rat = mm.in("rat"); const auto& Xm = mm.in("Xm");
      const auto& a = mm.val(6); const auto& b = mm.val(7);
      s.distr.form = Distr::MVNORMAL_ISO; s.distr.mu.resize(rat.size());
      for (size_t k = 0; k < rat.size(); ++k) { int i = (int)rat[k]; s.distr.mu[k] = a[i] + b[i] * Xm[k]; }
      s.distr.sigma = std::sqrt(mm.val(5)[0]);
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {
    // state record: mu_alpha, mu_beta, s2_alpha, s2_beta, s2_c, alpha[30], beta[30]
    const auto& rat = mm.in("rat"); const auto& Xm = mm.in("Xm"); const auto& y = mm.in("y");
    double mua = mm.val(0)[0], mub = mm.val(1)[0], s2a = mm.val(3)[0], s2b = mm.val(4)[0], s2c = mm.val(5)[0];
    const auto& a = mm.val(6); const auto& b = mm.val(7);
    for (int i = 0; i < 60; ++i) g[5 + i] = 0.0;
    double see = 0;
    for (size_t k = 0; k < rat.size(); ++k) {
      int i = (int)rat[k]; double e = y[k] - (a[i] + b[i] * Xm[k]);
      g[5 + i] += e / s2c; g[35 + i] += e * Xm[k] / s2c; see += e * e;
    }
    double sa = 0, saa = 0, sb = 0, sbb = 0;
    for (int i = 0; i < 30; ++i) {
      double da = a[i] - mua, db = b[i] - mub;
      g[5 + i] -= da / s2a; g[35 + i] -= db / s2b;
      sa += da; saa += da * da; sb += db; sbb += db * db;
    }
    g[0] = sa / s2a - mua / 1e6;
    g[1] = sb / s2b - mub / 1e6;
    g[2] = -15.0 / s2a + 0.5 * saa / (s2a * s2a) + ig_dlogpdf(0.001, 0.001, s2a);
    g[3] = -15.0 / s2b + 0.5 * sbb / (s2b * s2b) + ig_dlogpdf(0.001, 0.001, s2b);
    g[4] = -75.0 / s2c + 0.5 * see / (s2c * s2c) + ig_dlogpdf(0.001, 0.001, s2c);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
inline Model make_pumps() {
  Model m; m.template_id = TPL_PUMPS;
  m.inputs["y"] = {5, 1, 5, 14, 3, 19, 1, 1, 4, 22};
  m.inputs["t"] = {94.3, 15.7, 62.9, 126, 5.24, 31.4, 1.05, 1.05, 2.1, 10.5};
  { Node n = make_node("alpha", true, 1, true, true);   // Exponential(1.0)
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_EXPONENTIAL, 1.0, 0.0}; };
    m.nodes.push_back(n); }
  { Node n = make_node("beta", true, 1, true, true);    // Gamma(0.1, 1.0)
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_GAMMA, 0.1, 1.0}; };
    m.nodes.push_back(n); }
  { Node n = make_node("theta", true, 10, false, true); // (alpha, beta) -> Gamma(alpha, 1 / beta)
    n.sources = {0, 1};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_GAMMA, mm.val(0)[0], 1.0 / mm.val(1)[0]}; };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 10, false, false, true);  // Poisson(theta[i] * t[i])
    n.sources = {2};
    n.eval = [](const Model& mm, Node& s) {
      const auto& t = mm.in("t"); const auto& th = mm.val(2);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(t.size());
      for (size_t i = 0; i < t.size(); ++i) s.distr.arr[i] = {D_POISSON, th[i] * t[i], 0.0};
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {
    const auto& t = mm.in("t"); const auto& y = mm.in("y");
    double al = mm.val(0)[0], be = mm.val(1)[0]; const auto& th = mm.val(2);
    double slog = 0, sth = 0; double N = (double)t.size();
    for (size_t i = 0; i < t.size(); ++i) {
      g[2 + i] = y[i] / th[i] - t[i] + (al - 1.0) / th[i] - be;
      slog += std::log(th[i]); sth += th[i];
    }
    g[0] = N * std::log(be) + slog - N * digamma(al) - 1.0;
    g[1] = N * al / be - sth + (0.1 - 1.0) / be - 1.0;
  };
  // conjugate full conditionals (BASELINE.json configs[4] "Gibbs + AMWG"):
  //   theta_i | . ~ Gamma(alpha + y_i, 1 / (beta + t_i)),   beta | . ~ Gamma(0.1 + N alpha, 1 / (1 + sum_i theta_i))
  m.gibbs = [](Model& mm, int node, Rng& rng) {
    const auto& t = mm.in("t"); const auto& y = mm.in("y");
    const double al = mm.val(0)[0], be = mm.val(1)[0];
    if (node == 2) {
      for (size_t i = 0; i < t.size(); ++i) mm.nodes[2].value[i] = rgamma_mt(al + y[i], rng) / (be + t[i]);
      return true;
    }
    if (node == 1) {
      double sth = 0; for (double v : mm.val(2)) sth += v;
      mm.nodes[1].value[0] = rgamma_mt(0.1 + (double)t.size() * al, rng) / (1.0 + sth);
      return true;
    }
    return false;
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// y_i ~ Bernoulli(invlogit(X[i,:] . beta)), beta ~ MvNormal(d, sqrt(1000)).  Inputs "X" (N*d,
// row-major) and "y" must be supplied; d is taken from inputs["d"].
inline Model make_glm(int d) {
  Model m; m.template_id = TPL_GLM;
  m.inputs["d"] = {(double)d};
  { Node n = make_node("beta", true, d, false, true);
    n.eval = [d](const Model&, Node& s) { s.distr.form = Distr::MVNORMAL_ISO; s.distr.mu.assign(d, 0.0); s.distr.sigma = std::sqrt(1000.0); };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 0, false, false, true);
    n.sources = {0};
    n.eval = [d](const Model& mm, Node& s) {
      const auto& X = mm.in("X"); const auto& be = mm.val(0);
      size_t N = X.size() / d;
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(N);
      // GLM family (north_star "a GLM family"): inputs["family"] 0 = Bernoulli / logit (default), 1 = Poisson / log,
      // 2 = Normal / identity with known sd inputs["sigma"] (default 1)
      const int fam = mm.inputs.count("family") ? (int)mm.in("family")[0] : 0;
      const double sg = mm.inputs.count("sigma") ? mm.in("sigma")[0] : 1.0;
      for (size_t i = 0; i < N; ++i) {
        double eta = 0; for (int j = 0; j < d; ++j) eta += X[i * d + j] * be[j];
        if (fam == 1) s.distr.arr[i] = {D_POISSON, std::exp(eta), 0.0};
        else if (fam == 2) s.distr.arr[i] = {D_NORMAL, eta, sg};
        else s.distr.arr[i] = {D_BERNOULLI, invlogit(eta), 0.0};
      }
    };
    m.nodes.push_back(n); }
  m.joint_grad = [d](const Model& mm, std::vector<double>& g) {
    const auto& X = mm.in("X"); const auto& y = mm.in("y"); const auto& be = mm.val(0);
    size_t N = X.size() / d;
    for (int j = 0; j < d; ++j) g[j] = -be[j] / 1000.0;
    for (size_t i = 0; i < N; ++i) {
      double eta = 0; for (int j = 0; j < d; ++j) eta += X[i * d + j] * be[j];
      const int fam = mm.inputs.count("family") ? (int)mm.in("family")[0] : 0;
      const double sg = mm.inputs.count("sigma") ? mm.in("sigma")[0] : 1.0;
      double r = fam == 1 ? y[i] - std::exp(eta) : fam == 2 ? (y[i] - eta) / (sg * sg) : y[i] - invlogit(eta);
      for (int j = 0; j < d; ++j) g[j] += r * X[i * d + j];
    }
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// magnesium: doc/examples/magnesium.jl:21-82 (data :4-17) — meta-analysis of 8 trials under six priors for the between-trial sd tau.
// The one example whose parameter nodes carry BOUNDED distributions: pc ~ Uniform(0, 1), mu ~ Uniform(-10, 10) (sampled by AMWG on the
// two-sided link logit((x - a) / (b - a)), transformdistribution.jl:6-48), priors = [InverseGamma, Uniform(0, 50) x 2, Uniform(0, 1) x 2,
// Truncated(Normal(0, sqrt(s2_0 / erf(0.75))), 0, Inf)].  6 x 8 matrices are stored column-major (prior index i fastest), as Julia does.
// Node order: priors, mu, tau (Logical), OR (Logical), theta, pc, rcx (observed), rtx (observed); monitored: tau[6], OR[6].
inline Model make_magnesium() {
  Model m; m.template_id = TPL_MAGNESIUM;
  m.inputs["rt"] = {1, 9, 2, 1, 10, 1, 1, 90};
  m.inputs["nt"] = {40, 135, 200, 48, 150, 59, 25, 1159};
  m.inputs["rc"] = {2, 23, 7, 1, 8, 9, 3, 118};
  m.inputs["nc"] = {36, 135, 200, 46, 148, 56, 23, 1157};
  {
    const auto &rt = m.inputs["rt"], &nt = m.inputs["nt"], &rc = m.inputs["rc"], &nc = m.inputs["nc"];
    std::vector<double> rtx(48), rcx(48);
    double sinv = 0.0;
    for (int j = 0; j < 8; ++j) {
      for (int i = 0; i < 6; ++i) { rtx[i + 6 * j] = rt[j]; rcx[i + 6 * j] = rc[j]; }                      // magnesium.jl:11-12
      const double s2 = 1.0 / (rt[j] + 0.5) + 1.0 / (nt[j] - rt[j] + 0.5) + 1.0 / (rc[j] + 0.5) + 1.0 / (nc[j] - rc[j] + 0.5);   // :13-16
      sinv += 1.0 / s2;
    }
    m.inputs["rtx"] = rtx; m.inputs["rcx"] = rcx;
    m.inputs["s2_0"] = {1.0 / (sinv / 8.0)};                                                               // :17
  }
  { Node n = make_node("priors", true, 6, false, false);                                                   // 0
    n.eval = [](const Model& mm, Node& s) {
      const double s2_0 = mm.in("s2_0")[0];
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(6);
      s.distr.arr[0] = {D_INVGAMMA, 0.001, 0.001};
      s.distr.arr[1] = {D_UNIFORM, 0.0, 50.0}; s.distr.arr[2] = {D_UNIFORM, 0.0, 50.0};
      s.distr.arr[3] = {D_UNIFORM, 0.0, 1.0}; s.distr.arr[4] = {D_UNIFORM, 0.0, 1.0};
      UDist t; t.k = D_TRUNCNORMAL; t.a = 0.0; t.b = std::sqrt(s2_0 / std::erf(0.75)); t.lo = 0.0; t.hi = INFINITY;
      s.distr.arr[5] = t;
    };
    m.nodes.push_back(n); }
  { Node n = make_node("mu", true, 6, false, false);                                                       // 1
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_UNIFORM, -10.0, 10.0}; };
    m.nodes.push_back(n); }
  { Node n = make_node("tau", false, 6, false, true);                                                      // 2: Logical
    n.sources = {0};
    n.eval = [](const Model& mm, Node& l) {
      const auto& p = mm.val(0); const double s2_0 = mm.in("s2_0")[0];
      l.value = {std::sqrt(p[0]), std::sqrt(p[1]), p[2], std::sqrt(s2_0 * (1.0 / p[3] - 1.0)), std::sqrt(s2_0) * (1.0 / p[4] - 1.0), std::sqrt(p[5])};
    };
    m.nodes.push_back(n); }
  { Node n = make_node("OR", false, 6, false, true);                                                       // 3: Logical, exp(mu)
    n.sources = {1};
    n.eval = [](const Model& mm, Node& l) { l.value.resize(6); for (int i = 0; i < 6; ++i) l.value[i] = std::exp(mm.val(1)[i]); };
    m.nodes.push_back(n); }
  { Node n = make_node("theta", true, 48, false, false);                                                   // 4: Normal(mu[i], tau[i])
    n.sources = {1, 2};
    n.eval = [](const Model& mm, Node& s) {
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(48);
      for (int j = 0; j < 8; ++j) for (int i = 0; i < 6; ++i) s.distr.arr[i + 6 * j] = {D_NORMAL, mm.val(1)[i], mm.val(2)[i]};
    };
    m.nodes.push_back(n); }
  { Node n = make_node("pc", true, 48, false, false);                                                      // 5
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_UNIFORM, 0.0, 1.0}; };
    m.nodes.push_back(n); }
  { Node n = make_node("rcx", true, 48, false, false, true);                                               // 6: Binomial(nc[j], pc[i, j])
    n.sources = {5};
    n.eval = [](const Model& mm, Node& s) {
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(48);
      for (int j = 0; j < 8; ++j) for (int i = 0; i < 6; ++i) s.distr.arr[i + 6 * j] = {D_BINOMIAL, mm.in("nc")[j], mm.val(5)[i + 6 * j]};
    };
    m.nodes.push_back(n); }
  { Node n = make_node("rtx", true, 48, false, false, true);                                               // 7: Binomial(nt[j], invlogit(theta + logit(pc)))
    n.sources = {5, 4};
    n.eval = [](const Model& mm, Node& s) {
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(48);
      for (int e = 0; e < 48; ++e) {
        const double phi = logit(mm.val(5)[e]);
        s.distr.arr[e] = {D_BINOMIAL, mm.in("nt")[e / 6], invlogit(mm.val(4)[e] + phi)};
      }
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: priors[6], mu[6], theta[48], pc[48]
    const auto &rt = mm.in("rt"), &nt = mm.in("nt"), &rc = mm.in("rc"), &nc = mm.in("nc");
    const auto &p = mm.val(0), &mu = mm.val(1), &tau = mm.val(2), &th = mm.val(4), &pc = mm.val(5);
    const double s2_0 = mm.in("s2_0")[0];
    double dmu[6] = {0, 0, 0, 0, 0, 0}, dtau[6] = {0, 0, 0, 0, 0, 0};
    for (int e = 0; e < 48; ++e) {
      const int i = e % 6, j = e / 6;
      const double pt = invlogit(th[e] + logit(pc[e]));
      const double r = rt[j] - nt[j] * pt, dev = th[e] - mu[i], t2 = tau[i] * tau[i];
      g[12 + e] = -dev / t2 + r;
      g[60 + e] = rc[j] / pc[e] - (nc[j] - rc[j]) / (1.0 - pc[e]) + r / (pc[e] * (1.0 - pc[e]));
      dmu[i] += dev / t2;
      dtau[i] += -1.0 / tau[i] + dev * dev / (t2 * tau[i]);
    }
    for (int i = 0; i < 6; ++i) g[6 + i] = dmu[i];
    const double sg5 = std::sqrt(s2_0 / std::erf(0.75));
    g[0] = dtau[0] / (2.0 * tau[0]) + ig_dlogpdf(0.001, 0.001, p[0]);
    g[1] = dtau[1] / (2.0 * tau[1]);
    g[2] = dtau[2];
    g[3] = dtau[3] * (-s2_0 / (p[3] * p[3])) / (2.0 * tau[3]);
    g[4] = dtau[4] * (-std::sqrt(s2_0) / (p[4] * p[4]));
    g[5] = dtau[5] / (2.0 * tau[5]) - p[5] / (sg5 * sg5);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// oxford: doc/examples/oxford.jl:31-82 (data :4-28) — 120 strata of a case-control study.  244 unobserved elements:
// alpha, beta1, beta2 ~ Normal(0, 1000), s2 ~ InverseGamma(.001, .001), b[120] ~ Normal(0, sqrt(s2)), mu[120] ~ Normal(0, 1000);
// r0[i] ~ Binomial(n0[i], invlogit(mu[i])), r1[i] ~ Binomial(n1[i], invlogit(mu[i] + alpha + beta1 year[i] + beta2 (year[i]^2 - 22) + b[i])).
// Monitored: alpha, beta1, beta2, s2.
inline Model make_oxford() {
  Model m; m.template_id = TPL_OXFORD;
  m.inputs["r1"] = {3, 5, 2, 7, 7, 2, 5, 3, 5, 11, 6, 6, 11, 4, 4, 2, 8, 8, 6, 5, 15, 4, 9, 9, 4, 12, 8, 8, 6, 8,
      12, 4, 7, 16, 12, 9, 4, 7, 8, 11, 5, 12, 8, 17, 9, 3, 2, 7, 6, 5, 11, 14, 13, 8, 6, 4, 8, 4, 8, 7,
      15, 15, 9, 9, 5, 6, 3, 9, 12, 14, 16, 17, 8, 8, 9, 5, 9, 11, 6, 14, 21, 16, 6, 9, 8, 9, 8, 4, 11, 11,
      6, 9, 4, 4, 9, 9, 10, 14, 6, 3, 4, 6, 10, 4, 3, 3, 10, 4, 10, 5, 4, 3, 13, 1, 7, 5, 7, 6, 3, 7};
  m.inputs["n1"] = {28, 21, 32, 35, 35, 38, 30, 43, 49, 53, 31, 35, 46, 53, 61, 40, 29, 44, 52, 55, 61, 31, 48, 44, 42, 53, 56, 71, 43, 43,
      43, 40, 44, 70, 75, 71, 37, 31, 42, 46, 47, 55, 63, 91, 43, 39, 35, 32, 53, 49, 75, 64, 69, 64, 49, 29, 40, 27, 48, 43,
      61, 77, 55, 60, 46, 28, 33, 32, 46, 57, 56, 78, 58, 52, 31, 28, 46, 42, 45, 63, 71, 69, 43, 50, 31, 34, 54, 46, 58, 62,
      52, 41, 34, 52, 63, 59, 88, 62, 47, 53, 57, 74, 68, 61, 45, 45, 62, 73, 53, 39, 45, 51, 55, 41, 53, 51, 42, 46, 54, 32};
  m.inputs["r0"] = {0, 2, 2, 1, 2, 0, 1, 1, 1, 2, 4, 4, 2, 1, 7, 4, 3, 5, 3, 2, 4, 1, 4, 5, 2, 7, 5, 8, 2, 3,
      5, 4, 1, 6, 5, 11, 5, 2, 5, 8, 5, 6, 6, 10, 7, 5, 5, 2, 8, 1, 13, 9, 11, 9, 4, 4, 8, 6, 8, 6,
      8, 14, 6, 5, 5, 2, 4, 2, 9, 5, 6, 7, 5, 10, 3, 2, 1, 7, 9, 13, 9, 11, 4, 8, 2, 3, 7, 4, 7, 5,
      6, 6, 5, 6, 9, 7, 7, 7, 4, 2, 3, 4, 10, 3, 4, 2, 10, 5, 4, 5, 4, 6, 5, 3, 2, 2, 4, 6, 4, 1};
  m.inputs["n0"] = {28, 21, 32, 35, 35, 38, 30, 43, 49, 53, 31, 35, 46, 53, 61, 40, 29, 44, 52, 55, 61, 31, 48, 44, 42, 53, 56, 71, 43, 43,
      43, 40, 44, 70, 75, 71, 37, 31, 42, 46, 47, 55, 63, 91, 43, 39, 35, 32, 53, 49, 75, 64, 69, 64, 49, 29, 40, 27, 48, 43,
      61, 77, 55, 60, 46, 28, 33, 32, 46, 57, 56, 78, 58, 52, 31, 28, 46, 42, 45, 63, 71, 69, 43, 50, 31, 34, 54, 46, 58, 62,
      52, 41, 34, 52, 63, 59, 88, 62, 47, 53, 57, 74, 68, 61, 45, 45, 62, 73, 53, 39, 45, 51, 55, 41, 53, 51, 42, 46, 54, 32};
  m.inputs["year"] = {-10, -9, -9, -8, -8, -8, -7, -7, -7, -7, -6, -6, -6, -6, -6, -5, -5, -5, -5, -5, -5, -4, -4, -4, -4, -4, -4, -4, -3, -3,
      -3, -3, -3, -3, -3, -3, -2, -2, -2, -2, -2, -2, -2, -2, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0, 0,
      0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3,
      3, 3, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 9, 9, 10};
  auto vague = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
  { Node n = make_node("alpha", true, 1, true, true); n.eval = vague; m.nodes.push_back(n); }              // 0
  { Node n = make_node("beta1", true, 1, true, true); n.eval = vague; m.nodes.push_back(n); }              // 1
  { Node n = make_node("beta2", true, 1, true, true); n.eval = vague; m.nodes.push_back(n); }              // 2
  { Node n = make_node("s2", true, 1, true, true);                                                         // 3
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
    m.nodes.push_back(n); }
  { Node n = make_node("b", true, 120, false, false); n.sources = {3};                                     // 4
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, std::sqrt(mm.val(3)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("mu", true, 120, false, false); n.eval = vague; m.nodes.push_back(n); }             // 5
  { Node n = make_node("r0", true, 120, false, false, true); n.sources = {5};                              // 6
    n.eval = [](const Model& mm, Node& s) {
      const auto& n0 = mm.in("n0"); const auto& mu = mm.val(5);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(n0.size());
      for (size_t i = 0; i < n0.size(); ++i) s.distr.arr[i] = {D_BINOMIAL, n0[i], invlogit(mu[i])};
    };
    m.nodes.push_back(n); }
  { Node n = make_node("r1", true, 120, false, false, true); n.sources = {5, 0, 1, 2, 4};                  // 7
    n.eval = [](const Model& mm, Node& s) {
      const auto& n1 = mm.in("n1"); const auto& yr = mm.in("year"); const auto& mu = mm.val(5); const auto& b = mm.val(4);
      const double al = mm.val(0)[0], b1 = mm.val(1)[0], b2 = mm.val(2)[0];
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(n1.size());
      for (size_t i = 0; i < n1.size(); ++i)
        s.distr.arr[i] = {D_BINOMIAL, n1[i], invlogit(mu[i] + al + b1 * yr[i] + b2 * (yr[i] * yr[i] - 22.0) + b[i])};
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: alpha, beta1, beta2, s2, b[120], mu[120]
    const auto &r0 = mm.in("r0"), &n0 = mm.in("n0"), &r1 = mm.in("r1"), &n1 = mm.in("n1"), &yr = mm.in("year");
    const auto &b = mm.val(4), &mu = mm.val(5);
    const double al = mm.val(0)[0], b1 = mm.val(1)[0], b2 = mm.val(2)[0], s2 = mm.val(3)[0];
    double ga = 0, g1 = 0, g2 = 0, sbb = 0;
    for (int i = 0; i < 120; ++i) {
      const double q = yr[i] * yr[i] - 22.0;
      const double res1 = r1[i] - n1[i] * invlogit(mu[i] + al + b1 * yr[i] + b2 * q + b[i]);
      const double res0 = r0[i] - n0[i] * invlogit(mu[i]);
      ga += res1; g1 += res1 * yr[i]; g2 += res1 * q;
      g[4 + i] = res1 - b[i] / s2;
      g[124 + i] = res0 + res1 - mu[i] / 1e6;
      sbb += b[i] * b[i];
    }
    g[0] = ga - al / 1e6; g[1] = g1 - b1 / 1e6; g[2] = g2 - b2 / 1e6;
    g[3] = -60.0 / s2 + 0.5 * sbb / (s2 * s2) + ig_dlogpdf(0.001, 0.001, s2);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// epil: doc/examples/epil.jl:33-111 (data :4-30) — Poisson GLMM of seizure counts, 59 patients x 4 visits.  303 unobserved elements:
// a0, alpha_Base, alpha_Trt, alpha_BT, alpha_Age, alpha_V4 ~ Normal(0, 100), s2_b1, s2_b ~ InverseGamma(.001, .001),
// b1[59] ~ Normal(0, sqrt(s2_b1)), b[59 x 4] ~ Normal(0, sqrt(s2_b)) (column-major, patient fastest);
// y[i, j] ~ Poisson(exp(a0 + alpha_Base (logBase4_i - mean) + alpha_Trt (Trt_i - mean) + alpha_BT (BT_i - mean) + alpha_Age (logAge_i - mean)
//                       + alpha_V4 (V4_j - mean) + b1[i] + b[i, j])).
// Monitored: alpha_Base, alpha_Trt, alpha_BT, alpha_Age, alpha_V4, alpha0 (Logical), s2_b1, s2_b.
struct EpilCov { std::vector<double> lb, trt, bt, la; double v4[4]; double lbbar, trtbar, btbar, labar, v4bar; };
inline EpilCov epil_cov(const Model& mm) {     // epil.jl:25-30
  const auto &Base = mm.in("Base"), &Trt = mm.in("Trt"), &Age = mm.in("Age"), &V4 = mm.in("V4");
  EpilCov c; const size_t N = Base.size();
  c.lb.resize(N); c.trt = Trt; c.bt.resize(N); c.la.resize(N);
  double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (size_t i = 0; i < N; ++i) { c.lb[i] = std::log(Base[i] / 4.0); c.bt[i] = c.lb[i] * Trt[i]; c.la[i] = std::log(Age[i]); s1 += c.lb[i]; s2 += Trt[i]; s3 += c.bt[i]; s4 += c.la[i]; }
  c.lbbar = s1 / N; c.trtbar = s2 / N; c.btbar = s3 / N; c.labar = s4 / N;
  double sv = 0; for (int j = 0; j < 4; ++j) { c.v4[j] = V4[j]; sv += V4[j]; }
  c.v4bar = sv / 4.0;
  return c;
}
inline Model make_epil() {
  Model m; m.template_id = TPL_EPIL;
  m.inputs["y"] = {5, 3, 2, 4, 7, 5, 6, 40, 5, 14, 26, 12, 4, 7, 16, 11, 0, 37, 3, 3, 3, 3, 2, 8, 18, 2, 3, 13, 11, 8,
      0, 3, 2, 4, 22, 5, 2, 3, 4, 2, 0, 5, 11, 10, 19, 1, 6, 2, 102, 4, 8, 1, 18, 6, 3, 1, 2, 0, 1, 3,
      5, 4, 4, 18, 2, 4, 20, 6, 13, 12, 6, 4, 9, 24, 0, 0, 29, 5, 0, 4, 4, 3, 12, 24, 1, 1, 15, 14, 7, 4,
      6, 6, 3, 17, 4, 4, 7, 18, 1, 2, 4, 14, 5, 7, 1, 10, 1, 65, 3, 6, 3, 11, 3, 5, 23, 3, 0, 4, 3, 3,
      0, 1, 9, 8, 0, 21, 6, 6, 6, 8, 6, 12, 10, 0, 3, 28, 2, 6, 3, 3, 3, 2, 76, 2, 4, 13, 9, 9, 3, 1,
      7, 1, 19, 7, 0, 7, 2, 1, 4, 0, 25, 3, 6, 2, 8, 0, 72, 2, 5, 1, 28, 4, 4, 19, 0, 0, 3, 3, 3, 5,
      4, 21, 7, 2, 12, 5, 0, 22, 4, 2, 14, 9, 5, 3, 29, 5, 7, 4, 4, 5, 8, 25, 1, 2, 12, 8, 4, 0, 3, 4,
      3, 16, 4, 4, 7, 5, 0, 0, 3, 15, 8, 7, 3, 8, 0, 63, 4, 7, 5, 13, 0, 3, 8, 1, 0, 2};   // 59 x 4, column-major (patient fastest)
  m.inputs["Trt"] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,
      1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  m.inputs["Base"] = {11, 11, 6, 8, 66, 27, 12, 52, 23, 10, 52, 33, 18, 42, 87, 50, 18, 111, 18, 20, 12, 9, 17, 28, 55, 9, 10, 47, 76, 38,
      19, 10, 19, 24, 31, 14, 11, 67, 41, 7, 22, 13, 46, 36, 38, 7, 36, 11, 151, 22, 41, 32, 56, 24, 16, 22, 25, 13, 12};
  m.inputs["Age"] = {31, 30, 25, 36, 22, 29, 31, 42, 37, 28, 36, 24, 23, 36, 26, 26, 28, 31, 32, 21, 29, 21, 32, 25, 30, 40, 19, 22, 18, 32,
      20, 30, 18, 24, 30, 35, 27, 20, 22, 28, 23, 40, 33, 21, 35, 25, 26, 25, 22, 32, 25, 35, 21, 41, 32, 26, 21, 36, 37};
  m.inputs["V4"] = {0, 0, 0, 1};
  auto coef = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 100.0}; };
  { Node n = make_node("a0", true, 1, true, false); n.eval = coef; m.nodes.push_back(n); }                  // 0
  const char* names[5] = {"alpha_Base", "alpha_Trt", "alpha_BT", "alpha_Age", "alpha_V4"};
  for (int k = 0; k < 5; ++k) { Node n = make_node(names[k], true, 1, true, true); n.eval = coef; m.nodes.push_back(n); }   // 1..5
  { Node n = make_node("alpha0", false, 1, true, true); n.sources = {0, 1, 2, 3, 4, 5};                     // 6: Logical, epil.jl:85-91
    n.eval = [](const Model& mm, Node& l) {
      const EpilCov c = epil_cov(mm);
      l.value.assign(1, mm.val(0)[0] - mm.val(1)[0] * c.lbbar - mm.val(2)[0] * c.trtbar - mm.val(3)[0] * c.btbar - mm.val(4)[0] * c.labar - mm.val(5)[0] * c.v4bar);
    };
    m.nodes.push_back(n); }
  auto ig = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
  { Node n = make_node("s2_b1", true, 1, true, true); n.eval = ig; m.nodes.push_back(n); }                 // 7
  { Node n = make_node("s2_b", true, 1, true, true); n.eval = ig; m.nodes.push_back(n); }                  // 8
  { Node n = make_node("b1", true, 59, false, false); n.sources = {7};                                     // 9
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, std::sqrt(mm.val(7)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("b", true, 236, false, false); n.sources = {8};                                     // 10
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, std::sqrt(mm.val(8)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 236, false, false, true); n.sources = {0, 1, 2, 3, 4, 5, 9, 10};         // 11
    n.eval = [](const Model& mm, Node& s) {
      const EpilCov c = epil_cov(mm);
      const double a0 = mm.val(0)[0], aB = mm.val(1)[0], aT = mm.val(2)[0], aBT = mm.val(3)[0], aA = mm.val(4)[0], aV = mm.val(5)[0];
      const auto &b1 = mm.val(9), &b = mm.val(10);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(236);
      for (int j = 0; j < 4; ++j) for (int i = 0; i < 59; ++i) {
        const double eta = a0 + aB * (c.lb[i] - c.lbbar) + aT * (c.trt[i] - c.trtbar) + aBT * (c.bt[i] - c.btbar) + aA * (c.la[i] - c.labar) +
                           aV * (c.v4[j] - c.v4bar) + b1[i] + b[i + 59 * j];
        s.distr.arr[i + 59 * j] = {D_POISSON, std::exp(eta), 0.0};
      }
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: a0, alpha_Base, alpha_Trt, alpha_BT, alpha_Age, alpha_V4, s2_b1, s2_b, b1[59], b[236]
    const EpilCov c = epil_cov(mm);
    const auto& y = mm.in("y");
    const double a0 = mm.val(0)[0], aB = mm.val(1)[0], aT = mm.val(2)[0], aBT = mm.val(3)[0], aA = mm.val(4)[0], aV = mm.val(5)[0];
    const double s2b1 = mm.val(7)[0], s2b = mm.val(8)[0];
    const auto &b1 = mm.val(9), &b = mm.val(10);
    double ga[6] = {0, 0, 0, 0, 0, 0}, sb1 = 0, sb = 0;
    for (int i = 0; i < 59; ++i) g[8 + i] = 0.0;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 59; ++i) {
      const double x1 = c.lb[i] - c.lbbar, x2 = c.trt[i] - c.trtbar, x3 = c.bt[i] - c.btbar, x4 = c.la[i] - c.labar, x5 = c.v4[j] - c.v4bar;
      const double eta = a0 + aB * x1 + aT * x2 + aBT * x3 + aA * x4 + aV * x5 + b1[i] + b[i + 59 * j];
      const double res = y[i + 59 * j] - std::exp(eta);
      ga[0] += res; ga[1] += res * x1; ga[2] += res * x2; ga[3] += res * x3; ga[4] += res * x4; ga[5] += res * x5;
      g[8 + i] += res;
      g[67 + i + 59 * j] = res - b[i + 59 * j] / s2b;
      sb += b[i + 59 * j] * b[i + 59 * j];
    }
    for (int i = 0; i < 59; ++i) { g[8 + i] -= b1[i] / s2b1; sb1 += b1[i] * b1[i]; }
    const double co[6] = {a0, aB, aT, aBT, aA, aV};
    for (int k = 0; k < 6; ++k) g[k] = ga[k] - co[k] / 1e4;
    g[6] = -29.5 / s2b1 + 0.5 * sb1 / (s2b1 * s2b1) + ig_dlogpdf(0.001, 0.001, s2b1);
    g[7] = -118.0 / s2b + 0.5 * sb / (s2b * s2b) + ig_dlogpdf(0.001, 0.001, s2b);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// surgical: doc/examples/surgical.jl:11-43 (data :4-8).  Node order = topological order: mu, pop_mean, s2, b, p, r.
inline Model make_surgical() {
  Model m; m.template_id = TPL_SURGICAL;
  m.inputs["r"] = {0, 18, 8, 46, 8, 13, 9, 31, 14, 8, 29, 24};
  m.inputs["n"] = {47, 148, 119, 810, 211, 196, 148, 215, 207, 97, 256, 360};
  { Node n = make_node("mu", true, 1, true, true);                        // 0: Normal(0, 1000)
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
    m.nodes.push_back(n); }
  { Node n = make_node("pop_mean", false, 1, true, true);                 // 1: Logical, invlogit(mu)
    n.sources = {0};
    n.eval = [](const Model& mm, Node& l) { l.value.assign(1, invlogit(mm.val(0)[0])); };
    m.nodes.push_back(n); }
  { Node n = make_node("s2", true, 1, true, true);                        // 2: InverseGamma(0.001, 0.001)
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
    m.nodes.push_back(n); }
  { Node n = make_node("b", true, 12, false, false);                      // 3: Normal(mu, sqrt(s2))
    n.sources = {0, 2};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, mm.val(0)[0], std::sqrt(mm.val(2)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("p", false, 12, false, true);                      // 4: Logical, invlogit(b)
    n.sources = {3};
    n.eval = [](const Model& mm, Node& l) { const auto& b = mm.val(3); l.value.resize(b.size()); for (size_t i = 0; i < b.size(); ++i) l.value[i] = invlogit(b[i]); };
    m.nodes.push_back(n); }
  { Node n = make_node("r", true, 12, false, false, true);                // 5: Binomial(n[i], p[i])
    n.sources = {4};
    n.eval = [](const Model& mm, Node& s) {
      const auto& nn = mm.in("n"); const auto& p = mm.val(4);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(nn.size());
      for (size_t i = 0; i < nn.size(); ++i) s.distr.arr[i] = {D_BINOMIAL, nn[i], p[i]};
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: mu, s2, b[12]
    const auto& nn = mm.in("n"); const auto& r = mm.in("r"); const auto& b = mm.val(3);
    const double mu = mm.val(0)[0], s2 = mm.val(2)[0];
    double sd = 0, sdd = 0;
    for (size_t i = 0; i < nn.size(); ++i) {
      const double p = invlogit(b[i]), db = b[i] - mu;
      g[2 + i] = (r[i] - nn[i] * p) - db / s2;
      sd += db; sdd += db * db;
    }
    g[0] = sd / s2 - mu / 1e6;
    g[1] = -0.5 * (double)nn.size() / s2 + 0.5 * sdd / (s2 * s2) + ig_dlogpdf(0.001, 0.001, s2);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// dyes: doc/examples/dyes.jl:22-47 (data :4-17).  Node order: s2_between, theta, s2_within, mu, y (a valid topological order; the
// monitored columns come out in the order of doc/examples/dyes.rst).
inline Model make_dyes() {
  Model m; m.template_id = TPL_DYES;
  m.inputs["y"] = {1545, 1440, 1440, 1520, 1580, 1540, 1555, 1490, 1560, 1495, 1595, 1550, 1605, 1510, 1560,
                   1445, 1440, 1595, 1465, 1545, 1595, 1630, 1515, 1635, 1625, 1520, 1455, 1450, 1480, 1445};
  { std::vector<double> b(30); for (int k = 0; k < 30; ++k) b[k] = k / 5; m.inputs["batch"] = b; }
  auto prior_ig = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
  { Node n = make_node("s2_between", true, 1, true, true); n.eval = prior_ig; m.nodes.push_back(n); }          // 0
  { Node n = make_node("theta", true, 1, true, true);                                                          // 1
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
    m.nodes.push_back(n); }
  { Node n = make_node("s2_within", true, 1, true, true); n.eval = prior_ig; m.nodes.push_back(n); }           // 2
  { Node n = make_node("mu", true, 6, false, true);                                                            // 3
    n.sources = {1, 0};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, mm.val(1)[0], std::sqrt(mm.val(0)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 30, false, false, true);                                                     // 4: MvNormal(mu[batch], sqrt(s2_within))
    n.sources = {3, 2};
    n.eval = [](const Model& mm, Node& s) {
      const auto& b = mm.in("batch"); const auto& mu = mm.val(3);
      s.distr.form = Distr::MVNORMAL_ISO; s.distr.mu.resize(b.size());
      for (size_t k = 0; k < b.size(); ++k) s.distr.mu[k] = mu[(size_t)b[k]];
      s.distr.sigma = std::sqrt(mm.val(2)[0]);
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: s2_between, theta, s2_within, mu[6]
    const auto& y = mm.in("y"); const auto& b = mm.in("batch"); const auto& mu = mm.val(3);
    const double s2b = mm.val(0)[0], th = mm.val(1)[0], s2w = mm.val(2)[0];
    for (int i = 0; i < 6; ++i) g[3 + i] = 0.0;
    double see = 0;
    for (size_t k = 0; k < y.size(); ++k) { const size_t i = (size_t)b[k]; const double e = y[k] - mu[i]; g[3 + i] += e / s2w; see += e * e; }
    double sd = 0, sdd = 0;
    for (int i = 0; i < 6; ++i) { const double dm = mu[i] - th; g[3 + i] -= dm / s2b; sd += dm; sdd += dm * dm; }
    g[1] = sd / s2b - th / 1e6;
    g[0] = -3.0 / s2b + 0.5 * sdd / (s2b * s2b) + ig_dlogpdf(0.001, 0.001, s2b);
    g[2] = -0.5 * (double)y.size() / s2w + 0.5 * see / (s2w * s2w) + ig_dlogpdf(0.001, 0.001, s2w);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// salm: doc/examples/salm.jl:16-53 (data :4-11).  Node order s2, gamma, beta, alpha, lambda, y: a valid topological order that puts
// the monitored columns in the order of doc/examples/salm.rst.  The 3 x 6 matrices are flattened column-major (plate fastest).
inline Model make_salm() {
  Model m; m.template_id = TPL_SALM;
  m.inputs["y"] = {15, 21, 29, 16, 18, 21, 16, 26, 33, 27, 41, 60, 33, 38, 41, 20, 27, 42};
  m.inputs["x"] = {0, 10, 33, 100, 333, 1000};
  auto prior_n = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
  { Node n = make_node("s2", true, 1, true, true);                                                             // 0
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
    m.nodes.push_back(n); }
  { Node n = make_node("gamma", true, 1, true, true); n.eval = prior_n; m.nodes.push_back(n); }                // 1
  { Node n = make_node("beta", true, 1, true, true); n.eval = prior_n; m.nodes.push_back(n); }                 // 2
  { Node n = make_node("alpha", true, 1, true, true); n.eval = prior_n; m.nodes.push_back(n); }                // 3
  { Node n = make_node("lambda", true, 18, false, false);                                                      // 4: Normal(0, sqrt(s2))
    n.sources = {0};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, std::sqrt(mm.val(0)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 18, false, false, true);                                                     // 5: Poisson(mu_ij)
    n.sources = {3, 2, 1, 4};
    n.eval = [](const Model& mm, Node& s) {
      const auto& x = mm.in("x"); const auto& lam = mm.val(4);
      const double alpha = mm.val(3)[0], beta = mm.val(2)[0], gamma = mm.val(1)[0];
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(18);
      for (int e = 0; e < 18; ++e) { const double xj = x[e / 3]; s.distr.arr[e] = {D_POISSON, std::exp(alpha + beta * std::log(xj + 10.0) + gamma * xj + lam[e]), 0.0}; }
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: s2, gamma, beta, alpha, lambda[18]
    const auto& x = mm.in("x"); const auto& y = mm.in("y"); const auto& lam = mm.val(4);
    const double s2 = mm.val(0)[0], gamma = mm.val(1)[0], beta = mm.val(2)[0], alpha = mm.val(3)[0];
    double ga = 0, gb = 0, gg = 0, sll = 0;
    for (int e = 0; e < 18; ++e) {
      const double xj = x[e / 3];
      const double r = y[e] - std::exp(alpha + beta * std::log(xj + 10.0) + gamma * xj + lam[e]);
      ga += r; gb += r * std::log(xj + 10.0); gg += r * xj;
      g[4 + e] = r - lam[e] / s2; sll += lam[e] * lam[e];
    }
    g[3] = ga - alpha / 1e6; g[2] = gb - beta / 1e6; g[1] = gg - gamma / 1e6;
    g[0] = -9.0 / s2 + 0.5 * sll / (s2 * s2) + ig_dlogpdf(0.001, 0.001, s2);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// equiv: doc/examples/equiv.jl:25-75 (data :4-22).  Node order s2_2, s2_1, pi, phi, theta, equiv, mu, delta, y (topological; the
// monitored columns come out in the order of doc/examples/equiv.rst).  10 x 2 matrices flattened column-major (subject fastest).
inline Model make_equiv() {
  Model m; m.template_id = TPL_EQUIV;
  m.inputs["group"] = {1, 1, 2, 2, 2, 1, 1, 1, 2, 2};
  m.inputs["y"] = {1.40, 1.64, 1.44, 1.36, 1.65, 1.08, 1.09, 1.25, 1.25, 1.30, 1.65, 1.57, 1.58, 1.68, 1.69, 1.31, 1.43, 1.44, 1.39, 1.52};
  auto prior_n = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
  auto prior_ig = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
  { Node n = make_node("s2_2", true, 1, true, true); n.eval = prior_ig; m.nodes.push_back(n); }               // 0
  { Node n = make_node("s2_1", true, 1, true, true); n.eval = prior_ig; m.nodes.push_back(n); }               // 1
  { Node n = make_node("pi", true, 1, true, true); n.eval = prior_n; m.nodes.push_back(n); }                   // 2
  { Node n = make_node("phi", true, 1, true, true); n.eval = prior_n; m.nodes.push_back(n); }                  // 3
  { Node n = make_node("theta", false, 1, true, true);                                                         // 4: Logical, exp(phi)
    n.sources = {3};
    n.eval = [](const Model& mm, Node& l) { l.value.assign(1, std::exp(mm.val(3)[0])); };
    m.nodes.push_back(n); }
  { Node n = make_node("equiv", false, 1, true, true);                                                         // 5: Logical, Int(0.8 < theta < 1.2)
    n.sources = {4};
    n.eval = [](const Model& mm, Node& l) { const double th = mm.val(4)[0]; l.value.assign(1, (0.8 < th && th < 1.2) ? 1.0 : 0.0); };
    m.nodes.push_back(n); }
  { Node n = make_node("mu", true, 1, true, true); n.eval = prior_n; m.nodes.push_back(n); }                   // 6
  { Node n = make_node("delta", true, 20, false, false);                                                       // 7: Normal(0, sqrt(s2_2))
    n.sources = {0};
    n.eval = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, std::sqrt(mm.val(0)[0])}; };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 20, false, false, true);                                                     // 8: Normal(m_ij, sqrt(s2_1))
    n.sources = {7, 6, 3, 2, 1};
    n.eval = [](const Model& mm, Node& s) {
      const auto& grp = mm.in("group"); const auto& dl = mm.val(7);
      const double mu = mm.val(6)[0], phi = mm.val(3)[0], pi = mm.val(2)[0], sigma = std::sqrt(mm.val(1)[0]);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(20);
      for (int e = 0; e < 20; ++e) {
        const int i = e % 10, j = e / 10;
        const double T = j == 0 ? grp[i] : 3.0 - grp[i];
        const double mean = mu + (T == 1.0 ? 1.0 : -1.0) * phi / 2.0 + (j == 0 ? 1.0 : -1.0) * pi / 2.0 + dl[e];
        s.distr.arr[e] = {D_NORMAL, mean, sigma};
      }
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: s2_2, s2_1, pi, phi, mu, delta[20]
    const auto& grp = mm.in("group"); const auto& y = mm.in("y"); const auto& dl = mm.val(7);
    const double s22 = mm.val(0)[0], s21 = mm.val(1)[0], pi = mm.val(2)[0], phi = mm.val(3)[0], mu = mm.val(6)[0];
    double gm = 0, gp = 0, gq = 0, see = 0, sdd = 0;
    for (int e = 0; e < 20; ++e) {
      const int i = e % 10, j = e / 10;
      const double T = j == 0 ? grp[i] : 3.0 - grp[i];
      const double sp = T == 1.0 ? 1.0 : -1.0, sq = j == 0 ? 1.0 : -1.0;
      const double res = y[e] - (mu + sp * phi / 2.0 + sq * pi / 2.0 + dl[e]);
      const double r = res / s21;
      gm += r; gp += r * sp / 2.0; gq += r * sq / 2.0;
      g[5 + e] = r - dl[e] / s22; see += res * res; sdd += dl[e] * dl[e];
    }
    g[4] = gm - mu / 1e6; g[3] = gp - phi / 1e6; g[2] = gq - pi / 1e6;
    g[1] = -10.0 / s21 + 0.5 * see / (s21 * s21) + ig_dlogpdf(0.001, 0.001, s21);
    g[0] = -10.0 / s22 + 0.5 * sdd / (s22 * s22) + ig_dlogpdf(0.001, 0.001, s22);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// blocker: doc/examples/blocker.jl:22-69 (data :4-18).  Node order s2, d, delta_new, mu, delta, rc, rt (topological; monitored columns
// in the order of doc/examples/blocker.rst).  Two observed nodes.
inline Model make_blocker() {
  Model m; m.template_id = TPL_BLOCKER;
  m.inputs["rt"] = {3, 7, 5, 102, 28, 4, 98, 60, 25, 138, 64, 45, 9, 57, 25, 33, 28, 8, 6, 32, 27, 22};
  m.inputs["nt"] = {38, 114, 69, 1533, 355, 59, 945, 632, 278, 1916, 873, 263, 291, 858, 154, 207, 251, 151, 174, 209, 391, 680};
  m.inputs["rc"] = {3, 14, 11, 127, 27, 6, 152, 48, 37, 188, 52, 47, 16, 45, 31, 38, 12, 6, 3, 40, 43, 39};
  m.inputs["nc"] = {39, 116, 93, 1520, 365, 52, 939, 471, 282, 1921, 583, 266, 293, 883, 147, 213, 122, 154, 134, 218, 364, 674};
  { Node n = make_node("s2", true, 1, true, true);                                                             // 0
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
    m.nodes.push_back(n); }
  { Node n = make_node("d", true, 1, true, true);                                                              // 1
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
    m.nodes.push_back(n); }
  auto effect = [](const Model& mm, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, mm.val(1)[0], std::sqrt(mm.val(0)[0])}; };
  { Node n = make_node("delta_new", true, 1, true, true); n.sources = {1, 0}; n.eval = effect; m.nodes.push_back(n); }   // 2
  { Node n = make_node("mu", true, 22, false, false);                                                          // 3
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
    m.nodes.push_back(n); }
  { Node n = make_node("delta", true, 22, false, false); n.sources = {1, 0}; n.eval = effect; m.nodes.push_back(n); }    // 4
  { Node n = make_node("rc", true, 22, false, false, true);                                                    // 5: Binomial(nc[i], invlogit(mu[i]))
    n.sources = {3};
    n.eval = [](const Model& mm, Node& s) {
      const auto& nc = mm.in("nc"); const auto& mu = mm.val(3);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(nc.size());
      for (size_t i = 0; i < nc.size(); ++i) s.distr.arr[i] = {D_BINOMIAL, nc[i], invlogit(mu[i])};
    };
    m.nodes.push_back(n); }
  { Node n = make_node("rt", true, 22, false, false, true);                                                    // 6: Binomial(nt[i], invlogit(mu[i] + delta[i]))
    n.sources = {3, 4};
    n.eval = [](const Model& mm, Node& s) {
      const auto& nt = mm.in("nt"); const auto& mu = mm.val(3); const auto& dl = mm.val(4);
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(nt.size());
      for (size_t i = 0; i < nt.size(); ++i) s.distr.arr[i] = {D_BINOMIAL, nt[i], invlogit(mu[i] + dl[i])};
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: s2, d, delta_new, mu[22], delta[22]
    const auto& rc = mm.in("rc"); const auto& nc = mm.in("nc"); const auto& rt = mm.in("rt"); const auto& nt = mm.in("nt");
    const auto& mu = mm.val(3); const auto& dl = mm.val(4);
    const double s2 = mm.val(0)[0], d = mm.val(1)[0], dn = mm.val(2)[0];
    double sd = dn - d, sdd = sd * sd;
    g[2] = -(dn - d) / s2;
    for (int i = 0; i < 22; ++i) {
      const double pc = invlogit(mu[i]), pt = invlogit(mu[i] + dl[i]);
      const double r = rt[i] - nt[i] * pt;
      g[3 + i] = (rc[i] - nc[i] * pc) + r - mu[i] / 1e6;
      g[25 + i] = r - (dl[i] - d) / s2;
      sd += dl[i] - d; sdd += (dl[i] - d) * (dl[i] - d);
    }
    g[1] = sd / s2 - d / 1e6;
    g[0] = -11.5 / s2 + 0.5 * sdd / (s2 * s2) + ig_dlogpdf(0.001, 0.001, s2);
  };
  m.finalize();
  return m;
}

// ------------------------------------------------------------------------------------------
// stacks: doc/examples/stacks.jl:41-94 (data :4-38): stack-loss regression on standardised covariates with a Laplace likelihood; every
// monitored quantity is a Logical node (b, b0, sigma, outlier[1, 3, 4, 21]).  Node order beta0, beta, s2, b, b0, sigma, mu, outlier, y.
inline std::vector<double> stacks_x() {   // 21 x 3, row-major
  return {80, 27, 89, 80, 27, 88, 75, 25, 90, 62, 24, 87, 62, 22, 87, 62, 23, 87, 62, 24, 93, 62, 24, 93, 58, 23, 87, 58, 18, 80, 58, 18, 89,
          58, 17, 88, 58, 18, 82, 58, 19, 93, 50, 18, 89, 50, 18, 86, 50, 19, 72, 50, 19, 79, 50, 20, 80, 56, 20, 82, 70, 20, 91};
}
inline Model make_stacks() {
  Model m; m.template_id = TPL_STACKS;
  m.inputs["y"] = {42, 37, 37, 28, 18, 18, 19, 20, 15, 14, 14, 13, 11, 12, 8, 7, 8, 8, 9, 15, 15};
  m.inputs["x"] = stacks_x();
  {   // meanx, sdx (sample sd), z = (x - meanx) / sdx: stacks.jl:32-37
    const auto& x = m.inputs["x"]; const int N = 21;
    std::vector<double> mean(3, 0.0), sd(3, 0.0), z(N * 3);
    for (int j = 0; j < 3; ++j) { for (int i = 0; i < N; ++i) mean[j] += x[i * 3 + j]; mean[j] /= N; }
    for (int j = 0; j < 3; ++j) { double s = 0; for (int i = 0; i < N; ++i) { const double e = x[i * 3 + j] - mean[j]; s += e * e; } sd[j] = std::sqrt(s / (N - 1)); }
    for (int i = 0; i < N; ++i) for (int j = 0; j < 3; ++j) z[i * 3 + j] = (x[i * 3 + j] - mean[j]) / sd[j];
    m.inputs["meanx"] = mean; m.inputs["sdx"] = sd; m.inputs["z"] = z;
  }
  auto prior_n = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_NORMAL, 0.0, 1000.0}; };
  { Node n = make_node("beta0", true, 1, true, false); n.eval = prior_n; m.nodes.push_back(n); }              // 0
  { Node n = make_node("beta", true, 3, false, false); n.eval = prior_n; m.nodes.push_back(n); }              // 1
  { Node n = make_node("s2", true, 1, true, false);                                                            // 2
    n.eval = [](const Model&, Node& s) { s.distr.form = Distr::UNI; s.distr.u = {D_INVGAMMA, 0.001, 0.001}; };
    m.nodes.push_back(n); }
  { Node n = make_node("b", false, 3, false, true);                                                            // 3: beta ./ sdx
    n.sources = {1};
    n.eval = [](const Model& mm, Node& l) { const auto& be = mm.val(1); const auto& sd = mm.in("sdx"); l.value.resize(3); for (int j = 0; j < 3; ++j) l.value[j] = be[j] / sd[j]; };
    m.nodes.push_back(n); }
  { Node n = make_node("b0", false, 1, true, true);                                                            // 4: beta0 - dot(b, meanx)
    n.sources = {0, 3};
    n.eval = [](const Model& mm, Node& l) { const auto& b = mm.val(3); const auto& mx = mm.in("meanx"); double s = 0; for (int j = 0; j < 3; ++j) s += b[j] * mx[j]; l.value.assign(1, mm.val(0)[0] - s); };
    m.nodes.push_back(n); }
  { Node n = make_node("sigma", false, 1, true, true);                                                         // 5: sqrt(2) * s2
    n.sources = {2};
    n.eval = [](const Model& mm, Node& l) { l.value.assign(1, std::sqrt(2.0) * mm.val(2)[0]); };
    m.nodes.push_back(n); }
  { Node n = make_node("mu", false, 21, false, false);                                                         // 6: beta0 + z * beta
    n.sources = {0, 1};
    n.eval = [](const Model& mm, Node& l) {
      const auto& z = mm.in("z"); const auto& be = mm.val(1); const double b0 = mm.val(0)[0];
      l.value.resize(21);
      for (int i = 0; i < 21; ++i) { double s = 0; for (int j = 0; j < 3; ++j) s += z[i * 3 + j] * be[j]; l.value[i] = b0 + s; }
    };
    m.nodes.push_back(n); }
  { Node n = make_node("outlier", false, 21, false, false);                                                    // 7: |y - mu| / sigma > 2.5, monitor [1, 3, 4, 21]
    n.monitor = {0, 2, 3, 20};
    n.sources = {6, 5};
    n.eval = [](const Model& mm, Node& l) {
      const auto& y = mm.in("y"); const auto& mu = mm.val(6); const double sg = mm.val(5)[0];
      l.value.resize(21);
      for (int i = 0; i < 21; ++i) l.value[i] = std::fabs((y[i] - mu[i]) / sg) > 2.5 ? 1.0 : 0.0;
    };
    m.nodes.push_back(n); }
  { Node n = make_node("y", true, 21, false, false, true);                                                     // 8: Laplace(mu[i], s2)
    n.sources = {6, 2};
    n.eval = [](const Model& mm, Node& s) {
      const auto& mu = mm.val(6); const double th = mm.val(2)[0];
      s.distr.form = Distr::UNI_ARRAY; s.distr.arr.resize(21);
      for (int i = 0; i < 21; ++i) s.distr.arr[i] = {D_LAPLACE, mu[i], th};
    };
    m.nodes.push_back(n); }
  m.joint_grad = [](const Model& mm, std::vector<double>& g) {   // state order: beta0, beta[3], s2
    const auto& y = mm.in("y"); const auto& z = mm.in("z"); const auto& mu = mm.val(6); const auto& be = mm.val(1);
    const double b0 = mm.val(0)[0], th = mm.val(2)[0];
    double g0 = 0, gb[3] = {0, 0, 0}, sabs = 0;
    for (int i = 0; i < 21; ++i) {
      const double e = y[i] - mu[i];
      const double sgn = e > 0 ? 1.0 : (e < 0 ? -1.0 : 0.0);
      g0 += sgn; for (int j = 0; j < 3; ++j) gb[j] += sgn * z[i * 3 + j];
      sabs += std::fabs(e);
    }
    g[0] = g0 / th - b0 / 1e6;
    for (int j = 0; j < 3; ++j) g[1 + j] = gb[j] / th - be[j] / 1e6;
    g[4] = -21.0 / th + sabs / (th * th) + ig_dlogpdf(0.001, 0.001, th);
  };
  m.finalize();
  return m;
}

inline Model make_template(int id, int glm_d = 0) {
  switch (id) {
    case TPL_LINE: return make_line();
    case TPL_SEEDS: return make_seeds();
    case TPL_RATS: return make_rats();
    case TPL_PUMPS: return make_pumps();
    case TPL_GLM: return make_glm(glm_d > 0 ? glm_d : 1);
    case TPL_SURGICAL: return make_surgical();
    case TPL_DYES: return make_dyes();
    case TPL_SALM: return make_salm();
    case TPL_BLOCKER: return make_blocker();
    case TPL_MAGNESIUM: return make_magnesium();
    case TPL_OXFORD: return make_oxford();
    case TPL_EPIL: return make_epil();
    case TPL_STACKS: return make_stacks();
    case TPL_EQUIV: return make_equiv();
    default: throw std::runtime_error("unknown template");
  }
}

}  // namespace orc
