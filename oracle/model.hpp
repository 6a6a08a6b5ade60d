// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
//
// model.hpp — the reference's Model / node / block-density layer, restated.
//   Model ctor + targets       : src/model/model.jl:5-27, src/model/graph.jl:93-103
//   keys(m, :block/:target)    : src/model/model.jl:98-110,185-205
//   setinits! / setsamplers!   : src/model/initialization.jl:3-28,42-48
//   logpdf! / logpdf / relist / unlist / update! / gradlogpdf! : src/model/simulation.jl:47-176
//   node-level logpdf/unlist/relist : src/model/dependent.jl:98-101,176-213
//   logpdfgrad!                : src/samplers/sampler.jl:106-111
// Like the reference, the density is INTERPRETED: every logpdf! call re-runs each target node's
// closure and re-builds its distribution (dependent.jl:176-179).
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "dist.hpp"

namespace orc {

struct Model;

struct Node {
  std::string name;
  bool stochastic = true;
  bool observed = false;            // stochastic node whose value is data (never in a sampler)
  int len = 1;
  bool scalar = true;
  std::vector<int> monitor;         // monitored element indices (0-based); empty = unmonitored
  std::vector<int> sources;         // dependent-node sources only (inputs are captured by eval)
  std::function<void(const Model&, Node&)> eval;  // Stochastic: sets distr; Logical: sets value
  std::vector<double> value;
  Distr distr;
  std::vector<int> targets;         // model.jl:17-25
};

enum SamplerKind { S_AMWG = 0, S_SLICE_UNI, S_SLICE_MULTI, S_RWM, S_NUTS, S_HMC, S_AMM, S_GIBBS, S_MALA };

struct Tune {  // union of the reference's *Tune types (amwg.jl:5-21, slice.jl:7-26, rwm.jl:5-22, nuts.jl:5-27, hmc.jl:5-28, amm.jl:5-24)
  bool init = false;
  // AMWG
  bool adapt = false; std::vector<long> accept; int batchsize = 50; long m = 0;
  std::vector<double> sigma; double target = 0.44;
  // NUTS
  double alpha = 0, epsilon = 0, epsilonbar = 1, gamma = 0.05, Hbar = 0, kappa = 0.75, mu = NAN, t0 = 10;
  long nalpha = 0;
  // AMM
  double beta = 0.05, scale = 2.38; std::vector<double> Mv, Mvv, SigmaL, SigmaLm;
};

struct SamplerSpec {
  int kind = S_AMWG;
  std::vector<int> params;      // node indices, in the user's order
  bool transform = true;
  int adapt = 0;                // 0 all, 1 burnin, 2 none
  int batchsize = 50;
  int proposal = 0;
  int L = 1;
  int grad = 0;                 // 0 analytic, 1 forward, 2 central
  int max_depth = 0;            // 0 = unbounded, as the reference
  double target = 0.0;
  double epsilon = 0.0;
  double beta = 0.05, amm_scale = 2.38;
  std::vector<double> scale;    // sigma / width / scale (1 or k) ; Sigma k*k
  std::vector<int> targets;     // initialization.jl:42-48
  Tune tune;
};

struct Model {
  int template_id = -1;
  std::vector<Node> nodes;                       // in a valid topological order == keys(m, :dependent)
  std::map<std::string, std::vector<double>> inputs;
  std::vector<SamplerSpec> samplers;
  long iter = 0, burnin = 0;
  // analytic gradient of the joint log density w.r.t. every unobserved stochastic element
  // (constrained scale), written to g[state offset].  Hand-derived per template (templates.hpp).
  std::function<void(const Model&, std::vector<double>&)> joint_grad;
  // user-defined Gibbs samplers of the template (Sampler(params, f) closures, sampler.jl:20-24): draws the node's value from
  // its full conditional; returns false when the template has no conjugate form for that node
  std::function<bool(Model&, int /*node*/, struct Rng&)> gibbs;

  int idx(const std::string& s) const {
    for (size_t i = 0; i < nodes.size(); ++i) if (nodes[i].name == s) return (int)i;
    throw std::runtime_error("unknown node " + s);
  }
  const std::vector<double>& in(const std::string& s) const {
    auto it = inputs.find(s);
    if (it == inputs.end()) throw std::runtime_error("missing inputs for node : " + s);
    return it->second;
  }
  const std::vector<double>& val(int i) const { return nodes[i].value; }

  // ---- state record: unobserved stochastic elements in node order -------------------------
  std::vector<int> state_nodes() const {
    std::vector<int> r;
    for (size_t i = 0; i < nodes.size(); ++i) if (nodes[i].stochastic && !nodes[i].observed) r.push_back((int)i);
    return r;
  }
  int state_dim() const { int d = 0; for (int i : state_nodes()) d += nodes[i].len; return d; }
  int state_offset(int node) const {
    int off = 0;
    for (int i : state_nodes()) { if (i == node) return off; off += nodes[i].len; }
    return -1;
  }
  void get_state(double* x) const {
    int o = 0;
    for (int i : state_nodes()) for (double v : nodes[i].value) x[o++] = v;
  }
  int n_monitor() const { int p = 0; for (auto& n : nodes) p += (int)n.monitor.size(); return p; }
  // unlist(m, true): simulation.jl:114-121
  void get_monitor(double* out) const {
    int o = 0;
    for (auto& n : nodes) for (int e : n.monitor) out[o++] = n.value[e];
  }
  std::vector<std::string> names(bool monitoronly) const {  // model.jl:231-240
    std::vector<std::string> r;
    for (auto& n : nodes) {
      auto nm = [&](int e) { return n.scalar ? n.name : n.name + "[" + std::to_string(e + 1) + "]"; };
      if (monitoronly) for (int e : n.monitor) r.push_back(nm(e));
      else for (int e = 0; e < n.len; ++e) r.push_back(nm(e));
    }
    return r;
  }

  // ---- DAG: gettargets (graph.jl:93-103) applied in the Model ctor (model.jl:17-25) --------
  void gettargets_rec(int v, std::vector<int>& values) const {
    for (size_t t = 0; t < nodes.size(); ++t) {
      const Node& tn = nodes[t];
      if (std::find(tn.sources.begin(), tn.sources.end(), v) == tn.sources.end()) continue;
      if (std::find(values.begin(), values.end(), (int)t) == values.end()) values.push_back((int)t);
      if (!tn.stochastic) gettargets_rec((int)t, values);   // terminalkeys = stochastic nodes
    }
  }
  void finalize() {
    // node order must be topological (stands in for tsort, graph.jl:105-108)
    for (size_t i = 0; i < nodes.size(); ++i)
      for (int s : nodes[i].sources)
        if (s >= (int)i) throw std::runtime_error("template node order is not topological");
    for (size_t v = 0; v < nodes.size(); ++v) {
      std::vector<int> t;
      gettargets_rec((int)v, t);
      std::sort(t.begin(), t.end());   // intersect(dependentkeys, ...) keeps dependent (topological) order
      nodes[v].targets = t;
    }
  }
  // setsamplers!: initialization.jl:42-48 ; keys_target(m, nodekeys): model.jl:199-205
  void setsamplers(const std::vector<SamplerSpec>& s) {
    samplers = s;
    for (auto& sp : samplers) {
      std::vector<int> t;
      for (int p : sp.params) for (int q : nodes[p].targets) if (std::find(t.begin(), t.end(), q) == t.end()) t.push_back(q);
      std::sort(t.begin(), t.end());
      sp.targets = t;
      sp.tune = Tune();
    }
  }

  // update!(node, m): dependent.jl:98-101,176-179
  void update(int i) { nodes[i].eval(*this, nodes[i]); }
  void update_all() { for (size_t i = 0; i < nodes.size(); ++i) update((int)i); }
  void update_block(int b) { for (int t : samplers[b].targets) update(t); }  // simulation.jl:166-176

  // setinits!(m, inits): initialization.jl:3-18.  `x` is the state record; observed nodes take
  // their value from the input of the same name (the reference passes them inside `inits`).
  void setinits(const double* x) {
    iter = 0;
    int o = 0;
    for (size_t i = 0; i < nodes.size(); ++i) {
      Node& n = nodes[i];
      if (n.stochastic) {
        if (n.observed) n.value = in(n.name);
        else { n.value.assign(x + o, x + o + n.len); o += n.len; }
        n.eval(*this, n);     // s.distr = s.eval(m): dependent.jl:157-170
      } else {
        n.eval(*this, n);     // dependent.jl:93-96
      }
    }
  }
  void set_state(const double* x) {   // relist!(m, state.value); mcmc.jl:66
    int o = 0;
    for (int i : state_nodes()) { nodes[i].value.assign(x + o, x + o + nodes[i].len); o += nodes[i].len; }
    update_all();
  }

  // node-level logpdf(s, transform): dependent.jl:207-213
  double node_logpdf(int i, bool transform) const {
    const Node& n = nodes[i];
    if (!n.stochastic) return 0.0;   // dependent.jl:57-59
    return logpdf_sub(n.distr, n.value, transform);
  }

  int block_dim(int b) const { int k = 0; for (int p : samplers[b].params) k += nodes[p].len; return k; }

  // unlist(m, block, transform): simulation.jl:110-125 → dependent.jl:192-195
  std::vector<double> unlist_block(int b, bool transform) const {
    std::vector<double> x, y;
    for (int p : samplers[b].params) {
      const Node& n = nodes[p];
      if (transform) { link_sub(n.distr, n.value, y); x.insert(x.end(), y.begin(), y.end()); }
      else x.insert(x.end(), n.value.begin(), n.value.end());
    }
    return x;
  }
  // m[params] = relist(m, x, params, transform): simulation.jl:133-146 → dependent.jl:201-205
  void relist_block(int b, const std::vector<double>& x, bool transform) {
    size_t o = 0;
    std::vector<double> y;
    for (int p : samplers[b].params) {
      Node& n = nodes[p];
      if (transform) { invlink_sub(n.distr, &x[o], n.len, y); n.value = y; }
      else n.value.assign(x.begin() + o, x.begin() + o + n.len);
      o += n.len;
    }
    if (o != x.size()) throw std::runtime_error("incompatible number of values to put in nodes");
  }

  // logpdf!(m, x, block, transform): simulation.jl:77-90 (mutates the model to x)
  double logpdf_block(int b, const std::vector<double>& x, bool transform) {
    const SamplerSpec& sp = samplers[b];
    relist_block(b, x, transform);
    double lp = 0.0;
    // logpdf(m, setdiff(params, targets), transform): simulation.jl:60-67
    for (int p : sp.params) {
      if (std::find(sp.targets.begin(), sp.targets.end(), p) != sp.targets.end()) continue;
      lp += node_logpdf(p, transform);
      if (!std::isfinite(lp)) break;
    }
    for (int t : sp.targets) {
      if (!std::isfinite(lp)) break;
      update(t);
      bool inparams = std::find(sp.params.begin(), sp.params.end(), t) != sp.params.end();
      lp += inparams ? node_logpdf(t, transform) : node_logpdf(t, false);
    }
    return lp;
  }

  // gradlogpdf!(m, x, block, transform; dtype): simulation.jl:47-51 via Calculus.gradient
  // (Calculus.jl >= 0.1.13, REQUIRE:4, not in tree): forward: eps_i = sqrt(eps)*max(1,|x_i|),
  // central: eps_i = cbrt(eps)*max(1,|x_i|).
  std::vector<double> gradlogpdf_fd(int b, const std::vector<double>& x0, bool transform, int dtype) {
    std::vector<double> x = x0, g(x.size());
    const double EPS = 2.220446049250313e-16;
    if (dtype == 1) {
      double f0 = logpdf_block(b, x, transform);
      for (size_t i = 0; i < x.size(); ++i) {
        double h = std::sqrt(EPS) * std::max(1.0, std::fabs(x[i]));
        double old = x[i]; x[i] = old + h;
        double f1 = logpdf_block(b, x, transform);
        g[i] = (f1 - f0) / h; x[i] = old;
      }
    } else {
      for (size_t i = 0; i < x.size(); ++i) {
        double h = std::cbrt(EPS) * std::max(1.0, std::fabs(x[i]));
        double old = x[i];
        x[i] = old + h; double f1 = logpdf_block(b, x, transform);
        x[i] = old - h; double f2 = logpdf_block(b, x, transform);
        g[i] = (f1 - f2) / (2.0 * h); x[i] = old;
      }
    }
    return g;
  }
  // analytic gradient of the block density on the sampler's scale (engine mode; SURVEY.md §7 hard part 3)
  std::vector<double> gradlogpdf_analytic(int b, const std::vector<double>& x, bool transform) {
    relist_block(b, x, transform);
    for (int t : samplers[b].targets) update(t);
    std::vector<double> gj(state_dim(), 0.0), g(x.size());
    joint_grad(*this, gj);
    size_t o = 0;
    for (int p : samplers[b].params) {
      const Node& n = nodes[p];
      int so = state_offset(p);
      for (int e = 0; e < n.len; ++e, ++o) {
        const UDist& d = n.distr.form == Distr::UNI_ARRAY ? n.distr.arr[e] : n.distr.u;
        // theta = invlink(x): d/dx [lp(theta) + log|dtheta/dx|] = dlp/dtheta * dtheta/dx + d(log-Jacobian)/dx
        if (transform && n.distr.form != Distr::MVNORMAL_ISO) {
          double dth, dj; link_chain(d, n.value[e], dth, dj);
          g[o] = gj[so + e] * dth + dj;
        } else g[o] = gj[so + e];
      }
    }
    return g;
  }
  // logpdfgrad!(block, x, dtype): sampler.jl:106-111
  double logpdfgrad(int b, const std::vector<double>& x, bool transform, int dtype, std::vector<double>& grad) {
    grad = dtype == 0 ? gradlogpdf_analytic(b, x, transform) : gradlogpdf_fd(b, x, transform, dtype);
    double logf = logpdf_block(b, x, transform);
    for (double& gi : grad) if (!std::isfinite(gi)) gi = 0.0;
    return logf;
  }
};

}  // namespace orc
