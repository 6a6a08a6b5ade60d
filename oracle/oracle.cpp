// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's sampler hot path
// (jamesonquinn/Mamba.jl, Julia 0.5; cannot run in this environment).  Nothing here is linked
// into or executed by the product; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker or as the CPU baseline.
//
// PARITY STATUS: "parity unpinned" for per-step values — the reference holds no test that
// asserts a logpdf!, gradient, tune value, accept decision, PSRF or ESS number (SURVEY.md §4,
// §8c).  What pins this oracle: (i) the closed-form log posterior and gradient of the line
// model written out in the reference's own doc/samplers/amwg.jl:17-25 and
// doc/samplers/nuts.jl:17-31 (tests/test_oracle_kat.py), (ii) the posterior tables in
// doc/tutorial.rst:427-436, doc/examples/{seeds,rats,pumps,line_amwg_slice}.rst (statistical,
// tests/test_oracle_posterior.py), (iii) scipy for the third-party special functions.
//
// oracle.cpp — engine + C API.
//   mcmc / mcmc_master! / mcmc_worker!      : src/model/mcmc.jl:19-83
//   sample!(m) Gibbs sweep                  : src/model/simulation.jl:93-107
//   model-based sampler closures            : src/samplers/{amwg.jl:47-61, slice.jl:47-58,
//                                             rwm.jl:49-58, nuts.jl:47-56, hmc.jl:62-65, amm.jl:45-59}
//   SamplerVariate(block, ...) tune persistence (iter == 1 ⇒ new tune) : src/samplers/sampler.jl:31-47
#include <cstring>
#include <memory>
#include <string>
#include <thread>

#include "../include/mambacuda.h"
#include "output.hpp"
#include "samplers.hpp"
#include "templates.hpp"

using namespace orc;

namespace {

Vec expand_scale(const SamplerSpec& sp, size_t k) {
  Vec s(k);
  for (size_t i = 0; i < k; ++i) s[i] = sp.scale.size() == 1 ? sp.scale[0] : sp.scale[i];
  return s;
}

// One block update == sampler.eval(m, b) followed by m[params] = value; update!(m, b)
// (simulation.jl:97-103).
void block_update(Model& m, int b, Rng& rng) {
  SamplerSpec& sp = m.samplers[b];
  Tune& t = sp.tune;
  bool tr = sp.transform;
  Vec v = m.unlist_block(b, tr);
  size_t k = v.size();
  bool fresh = (m.iter == 1) || !t.init;    // sampler.jl:40-45
  LogF logf = [&](const Vec& x) { return m.logpdf_block(b, x, tr); };
  LogFGrad logfgrad = [&](const Vec& x, Vec& g) { return m.logpdfgrad(b, x, tr, sp.grad, g); };
  switch (sp.kind) {
    case S_AMWG: {
      if (fresh) { t = Tune(); t.init = true; t.accept.assign(k, 0); t.sigma = expand_scale(sp, k);
                   t.batchsize = sp.batchsize > 0 ? sp.batchsize : 50; t.target = sp.target > 0 ? sp.target : 0.44; }
      bool isadapt = sp.adapt == 1 ? m.iter <= m.burnin : sp.adapt == 0;   // amwg.jl:55-56
      amwg_sample(v, t, logf, isadapt, rng);
      break;
    }
    case S_SLICE_UNI: slice_uni_sample(v, expand_scale(sp, k), logf, rng); break;
    case S_SLICE_MULTI: slice_multi_sample(v, expand_scale(sp, k), logf, rng); break;
    case S_RWM: rwm_sample(v, expand_scale(sp, k), sp.proposal, logf, rng); break;
    case S_NUTS: {
      if (fresh) {   // NUTSTune(x, nutsepsilon(x, f)): nuts.jl:29-30
        t = Tune(); t.init = true; t.target = sp.target > 0 ? sp.target : 0.6;
        t.epsilon = sp.epsilon > 0 ? sp.epsilon : nutsepsilon(v, logfgrad, rng);
      }
      nuts_sample(v, t, logfgrad, m.iter <= m.burnin, rng, sp.max_depth);   // nuts.jl:52
      break;
    }
    case S_HMC: {
      Vec SL;
      if (!sp.scale.empty() && sp.scale.size() == k * k) {   // HMC(params, epsilon, L, Sigma): hmc.jl:19-27
        if (!chol_lower(sp.scale, k, SL)) SL.clear();
      }
      hmc_sample(v, sp.epsilon, sp.L, SL, logfgrad, rng);
      break;
    }
    case S_MALA: {
      Vec SL;
      if (!sp.scale.empty() && sp.scale.size() == k * k) { if (!chol_lower(sp.scale, k, SL)) SL.clear(); }   // MALA(params, epsilon, Sigma): mala.jl:17-23
      mala_sample(v, sp.epsilon, SL, logfgrad, rng);
      break;
    }
    case S_GIBBS: {   // user-defined sampler: f(model) returns the new value of its node, written back below like any other (simulation.jl:99-103)
      if (!m.gibbs || sp.params.size() != 1 || !m.gibbs(m, sp.params[0], rng)) throw std::runtime_error("no Gibbs full conditional for this node");
      v = m.unlist_block(b, tr);
      break;
    }
    case S_AMM: {
      if (fresh) { t = Tune(); t.init = true; t.beta = sp.beta > 0 ? sp.beta : 0.05; t.scale = sp.amm_scale > 0 ? sp.amm_scale : 2.38;
                   chol_lower(sp.scale, k, t.SigmaL); }
      bool isadapt = sp.adapt == 1 ? m.iter <= m.burnin : sp.adapt == 0;
      amm_sample(v, t, logf, isadapt, rng);
      break;
    }
  }
  m.relist_block(b, v, tr);   // relist(block, v) → m[sampler.params] = value
  m.update_block(b);          // update!(m, b)
}

void sweep(Model& m, Rng& rng) {   // sample!(m): simulation.jl:93-107
  m.iter += 1;
  for (size_t b = 0; b < m.samplers.size(); ++b) {
    rng.seek((uint32_t)m.iter, (uint32_t)b, 0);
    block_update(m, (int)b, rng);
  }
}

struct Ctx {
  Model model;
  std::string err;
  int glm_d = 0;
};

SamplerSpec spec_from_desc(const Model& m, const mcu_block_desc& d) {
  SamplerSpec s;
  s.kind = d.kind;
  size_t k = 0;
  auto sn = m.state_nodes();
  for (int i = 0; i < d.n_nodes; ++i) {
    if (d.nodes[i] < 0 || d.nodes[i] >= (int)sn.size()) throw std::runtime_error("bad node id");
    s.params.push_back(sn[d.nodes[i]]); k += m.nodes[sn[d.nodes[i]]].len;
  }
  s.transform = d.kind == MCU_SLICE_UNI || d.kind == MCU_SLICE_MULTI ? d.transform != 0 : d.kind != MCU_GIBBS;   // MALA, like NUTS / HMC / AMWG / AMM: SamplingBlock(model, block, true)
  s.adapt = d.adapt; s.batchsize = d.batchsize; s.proposal = d.proposal; s.L = d.L; s.grad = d.grad;
  s.max_depth = d.max_depth; s.target = d.target; s.epsilon = d.epsilon;
  s.beta = d.beta; s.amm_scale = d.amm_scale;
  if (d.scale) s.scale.assign(d.scale, d.scale + d.n_scale);
  (void)k;
  return s;
}

}  // namespace

extern "C" {

void* orc_create(int template_id, int glm_d) {
  try {
    Ctx* c = new Ctx();
    c->glm_d = glm_d;
    c->model = make_template(template_id, glm_d);
    return c;
  } catch (...) { return nullptr; }
}
void orc_destroy(void* h) { delete (Ctx*)h; }
const char* orc_last_error(void* h) { return ((Ctx*)h)->err.c_str(); }

int orc_set_data(void* h, const char* name, const double* v, int64_t n) {
  Ctx* c = (Ctx*)h;
  c->model.inputs[name].assign(v, v + n);
  return 0;
}
int orc_set_scheme(void* h, int nb, const mcu_block_desc* d) {
  Ctx* c = (Ctx*)h;
  try {
    std::vector<SamplerSpec> s;
    for (int i = 0; i < nb; ++i) s.push_back(spec_from_desc(c->model, d[i]));
    c->model.setsamplers(s);
    return 0;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
int orc_dims(void* h, int* D, int* p) {
  Ctx* c = (Ctx*)h;
  // the GLM's observed node has data-dependent length; state/monitor dims don't depend on it
  *D = c->model.state_dim(); *p = c->model.n_monitor();
  return 0;
}
int orc_names(void* h, int monitoronly, char* buf, size_t buflen) {
  Ctx* c = (Ctx*)h;
  std::string s;
  Model tmp = c->model;
  for (auto& n : tmp.nodes) if (n.observed) n.len = 0;
  std::vector<std::string> nm;
  if (monitoronly) nm = tmp.names(true);
  else for (int i : tmp.state_nodes()) { const Node& n = tmp.nodes[i]; for (int e = 0; e < n.len; ++e) nm.push_back(n.scalar ? n.name : n.name + "[" + std::to_string(e + 1) + "]"); }
  for (size_t i = 0; i < nm.size(); ++i) { if (i) s += "\n"; s += nm[i]; }
  if (s.size() + 1 > buflen) return (int)s.size() + 1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}

// logpdf!(block, x) for B independent states.  state [B×D]; x [B×k] or NULL.
int orc_logpdf(void* h, int block, int64_t B, const double* state, const double* x, double* lp) {
  Ctx* c = (Ctx*)h;
  try {
    Model m = c->model;
    int D = m.state_dim(); int k = m.block_dim(block); bool tr = m.samplers[block].transform;
    for (int64_t i = 0; i < B; ++i) {
      m.setinits(state + i * D);
      Vec xv = x ? Vec(x + i * k, x + (i + 1) * k) : m.unlist_block(block, tr);
      lp[i] = m.logpdf_block(block, xv, tr);
    }
    return 0;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
// logpdf(mc, nodekeys) / logpdf(m, nodekeys) (src/output/modelstats.jl:16-58, src/model/simulation.jl:60-67 without the early exit):
// sum of the selected stochastic nodes' log densities (constrained scale) at B states.  Bit f of `mask` selects the f-th
// unobserved stochastic node in state-record order, bits from the number of such nodes upwards the observed ones (keys(m, :output)).
int orc_logpdf_nodes(void* h, uint32_t mask, int64_t B, const double* state, double* lp) {
  Ctx* c = (Ctx*)h;
  try {
    Model m = c->model;
    const int D = m.state_dim();
    std::vector<int> order = m.state_nodes();
    for (size_t i = 0; i < m.nodes.size(); ++i) if (m.nodes[i].stochastic && m.nodes[i].observed) order.push_back((int)i);
    for (int64_t i = 0; i < B; ++i) {
      m.setinits(state + i * D);
      double s = 0.0;
      for (size_t f = 0; f < order.size(); ++f) if ((mask >> f) & 1u) s += m.node_logpdf(order[f], false);
      lp[i] = s;
    }
    return 0;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
// predict(mc, nodekeys = keys(m, :output)) (src/output/modelstats.jl:63-96): rand(m[key]) for every element of the observed nodes at
// B states.  Engine RNG contract for this call: stream (seed, chain = stream_id, iteration = record index, block 0, kind 15); a Normal
// element takes one normal draw, a discrete element one uniform and inverts the CDF by sequential search from 0.
static double rand_udist(const UDist& d, Rng& rng) {
  if (d.k == D_NORMAL) return d.a + d.b * rng.normal();
  const double u = rng.uniform();
  if (d.k == D_BERNOULLI) return u < d.a ? 1.0 : 0.0;
  if (d.k == D_LAPLACE) { const double c = u - 0.5; return d.a - d.b * (c < 0 ? -1.0 : 1.0) * std::log(1.0 - 2.0 * std::fabs(c)); }   // inverse CDF
  if (d.k == D_BINOMIAL) {
    const double n = d.a, p = d.b, q = 1.0 - p;
    if (!(p > 0.0)) return 0.0;
    if (!(q > 0.0)) return n;
    const double ratio = p / q;
    double pmf = std::exp(n * std::log(q)), cdf = pmf, k = 0.0;
    while (u >= cdf && k < n) { pmf *= (n - k) / (k + 1.0) * ratio; k += 1.0; cdf += pmf; }
    return k;
  }
  if (d.k == D_POISSON) {
    double pmf = std::exp(-d.a), cdf = pmf, k = 0.0;
    while (u >= cdf && k < 100000.0) { k += 1.0; pmf *= d.a / k; cdf += pmf; }
    return k;
  }
  throw std::runtime_error("predict: no sampler for this observed distribution");
}
int orc_predict(void* h, uint64_t seed, uint32_t stream_id, int64_t B, const double* state, double* out, int64_t* n_out) {
  Ctx* c = (Ctx*)h;
  try {
    Model m = c->model;
    const int D = m.state_dim();
    std::vector<int> obs;
    int64_t L = 0;
    for (size_t i = 0; i < m.nodes.size(); ++i) if (m.nodes[i].stochastic && m.nodes[i].observed) { obs.push_back((int)i); L += m.nodes[i].len; }
    if (n_out) *n_out = L;
    if (!out) return 0;
    for (int64_t i = 0; i < B; ++i) {
      m.setinits(state + i * D);
      PhiloxRng rng(seed, stream_id); rng.seek((uint32_t)i, 0, 15);
      int64_t o = 0;
      for (int q : obs) {
        const Node& n = m.nodes[q];
        for (int e = 0; e < n.len; ++e) {
          double v;
          if (n.distr.form == Distr::UNI) v = rand_udist(n.distr.u, rng);
          else if (n.distr.form == Distr::UNI_ARRAY) v = rand_udist(n.distr.arr[e], rng);
          else if (n.distr.form == Distr::MVNORMAL_ISO) v = n.distr.mu[e] + n.distr.sigma * rng.normal();
          else throw std::runtime_error("predict: node without a distribution");
          out[i * L + o++] = v;
        }
      }
    }
    return 0;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
// n draws of rand(proposal(0, 1)) from one Philox stream (tests: moments and CDF of the RWM proposal kernels)
void orc_rwm_draws(int proposal, uint64_t seed, int64_t n, double* out) {
  PhiloxRng rng(seed, 0); rng.seek(1, 0, 3);
  for (int64_t i = 0; i < n; ++i) out[i] = rwm_draw(proposal, rng);
}
int orc_gradlogpdf(void* h, int block, int grad_mode, int64_t B, const double* state, const double* x, double* lp, double* g) {
  Ctx* c = (Ctx*)h;
  try {
    Model m = c->model;
    int D = m.state_dim(); int k = m.block_dim(block); bool tr = m.samplers[block].transform;
    for (int64_t i = 0; i < B; ++i) {
      m.setinits(state + i * D);
      Vec xv = x ? Vec(x + i * k, x + (i + 1) * k) : m.unlist_block(block, tr);
      Vec gv;
      double l = m.logpdfgrad(block, xv, tr, grad_mode, gv);
      if (lp) lp[i] = l;
      std::memcpy(g + i * k, gv.data(), sizeof(double) * k);
    }
    return 0;
  } catch (std::exception& e) { c->err = e.what(); return -1; }
}
// unlist(block) of a state: the block vector on the sampler's scale.
int orc_unlist(void* h, int block, const double* state, double* x) {
  Ctx* c = (Ctx*)h;
  Model m = c->model;
  m.setinits(state);
  Vec v = m.unlist_block(block, m.samplers[block].transform);
  std::memcpy(x, v.data(), sizeof(double) * v.size());
  return (int)v.size();
}

// Tune blob layout per chain (shared by specification with mcu_get_state), blocks in scheme order:
//   AMWG : m, adapt, sigma[k], accept[k]            NUTS : adapt, alpha, epsilon, epsilonbar, Hbar, m, mu, nalpha
//   AMM  : adapt, m, Mv[k], Mvv[k*k], SigmaLm[k*k]  Slice / RWM / HMC : nothing
static int64_t tune_size(const Model& m) {
  int64_t n = 0;
  for (size_t b = 0; b < m.samplers.size(); ++b) {
    int64_t k = m.block_dim((int)b);
    switch (m.samplers[b].kind) {
      case S_AMWG: n += 2 + 2 * k; break;
      case S_NUTS: n += 8; break;
      case S_AMM: n += 2 + k + 2 * k * k; break;
      default: break;
    }
  }
  return n;
}
static void write_tune(const Model& m, double* o) {
  for (size_t b = 0; b < m.samplers.size(); ++b) {
    const Tune& t = m.samplers[b].tune; size_t k = (size_t)m.block_dim((int)b);
    switch (m.samplers[b].kind) {
      case S_AMWG:
        *o++ = (double)t.m; *o++ = t.adapt ? 1.0 : 0.0;
        for (size_t i = 0; i < k; ++i) *o++ = t.sigma[i];
        for (size_t i = 0; i < k; ++i) *o++ = (double)t.accept[i];
        break;
      case S_NUTS:
        *o++ = t.adapt ? 1.0 : 0.0; *o++ = t.alpha; *o++ = t.epsilon; *o++ = t.epsilonbar; *o++ = t.Hbar;
        *o++ = (double)t.m; *o++ = t.mu; *o++ = (double)t.nalpha;
        break;
      case S_AMM:
        *o++ = t.adapt ? 1.0 : 0.0; *o++ = (double)t.m;
        for (size_t i = 0; i < k; ++i) *o++ = t.Mv.size() == k ? t.Mv[i] : 0.0;
        for (size_t i = 0; i < k * k; ++i) *o++ = t.Mvv.size() == k * k ? t.Mvv[i] : 0.0;
        for (size_t i = 0; i < k * k; ++i) *o++ = t.SigmaLm.size() == k * k ? t.SigmaLm[i] : 0.0;
        break;
      default: break;
    }
  }
}

int64_t orc_tune_size(void* h) { return tune_size(((Ctx*)h)->model); }
int64_t orc_kept(int64_t iters, int64_t burnin, int64_t thin) { return iters > burnin ? (iters - burnin) / thin : 0; }
// number of i in (first, first + iters] with i > burnin and (i - burnin) % thin == 0 (restart segments, mcmc.jl:3-16)
int64_t orc_kept2(int64_t first, int64_t iters, int64_t burnin, int64_t thin) {
  auto upto = [&](int64_t i) { return i > burnin ? (i - burnin) / thin : 0; };
  return upto(first + iters) - upto(first);
}

// inverse of write_tune: the tune records of a restarted chain (mcmc(mc, iters) keeps the Sampler.tune objects: mcmc.jl:8, sampler.jl:40-45);
// the fields that are not part of the blob are the constructor constants the fresh branch of block_update sets
static void read_tune(Model& m, const double* o) {
  for (size_t b = 0; b < m.samplers.size(); ++b) {
    SamplerSpec& sp = m.samplers[b]; Tune& t = sp.tune; const size_t k = (size_t)m.block_dim((int)b);
    switch (sp.kind) {
      case S_AMWG:
        t = Tune(); t.init = true; t.batchsize = sp.batchsize > 0 ? sp.batchsize : 50; t.target = sp.target > 0 ? sp.target : 0.44;
        t.m = (long)*o++; t.adapt = *o++ != 0.0;
        t.sigma.assign(o, o + k); o += k;
        t.accept.resize(k); for (size_t i = 0; i < k; ++i) t.accept[i] = (long)*o++;
        break;
      case S_NUTS:
        t = Tune(); t.init = true; t.target = sp.target > 0 ? sp.target : 0.6;
        t.adapt = *o++ != 0.0; t.alpha = *o++; t.epsilon = *o++; t.epsilonbar = *o++; t.Hbar = *o++; t.m = (long)*o++; t.mu = *o++; t.nalpha = (long)*o++;
        break;
      case S_AMM:
        t = Tune(); t.init = true; t.beta = sp.beta > 0 ? sp.beta : 0.05; t.scale = sp.amm_scale > 0 ? sp.amm_scale : 2.38;
        chol_lower(sp.scale, k, t.SigmaL);
        t.adapt = *o++ != 0.0; t.m = (long)*o++;
        t.Mv.assign(o, o + k); o += k; t.Mvv.assign(o, o + k * k); o += k * k; t.SigmaLm.assign(o, o + k * k); o += k * k;
        break;
      default: break;
    }
  }
}

// mcmc(model, data, inits, iters; burnin, thin, chains): mcmc.jl:19-83.
//  inits [n_inits × D]; chain c (global id g = chain_ids ? chain_ids[c] : chain_offset + c) starts from record g % n_inits,
//  plus optional N(0, jitter_sd²) jitter on the unconstrained scale (Philox stream kind 1, iter 0,
//  block 0, draw j = state element index).
//  Restart (mcmc(mc, iters): mcmc.jl:3-16) when iter0 > 0: inits is [n_chains × D] (record c = state of chain c after iteration
//  iter0), tune_in [n_chains × orc_tune_size] the tune records written by an earlier call; iterations iter0 + 1 .. iter0 + iters run.
//  out [kept × p × n_chains] column-major (may be NULL), kept = orc_kept2(iter0, iters, burnin, thin); final_state [n_chains × D]
//  (may be NULL); tune_out [n_chains × orc_tune_size] (may be NULL);
//  margins [n_chains × iters] (may be NULL): the smallest decision margin of each iteration (samplers.hpp note_margin);
//  ext_u: NULL for Philox mode, else [n_chains × n_per_chain] uniforms consumed sequentially.
//  nthreads: chains are distributed over this many std::threads (the reference would use pmap
//  over worker processes, utils.jl:91-98 — disabled in this version).
int orc_run2(void* h, int64_t n_chains, int64_t chain_offset, const int64_t* chain_ids, uint64_t seed, const double* inits, int64_t n_inits,
             double jitter_sd, int64_t iter0, const double* tune_in, int64_t iters, int64_t burnin, int64_t thin, double* out,
             double* final_state, double* tune_out, double* margins, const double* ext_u, int64_t n_per_chain, int nthreads, int partial) {
  Ctx* c = (Ctx*)h;
  // partial != 0: the call is one segment of a longer run whose burn-in may extend past it (the caller continues with iter0 > 0)
  if (iter0 == 0 && !partial && iters <= burnin) { c->err = "burnin is greater than or equal to iters"; return MCU_ERR_ARG; }   // mcmc.jl:22-23
  if (n_inits < 1) { c->err = "fewer initial values than chains"; return MCU_ERR_ARG; }              // mcmc.jl:24-25
  if (thin < 1) { c->err = "thin must be positive"; return MCU_ERR_ARG; }
  if (iter0 > 0 && n_inits != n_chains) { c->err = "a restart needs one state record per chain"; return MCU_ERR_ARG; }
  const int D = c->model.state_dim(); const int p = c->model.n_monitor();
  const int64_t kept = orc_kept2(iter0, iters, burnin, thin);
  const int64_t row0 = iter0 > burnin ? (iter0 - burnin) / thin : 0;
  const int64_t nt = tune_size(c->model);
  if (nthreads < 1) nthreads = 1;
  std::vector<std::string> errs(nthreads);
  auto worker = [&](int tid) {
    try {
      for (int64_t k = tid; k < n_chains; k += nthreads) {   // mcmc_worker!: mcmc.jl:62-83
        Model m = c->model;                                    // deepcopy(m)
        m.burnin = burnin;
        int64_t g = chain_ids ? chain_ids[k] : chain_offset + k;
        std::unique_ptr<Rng> rng;
        if (ext_u) rng.reset(new ExternalRng(ext_u + k * n_per_chain, (size_t)n_per_chain));
        else rng.reset(new PhiloxRng(seed, (uint32_t)g));
        const int64_t rec = iter0 > 0 ? k : g % n_inits;
        std::vector<double> x0(inits + rec * D, inits + (rec + 1) * D);
        m.setinits(x0.data());
        if (jitter_sd > 0 && iter0 == 0) {
          PhiloxRng jr(seed, (uint32_t)g); jr.seek(0, 0, 1);
          for (int i : m.state_nodes()) {
            Node& n = m.nodes[i]; Vec y, z;
            link_sub(n.distr, n.value, y);
            for (double& yi : y) yi += jitter_sd * jr.normal();
            invlink_sub(n.distr, y.data(), y.size(), z);
            n.value = z;
          }
          std::vector<double> xs(D); m.get_state(xs.data()); m.setinits(xs.data());
        }
        for (auto& sp : m.samplers) sp.tune = Tune();
        if (iter0 > 0) { m.iter = (long)iter0; if (tune_in && nt > 0) read_tune(m, tune_in + k * nt); }
        std::vector<double> mon(p);
        for (int64_t i = iter0 + 1; i <= iter0 + iters; ++i) {
          double margin = INFINITY;
          g_margin_slot = margins ? &margin : nullptr;
          sweep(m, *rng);
          g_margin_slot = nullptr;
          if (margins) margins[k * iters + (i - iter0 - 1)] = margin;
          if (i > burnin && (i - burnin) % thin == 0 && out) {          // mcmc.jl:76-78
            m.get_monitor(mon.data());
            int64_t row = (i - burnin) / thin - 1 - row0;                // iters2inds: chains.jl:66-69,81-87
            for (int j = 0; j < p; ++j) out[row + kept * (j + (int64_t)p * k)] = mon[j];
          }
        }
        if (final_state) m.get_state(final_state + k * D);
        if (tune_out) write_tune(m, tune_out + k * nt);
      }
    } catch (std::exception& e) { errs[tid] = e.what(); g_margin_slot = nullptr; }
  };
  if (nthreads == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  for (auto& e : errs) if (!e.empty()) { c->err = e; return -1; }
  return 0;
}
int orc_run(void* h, int64_t n_chains, int64_t chain_offset, uint64_t seed, const double* inits, int64_t n_inits,
            double jitter_sd, int64_t iters, int64_t burnin, int64_t thin, double* out, double* final_state,
            double* tune_out, const double* ext_u, int64_t n_per_chain, int nthreads) {
  return orc_run2(h, n_chains, chain_offset, nullptr, seed, inits, n_inits, jitter_sd, 0, nullptr, iters, burnin, thin, out, final_state,
                  tune_out, nullptr, ext_u, n_per_chain, nthreads, 0);
}

// gelmandiag(c; alpha, transform): linkcode per column (-1 heuristic / 0 identity / 1 log) or NULL = no transform
int orc_gelmandiag(const double* chains, int64_t n, int64_t p, int64_t m, double alpha, const int* linkcode, double* psrf) {
  if (m < 2) return MCU_ERR_ARG;   // gelmandiag.jl:6-7
  if (linkcode) {
    std::vector<double> cc;
    link_chains(chains, (size_t)n, (size_t)p, (size_t)m, linkcode, cc);
    gelmandiag(cc.data(), (size_t)n, (size_t)p, (size_t)m, alpha, psrf);
  } else gelmandiag(chains, (size_t)n, (size_t)p, (size_t)m, alpha, psrf);
  return 0;
}
int orc_summarystats(const double* chains, int64_t n, int64_t p, int64_t m, int etype, int64_t batch, double* out) {
  summarystats(chains, (size_t)n, (size_t)p, (size_t)m, etype, (size_t)batch, out);
  return 0;
}
// link layer probe (transformdistribution.jl:6-93) for one univariate distribution: out = { link(x), invlink(link(x)), logpdf(x, true) - logpdf(x, false) }
// kind: DKind of dist.hpp (9 Uniform(a, b), 10 Beta(a, b), 11 Truncated(Normal(a, b), lo, hi), 2 InverseGamma, ...)
void orc_link(int kind, double a, double b, double lo, double hi, double x, double* out) {
  UDist d; d.k = (DKind)kind; d.a = a; d.b = b; d.lo = lo; d.hi = hi;
  out[0] = link(d, x); out[1] = invlink(d, out[0]); out[2] = logpdf(d, x, true) - logpdf(d, x, false);
}
double orc_udist_logpdf(int kind, double a, double b, double lo, double hi, double x) {
  UDist d; d.k = (DKind)kind; d.a = a; d.b = b; d.lo = lo; d.hi = hi;
  return logpdf_sub(d, x, false);
}
double orc_fquantile(double q, double d1, double d2) { return fquantile(q, d1, d2); }
double orc_digamma(double x) { return digamma(x); }
void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
// stream probe: the first n draws of (seed, chain, iter, block, kind); kinds[i] 0 = uniform, 1 = normal
void orc_draws(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t block, uint32_t kind, int n, const int* kinds, double* out) {
  PhiloxRng r(seed, chain); r.seek(iter, block, kind);
  for (int i = 0; i < n; ++i) out[i] = kinds[i] ? r.normal() : r.uniform();
}
// stand-alone sampler faces on the reference's closed-form line logf (doc/samplers/amwg.jl:17-25,
// doc/samplers/nuts.jl:17-31): theta = (b0, b1, log s2).  Used by the posterior pin test.
static double line_logf(const Vec& x, Vec* grad) {
  const double X[5] = {1, 2, 3, 4, 5}, Y[5] = {1, 3, 3, 3, 5};
  double b0 = x[0], b1 = x[1], logs2 = x[2];
  double rr = 0, sr = 0, sxr = 0;
  for (int i = 0; i < 5; ++i) { double r = Y[i] - b0 - b1 * X[i]; rr += r * r; sr += r; sxr += X[i] * r; }
  double logf = (-0.5 * 5 - 0.001) * logs2 - (0.5 * rr + 0.001) / std::exp(logs2) - 0.5 * b0 * b0 / 1000 - 0.5 * b1 * b1 / 1000;
  if (grad) {
    grad->resize(3);
    (*grad)[0] = sr / std::exp(logs2) - b0 / 1000;
    (*grad)[1] = sxr / std::exp(logs2) - b1 / 1000;
    (*grad)[2] = -0.5 * 5 - 0.001 + (0.5 * rr + 0.001) / std::exp(logs2);
  }
  return logf;
}
double orc_line_logf(const double* x, double* grad) {
  Vec xv(x, x + 3), g;
  double l = line_logf(xv, grad ? &g : nullptr);
  if (grad) for (int i = 0; i < 3; ++i) grad[i] = g[i];
  return l;
}
// which: 0 AMWG(1.0) [doc/samplers/amwg.jl:28-35], 1 NUTS [doc/samplers/nuts.jl:34-43],
//        2 SliceUnivariate(width 1,1,2), 3 SliceMultivariate [doc/samplers/slice.jl:31-40],
//        4 AMWG(beta)+SliceMultivariate(log s2; 5.0) [doc/examples/line_amwg_slice.jl:35-43]
//        5 AMM(eye(3)) [doc/samplers/amm.jl:28-35], 6 / 7 HMC(0.1, 50) without / with Sigma = eye(3) [doc/samplers/hmc.jl:37-51],
//        8 / 9 MALA(0.1) without / with Sigma = eye(3) [doc/samplers/mala.jl:37-50], 10 RWM([0.5, 0.25, 1.0], SymUniform) [doc/samplers/rwm.jl:28-35]
// out [n × 3] column-major, columns b0, b1, s2 = exp(theta3).
int orc_standalone_line(int which, uint64_t seed, int64_t n, int64_t burnin, double* out) {
  PhiloxRng rng(seed, 0);
  Vec theta = {0.0, 0.0, 0.0};
  LogF logf = [](const Vec& x) { return line_logf(x, nullptr); };
  LogFGrad lfg = [](const Vec& x, Vec& g) { return line_logf(x, &g); };
  Tune t; t.accept.assign(3, 0); t.sigma.assign(3, 1.0);
  Tune tn;
  if (which == 1) { rng.seek(0, 0, 0); tn.epsilon = nutsepsilon(theta, lfg, rng); tn.target = 0.6; }
  Tune tb; tb.accept.assign(2, 0); tb.sigma.assign(2, 1.0);
  Tune tm; tm.beta = 0.05; tm.scale = 2.38; chol_lower({1, 0, 0, 0, 1, 0, 0, 0, 1}, 3, tm.SigmaL);   // AMMVariate(x, eye(3), logf)
  const Vec eye3 = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  Vec eyeL; chol_lower(eye3, 3, eyeL);
  for (int64_t i = 1; i <= n; ++i) {
    rng.seek((uint32_t)i, 0, 0);
    switch (which) {
      case 0: amwg_sample(theta, t, logf, i <= burnin, rng); break;
      case 1: nuts_sample(theta, tn, lfg, i <= burnin, rng, 0); break;
      case 2: slice_uni_sample(theta, {1.0, 1.0, 2.0}, logf, rng); break;
      case 3: slice_multi_sample(theta, {1.0, 1.0, 2.0}, logf, rng); break;
      case 4: {
        Vec beta = {theta[0], theta[1]}; double ls2 = theta[2];
        amwg_sample(beta, tb, [&](const Vec& x) { return line_logf({x[0], x[1], ls2}, nullptr); }, i <= burnin, rng);
        Vec l = {ls2};
        rng.seek((uint32_t)i, 1, 0);
        slice_multi_sample(l, {5.0}, [&](const Vec& x) { return line_logf({beta[0], beta[1], x[0]}, nullptr); }, rng);
        theta = {beta[0], beta[1], l[0]};
        break;
      }
      case 5: amm_sample(theta, tm, logf, i <= burnin, rng); break;
      case 6: hmc_sample(theta, 0.1, 50, Vec(), lfg, rng); break;
      case 7: hmc_sample(theta, 0.1, 50, eyeL, lfg, rng); break;
      case 8: mala_sample(theta, 0.1, Vec(), lfg, rng); break;
      case 9: mala_sample(theta, 0.1, eyeL, lfg, rng); break;
      case 10: rwm_sample(theta, {0.5, 0.25, 1.0}, 1, logf, rng); break;
      default: return -1;
    }
    out[(i - 1)] = theta[0]; out[(i - 1) + n] = theta[1]; out[(i - 1) + 2 * n] = std::exp(theta[2]);
  }
  return 0;
}

}  // extern "C"
