// ORACLE (test infrastructure, never shipped, never on the product path).
//
// C entry points over the CPU restatement, loaded with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs only.
// Nothing under quicksand_b200/ may load this library.
//
// Parity pinning: the density functions are checked against all 26 known-answer values
// of test/testsuite.t:159-187 (tests/test_oracle_goldens.py) and the kernels against the
// expectation truths of testsuite.t:621-691, 719-730, 776-787, 812-823.  The reference
// itself cannot run here (no Terra), so per-step kernel behaviour is pinned only by this
// restatement: "parity pinned for densities, unpinned for kernels" (SURVEY.md 8c).
#include <chrono>
#include <cstring>
#include "infer.hpp"

using namespace qso;

extern "C" {

typedef struct {
  int32_t kind;        // 1 HMC, 2 DRIFT, 3 HARM
  int32_t doAdapt;     // doStepSizeAdapt / doScaleAdapt
  double stepSize;     // HMC
  uint64_t numSteps;   // HMC
  double scale;        // Drift / HARM
  double weight;       // MixtureKernel weight (ignored for a single kernel)
  int32_t lexicalScaleSharing, reserved;   // Drift (drift.t:58-60, 71-73)
} qso_kernel_spec;

typedef struct {
  uint64_t grad_evals, ad_nodes, rng_draws;
  uint64_t props_made, props_accepted;
  double seconds;
} qso_stats;

// ---- densities: which = 0 bernoulli(val,p) 1 uniform(x,lo,hi) 2 gaussian(x,mu,sigma) 3 gamma(x,k,theta)
//      4 beta(x,a,b) 5 binomial(s,p,n) 6 poisson(k,mu) 7 categorical(val, p0..p{n-1}) 8 dirichlet(theta[n], alpha[n]) 9 log_gamma(x)
double qso_logprob(int which, const double* a, int n) {
  switch (which) {
    case 0: return D::bernoulli_logprob<double>(a[0] != 0.0, a[1]);
    case 1: return D::uniform_logprob<double>(a[0], a[1], a[2]);
    case 2: return D::gaussian_logprob<double>(a[0], a[1], a[2]);
    case 3: return D::gamma_logprob<double>(a[0], a[1], a[2]);
    case 4: return D::beta_logprob<double>(a[0], a[1], a[2]);
    case 5: return D::binomial_logprob((int)a[0], a[1], (int)a[2]);
    case 6: return D::poisson_logprob((int)a[0], a[1]);
    case 7: return D::categorical_logprob((int)a[0], a + 1, n - 1);
    case 8: return D::dirichlet_logprob<double>(a, a + n / 2, n / 2);
    case 10: { std::vector<double> x(a, a + n); return simplex_lpincr<double>(x); }
    case 9: return D::log_gamma<double>(a[0]);
  }
  return NAN;
}

// d/dx of a density through the tape AD (checks the AD restatement itself against finite differences in tests)
double qso_logprob_grad(int which, const double* a, int wrt, double* grad_out) {
  ad::recoverMemory();
  ad::num v[3] = {ad::num(a[0]), ad::num(a[1]), ad::num(a[2])};
  ad::num lp;
  switch (which) {
    case 1: lp = D::uniform_logprob<ad::num>(v[0], v[1], v[2]); break;
    case 2: lp = D::gaussian_logprob<ad::num>(v[0], v[1], v[2]); break;
    case 3: lp = D::gamma_logprob<ad::num>(v[0], v[1], v[2]); break;
    case 4: lp = D::beta_logprob<ad::num>(v[0], v[1], v[2]); break;
    default: return NAN;
  }
  double r = lp.val();
  ad::grad(lp);
  *grad_out = v[wrt].adj();
  ad::recoverMemory();
  return r;
}

// ---- the injected uniform stream
void qso_uniforms(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot, int n, double* out) {
  Rng& r = rng();
  r.seed = seed; r.chain = chain; r.at(iter, slot);
  for (int i = 0; i < n; ++i) out[i] = r.random();
}
void qso_gaussians(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot0, int n, double mu, double sigma, double* out) {
  Rng& r = rng();
  r.seed = seed; r.chain = chain;
  for (int i = 0; i < n; ++i) { r.at(iter, slot0 + i); out[i] = D::gaussian_sample(mu, sigma); }
}
void qso_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

// ---- programs
void* qso_program_indep(int K, const int* kind, const double* p0, const double* p1, int ret_mode, double ret_div) {
  IndepErps* p = new IndepErps();
  p->kind.assign(kind, kind + K); p->p0.assign(p0, p0 + K); p->p1.assign(p1, p1 + K);
  p->ret_mode = ret_mode; p->ret_div = ret_div; p->retdim = ret_mode == 0 ? 1 : K;
  return p;
}
void qso_program_indep_lexids(void* prog, int K, const int* lexid) { ((IndepErps*)prog)->lexid.assign(lexid, lexid + K); }
void qso_program_indep_factors(void* prog, int K, const double* fmu, const double* fsigma) {
  ((IndepErps*)prog)->fmu.assign(fmu, fmu + K); ((IndepErps*)prog)->fsigma.assign(fsigma, fsigma + K);
}
void qso_program_indep_latent(void* prog, int K, const int* ref, const double* b, const int* tf) {
  IndepErps* p = (IndepErps*)prog;
  p->lat_ref.assign(ref, ref + 2 * K); p->lat_b.assign(b, b + 2 * K); p->lat_tf.assign(tf, tf + 2 * K);
}
void qso_program_indep_conditions(void* prog, int K, const double* lo, const double* hi) {
  IndepErps* p = (IndepErps*)prog;
  p->cond_lo.assign(lo, lo + K); p->cond_hi.assign(hi, hi + K);
}
void* qso_program_gauss_prec(int d, const double* A) {
  GaussPrecision* p = new GaussPrecision();
  p->d = d; p->A.assign(A, A + (size_t)d * d); p->retdim = d;
  return p;
}
void* qso_program_logistic(int N, int d, const double* X, const unsigned char* y, double prior_sigma) {
  Logistic* p = new Logistic();
  p->N = N; p->d = d; p->X.assign(X, X + (size_t)N * d); p->y.assign(y, y + N); p->prior_sigma = prior_sigma; p->retdim = d;
  return p;
}
void* qso_program_eight_schools(int J, const double* y, const double* sigma) {
  EightSchools* p = new EightSchools();
  p->J = J; p->y.assign(y, y + J); p->sigma.assign(sigma, sigma + J); p->retdim = J + 2;
  return p;
}
void* qso_program_funnel(int d) {
  Funnel* p = new Funnel();
  p->d = d; p->retdim = d;
  return p;
}
void qso_program_free(void* p) { delete (Program*)p; }
int qso_program_retdim(void* p) { return ((Program*)p)->retdim; }

// AnnealingKernel schedule applied by the following qso_run calls of this thread (n = 0: none)
static thread_local std::vector<double> g_temps;
void qso_set_temperatures(const double* temps, uint64_t n) { g_temps.clear(); if (temps && n) g_temps.assign(temps, temps + n); }

// Teacher-forcing record of a single HMC kernel, filled by the following qso_run call of this thread (all NULL: off).
//   x0, g0, p0 [nchains][iters][n]   lp0 [nchains][iters]   (state entering each trajectory: positions, cached gradient,
//   sampled momenta, currTrace.logprob)      leap_mom, leap_grad [nchains][iters][L][n]   leap_lp [nchains][iters][L]
static thread_local double *g_tx0 = nullptr, *g_tg0 = nullptr, *g_tp0 = nullptr, *g_tlp0 = nullptr, *g_tmom = nullptr, *g_tgrad = nullptr, *g_tlp = nullptr;
void qso_set_traj_record(double* x0, double* g0, double* p0, double* lp0, double* leap_mom, double* leap_grad, double* leap_lp) {
  g_tx0 = x0; g_tg0 = g0; g_tp0 = p0; g_tlp0 = lp0; g_tmom = leap_mom; g_tgrad = leap_grad; g_tlp = leap_lp;
}

static Kernel* make_kernel(const qso_kernel_spec* specs, int nspecs) {
  auto one = [](const qso_kernel_spec& s) -> Kernel* {
    switch (s.kind) {
      case 1: return new HMCKernel(s.stepSize, s.numSteps, s.doAdapt != 0);
      case 2: return new DriftKernel(s.scale, s.doAdapt != 0, s.lexicalScaleSharing != 0);
      case 3: return new HARMKernel(s.scale, s.doAdapt != 0);
      case 4: return new TraceMHKernel();
    }
    return nullptr;
  };
  if (nspecs == 1) return one(specs[0]);
  MixtureKernel* m = new MixtureKernel();
  for (int i = 0; i < nspecs; ++i) { m->kernels.emplace_back(one(specs[i])); m->weights.push_back(specs[i].weight); }
  return m;
}

// Number of unconstrained real components of the program (one forward run).
int qso_program_dim(void* prog) {
  Rng& r = rng(); r.seed = 0; r.chain = 0; r.at(ITER_INIT, 0);
  Trace<double> t; t.init((Program*)prog, true);
  return (int)t.countRealComps();
}

// Run chains [chain_begin, chain_begin+nchains) one after the other, like nchains runs of doMCMC.
//   samples  [nchains][nsamps][retdim]   logprobs [nchains][nsamps]
// Optional records (NULL to skip), iters = burnin + nsamps*lag, n = program dim:
//   rec_state [nchains][iters][n]  rec_accept/rec_kidx [nchains][iters] (int32)  rec_dh/rec_step [nchains][iters]
//   rec_leap  [nchains][iters][L][n]  (single HMC kernel only)   rec_init [nchains][n]
int qso_run(void* prog, const qso_kernel_spec* specs, int nspecs, uint64_t seed, uint32_t chain_begin, uint32_t nchains,
            uint64_t nsamps, uint64_t burnin, uint64_t lag, double* samples, double* logprobs,
            double* rec_state, int32_t* rec_accept, double* rec_dh, double* rec_step, int32_t* rec_kidx, double* rec_leap,
            double* rec_init, qso_stats* stats) {
  Program* program = (Program*)prog;
  uint64_t iters = burnin + nsamps * lag;
  if (iters >= ITER_SEARCH) return 1;
  bool want_rec = rec_state || rec_accept || rec_dh || rec_step || rec_kidx || rec_leap;
  uint64_t ge = 0, pm = 0, pa = 0;
  uint64_t nodes0 = ad::G().nodes;
  Rng& r = rng();
  r.ndraws = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (uint32_t c = 0; c < nchains; ++c) {
    r.seed = seed; r.chain = chain_begin + c;
    ad::recoverMemory();
    Kernel* inner = make_kernel(specs, nspecs);
    if (!inner) return 2;
    std::unique_ptr<Kernel> k;
    if (g_temps.empty()) k.reset(inner);
    else { AnnealingKernel* a = new AnnealingKernel(); a->k.reset(inner); a->temps = g_temps; k.reset(a); }
    Recorder rec;
    const bool want_traj = g_tx0 && nspecs == 1 && specs[0].kind == 1;
    rec.want_traj = want_traj;
    recorder() = (want_rec || want_traj) ? &rec : nullptr;
    std::vector<Sample> samps;
    samps.reserve(nsamps);
    if (rec_init) {
      // replay the initial forward sample to report it (same sub-streams, no side effect on the run)
      r.at(ITER_INIT, 0);
      Trace<double> t; t.init(program, true);
      std::vector<double> x0;
      for (auto& rc : t.choices) rc.getUnboundedRealComps(x0);
      std::memcpy(rec_init + (size_t)c * x0.size(), x0.data(), sizeof(double) * x0.size());
    }
    doMCMC(program, k.get(), samps, nsamps, burnin, lag);
    recorder() = nullptr;
    int retdim = program->retdim;
    for (size_t s = 0; s < samps.size(); ++s) {
      if (samples) std::memcpy(samples + ((size_t)c * nsamps + s) * retdim, samps[s].value.data(), sizeof(double) * retdim);
      if (logprobs) logprobs[(size_t)c * nsamps + s] = samps[s].logprob;
    }
    if (want_rec) {
      size_t n = iters ? rec.state.size() / iters : 0;
      if (rec_state) std::memcpy(rec_state + (size_t)c * iters * n, rec.state.data(), sizeof(double) * rec.state.size());
      if (rec_accept) for (size_t i = 0; i < rec.accepted.size(); ++i) rec_accept[(size_t)c * iters + i] = rec.accepted[i];
      if (rec_dh) std::memcpy(rec_dh + (size_t)c * iters, rec.dH.data(), sizeof(double) * rec.dH.size());
      if (rec_step) std::memcpy(rec_step + (size_t)c * iters, rec.stepsize.data(), sizeof(double) * rec.stepsize.size());
      if (rec_kidx) for (size_t i = 0; i < rec.kernel_index.size(); ++i) rec_kidx[(size_t)c * iters + i] = rec.kernel_index[i];
      if (rec_leap && nspecs == 1 && specs[0].kind == 1)
        std::memcpy(rec_leap + (size_t)c * iters * specs[0].numSteps * n, rec.leapfrog_pos.data(), sizeof(double) * rec.leapfrog_pos.size());
    }
    if (want_traj) {
      const size_t n = iters ? rec.traj_x0.size() / iters : 0, L = specs[0].numSteps;
      std::memcpy(g_tx0 + (size_t)c * iters * n, rec.traj_x0.data(), sizeof(double) * rec.traj_x0.size());
      if (g_tg0) std::memcpy(g_tg0 + (size_t)c * iters * n, rec.traj_g0.data(), sizeof(double) * rec.traj_g0.size());
      if (g_tp0) std::memcpy(g_tp0 + (size_t)c * iters * n, rec.traj_p0.data(), sizeof(double) * rec.traj_p0.size());
      if (g_tlp0) std::memcpy(g_tlp0 + (size_t)c * iters, rec.traj_lp0.data(), sizeof(double) * rec.traj_lp0.size());
      if (g_tmom) std::memcpy(g_tmom + (size_t)c * iters * L * n, rec.leapfrog_mom.data(), sizeof(double) * rec.leapfrog_mom.size());
      if (g_tgrad) std::memcpy(g_tgrad + (size_t)c * iters * L * n, rec.leapfrog_grad.data(), sizeof(double) * rec.leapfrog_grad.size());
      if (g_tlp) std::memcpy(g_tlp + (size_t)c * iters * L, rec.leapfrog_lp.data(), sizeof(double) * rec.leapfrog_lp.size());
    }
    pm += k->proposalsMade(); pa += k->proposalsAccepted();
    if (nspecs == 1 && specs[0].kind == 1) ge += ((HMCKernel*)inner)->gradEvals;
    else if (nspecs > 1) {
      MixtureKernel* m = (MixtureKernel*)inner;
      for (int i = 0; i < nspecs; ++i) if (specs[i].kind == 1) ge += ((HMCKernel*)m->kernels[i].get())->gradEvals;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    stats->grad_evals = ge; stats->ad_nodes = ad::G().nodes - nodes0; stats->rng_draws = r.ndraws;
    stats->props_made = pm; stats->props_accepted = pa;
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
  }
  return 0;
}

// log-density and gradient at a given unconstrained point x[n], as the HMC kernel sees them:
//   *lp_dual  = dual-trace logprob (with log-Jacobians), grad[n] = its gradient (tape AD)
//   *lp_float = float-trace logprob of the constrained values (no Jacobian)  (erp.t:361)
int qso_logp_grad(void* prog, const double* x, int n, double* lp_dual, double* grad, double* lp_float) {
  Program* program = (Program*)prog;
  Rng& r = rng(); r.seed = 0; r.chain = 0; r.at(ITER_INIT, 0);
  ad::recoverMemory();
  Trace<double> t; t.init(program, true);
  if ((int)t.countRealComps() != n) return 1;
  { std::vector<double> xv(x, x + n); size_t index = 0; for (auto& rc : t.choices) index += rc.setUnboundedRealComps(xv, index); }
  t.update(false);
  if (lp_float) *lp_float = t.logprob;
  HMCKernel k(1.0, 1, false);
  // the kernel's own gather applies inv(fwd(x)); overwrite with the exact x afterwards
  k.checkForChanges(&t);
  k.positions.assign(x, x + n);
  double lp = k.traceUpdate(k.positions, k.gradient);
  if (lp_dual) *lp_dual = lp;
  if (grad) std::memcpy(grad, k.gradient.data(), sizeof(double) * n);
  return 0;
}

// Expectation / Autocorrelation over a sample matrix [nsamps][dim] (loglikelihood = 0 as doMCMC stores it)
void qso_expectation(const double* values, int nsamps, int dim, int doVariance, double* mean, double* var) {
  std::vector<Sample> s(nsamps);
  for (int i = 0; i < nsamps; ++i) s[i].value.assign(values + (size_t)i * dim, values + (size_t)(i + 1) * dim);
  std::vector<double> m; double v;
  Expectation(s, doVariance != 0, m, v);
  std::memcpy(mean, m.data(), sizeof(double) * dim);
  if (var) *var = v;
}
void qso_autocorrelation(const double* values, int nsamps, int dim, double* ac) {
  std::vector<Sample> s(nsamps);
  for (int i = 0; i < nsamps; ++i) s[i].value.assign(values + (size_t)i * dim, values + (size_t)(i + 1) * dim);
  std::vector<double> m; double v;
  Expectation(s, true, m, v);
  std::vector<double> a;
  Autocorrelation(s, m, v, a);
  std::memcpy(ac, a.data(), sizeof(double) * nsamps);
}

}  // extern "C"
