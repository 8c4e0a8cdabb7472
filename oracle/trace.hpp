// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the fixed-structure part of qs/trace.t and qs/erp.t:
//   * RandomChoice(real): stored value + params + cached logprob, bounded <-> unbounded
//     accessors, rescore = logprob(fwd(x), params) + logJ(x) in dual traces only
//     (erp.t:286-311, 361, 369-429, 546-640)
//   * Bounds.None / Lower / LowerUpper transforms and log-Jacobians (erp.t:128-198)
//   * ERP bindings gaussian->None, gamma->Lower(0), beta->LowerUpper(0,1),
//     uniform->LowerUpper(lo,hi) (erpdefs.t:28-36, 46-57, 61-69, 92-100)
//   * RandExecTrace: rejection init by forward sampling, update() that re-executes the
//     program, addPrior/addLikelihood/checkCondition, temperature (trace.t:592-624,
//     711-773, 794-824), factor (trace.t:1099-1107), observe (erp.t:687-696)
// Everything that exists only for structure-changing programs (address stacks, hash-map
// rdbs, lpDiff, newlogprob/oldlogprob) is out of scope.
//
// Deviation, documented: the flattened order of real components is program order.  The
// reference orders by ERP *type* first, in Lua pairs() order, which is unspecified
// (trace.t:560-568); any fixed order is a valid reading.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>
#include "ad.hpp"
#include "distrib.hpp"
#include "philox.hpp"

namespace qso {

// ERP_NEGGAMMA: a user-defined ERP built with makeRandomChoice over Bounds.Upper (erp.t:158-173) -- the reference binds Upper
// to none of its own ERPs: value = -gamma(shape, scale) < 0, upper bound 0.
enum ErpKind { ERP_GAUSSIAN = 0, ERP_GAMMA = 1, ERP_BETA = 2, ERP_UNIFORM = 3, ERP_DIRICHLET = 4, ERP_NEGGAMMA = 5 };

template <class real> struct is_dual : std::false_type {};
template <> struct is_dual<ad::num> : std::true_type {};

// erp.t:91-96
template <class real> inline real logistic(real x) { return 1.0 / (1.0 + tmath::exp(-x)); }
template <class real> inline real invlogistic(real y) { return -tmath::log(1.0 / y - 1.0); }

static const double POS_INF = std::numeric_limits<double>::infinity();

// ---- Bounds.Lower (erp.t:141-156); lo is a plain double (erpdefs.t:64-68 returns 0.0)
template <class real> inline real lower_fwd(real x, double lo) { return tmath::fmax(tmath::exp(x) + lo, lo + 1e-15); }
template <class real> inline real lower_inv(real y, double lo) { real z = tmath::fmax(y, lo + 1e-15); return tmath::log(z - lo); }
template <class real> inline real lower_lpincr(real x, double) { return x; }
// ---- Bounds.Upper (erp.t:158-173)
template <class real> inline real upper_fwd(real x, double hi) { return tmath::fmin(hi - tmath::exp(x), hi - 1e-15); }
template <class real> inline real upper_inv(real y, double hi) { real z = tmath::fmin(y, hi - 1e-15); return tmath::log(hi - z); }
template <class real> inline real upper_lpincr(real x, double) { return x; }
// ---- Bounds.LowerUpper (erp.t:175-198); LO/HI are double (beta) or real (uniform)
template <class real, class BT> inline real lowerupper_fwd(real x, BT lo, BT hi) {
  real logit = logistic(x);
  if (x > D::NEG_INF && logit == 0.0) logit = real(1e-15);
  if (x < POS_INF && logit == 1.0) logit = real(1.0 - 1e-15);
  real y = lo + (hi - lo) * logit;
  return y;
}
template <class real, class BT> inline real lowerupper_inv(real y, BT lo, BT hi) {
  real z = tmath::fmax(tmath::fmin(y, hi - 1e-15), lo + 1e-15);
  real t = (z - lo) / (hi - lo);
  return invlogistic(t);
}
template <class real, class BT> inline real lowerupper_lpincr(real x, BT lo, BT hi) {
  return tmath::log(hi - lo) - x - 2.0 * tmath::log(1.0 + tmath::exp(-x));
}

// ---- Bounds.UnitSimplex (erp.t:200-254): stick breaking on a K-vector; component K-1 of x is unused
template <class real> inline std::vector<real> simplex_fwd(const std::vector<real>& x) {
  std::vector<real> y(x.size(), real(0.0));
  int Km1 = (int)x.size() - 1;
  real sticklen = real(1.0);
  for (int k = 0; k < Km1; ++k) {
    real z = logistic<real>(x[k] - tmath::log((double)(Km1 - k)));
    y[k] = sticklen * z;
    sticklen = sticklen - y[k];
  }
  y[Km1] = sticklen;
  return y;
}
template <class real> inline std::vector<real> simplex_inv(const std::vector<real>& y) {
  std::vector<real> x(y.size(), real(0.0));   // makeArrayLikeFrom: zeros (erp.t:97-103); x[K-1] stays 0
  int Km1 = (int)y.size() - 1;
  real sticklen = y[Km1];
  for (int k = Km1 - 1; k >= 0; --k) {
    sticklen = sticklen + y[k];
    real z = y[k] / sticklen;
    x[k] = invlogistic<real>(z) + tmath::log((double)(Km1 - k));
  }
  return x;
}
template <class real> inline real simplex_lpincr(const std::vector<real>& x) {
  real lp = real(0.0);
  int Km1 = (int)x.size() - 1;
  real sticklen = real(1.0);
  for (int k = 0; k < Km1; ++k) {
    real z = logistic<real>(x[k] - tmath::log((double)(Km1 - k)));
    lp = lp + tmath::log(sticklen);
    lp = lp + tmath::log(z);
    lp = lp + tmath::log(1.0 - z);
    sticklen = sticklen - (sticklen * z);
  }
  return lp;
}

template <class real> struct RandomChoice {   // RandomChoiceT, erp.t:286-311
  // vector-valued choice (dirichlet, erpdefs.t:179-212): value and params are K-vectors
  std::vector<real> vvalue, vparams;
  bool is_vector() const { return kind == ERP_DIRICHLET; }
  std::vector<real> getValueV() { if (applyingBounds) return simplex_fwd<real>(vvalue); return vvalue; }
  void setValueV(const std::vector<real>& v) { vvalue = applyingBounds ? simplex_inv<real>(v) : v; needsRescore = true; }
  std::vector<real> getUnboundedValueV() { if (applyingBounds) return vvalue; return simplex_inv<real>(vvalue); }
  void setUnboundedValueV(const std::vector<real>& v) { vvalue = applyingBounds ? v : simplex_fwd<real>(v); needsRescore = true; }
  // real-component accessors (erp.t:436-526): how the kernels read and write a choice
  size_t getUnboundedRealComps(std::vector<real>& comps) {
    if (is_vector()) { std::vector<real> v = getUnboundedValueV(); comps.insert(comps.end(), v.begin(), v.end()); return v.size(); }
    comps.push_back(getUnboundedValue()); return 1;
  }
  size_t setUnboundedRealComps(const std::vector<real>& comps, size_t start) {
    if (is_vector()) { std::vector<real> v(comps.begin() + start, comps.begin() + start + vvalue.size()); setUnboundedValueV(v); return v.size(); }
    setUnboundedValue(comps[start]); return 1;
  }
  size_t setStoredRealComps(const std::vector<real>& comps, size_t start) {
    if (is_vector()) { for (size_t i = 0; i < vvalue.size(); ++i) vvalue[i] = comps[start + i]; needsRescore = true; return vvalue.size(); }
    setStoredValue(comps[start]); return 1;
  }
  size_t ncomps() const { return is_vector() ? vvalue.size() : 1; }
  void init_vector(const std::vector<real>& params, unsigned lexid_, const std::vector<real>& initval) {   // erp.t:546-563
    kind = ERP_DIRICHLET; vparams = params;
    vvalue = applyingBounds ? simplex_inv<real>(initval) : initval;
    rescore();
    active = true; lexid = lexid_;
  }
  void update_vector(const std::vector<real>& params) {   // erp.t:576-610
    bool needs = needsRescore;
    if (is_dual<real>::value) { needs = true; vparams = params; }
    else for (size_t i = 0; i < params.size(); ++i) if (!(val(vparams[i]) == val(params[i]))) { needs = true; vparams[i] = params[i]; }
    if (needs) rescore();
    active = true;
  }

  int kind = ERP_GAUSSIAN;
  real value;        // stored value: constrained in float traces, unconstrained in dual traces
  real param0, param1;
  real logprob;
  bool active = false, needsRescore = false;
  unsigned lexid = 0;
  static constexpr bool applyingBounds = is_dual<real>::value;   // erp.t:361

  real fwd(real x) {
    switch (kind) {
      case ERP_GAMMA: return lower_fwd<real>(x, 0.0);
      case ERP_NEGGAMMA: return upper_fwd<real>(x, 0.0);
      case ERP_BETA: return lowerupper_fwd<real, double>(x, 0.0, 1.0);
      case ERP_UNIFORM: return lowerupper_fwd<real, real>(x, param0, param1);
      default: return x;
    }
  }
  real inv(real y) {
    switch (kind) {
      case ERP_GAMMA: return lower_inv<real>(y, 0.0);
      case ERP_NEGGAMMA: return upper_inv<real>(y, 0.0);
      case ERP_BETA: return lowerupper_inv<real, double>(y, 0.0, 1.0);
      case ERP_UNIFORM: return lowerupper_inv<real, real>(y, param0, param1);
      default: return y;
    }
  }
  real lpincr(real x) {
    switch (kind) {
      case ERP_GAMMA: return lower_lpincr<real>(x, 0.0);
      case ERP_NEGGAMMA: return upper_lpincr<real>(x, 0.0);
      case ERP_BETA: return lowerupper_lpincr<real, double>(x, 0.0, 1.0);
      case ERP_UNIFORM: return lowerupper_lpincr<real, real>(x, param0, param1);
      default: return real(0.0);
    }
  }
  real prior_logprob(real val) {
    switch (kind) {
      case ERP_GAMMA: return D::gamma_logprob<real>(val, param0, param1);
      case ERP_NEGGAMMA: return D::gamma_logprob<real>(-val, param0, param1);
      case ERP_BETA: return D::beta_logprob<real>(val, param0, param1);
      case ERP_UNIFORM: return D::uniform_logprob<real>(val, param0, param1);
      default: return D::gaussian_logprob<real>(val, param0, param1);
    }
  }
  double sample(double p0, double p1) {
    switch (kind) {
      case ERP_GAMMA: return D::gamma_sample(p0, p1);
      case ERP_NEGGAMMA: return -D::gamma_sample(p0, p1);
      case ERP_BETA: return D::beta_sample(p0, p1);
      case ERP_UNIFORM: return D::uniform_sample(p0, p1);
      default: return D::gaussian_sample(p0, p1);
    }
  }
  // RandomChoiceT:proposal (erp.t:643-657): the gaussian ERP drifts around the current value with its own sigma
  // (erpdefs.t:46-57); every other ERP resamples from its prior (makeDefaultProposal, erp.t:18-38).  Float traces only
  // (TraceMH never runs on a dual trace).  Returns {fwdlp, rvslp}.
  std::pair<double, double> proposal() {
    static_assert(true, "");
    double fwdlp, rvslp;
    if (is_vector()) {
      std::vector<real> cur = getValueV();
      std::vector<double> pv(vparams.size()), out(vparams.size()), cv(vparams.size());
      for (size_t i = 0; i < pv.size(); ++i) { pv[i] = val(vparams[i]); cv[i] = val(cur[i]); }
      D::dirichlet_sample(pv.data(), (int)pv.size(), out.data());
      fwdlp = D::dirichlet_logprob<double>(out.data(), pv.data(), (int)pv.size());
      rvslp = D::dirichlet_logprob<double>(cv.data(), pv.data(), (int)pv.size());
      setValueV(std::vector<real>(out.begin(), out.end()));
      return {fwdlp, rvslp};
    }
    double currval = val(getValue()), p0 = val(param0), p1 = val(param1), newval;
    if (kind == ERP_GAUSSIAN) {
      newval = D::gaussian_sample(currval, p1);
      fwdlp = D::gaussian_logprob<double>(newval, currval, p1);
      rvslp = D::gaussian_logprob<double>(currval, newval, p1);
    } else {
      newval = sample(p0, p1);
      RandomChoice<double> tmp; tmp.kind = kind; tmp.param0 = p0; tmp.param1 = p1;
      fwdlp = tmp.prior_logprob(newval);
      rvslp = tmp.prior_logprob(currval);
    }
    setValue(real(newval));
    return {fwdlp, rvslp};
  }
  // erp.t:369-429
  real getValue() { if (applyingBounds) return fwd(value); return value; }
  void setValue(real newval) { if (applyingBounds) newval = inv(newval); value = newval; needsRescore = true; }
  real getUnboundedValue() { if (applyingBounds) return value; return inv(value); }
  void setUnboundedValue(real newval) { if (!applyingBounds) newval = fwd(newval); value = newval; needsRescore = true; }
  real getStoredValue() { return value; }
  void setStoredValue(real newval) { value = newval; needsRescore = true; }
  // erp.t:623-640
  void rescore() {
    if (is_vector()) {
      std::vector<real> v = getValueV();
      logprob = D::dirichlet_logprob<real>(v.data(), vparams.data(), (int)v.size());
      if (applyingBounds) logprob = logprob + simplex_lpincr<real>(vvalue);
      needsRescore = false;
      return;
    }
    real val = getValue();
    logprob = prior_logprob(val);
    if (applyingBounds) {
      if (kind != ERP_GAUSSIAN) logprob = logprob + lpincr(getStoredValue());
      else logprob = logprob + 0.0;   // Bounds.None.logprobIncrement is `0.0 (erp.t:135): still one AD add
    }
    needsRescore = false;
  }
  // erp.t:546-563 (constructor with an initial value)
  void init(int kind_, real p0, real p1, unsigned lexid_, real initval) {
    kind = kind_; param0 = p0; param1 = p1;
    if (applyingBounds) value = inv(initval); else value = initval;
    rescore();
    active = true; lexid = lexid_;
  }
  // erp.t:576-610
  void update(real p0, real p1) {
    bool needs = needsRescore;
    if (is_dual<real>::value) {
      needs = true; param0 = p0; param1 = p1;
    } else {
      if (!(val(param0) == val(p0))) { needs = true; param0 = p0; }
      if (!(val(param1) == val(p1))) { needs = true; param1 = p1; }
    }
    if (needs) rescore();
    active = true;
  }
};

template <class real> struct Trace;

struct Program {
  int retdim = 1;
  virtual ~Program() {}
  virtual void run(Trace<double>& t) = 0;
  virtual void run(Trace<ad::num>& t) = 0;
};
#define QSO_PROGRAM_RUNNERS                                   \
  void run(Trace<double>& t) override { body<double>(t); }    \
  void run(Trace<ad::num>& t) override { body<ad::num>(t); }

template <class real> struct Trace {   // RandExecTraceT, trace.t:523-540
  Program* program = nullptr;
  std::vector<RandomChoice<real>> choices;   // the flat non-structural choice list (trace.t:314-320)
  size_t counter = 0;
  bool creating = false;
  // lexical id of the ERP call site being executed (erp.t:674-679 passes a per-call-site id to the constructor;
  // getLexicalID erp.t:318-321).  A program sets it before each ERP call: calls inside one loop share one id.
  unsigned site = 0;
  real logprob, loglikelihood;
  double logprob_val = 0.0, loglik_val = 0.0; // val(logprob), val(loglikelihood): kept because dual nodes die with the tape
  double temperature = 1.0;
  bool conditionsSatisfied = false, hasReturnValue = false;
  bool canStructureChange = true, doEvalLikelihoodsAndConditions = true;
  int64_t numUpdates = 0;
  std::vector<real> returnValue;

  // trace.t:592-624
  void init(Program* p, bool doRejectionInit) {
    program = p;
    logprob = real(0.0); loglikelihood = real(0.0); logprob_val = 0.0; loglik_val = 0.0;
    temperature = 1.0; conditionsSatisfied = false; hasReturnValue = false;
    canStructureChange = true; doEvalLikelihoodsAndConditions = true; numUpdates = 0;
    choices.clear();
    if (doRejectionInit) {
      // every rejection-init attempt reads fresh sub-streams: attempt a samples under iteration code ITER_INIT - a
      // (the reference's rand() stream simply continues, trace.t:612-623)
      uint32_t attempt = 0;
      while (!conditionsSatisfied) {
        choices.clear();
        rng().iter = ITER_INIT - attempt; ++attempt;
        update(true);
      }
    }
  }
  // trace.t:711-773
  void update(bool canStructureChange_, bool doEval = true) {
    canStructureChange = canStructureChange_;
    doEvalLikelihoodsAndConditions = doEval;
    logprob = real(0.0); loglikelihood = real(0.0);
    conditionsSatisfied = true;
    counter = 0;
    creating = canStructureChange_ && choices.empty();
    returnValue.clear();
    program->run(*this);
    hasReturnValue = true;
    logprob = logprob / temperature;
    loglikelihood = loglikelihood / temperature;
    logprob_val = val(logprob); loglik_val = val(loglikelihood);
    if (canStructureChange_ && choices.empty()) {
      std::fprintf(stderr, "Execution of probabilistic program resulted in zero random choices\n");
      std::abort();
    }
    numUpdates = numUpdates + 1;
  }
  void setTemperature(double t) { temperature = t; }
  // trace.t:794-808
  void checkCondition(bool c) { conditionsSatisfied = conditionsSatisfied && c; }
  void addPrior(real x) { logprob = logprob + x; }
  void addLikelihood(real x) { logprob = logprob + x; loglikelihood = loglikelihood + x; }
  // trace.t:817-824
  template <class R2> void copyProbabilities(Trace<R2>& other) {
    logprob = real(other.logprob_val);
    logprob_val = other.logprob_val;
    loglikelihood = real(other.loglik_val);
    loglik_val = other.loglik_val;
  }
  // lookupRandomChoiceValue, trace.t:948-994 (non-structural branch, trace.t:314-320)
  real erp(int kind, real p0, real p1) {
    if (creating) {
      // constructor 2 (erp.t:567-572): draw a sample, then constructor 1
      rng().slot = (uint32_t)choices.size(); rng().k = 0;
      RandomChoice<real> rc;
      rc.kind = kind;
      double sampled = rc.sample(val(p0), val(p1));
      rc.init(kind, p0, p1, site, real(sampled));
      choices.push_back(rc);
      RandomChoice<real>& r = choices.back();
      addPrior(r.logprob);
      return r.getValue();
    }
    RandomChoice<real>& rc = choices[counter];
    counter = counter + 1;
    rc.update(p0, p1);
    addPrior(rc.logprob);
    return rc.getValue();
  }
  // vector-valued ERP site (dirichlet): same lookup, K real components
  std::vector<real> dirichlet(const std::vector<real>& params) {
    if (creating) {
      rng().slot = (uint32_t)choices.size(); rng().k = 0;
      std::vector<double> pv(params.size()), out(params.size());
      for (size_t i = 0; i < params.size(); ++i) pv[i] = val(params[i]);
      D::dirichlet_sample(pv.data(), (int)pv.size(), out.data());
      std::vector<real> iv(out.begin(), out.end());
      RandomChoice<real> rc;
      rc.init_vector(params, site, iv);
      choices.push_back(rc);
      RandomChoice<real>& r = choices.back();
      addPrior(r.logprob);
      return r.getValueV();
    }
    RandomChoice<real>& rc = choices[counter];
    counter = counter + 1;
    rc.update_vector(params);
    addPrior(rc.logprob);
    return rc.getValueV();
  }
  real gaussian(real mu, real sigma) { return erp(ERP_GAUSSIAN, mu, sigma); }
  real gamma(real shape, real scale) { return erp(ERP_GAMMA, shape, scale); }
  real beta(real a, real b) { return erp(ERP_BETA, a, b); }
  real uniform(real lo, real hi) { return erp(ERP_UNIFORM, lo, hi); }
  // trace.t:1099-1107
  void factor(real x) { if (doEvalLikelihoodsAndConditions) addLikelihood(x); }
  void condition(bool c) { if (doEvalLikelihoodsAndConditions) checkCondition(c); }
  // erp.observe (erp.t:687-696)
  void gaussian_observe(real value, real mu, real sigma) { if (doEvalLikelihoodsAndConditions) addLikelihood(D::gaussian_logprob<real>(value, mu, sigma)); }
  void flip_observe(bool value, real p) { if (doEvalLikelihoodsAndConditions) addLikelihood(D::bernoulli_logprob<real>(value, p)); }

  // countChoices/getChoice({isStructural=false}) (trace.t:830-905): every choice here is non-structural and scalar
  size_t countChoices() const { return choices.size(); }
  size_t countRealComps() const { size_t n = 0; for (auto& rc : choices) n += rc.ncomps(); return n; }
};

// DualTraceType.copyFromRealType(qs.float) (trace.t:167-173; erp.t:531-540)
inline void copyFromRealType(Trace<ad::num>& self, Trace<double>& other) {
  self.program = other.program;
  self.choices.clear();
  for (size_t i = 0; i < other.choices.size(); ++i) {
    RandomChoice<double>& o = other.choices[i];
    RandomChoice<ad::num> rc;
    rc.kind = o.kind; rc.param0 = ad::num(o.param0); rc.param1 = ad::num(o.param1);
    rc.logprob = ad::num(o.logprob); rc.active = o.active; rc.lexid = o.lexid;
    if (o.is_vector()) {
      for (auto& p : o.vparams) rc.vparams.push_back(ad::num(p));
      std::vector<double> ov = o.getValueV();
      std::vector<ad::num> mv;
      for (double v : ov) mv.push_back(ad::num(v));
      rc.setValueV(mv);
      self.choices.push_back(rc);
      continue;
    }
    double otherval = o.getValue();
    rc.setValue(ad::num(otherval));
    self.choices.push_back(rc);
  }
  self.logprob = ad::num(other.logprob); self.loglikelihood = ad::num(other.loglikelihood);
  self.logprob_val = other.logprob_val; self.loglik_val = other.loglik_val;
  self.temperature = other.temperature;
  self.conditionsSatisfied = other.conditionsSatisfied;
  self.hasReturnValue = false;
  self.numUpdates = other.numUpdates;
}

}  // namespace qso
