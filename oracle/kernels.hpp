// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the reference's MCMC driver and transition kernels:
//   doMCMC / KernelPropStats / MixtureKernel   qs/mcmc.t:36-105, 116-125, 199-291
//   HMCKernel + DualAverage                    qs/hmc.t:14-48, 57-402
//   DriftKernel + RunningVar                   qs/drift.t:15-50, 63-300
//   HARMKernel                                 qs/harm.t:16-137
// Quirks are kept on purpose (Q3-Q12).  The only additions are
// (i) Rng::at(iter, slot) calls in front of each sampler call, which select the Philox
// sub-stream the injected random() reads (see philox.hpp), and (ii) an optional Recorder
// that copies out per-leapfrog positions for the parity tests.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <utility>
#include <vector>
#include "programs.hpp"

namespace qso {

struct Sample {                 // Sample(T), qs/infer.t:12-18
  std::vector<double> value;
  double logprob = 0.0, loglikelihood = 0.0;
};

// Optional per-step record for parity tests (not in the reference).
struct Recorder {
  std::vector<double> leapfrog_pos;   // every position vector after each HMCKernel:step of a `next`
  std::vector<double> leapfrog_lp;    // lp after each step
  std::vector<double> proposal;       // drift/HARM proposed unconstrained vector, per iteration
  std::vector<double> dH;             // HMC: H_new - H ; drift/HARM: acceptThresh
  std::vector<int> accepted;          // per iteration
  std::vector<double> stepsize;       // HMC stepSize / HARM scale used for the proposal of this iteration
  std::vector<int> kernel_index;      // mixture: which sub-kernel ran
  std::vector<double> state;          // unconstrained state after each iteration
  // teacher-forcing record of HMCKernel:next (tests replay every leapfrog step of a trajectory on the device from these):
  bool want_traj = false;
  std::vector<double> traj_x0, traj_g0, traj_p0;   // per HMC iteration: positions, cached gradient, sampled momenta  [n] each
  std::vector<double> traj_lp0;                    // per HMC iteration: currTrace.logprob entering the Hamiltonian
  std::vector<double> leapfrog_mom, leapfrog_grad; // after every HMCKernel:step: momenta, gradient at the new position [n] each
  bool in_next = false;
};
inline Recorder*& recorder() { static thread_local Recorder* r = nullptr; return r; }

struct Kernel {
  uint64_t propsMade = 0, propsAccepted = 0;   // KernelPropStats, qs/mcmc.t:116-125
  virtual ~Kernel() {}
  virtual void next(Trace<double>* currTrace, uint32_t iter, uint32_t numiters) = 0;
  virtual uint64_t proposalsMade() { return propsMade; }
  virtual uint64_t proposalsAccepted() { return propsAccepted; }
  virtual void unconstrained_state(Trace<double>* t, std::vector<double>& out) {
    out.clear();
    for (auto& rc : t->choices) rc.getUnboundedRealComps(out);
  }
};

// ------------------------------------------------------------------ qs/hmc.t:14-48
struct DualAverage {
  double gbar = 0, xbar = 0, x0 = 0, lastx = 0;
  unsigned k = 0;
  double gamma = 0;
  void reinit(double x0_, double gamma_) { k = 0; x0 = x0_; lastx = x0_; gbar = 0.0; xbar = 0.0; gamma = gamma_; }
  double update(double g) {
    k = k + 1;
    double avgeta = 1.0 / (k + 10);
    double xbar_avgeta = std::pow((double)k, -0.75);
    double muk = 0.5 * std::sqrt((double)k) / gamma;
    gbar = avgeta * g + (1 - avgeta) * gbar;
    lastx = x0 - muk * gbar;
    xbar = xbar_avgeta * lastx + (1 - xbar_avgeta) * xbar;
    return lastx;
  }
};

// ------------------------------------------------------------------ qs/hmc.t:57-402
struct HMCKernel : Kernel {
  double stepSize;
  uint64_t numSteps;
  bool adapting;
  DualAverage adapter;
  double targetAcceptRate;
  Trace<ad::num> dualTrace;
  Trace<double>* lastTraceSeen = nullptr;
  int64_t lastNumUpdatesSeen = -1;
  std::vector<double> positions, positions_scratch, gradient, gradient_scratch, momenta, invMasses;
  std::vector<ad::num> positions_dual_scratch;
  uint32_t cur_iter = 0;      // which Philox sub-streams the momenta of this call read
  uint32_t search_probe = 0;
  uint64_t gradEvals = 0;

  HMCKernel(double stepSize_ = 1.0, uint64_t numSteps_ = 1, bool doStepSizeAdapt = true) {   // hmc.t:110-128
    stepSize = stepSize_; numSteps = numSteps_; adapting = doStepSizeAdapt;
    targetAcceptRate = (numSteps_ == 1) ? 0.574 : 0.65;
    adapter.reinit(0.0, 0.0);
  }

  void next(Trace<double>* currTrace, uint32_t iter, uint32_t) override {   // hmc.t:134-186
    propsMade = propsMade + 1;
    cur_iter = iter;
    checkForChanges(currTrace);
    cur_iter = iter;
    sampleMomenta();
    double H = hamiltonian(currTrace->logprob);
    double newlp = 0.0;
    positions_scratch = positions;
    gradient_scratch = gradient;
    Recorder* rec = recorder();
    if (rec) { rec->in_next = true; rec->stepsize.push_back(stepSize); }
    if (rec && rec->want_traj) {
      rec->traj_x0.insert(rec->traj_x0.end(), positions.begin(), positions.end());
      rec->traj_g0.insert(rec->traj_g0.end(), gradient.begin(), gradient.end());
      rec->traj_p0.insert(rec->traj_p0.end(), momenta.begin(), momenta.end());
      rec->traj_lp0.push_back(currTrace->logprob);
    }
    for (uint64_t i = 0; i < numSteps; ++i) newlp = step(positions_scratch, gradient_scratch);
    if (rec) rec->in_next = false;
    double H_new = hamiltonian(newlp);
    double dH = H_new - H;
    if (adapting) adaptStepSize(dH);
    rng().at(iter, (uint32_t)positions.size());
    bool accept = dualTrace.conditionsSatisfied && std::log(random_()) < dH;
    if (accept) {
      propsAccepted = propsAccepted + 1;
      std::swap(positions, positions_scratch);
      std::swap(gradient, gradient_scratch);
      { size_t index = 0; for (auto& rc : currTrace->choices) index += rc.setUnboundedRealComps(positions, index); }   // hmc.t:170-176
      currTrace->update(false, false);
      currTrace->copyProbabilities(dualTrace);
    }
    if (rec) { rec->dH.push_back(dH); rec->accepted.push_back(accept ? 1 : 0); }
    lastTraceSeen = currTrace;
    lastNumUpdatesSeen = currTrace->numUpdates;
  }

  void checkForChanges(Trace<double>* currTrace) {   // hmc.t:188-287
    if (currTrace != lastTraceSeen || currTrace->numUpdates != lastNumUpdatesSeen) {
      size_t oldn = positions.size();
      positions.clear();
      size_t numvars = currTrace->countChoices();
      if (numvars == 0) { std::printf("HMC: found 0 non-structural random choices.\n"); std::abort(); }
      for (size_t i = 0; i < numvars; ++i) currTrace->choices[i].getUnboundedRealComps(positions);
      size_t newn = positions.size();
      if (newn != oldn) {
        copyFromRealType(dualTrace, *currTrace);
        invMasses.assign(newn, 1.0);
      }
      traceUpdate(positions, gradient);
      if (newn != oldn && adapting && lastTraceSeen == nullptr) {
        searchForInitialStepSize();
        adapter.reinit(stepSize, 0.05);   // adaptRate, hmc.t:64, 261
      }
    }
    dualTrace.setTemperature(currTrace->temperature);
  }

  void sampleMomenta() {   // hmc.t:289-295
    size_t n = positions.size();
    momenta.clear(); momenta.reserve(n);
    for (size_t i = 0; i < n; ++i) {
      rng().at(cur_iter, (uint32_t)i);
      momenta.push_back(D::gaussian_sample(0.0, 1.0) * invMasses[i]);
    }
  }
  double hamiltonian(double logprob) {   // hmc.t:297-305
    double kinetic = 0.0;
    for (size_t i = 0; i < momenta.size(); ++i) { double m = momenta[i]; kinetic = kinetic + m * m * invMasses[i]; }
    kinetic = -0.5 * kinetic;
    return kinetic + logprob;
  }
  double step(std::vector<double>& pos, std::vector<double>& grad) {   // hmc.t:307-323
    for (size_t i = 0; i < momenta.size(); ++i) momenta[i] = momenta[i] + 0.5 * stepSize * grad[i];
    for (size_t i = 0; i < pos.size(); ++i) pos[i] = pos[i] + stepSize * momenta[i] * invMasses[i];
    double lp = traceUpdate(pos, grad);
    for (size_t i = 0; i < momenta.size(); ++i) momenta[i] = momenta[i] + 0.5 * stepSize * grad[i];
    Recorder* rec = recorder();
    if (rec && rec->in_next) {
      rec->leapfrog_pos.insert(rec->leapfrog_pos.end(), pos.begin(), pos.end()); rec->leapfrog_lp.push_back(lp);
      if (rec->want_traj) {
        rec->leapfrog_mom.insert(rec->leapfrog_mom.end(), momenta.begin(), momenta.end());
        rec->leapfrog_grad.insert(rec->leapfrog_grad.end(), grad.begin(), grad.end());
      }
    }
    return lp;
  }
  double traceUpdate(std::vector<double>& pos, std::vector<double>& grad) {   // hmc.t:325-342
    positions_dual_scratch.clear(); positions_dual_scratch.reserve(pos.size());
    for (size_t i = 0; i < pos.size(); ++i) positions_dual_scratch.push_back(ad::num(pos[i]));
    { size_t index = 0; for (auto& rc : dualTrace.choices) index += rc.setStoredRealComps(positions_dual_scratch, index); }   // hmc.t:329-334
    dualTrace.update(false);
    ad::num duallp = dualTrace.logprob;
    double lp = duallp.val();
    ad::grad(duallp, positions_dual_scratch, grad);
    ++gradEvals;
    return lp;
  }
  void searchForInitialStepSize() {   // hmc.t:345-391
    positions_scratch = positions; gradient_scratch = gradient;
    search_probe = 0;
    cur_iter = ITER_SEARCH + search_probe++;
    sampleMomenta();
    double lastlp = dualTrace.logprob_val;
    double lp = step(positions_scratch, gradient_scratch);
    double H = lp - lastlp;
    int direction = -1;
    if (H > std::log(0.5)) direction = 1;
    while (true) {
      positions_scratch = positions; gradient_scratch = gradient;
      cur_iter = ITER_SEARCH + search_probe++;
      sampleMomenta();
      double lastlp2 = dualTrace.logprob_val;   // Q5: the *previous probe's* lp
      double lp2 = step(positions_scratch, gradient_scratch);
      double H2 = lp2 - lastlp2;
      if (!(lp2 == lp2)) { std::printf("HMC step size search failed because logprob became NaN\n"); std::abort(); }
      if (direction == 1 && H2 <= std::log(0.5)) break;
      else if (direction == -1 && H2 >= std::log(0.5)) break;
      else if (direction == 1) stepSize = stepSize * 2.0;
      else stepSize = stepSize * 0.5;
      if (stepSize > 1e300) { std::printf("HMC step size search diverged to infinity (Is your probability distribution flat?)\n"); std::abort(); }
      if (stepSize == 0) { std::printf("HMC step size search collapsed to zero (Is your probability distribution discontinuous?)\n"); std::abort(); }
    }
  }
  void unconstrained_state(Trace<double>*, std::vector<double>& out) override { out = positions; }
  void adaptStepSize(double dH) {   // hmc.t:394-402
    double EdH = std::exp(dH);
    if (EdH > 1.0) EdH = 1.0;
    if (!(EdH == EdH)) EdH = 0.0;
    double adaptGrad = targetAcceptRate - EdH;
    stepSize = std::exp(adapter.update(adaptGrad));
  }
};

// ------------------------------------------------------------------ qs/drift.t:15-50
struct RunningVar {
  double mean = 0, variance = 0;
  unsigned n = 0;
  void update(double x) {
    n = n + 1;
    if (n == 1) { mean = x; variance = 0.0; }
    else {
      double newmean = mean + (x - mean) / n;
      variance = variance + (x - mean) * (x - newmean);
      mean = newmean;
    }
  }
  unsigned getN() { return n; }
  double getVariance() { if (n <= 1) return 0.0; return variance / (n - 1); }
  double getStdDev() { return std::sqrt(getVariance()); }
};

inline double lerp_(double a, double b, double t) { return (1.0 - t) * a + t * b; }

// ------------------------------------------------------------------ qs/drift.t:63-300
struct DriftKernel : Kernel {
  struct ScaleAdapter { double scale; RunningVar varEst; };
  double initScale;
  std::vector<std::vector<ScaleAdapter>> scaleGroups;
  bool adapting, lexicalScaleSharing;
  std::vector<unsigned> scaleGroupIndexForComp, indexWithinScaleGroupForComp;   // drift.t:101-105
  ScaleAdapter* getScale(size_t i) {   // drift.t:107-114
    if (lexicalScaleSharing) return &scaleGroups[scaleGroupIndexForComp[i]][indexWithinScaleGroupForComp[i]];
    return &scaleGroups[0][i];
  }
  std::vector<double> realcomps, realcomps_scratch;
  Trace<double>* lastTraceSeen = nullptr;
  int64_t lastNumUpdatesSeen = -1;

  DriftKernel(double scale = 1.0, bool doScaleAdapt = true, bool lexical = false) { initScale = scale; adapting = doScaleAdapt; lexicalScaleSharing = lexical; }

  void next(Trace<double>* currTrace, uint32_t iter, uint32_t) override {   // drift.t:133-172
    propsMade = propsMade + 1;
    checkForChanges(currTrace);
    size_t n = realcomps.size();
    for (size_t i = 0; i < n; ++i) {
      rng().at(iter, (uint32_t)i);
      realcomps_scratch[i] = D::gaussian_sample(realcomps[i], getScale(i)->scale);
    }
    Trace<double> nextTrace = *currTrace;   // TraceType.salloc():copy(currTrace)
    { size_t index = 0; for (auto& rc : nextTrace.choices) index += rc.setUnboundedRealComps(realcomps_scratch, index); }
    nextTrace.update(false);
    double acceptThresh = nextTrace.logprob - currTrace->logprob;
    rng().at(iter, (uint32_t)n);
    bool accept = nextTrace.conditionsSatisfied && std::log(random_()) < acceptThresh;
    Recorder* rec = recorder();
    if (rec) { rec->proposal.insert(rec->proposal.end(), realcomps_scratch.begin(), realcomps_scratch.end()); rec->dH.push_back(acceptThresh); rec->accepted.push_back(accept ? 1 : 0); }
    if (accept) {
      propsAccepted = propsAccepted + 1;
      std::swap(*currTrace, nextTrace);
      std::swap(realcomps, realcomps_scratch);
    }
    if (adapting) adapt();
    lastTraceSeen = currTrace;
    lastNumUpdatesSeen = currTrace->numUpdates;
  }
  void checkForChanges(Trace<double>* currTrace) {   // drift.t:174-271
    if (currTrace != lastTraceSeen || currTrace->numUpdates != lastNumUpdatesSeen) {
      realcomps.clear();
      size_t numvars = currTrace->countChoices();
      if (numvars == 0) { std::printf("DriftKernel: found 0 non-structural random choices\n"); std::abort(); }
      std::vector<unsigned> nCompsPerScaleGroup;
      std::vector<std::pair<unsigned, unsigned>> lexidToGroupIndex;   // HashMap(uint,uint) of drift.t:181
      if (lexicalScaleSharing) { scaleGroupIndexForComp.clear(); indexWithinScaleGroupForComp.clear(); }
      else if (scaleGroups.empty()) scaleGroups.emplace_back();
      for (size_t i = 0; i < numvars; ++i) {
        RandomChoice<double>& rc = currTrace->choices[i];
        size_t ncomps = rc.getUnboundedRealComps(realcomps);
        if (lexicalScaleSharing) {   // drift.t:208-235
          unsigned lexid = rc.lexid, g = 0;
          bool found = false;
          for (auto& kv : lexidToGroupIndex) if (kv.first == lexid) { g = kv.second; found = true; }
          if (!found) {
            g = (unsigned)nCompsPerScaleGroup.size();
            lexidToGroupIndex.emplace_back(lexid, g);
            nCompsPerScaleGroup.push_back(0);
            while (scaleGroups.size() < g + 1) scaleGroups.emplace_back();
          }
          for (size_t j = 0; j < ncomps; ++j) { scaleGroupIndexForComp.push_back(g); indexWithinScaleGroupForComp.push_back((unsigned)j); }
          if (ncomps > nCompsPerScaleGroup[g]) nCompsPerScaleGroup[g] = (unsigned)ncomps;
        }
      }
      size_t n = realcomps.size();
      realcomps_scratch.assign(n, 0.0);
      if (!lexicalScaleSharing) nCompsPerScaleGroup.push_back((unsigned)n);
      for (size_t i = 0; i < nCompsPerScaleGroup.size(); ++i)   // drift.t:262-269
        for (size_t j = scaleGroups[i].size(); j < nCompsPerScaleGroup[i]; ++j) { ScaleAdapter s; s.scale = initScale; scaleGroups[i].push_back(s); }
    }
  }
  void unconstrained_state(Trace<double>*, std::vector<double>& out) override { out = realcomps; }
  void adapt() {   // drift.t:274-300
    double ratio = (double)propsAccepted / propsMade;
    for (size_t i = 0; i < realcomps.size(); ++i) {
      if (ratio >= 0.05 && propsAccepted > 5) getScale(i)->varEst.update(realcomps[i]);
    }
    for (auto& scaleGroup : scaleGroups)
      for (auto& scale : scaleGroup) {
        if (scale.varEst.getN() > 100) scale.scale = scale.varEst.getStdDev();
        if (ratio < 0.05 && propsAccepted > 5) scale.scale = scale.scale * 0.99;
      }
  }
};

// ------------------------------------------------------------------ qs/harm.t:16-137
struct HARMKernel : Kernel {
  double scale;
  bool adapting;
  std::vector<double> realcomps, realcomps_scratch, direction;
  Trace<double>* lastTraceSeen = nullptr;
  int64_t lastNumUpdatesSeen = -1;

  HARMKernel(double scale_ = 1.0, bool doScaleAdapt = true) { scale = scale_; adapting = doScaleAdapt; }

  void next(Trace<double>* currTrace, uint32_t iter, uint32_t) override {   // harm.t:60-113
    propsMade = propsMade + 1;
    checkForChanges(currTrace);
    size_t n = realcomps.size();
    direction.clear(); direction.reserve(n);
    double norm = 0.0;
    for (size_t i = 0; i < n; ++i) {
      rng().at(iter, (uint32_t)i);
      double x = D::gaussian_sample(0.0, 1.0);
      norm = norm + x * x;
      direction.push_back(x);
    }
    norm = std::sqrt(norm);
    rng().at(iter, (uint32_t)n);
    Recorder* rec = recorder();
    if (rec) rec->stepsize.push_back(this->scale);
    double scale = random_() * this->scale;
    for (size_t i = 0; i < n; ++i) {
      direction[i] = direction[i] / norm;
      realcomps_scratch[i] = realcomps[i] + (scale * direction[i]);
    }
    Trace<double> nextTrace = *currTrace;
    { size_t index = 0; for (auto& rc : nextTrace.choices) index += rc.setUnboundedRealComps(realcomps_scratch, index); }
    nextTrace.update(false);
    double targetAcceptRatio = 0.234;
    if (n < 5) targetAcceptRatio = lerp_(0.44, 0.234, (n - 1) / 4.0);
    double acceptThresh = nextTrace.logprob - currTrace->logprob;
    rng().at(iter, (uint32_t)n + 1);
    bool accept = nextTrace.conditionsSatisfied && std::log(random_()) < acceptThresh;
    if (rec) { rec->proposal.insert(rec->proposal.end(), realcomps_scratch.begin(), realcomps_scratch.end()); rec->dH.push_back(acceptThresh); rec->accepted.push_back(accept ? 1 : 0); }
    if (accept) {
      propsAccepted = propsAccepted + 1;
      std::swap(*currTrace, nextTrace);
      std::swap(realcomps, realcomps_scratch);
      if (adapting) {
        double t = targetAcceptRatio;
        this->scale = this->scale + (this->scale / (t * (1.0 - t))) * (1.0 - t) / (iter + 1);
      }
    } else {
      if (adapting) {
        double t = targetAcceptRatio;
        this->scale = std::fabs(this->scale - (this->scale / (t * (1.0 - t))) * t / (iter + 1));
      }
    }
    lastTraceSeen = currTrace;
    lastNumUpdatesSeen = currTrace->numUpdates;
  }
  void unconstrained_state(Trace<double>*, std::vector<double>& out) override { out = realcomps; }
  void checkForChanges(Trace<double>* currTrace) {   // harm.t:115-137
    if (currTrace != lastTraceSeen || currTrace->numUpdates != lastNumUpdatesSeen) {
      realcomps.clear();
      size_t numvars = currTrace->countChoices();
      if (numvars == 0) { std::printf("HARMKernel: found 0 non-structural random choices\n"); std::abort(); }
      for (size_t i = 0; i < numvars; ++i) currTrace->choices[i].getUnboundedRealComps(realcomps);
      realcomps_scratch.assign(realcomps.size(), 0.0);
    }
  }
};

// ------------------------------------------------------------------ qs/mcmc.t:134-193, {doStruct=false}: every choice of a
// fixed-structure program is non-structural, so the filter keeps them all and no dimension-jumping terms arise
struct TraceMHKernel : Kernel {
  void next(Trace<double>* currTrace, uint32_t iter, uint32_t) override {
    propsMade = propsMade + 1;
    Trace<double> nextTrace = *currTrace;
    size_t numchoices = nextTrace.countChoices();
    if (numchoices == 0) { std::printf("TraceMHKernel: found 0 random choices satisfying doStruct=false and doNonstruct=true\n"); std::abort(); }
    rng().at(iter, SLOT_MH_INDEX);
    size_t randindex = (size_t)(random_() * numchoices);
    RandomChoice<double>& rc = nextTrace.choices[randindex];
    rng().at(iter, SLOT_MH_PROPOSAL);
    std::pair<double, double> pl = rc.proposal();
    double fwdPropLP = pl.first, rvsPropLP = pl.second;
    nextTrace.update(false);
    double acceptThresh = (nextTrace.logprob - currTrace->logprob) + rvsPropLP - fwdPropLP;
    rng().at(iter, SLOT_MH_ACCEPT);
    bool accept = nextTrace.conditionsSatisfied && std::log(random_()) < acceptThresh;
    Recorder* rec = recorder();
    if (rec) { rec->dH.push_back(acceptThresh); rec->accepted.push_back(accept ? 1 : 0); rec->stepsize.push_back((double)randindex); }
    if (accept) {
      propsAccepted = propsAccepted + 1;
      std::swap(*currTrace, nextTrace);
    }
  }
};

// ------------------------------------------------------------------ qs/mcmc.t:199-291
struct MixtureKernel : Kernel {
  std::vector<std::unique_ptr<Kernel>> kernels;
  std::vector<double> weights;
  void next(Trace<double>* currTrace, uint32_t iter, uint32_t numiters) override {   // mcmc.t:275-287
    rng().at(iter, SLOT_MIXTURE);
    int randindex = D::categorical_sample(weights.data(), (int)weights.size());
    Recorder* rec = recorder();
    if (rec) rec->kernel_index.push_back(randindex);
    last = randindex;
    kernels[randindex]->next(currTrace, iter, numiters);
  }
  int last = 0;
  void unconstrained_state(Trace<double>* t, std::vector<double>& out) override { kernels[last]->unconstrained_state(t, out); }
  uint64_t proposalsMade() override { uint64_t t = 0; for (auto& k : kernels) t += k->proposalsMade(); return t; }
  uint64_t proposalsAccepted() override { uint64_t t = 0; for (auto& k : kernels) t += k->proposalsAccepted(); return t; }
};

// ------------------------------------------------------------------ qs/mcmc.t:297-319 (next-row f1)
// annealSched(iter, numiters) is tabulated by the caller: temps[i] = annealSched(i, numiters).
struct AnnealingKernel : Kernel {
  std::unique_ptr<Kernel> k;
  std::vector<double> temps;
  double annealSched(uint32_t iter, uint32_t) const { return temps[iter < temps.size() ? iter : temps.size() - 1]; }
  void next(Trace<double>* currTrace, uint32_t iter, uint32_t numiters) override {
    currTrace->setTemperature(annealSched(iter, numiters));
    k->next(currTrace, iter, numiters);
  }
  uint64_t proposalsMade() override { return k->proposalsMade(); }
  uint64_t proposalsAccepted() override { return k->proposalsAccepted(); }
  void unconstrained_state(Trace<double>* t, std::vector<double>& out) override { k->unconstrained_state(t, out); }
};

// ------------------------------------------------------------------ doMCMC, qs/mcmc.t:50-90
inline void doMCMC(Program* program, Kernel* k, std::vector<Sample>& samples, uint64_t nsamps, uint64_t burnin, uint64_t lag) {
  uint64_t iters = burnin + (nsamps * lag);
  Trace<double> currTrace;
  rng().at(ITER_INIT, 0);
  currTrace.init(program, true);
  Recorder* rec = recorder();
  for (uint64_t i = 0; i < iters; ++i) {
    k->next(&currTrace, (uint32_t)i, (uint32_t)iters);
    if (rec) { std::vector<double> st; k->unconstrained_state(&currTrace, st); rec->state.insert(rec->state.end(), st.begin(), st.end()); }
    if (i >= burnin && i % lag == 0) {
      samples.emplace_back();
      Sample& s = samples.back();
      s.value = currTrace.returnValue;
      s.logprob = currTrace.logprob * currTrace.temperature;
      s.loglikelihood = 0.0;
    }
  }
}

}  // namespace qso
