// qsb200.hpp -- header-only C++17 host mirror of the reference's inference interface for the hot path, over the
// plain-C ABI of qsb200.h.  Same names, parameter meaning and defaults as the reference:
//
//   qs::infer(program, query, method)                         qs/infer.t:59-70
//   qs::MCMC(kernel, {numsamps, burnin, lag, verbose})        qs/mcmc.t:36-105   (+ numchains, seed, device)
//   qs::HMCKernel{stepSize, numSteps, doStepSizeAdapt}        qs/hmc.t:57-66
//   qs::DriftKernel{scale, doScaleAdapt, lexicalScaleSharing} qs/drift.t:63-73
//   qs::HARMKernel{scale, doScaleAdapt}                       qs/harm.t:16-22
//   qs::MixtureKernel({kernels}, {weights})                   qs/mcmc.t:199-291
//   qs::AnnealingKernel(kernel, annealSched)                  qs/mcmc.t:297-319
//   qs::Sample { value[retdim], logprob, loglikelihood }      qs/infer.t:12-18
//   qs::Expectation / qs::MAP / qs::Autocorrelation           qs/infer.t:146-264 (device queries over the sample vector)
//
// Errors: the reference prints a message and aborts (qs/lib/std.t:35-43); here a qs::Error is thrown with the message.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>
#include "qsb200.h"

namespace qs {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int rc) { if (rc != 0) throw Error(std::string("qsb200: ") + qsb_last_error()); }

// ---- programs: a fixed-structure qs.program tagged with its device descriptor
class Program {
 public:
  explicit Program(const qsb_model_desc& d) : desc_(d) {}
  static Program Logistic(const double* X, const uint8_t* y, int N, int d, double prior_sigma = 2.5) {
    qsb_model_desc m{}; m.family = QSB_FAM_LOGISTIC; m.dim = d; m.nobs = N; m.X = X; m.ybin = y; m.prior_sigma = prior_sigma; m.ret_mode = 1;
    return Program(m);
  }
  static Program GaussPrecision(const double* A, int d) { qsb_model_desc m{}; m.family = QSB_FAM_GAUSS_PREC; m.dim = d; m.A = A; m.ret_mode = 1; return Program(m); }
  static Program EightSchools(const double* y, const double* sigma, int J) {
    qsb_model_desc m{}; m.family = QSB_FAM_EIGHT_SCHOOLS; m.dim = J + 2; m.nobs = J; m.y = y; m.sigma = sigma; m.ret_mode = 1; return Program(m);
  }
  static Program Funnel(int d) { qsb_model_desc m{}; m.family = QSB_FAM_FUNNEL; m.dim = d; m.ret_mode = 1; return Program(m); }
  static Program Indep(const int32_t* kind, const double* p0, const double* p1, int K, bool return_sum = true, double ret_div = 1.0) {
    qsb_model_desc m{}; m.family = QSB_FAM_INDEP; m.dim = K; m.erp_kind = kind; m.erp_p0 = p0; m.erp_p1 = p1;
    m.ret_mode = return_sum ? 0 : 1; m.ret_div = ret_div; return Program(m);
  }
  // optional parts of an INDEP descriptor (arrays stay owned by the caller, like every pointer of qsb_model_desc)
  Program& LexicalIds(const int32_t* lexid) { desc_.erp_lexid = lexid; return *this; }                       // erp.t:318-321
  Program& Factors(const double* fmu, const double* fsigma) { desc_.erp_fmu = fmu; desc_.erp_fsigma = fsigma; return *this; }   // trace.t:1099-1107
  Program& LatentParams(const int32_t* pref, const double* pcoef, const int32_t* ptf = nullptr) {           // erp.t:576-610
    desc_.erp_pref = pref; desc_.erp_pcoef = pcoef; desc_.erp_ptf = ptf; return *this;
  }
  const qsb_model_desc& desc() const { return desc_; }
  int retdim() const { return desc_.family == QSB_FAM_INDEP && desc_.ret_mode == 0 ? 1 : desc_.dim; }
 private:
  qsb_model_desc desc_;
};

// ---- kernels
struct Kernel {
  std::vector<qsb_kernel_spec> specs;
  std::function<double(uint32_t, uint32_t)> annealSched;   // AnnealingKernel: temperature of (iter, numiters); empty = none
};
struct HMCParams { double stepSize = 1.0; uint64_t numSteps = 1; bool doStepSizeAdapt = true; };
struct DriftParams { double scale = 1.0; bool doScaleAdapt = true; bool lexicalScaleSharing = false; };
struct HARMParams { double scale = 1.0; bool doScaleAdapt = true; };
inline Kernel HMCKernel(HMCParams p = {}) { return Kernel{{qsb_kernel_spec{QSB_KERNEL_HMC, p.doStepSizeAdapt ? 1 : 0, p.stepSize, p.numSteps, 1.0, 1.0, 0, 0}}}; }
inline Kernel DriftKernel(DriftParams p = {}) { return Kernel{{qsb_kernel_spec{QSB_KERNEL_DRIFT, p.doScaleAdapt ? 1 : 0, 1.0, 1, p.scale, 1.0, p.lexicalScaleSharing ? 1 : 0, 0}}}; }
// TraceMHKernel{doStruct=false} (mcmc.t:134-193): fixed-structure programs have non-structural choices only
inline Kernel TraceMHKernel() { return Kernel{{qsb_kernel_spec{QSB_KERNEL_TRACEMH, 0, 1.0, 1, 1.0, 1.0, 0, 0}}, {}}; }
inline Kernel HARMKernel(HARMParams p = {}) { return Kernel{{qsb_kernel_spec{QSB_KERNEL_HARM, p.doScaleAdapt ? 1 : 0, 1.0, 1, p.scale, 1.0, 0, 0}}}; }
inline Kernel MixtureKernel(const std::vector<Kernel>& kernels, const std::vector<double>& weights) {
  if (kernels.size() != weights.size()) throw Error("MixtureKernel: number of kernels must equal number of weights");   // mcmc.t:200-201
  Kernel k;
  for (size_t i = 0; i < kernels.size(); ++i) {
    if (kernels[i].specs.size() != 1) throw Error("MixtureKernel: nested mixtures are not supported");
    if (kernels[i].annealSched) throw Error("MixtureKernel: anneal the mixture, not its parts");
    qsb_kernel_spec s = kernels[i].specs[0]; s.weight = weights[i]; k.specs.push_back(s);
  }
  return k;
}
// AnnealingKernel(kernel, annealSched) (mcmc.t:297-319): the schedule stays host code and is tabulated once per run
inline Kernel AnnealingKernel(Kernel kernel, std::function<double(uint32_t, uint32_t)> annealSched) {
  if (kernel.annealSched) throw Error("AnnealingKernel: nested schedules are not supported");
  kernel.annealSched = std::move(annealSched);
  return kernel;
}

// ---- samples: contiguous storage laid out like S.Vector(Sample(T))._data, [chain][sample]{value[retdim], logprob, loglikelihood}
struct SampleVector {
  int retdim = 0; uint32_t numchains = 0; uint64_t numsamps = 0;
  std::vector<double> data;
  size_t stride() const { return (size_t)retdim + 2; }
  const double* value(uint32_t chain, uint64_t i) const { return &data[((size_t)chain * numsamps + i) * stride()]; }
  double logprob(uint32_t chain, uint64_t i) const { return value(chain, i)[retdim]; }
  double loglikelihood(uint32_t chain, uint64_t i) const { return value(chain, i)[retdim + 1]; }
  size_t size() const { return (size_t)numchains * numsamps; }
};

struct MCMCParams { uint64_t numsamps = 1000, burnin = 0, lag = 1; bool verbose = false; uint32_t numchains = 1; uint64_t seed = 0; int device = -1; };
using Method = std::function<void(const Program&, SampleVector&)>;

// MCMC(kernel, params) -> method(program, samples): the replacement of doMCMC (mcmc.t:50-90)
inline Method MCMC(Kernel kernel, MCMCParams p = {}) {
  return [kernel, p](const Program& program, SampleVector& samples) {
    qsb_model* model = nullptr; qsb_sampler* s = nullptr;
    check(qsb_model_create(&program.desc(), p.device, &model));
    qsb_run_config cfg{}; cfg.seed = p.seed; cfg.numchains = p.numchains; cfg.device = p.device;
    try {
      check(qsb_sampler_create(model, kernel.specs.data(), (int32_t)kernel.specs.size(), &cfg, &s));
      check(qsb_sampler_init(s));
      if (kernel.annealSched) {
        const uint64_t iters = p.burnin + p.numsamps * p.lag;
        std::vector<double> temps(iters);
        for (uint64_t i = 0; i < iters; ++i) temps[i] = kernel.annealSched((uint32_t)i, (uint32_t)iters);
        check(qsb_sampler_set_temperatures(s, temps.data(), iters));
      }
      check(qsb_sampler_run(s, p.numsamps, p.burnin, p.lag));
      samples.retdim = program.retdim(); samples.numchains = p.numchains; samples.numsamps = qsb_sampler_num_samples(s);
      samples.data.resize(samples.size() * samples.stride());
      check(qsb_sampler_fetch_samples(s, 0, p.numchains, 0, samples.numsamps, samples.data.data(), samples.stride() * sizeof(double)));
      if (p.verbose) {   // same line formats as mcmc.t:81,88
        qsb_stats st; check(qsb_sampler_stats(s, &st));
        std::printf("Acceptance Ratio: %llu/%llu (%g%%)\n", (unsigned long long)st.props_accepted, (unsigned long long)st.props_made,
                    100.0 * (double)st.props_accepted / (double)st.props_made);
        std::printf("Time: %g\n", st.seconds);
      }
    } catch (...) { if (s) qsb_sampler_destroy(s); qsb_model_destroy(model); throw; }
    qsb_sampler_destroy(s); qsb_model_destroy(model);
  };
}

// Queries (infer.t:146-264), evaluated by the library's CUDA kernels over the caller-owned sample vector.
// Expectation: [numchains][retdim] means (mean starts at value[0], weights exp(loglikelihood) summed from i = 1) and,
// if `variance` is given, the per-chain variances (sum |v - mean|^2 / N).
inline std::vector<double> Expectation(const SampleVector& v, std::vector<double>* variance = nullptr) {
  std::vector<double> m((size_t)v.numchains * v.retdim);
  if (variance) variance->resize(v.numchains);
  check(qsb_query_expectation_host(v.data.data(), v.stride() * sizeof(double), v.numchains, v.numsamps, v.retdim, m.data(),
                                   variance ? variance->data() : nullptr));
  return m;
}
inline std::vector<double> MAP(const SampleVector& v) {   // [numchains][retdim]
  std::vector<double> val((size_t)v.numchains * v.retdim);
  check(qsb_query_map_host(v.data.data(), v.stride() * sizeof(double), v.numchains, v.numsamps, v.retdim, val.data(), nullptr));
  return val;
}
inline std::vector<double> Autocorrelation(const SampleVector& v, const std::vector<double>* mean = nullptr, double variance = 0.0) {   // [numchains][numsamps]
  std::vector<double> ac((size_t)v.numchains * v.numsamps);
  check(qsb_query_autocorrelation_host(v.data.data(), v.stride() * sizeof(double), v.numchains, v.numsamps, v.retdim,
                                       mean ? mean->data() : nullptr, variance, ac.data()));
  return ac;
}

// infer(program, query, method) (infer.t:59-70)
template <class Query> auto infer(const Program& program, Query query, const Method& method) {
  return [&program, query, method]() { SampleVector s; method(program, s); return query(s); };
}

}  // namespace qs
