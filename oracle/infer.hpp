// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the queries of qs/infer.t that the hot path feeds:
//   Expectation(doVariance)  infer.t:146-187  (mean starts at value[0], weights summed from i=1:
//                                              mean = sum(v)/(N-1) for MCMC samples; variance / N)
//   Autocorrelation          infer.t:213-264
// Vector-valued returns are treated component-wise for the mean; the variance uses the
// inner product diff*diff like the reference's "*" on vector types.
#pragma once
#include <cmath>
#include <vector>
#include "kernels.hpp"

namespace qso {

inline void Expectation(const std::vector<Sample>& samples, bool doVariance, std::vector<double>& m, double& v) {
  size_t dim = samples[0].value.size();
  m = samples[0].value;
  double totalw = 0.0;
  for (size_t i = 1; i < samples.size(); ++i) {
    const Sample& s = samples[i];
    double w = std::exp(s.loglikelihood);
    totalw = totalw + w;
    for (size_t j = 0; j < dim; ++j) m[j] = m[j] + w * s.value[j];
  }
  for (size_t j = 0; j < dim; ++j) m[j] = m[j] / totalw;
  v = 0.0;
  if (doVariance) {
    for (const Sample& s : samples) {
      double dd = 0.0;
      for (size_t j = 0; j < dim; ++j) { double diff = s.value[j] - m[j]; dd += diff * diff; }
      v = v + dd;
    }
    v = v / samples.size();
  }
}

inline void Autocorrelation(const std::vector<Sample>& samples, const std::vector<double>& m, double v, std::vector<double>& ac) {
  size_t dim = samples[0].value.size();
  ac.clear();
  for (size_t t = 0; t < samples.size(); ++t) {
    double act = 0.0;
    size_t n = samples.size() - t;
    for (size_t i = 0; i < n; ++i) {
      double dd = 0.0;
      for (size_t j = 0; j < dim; ++j) dd += (samples[i].value[j] - m[j]) * (samples[i + t].value[j] - m[j]);
      act = act + dd;
    }
    if (n > 0) act = act / (n * v);
    ac.push_back(act);
  }
}

}  // namespace qso
