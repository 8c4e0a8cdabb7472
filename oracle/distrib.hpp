// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of qs/distrib.t: sample + logprob pairs, generic over `real`
// (double for ordinary traces, ad::num for HMC's dual trace).  Every quirk the
// reference has is kept, notably the log_gamma loop that uses only 5 of the 6
// Lanczos coefficients (distrib.t:100; pinned by the gamma/beta goldens of
// test/testsuite.t:165-170).
#pragma once
#include <cmath>
#include <limits>
#include <vector>
#include "ad.hpp"
#include "philox.hpp"

namespace qso {
namespace D {

static const double NEG_INF = -std::numeric_limits<double>::infinity();

// ---- bernoulli (distrib.t:11-27)
inline bool bernoulli_sample(double p) { double randval = random_(); return randval < p; }
template <class real> inline real bernoulli_logprob(bool val, real p) {
  real prob = p;
  if (!val) prob = 1.0 - p;
  return tmath::log(prob);
}

// ---- uniform (distrib.t:31-42)
inline double uniform_sample(double lo, double hi) { double u = random_(); return (1.0 - u) * lo + u * hi; }
template <class real> inline real uniform_logprob(real val, real lo, real hi) {
  if (val < lo || val > hi) return real(NEG_INF);
  return -tmath::log(hi - lo);
}

// ---- uniformInt logprob (distrib.t:52-55)
inline double uniformInt_logprob(int val, int lo, int hi) {
  if (val < lo || val >= hi) return NEG_INF;
  return -std::log((double)(hi - lo));
}

// ---- gaussian (distrib.t:61-80): Leva's ratio-of-uniforms sampler
inline double gaussian_sample(double mu, double sigma) {
  double u, v, x, y, q;
  do {
    u = 1.0 - random_();
    v = 1.7156 * (random_() - 0.5);
    x = u - 0.449871;
    y = std::fabs(v) + 0.386595;
    q = x * x + y * (0.196 * y - 0.25472 * x);
  } while (q >= 0.27597 && (q > 0.27846 || v * v > -4 * u * u * std::log(u)));
  return mu + sigma * v / u;
}
template <class real> inline real gaussian_logprob(real x, real mu, real sigma) {
  real xminusmu = x - mu;
  return -.5 * (1.8378770664093453 + 2 * tmath::log(sigma) + xminusmu * xminusmu / (sigma * sigma));
}

// ---- log_gamma (distrib.t:84-106) -- `for j=0,5` is half-open: coefficients 0..4 only
static const double gamma_cof[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                                    -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
template <class real> inline real log_gamma(real xx) {
  real x = xx - 1.0;
  real tmp = x + 5.5;
  tmp = tmp - (x + 0.5) * tmath::log(tmp);
  real ser = real(1.000000000190015);
  for (int j = 0; j < 5; ++j) {
    x = x + 1.0;
    ser = ser + gamma_cof[j] / x;
  }
  return -tmp + tmath::log(2.5066282746310005 * ser);
}

// ---- gamma (distrib.t:109-135): Marsaglia-Tsang
inline double gamma_sample(double shape, double scale) {
  if (shape < 1.0) {   // left operand first (Terra emits operands in source order); C++ leaves the order unspecified
    double g = gamma_sample(1.0 + shape, scale);
    return g * std::pow(random_(), 1.0 / shape);
  }
  double x, v, u;
  double d = shape - 1.0 / 3.0;
  double c = 1.0 / std::sqrt(9.0 * d);
  while (true) {
    do {
      x = gaussian_sample(0.0, 1.0);
      v = 1.0 + c * x;
    } while (!(v > 0.0));
    v = v * v * v;
    u = random_();
    if ((u < 1.0 - .331 * x * x * x * x) || (std::log(u) < .5 * x * x + d * (1.0 - v + std::log(v)))) return scale * d * v;
  }
}
template <class real> inline real gamma_logprob(real x, real shape, real scale) {
  if (x <= 0.0) return real(NEG_INF);
  return (shape - 1.0) * tmath::log(x) - x / scale - log_gamma(shape) - shape * tmath::log(scale);
}

// ---- beta (distrib.t:139-160)
template <class real> inline real log_beta(real a, real b) { return log_gamma(a) + log_gamma(b) - log_gamma(a + b); }
inline double beta_sample(double a, double b) {
  double x = gamma_sample(a, 1.0);
  return x / (x + gamma_sample(b, 1.0));
}
template <class real> inline real beta_logprob(real x, real a, real b) {
  if (x > 0.0 && x < 1.0) return (a - 1.0) * tmath::log(x) + (b - 1.0) * tmath::log(1.0 - x) - log_beta(a, b);
  return real(NEG_INF);
}

// ---- binomial logprob (distrib.t:164-171, 203-217) -- discrete, out of scope for the
// kernels, restated only so that all 26 density goldens of testsuite.t:159-187 are checked.
inline double g_(double x) {
  if (x == 0.0) return 1.0;
  if (x == 1.0) return 0.0;
  double d = 1.0 - x;
  return (1.0 - (x * x) + (2.0 * x * std::log(x))) / (d * d);
}
inline double binomial_logprob(int s, double p, int n) {
  const double inv2 = 1.0 / 2, inv3 = 1.0 / 3, inv6 = 1.0 / 6;
  if (s < 0 || s >= n) return NEG_INF;
  double q = 1.0 - p;
  double S = s + inv2;
  double T = n - s - inv2;
  double d1 = s + inv6 - (n + inv3) * p;
  double d2 = q / (s + inv2) - p / (T + inv2) + (q - inv2) / (n + 1);
  d2 = d1 + 0.02 * d2;
  double num = 1.0 + q * g_(S / (n * p)) + p * g_(T / (n * q));
  double den = (n + inv6) * p * q;
  double z = num / den;
  double invsd = std::sqrt(z);
  z = d2 * invsd;
  return gaussian_logprob<double>(z, 0.0, 1.0) + std::log(invsd);
}

// ---- poisson logprob (distrib.t:223-245, 269-272)
inline int fact_(int x) { int t = 1; while (x > 1) { t = t * x; x = x - 1; } return t; }
inline double lnfact_(int x) {
  if (x < 1) x = 1;
  if (x < 12) return std::log((double)fact_(x));
  double invx = 1.0 / x, invx2 = invx * invx, invx3 = invx2 * invx, invx5 = invx3 * invx2, invx7 = invx5 * invx2;
  double ssum = ((x + 0.5) * std::log((double)x)) - x;
  ssum = ssum + std::log(2 * M_PI) / 2.0;
  ssum = ssum + (invx / 12) - (invx / 360);
  ssum = ssum + (invx5 / 1260) - (invx7 / 1680);
  return ssum;
}
inline double poisson_logprob(int k, double mu) {
  if (k < 0) return NEG_INF;
  return k * std::log(mu) - mu - lnfact_(k);
}

// ---- categorical (distrib.t:278-332)
inline int categorical_sample(const double* params, int N) {
  double sum = 0.0;
  for (int i = 0; i < N; ++i) sum = sum + params[i];
  int result = 0;
  double x = random_() * sum;
  double probAccum = 0.0;
  do {
    probAccum = probAccum + params[result];
    result = result + 1;
  } while (!(probAccum > x || result == N));
  return result - 1;
}
inline double categorical_logprob(int val, const double* params, int N) {
  if (val < 0 || val >= N) return NEG_INF;
  double sum = 0.0;
  for (int i = 0; i < N; ++i) sum = sum + params[i];
  return std::log(params[val] / sum);
}

// ---- dirichlet (distrib.t:336-398)
inline void dirichlet_sample(const double* params, int N, double* out) {
  double ssum = 0.0;
  for (int i = 0; i < N; ++i) { double t = gamma_sample(params[i], 1.0); out[i] = t; ssum = ssum + t; }
  for (int i = 0; i < N; ++i) out[i] = out[i] / ssum;
}
template <class real> inline real dirichlet_logprob(const real* theta, const real* params, int N) {
  real sum = real(0.0);
  for (int i = 0; i < N; ++i) sum = sum + params[i];
  real logp = log_gamma(sum);
  for (int i = 0; i < N; ++i) {
    real a = params[i];
    logp = logp + (a - 1.0) * tmath::log(theta[i]);
    logp = logp - log_gamma(a);
  }
  return logp;
}

}  // namespace D
}  // namespace qso
