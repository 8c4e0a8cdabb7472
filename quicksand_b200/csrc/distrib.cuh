// Hand-written fp64 log-densities, their value-gradients, the bounding transforms and the
// forward samplers of the continuous ERPs -- the device replacement of qs/distrib.t, the Bounds
// table of qs/erp.t:128-198 and of the tape AD (qs/lib/ad.t) that differentiated them.
// Closed forms: SURVEY.md Appendix A.  Every quirk the reference has is kept:
//   * log_gamma uses 5 of the 6 Lanczos coefficients (distrib.t:100, `for j=0,5` is half-open);
//     pinned by the gamma/beta known-answer values of test/testsuite.t:165-170
//   * clamped branches of fmax/fmin and of the logistic saturation are constants: zero gradient
//     (ad.t:545-575; erp.t:181-182)
#pragma once
#include <cmath>
#include <cstdint>
#include "philox.cuh"

namespace qsb {

#define QSB_HD __host__ __device__ __forceinline__
#define QSB_D __device__ __forceinline__

#ifdef __CUDA_ARCH__
#define QSB_INF (__longlong_as_double(0x7ff0000000000000LL))
#else
#define QSB_INF (INFINITY)
#endif

// ---- gaussian (distrib.t:75-78)
QSB_D double gaussian_lp(double x, double mu, double sigma) {
  double xm = x - mu;
  return -.5 * (1.8378770664093453 + 2.0 * log(sigma) + xm * xm / (sigma * sigma));
}
QSB_D double gaussian_dlp_dx(double x, double mu, double sigma) { return -(x - mu) / (sigma * sigma); }

// ---- log_gamma (distrib.t:84-106), 5 coefficients
QSB_D double log_gamma5(double xx) {
  const double cof[5] = {76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155, 0.1208650973866179e-2};
  double x = xx - 1.0;
  double tmp = x + 5.5;
  tmp = tmp - (x + 0.5) * log(tmp);
  double ser = 1.000000000190015;
#pragma unroll
  for (int j = 0; j < 5; ++j) { x = x + 1.0; ser = ser + cof[j] / x; }
  return -tmp + log(2.5066282746310005 * ser);
}

// d/dz of log_gamma5 as the reference's tape differentiates it (not the true digamma: they differ by ~2.6e-8 on [1,2])
QSB_D double digamma5(double xx) {
  const double cof[5] = {76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155, 0.1208650973866179e-2};
  const double t = (xx - 1.0) + 5.5;
  double ser = 1.000000000190015, dser = 0.0, x = xx - 1.0;
#pragma unroll
  for (int j = 0; j < 5; ++j) { x = x + 1.0; ser = ser + cof[j] / x; dser = dser - cof[j] / (x * x); }
  return log(t) + ((xx - 1.0) + 0.5) / t - 1.0 + dser / ser;
}

// ---- gamma (distrib.t:130-133)
QSB_D double gamma_lp(double x, double shape, double scale) {
  if (x <= 0.0) return -QSB_INF;
  return (shape - 1.0) * log(x) - x / scale - log_gamma5(shape) - shape * log(scale);
}
QSB_D double gamma_dlp_dx(double x, double shape, double scale) { return (shape - 1.0) / x - 1.0 / scale; }

// ---- beta (distrib.t:139-158)
QSB_D double log_beta5(double a, double b) { return log_gamma5(a) + log_gamma5(b) - log_gamma5(a + b); }
QSB_D double beta_lp(double x, double a, double b) {
  if (x > 0.0 && x < 1.0) return (a - 1.0) * log(x) + (b - 1.0) * log(1.0 - x) - log_beta5(a, b);
  return -QSB_INF;
}
QSB_D double beta_dlp_dx(double x, double a, double b) { return (a - 1.0) / x - (b - 1.0) / (1.0 - x); }

// ---- uniform (distrib.t:37-40)
QSB_D double uniform_lp(double x, double lo, double hi) {
  if (x < lo || x > hi) return -QSB_INF;
  return -log(hi - lo);
}

// ---- bernoulli (distrib.t:17-25)
QSB_D double bernoulli_lp(bool val, double p) { return log(val ? p : 1.0 - p); }

// ---- Bounds (erp.t:91-96, 141-198)
QSB_D double logistic_(double x) { return 1.0 / (1.0 + exp(-x)); }
QSB_D double invlogistic_(double y) { return -log(1.0 / y - 1.0); }

// ERP_NEGGAMMA: value = -gamma(shape, scale) < 0 under Bounds.Upper(0) (erp.t:158-173): the transform the reference offers
// to user-defined ERPs (makeRandomChoice) but binds to none of its own.
enum { ERP_GAUSSIAN = 0, ERP_GAMMA = 1, ERP_BETA = 2, ERP_UNIFORM = 3, ERP_DIRICHLET = 4, ERP_NEGGAMMA = 5 };

// ---- dirichlet (distrib.t:336-362) under Bounds.UnitSimplex (erp.t:200-254), one component at a time.
// A K-vector dirichlet choice occupies K consecutive real components; component K-1 of the unconstrained vector is
// unused by the transform (zero gradient).  The stick-breaking recurrence is carried from component to component:
//   z_k = logistic(x_k - log(K-1-k)),  y_k = stick*z_k,  stick -= y_k,  y_{K-1} = stick
//   logJ = sum_{k<K-1} log stick_k + log z_k + log(1-z_k)
//   lp   = lg5(sum alpha) + sum_k (alpha_k-1) log y_k - lg5(alpha_k)
// and d(lp + logJ)/dx_k = alpha_k (1-z_k) - z_k * sum_{m>k} alpha_m  (what the tape AD accumulates through the recurrence).
struct SimplexCarry { double stick, lp, lj; };
// returns y_k; adds the component's terms to the carry; *g = d(lp+logJ)/dx_k
QSB_D double dirichlet_component(SimplexCarry& c, int k, int K, double x, double alpha, double asum_after, double* g) {
  const int Km1 = K - 1;
  if (k == 0) { c.stick = 1.0; c.lp = log_gamma5(alpha + asum_after); c.lj = 0.0; }
  double y, gi = 0.0;
  if (k < Km1) {
    const double z = logistic_(x - log((double)(Km1 - k)));
    y = c.stick * z;
    c.lj = c.lj + log(c.stick); c.lj = c.lj + log(z); c.lj = c.lj + log(1.0 - z);
    c.stick = c.stick - y;
    gi = alpha * (1.0 - z) - z * asum_after;
  } else y = c.stick;
  c.lp = c.lp + (alpha - 1.0) * log(y);
  c.lp = c.lp - log_gamma5(alpha);
  if (g) *g = gi;
  return y;
}

// Forward transform x -> y of an ERP kind, with dy/dx as the tape AD would produce it.
QSB_D double erp_fwd(int kind, double x, double p0, double p1, double* dydx) {
  switch (kind) {
    case ERP_GAMMA: {                                  // Lower(0): fmax(exp(x)+lo, lo+1e-15)
      double e = exp(x) + 0.0;
      if (e >= 0.0 + 1e-15) { if (dydx) *dydx = e; return e; }
      if (dydx) *dydx = 0.0;
      return 0.0 + 1e-15;
    }
    case ERP_NEGGAMMA: {                               // Upper(0): fmin(hi - exp(x), hi - 1e-15); zero gradient on the clamp
      double ex = exp(x);
      double v = 0.0 - ex;
      if (v <= 0.0 - 1e-15) { if (dydx) *dydx = -ex; return v; }
      if (dydx) *dydx = 0.0;
      return 0.0 - 1e-15;
    }
    case ERP_BETA:
    case ERP_UNIFORM: {                                // LowerUpper(lo,hi)
      double lo = kind == ERP_BETA ? 0.0 : p0, hi = kind == ERP_BETA ? 1.0 : p1;
      double e = exp(-x);
      double ope = 1.0 + e;
      double s = 1.0 / ope;
      double ds = e / (ope * ope);
      if (x > -QSB_INF && s == 0.0) { s = 1e-15; ds = 0.0; }
      if (x < QSB_INF && s == 1.0) { s = 1.0 - 1e-15; ds = 0.0; }
      if (dydx) *dydx = (hi - lo) * ds;
      return lo + (hi - lo) * s;
    }
    default:
      if (dydx) *dydx = 1.0;
      return x;
  }
}
QSB_D double erp_inv(int kind, double y, double p0, double p1) {
  switch (kind) {
    case ERP_GAMMA: { double z = fmax(y, 0.0 + 1e-15); return log(z - 0.0); }
    case ERP_NEGGAMMA: { double z = fmin(y, 0.0 - 1e-15); return log(0.0 - z); }
    case ERP_BETA:
    case ERP_UNIFORM: {
      double lo = kind == ERP_BETA ? 0.0 : p0, hi = kind == ERP_BETA ? 1.0 : p1;
      double z = fmax(fmin(y, hi - 1e-15), lo + 1e-15);
      double t = (z - lo) / (hi - lo);
      return invlogistic_(t);
    }
    default: return y;
  }
}
// logprobIncrement (log-Jacobian) and its derivative
QSB_D double erp_logjac(int kind, double x, double p0, double p1, double* dldx) {
  switch (kind) {
    case ERP_GAMMA:
    case ERP_NEGGAMMA: if (dldx) *dldx = 1.0; return x;
    case ERP_BETA:
    case ERP_UNIFORM: {
      double lo = kind == ERP_BETA ? 0.0 : p0, hi = kind == ERP_BETA ? 1.0 : p1;
      double e = exp(-x);
      double ope = 1.0 + e;
      if (dldx) *dldx = e == QSB_INF ? -1.0 : -1.0 + 2.0 * (e / ope);   // ad.t:166-172: a -inf result has adjoint... see Appendix A
      return log(hi - lo) - x - 2.0 * log(ope);
    }
    default: if (dldx) *dldx = 0.0; return 0.0;
  }
}
QSB_D double erp_prior_lp(int kind, double y, double p0, double p1, double* dlpdy) {
  switch (kind) {
    case ERP_GAMMA: if (dlpdy) *dlpdy = gamma_dlp_dx(y, p0, p1); return gamma_lp(y, p0, p1);
    case ERP_NEGGAMMA: if (dlpdy) *dlpdy = -gamma_dlp_dx(-y, p0, p1); return gamma_lp(-y, p0, p1);
    case ERP_BETA: if (dlpdy) *dlpdy = beta_dlp_dx(y, p0, p1); return beta_lp(y, p0, p1);
    case ERP_UNIFORM: if (dlpdy) *dlpdy = 0.0; return uniform_lp(y, p0, p1);
    default: if (dlpdy) *dlpdy = gaussian_dlp_dx(y, p0, p1); return gaussian_lp(y, p0, p1);
  }
}

// d(prior logprob)/d(parameters) of a scalar ERP at value y: what the tape accumulates into the parameter expressions of
// a choice whose parameters depend on earlier choices
QSB_D void erp_dlp_dparams(int kind, double y, double p0, double p1, double& d0, double& d1) {
  switch (kind) {
    case ERP_NEGGAMMA: y = -y;   // the gamma density of -y; fall through
    case ERP_GAMMA:    // (k-1) ln y - y/theta - lg5(k) - k ln theta
      d0 = log(y) - digamma5(p0) - log(p1);
      d1 = y / (p1 * p1) - p0 / p1;
      break;
    case ERP_BETA: {   // (a-1) ln y + (b-1) ln(1-y) - lg5(a) - lg5(b) + lg5(a+b)
      const double dab = digamma5(p0 + p1);
      d0 = log(y) - digamma5(p0) + dab;
      d1 = log(1.0 - y) - digamma5(p1) + dab;
      break;
    }
    case ERP_UNIFORM: d0 = 1.0 / (p1 - p0); d1 = -1.0 / (p1 - p0); break;
    default: {         // -.5 (c + 2 ln sigma + (y-mu)^2 / sigma^2)
      const double r = y - p0, s2 = p1 * p1;
      d0 = r / s2;
      d1 = -1.0 / p1 + r * r / (s2 * p1);
    }
  }
}

// ---- forward samplers on a sequential sub-stream (initialisation only: trace.t:612-623)
QSB_D double uniform_sample_seq(SubStream& s, double lo, double hi) { double u = s.random(); return __dadd_rn(__dmul_rn(__dsub_rn(1.0, u), lo), __dmul_rn(u, hi)); }
// Marsaglia-Tsang for shape >= 1 (distrib.t:111-124).  Un-contracted arithmetic (no FMA): the draw is
// bit-identical to a CPU evaluation of the same expressions, like the Gaussian sampler.
QSB_D double gamma_mt_seq(SubStream& s, double shape, double scale) {
  double d = __dsub_rn(shape, 1.0 / 3.0), c = 1.0 / sqrt(__dmul_rn(9.0, d)), x, v, u;
  while (true) {
    do { x = gaussian_sample_seq(s, 0.0, 1.0); v = __dadd_rn(1.0, __dmul_rn(c, x)); } while (!(v > 0.0));
    v = __dmul_rn(__dmul_rn(v, v), v);
    u = s.random();
    double x2 = __dmul_rn(x, x);
    double squeeze = __dsub_rn(1.0, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(.331, x), x), x), x));
    if (u < squeeze) return __dmul_rn(__dmul_rn(scale, d), v);
    double rhs = __dadd_rn(__dmul_rn(__dmul_rn(.5, x), x), __dmul_rn(d, __dadd_rn(__dsub_rn(1.0, v), log(v))));
    (void)x2;
    if (log(u) < rhs) return __dmul_rn(__dmul_rn(scale, d), v);
  }
}
QSB_D double gamma_sample_seq(SubStream& s, double shape, double scale) {   // distrib.t:109-128
  if (shape < 1.0) {   // distrib.t:110: the boosted draw first, then the uniform of the power
    double g = gamma_mt_seq(s, __dadd_rn(1.0, shape), scale);
    return __dmul_rn(g, pow(s.random(), 1.0 / shape));
  }
  return gamma_mt_seq(s, shape, scale);
}
QSB_D double beta_sample_seq(SubStream& s, double a, double b) {   // distrib.t:150-153
  double x = gamma_sample_seq(s, a, 1.0);
  return x / (x + gamma_sample_seq(s, b, 1.0));
}
QSB_D double erp_sample_seq(int kind, SubStream& s, double p0, double p1) {
  switch (kind) {
    case ERP_GAMMA: return gamma_sample_seq(s, p0, p1);
    case ERP_NEGGAMMA: return -gamma_sample_seq(s, p0, p1);
    case ERP_BETA: return beta_sample_seq(s, p0, p1);
    case ERP_UNIFORM: return uniform_sample_seq(s, p0, p1);
    default: return gaussian_sample_seq(s, p0, p1);
  }
}

}  // namespace qsb
