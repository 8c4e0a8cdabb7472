// ORACLE (test infrastructure, never shipped, never on the product path).
//
// The fixed-structure probabilistic programs of BASELINE.json's configs, written the
// way a quicksand user would write them (ERP calls, qs.factor, erp.observe) and generic
// over `real` exactly like a Terra program compiled under qs.real = double / ad.num
// (qs/progmodule.t:11-46).  The oracle re-executes the whole body for every log-density
// and every gradient, as the reference does (qs/hmc.t:325-342, qs/trace.t:711-773).
//
//   F1 IndepErps       test/testsuite.t:621-691, 719-730, 776-787, 812-823  (config c1)
//   F2 GaussPrecision  10-D correlated Gaussian                              (config c2)
//   F3 Logistic        Bayesian logistic regression                          (config c3)
//   F4 EightSchools    hierarchical normal, J groups                         (config c4)
//   F5 Funnel          Neal's funnel                                         (config c5)
#pragma once
#include <vector>
#include "trace.hpp"

namespace qso {

enum Family { FAM_INDEP = 1, FAM_GAUSS_PREC = 2, FAM_LOGISTIC = 3, FAM_EIGHT_SCHOOLS = 4, FAM_FUNNEL = 5 };

// F1: K independent ERPs with constant parameters.
//   ret_mode 0: return (sum_i value_i) / ret_div    e.g. `return qs.gamma(2.0, 2.0)/10.0`, `return sum/10.0`
//   ret_mode 1: return the vector of values
struct IndepErps : Program {
  std::vector<int> kind;
  std::vector<double> p0, p1;
  std::vector<int> lexid;   // call-site id per entry (empty: every choice is its own call site)
  std::vector<double> fmu, fsigma;   // optional qs.factor(gaussian.logprob(value_i, fmu_i, fsigma_i)) after scalar choice i (fsigma <= 0: none)
  // latent parameters: parameter q of scalar entry i is f(p_q[i] + lat_b[2i+q] * value of the earlier scalar entry
  // lat_ref[2i+q]) with f = identity (lat_tf = 0) or exp (1); lat_ref < 0: the constant p_q[i].  Empty: all constant.
  // This is what a program like  mu = gaussian(0,1); x = gaussian(mu, exp(logtau))  does: the parameter expression is
  // re-evaluated on every run (trace.t:948-994 passes the current parameters to RandomChoiceT:update, erp.t:576-610), and in
  // a dual trace the tape differentiates the log-density with respect to it.
  std::vector<int> lat_ref, lat_tf;
  std::vector<double> lat_b;
  // optional hard constraint after scalar choice i: qs.condition(value_i > cond_lo[i] and value_i < cond_hi[i])
  // (trace.t:794-796, 1103-1107; +-inf = no bound).  Empty: none.
  std::vector<double> cond_lo, cond_hi;
  int ret_mode = 0;
  double ret_div = 1.0;
  // Entries are per real component.  A dirichlet choice of K components is K consecutive entries of kind
  // ERP_DIRICHLET with p0 = alpha_k and p1 = k (its index inside the choice); it is ONE ERP call site.
  template <class real> void body(Trace<real>& t) {
    size_t K = kind.size();
    real sum = real(0.0);
    std::vector<real> vals(K, real(0.0));
    auto par = [&](int q, size_t i) -> real {
      real a = real(q ? p1[i] : p0[i]);
      if (lat_ref.empty() || lat_ref[2 * i + q] < 0) return a;
      real z = a + lat_b[2 * i + q] * vals[(size_t)lat_ref[2 * i + q]];
      return lat_tf[2 * i + q] ? tmath::exp(z) : z;
    };
    for (size_t i = 0; i < K;) {
      t.site = lexid.empty() ? (unsigned)i : (unsigned)lexid[i];
      if (kind[i] == ERP_DIRICHLET) {
        size_t j = i + 1;
        while (j < K && kind[j] == ERP_DIRICHLET && p1[j] == (double)(j - i)) ++j;
        std::vector<real> params;
        for (size_t m = i; m < j; ++m) params.push_back(real(p0[m]));
        std::vector<real> v = t.dirichlet(params);
        for (auto& x : v) { if (ret_mode == 0) sum = sum + x; else t.returnValue.push_back(x); }
        i = j;
      } else {
        real x = t.erp(kind[i], par(0, i), par(1, i));
        vals[i] = x;
        if (!fsigma.empty() && fsigma[i] > 0.0) t.factor(D::gaussian_logprob<real>(x, real(fmu[i]), real(fsigma[i])));   // testsuite.t:394-442
        if (!cond_lo.empty() && (cond_lo[i] > -HUGE_VAL || cond_hi[i] < HUGE_VAL)) t.condition(val(x) > cond_lo[i] && val(x) < cond_hi[i]);
        if (ret_mode == 0) sum = sum + x; else t.returnValue.push_back(x);
        ++i;
      }
    }
    if (ret_mode == 0) t.returnValue.push_back(sum / ret_div);
  }
  QSO_PROGRAM_RUNNERS
};

// F2: x_i ~ gaussian(0,1), i<d;  factor(-0.5 * x' A x)  with A = P - I, so that log p = -0.5 x' P x + c.
struct GaussPrecision : Program {
  int d = 0;
  std::vector<double> A;   // d x d row-major
  template <class real> void body(Trace<real>& t) {
    std::vector<real> x(d);
    t.site = 0;   // one call site inside the loop
    for (int i = 0; i < d; ++i) x[i] = t.gaussian(real(0.0), real(1.0));
    real q = real(0.0);
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) q = q + x[i] * A[(size_t)i * d + j] * x[j];
    t.factor(-0.5 * q);
    for (int i = 0; i < d; ++i) t.returnValue.push_back(x[i]);
  }
  QSO_PROGRAM_RUNNERS
};

// F3: theta_j ~ gaussian(0, prior_sigma), j<d;  flip.observe(y_i, 1/(1+exp(-x_i . theta))), i<N.
struct Logistic : Program {
  int N = 0, d = 0;
  double prior_sigma = 2.5;
  std::vector<double> X;        // N x d row-major
  std::vector<unsigned char> y; // N
  template <class real> void body(Trace<real>& t) {
    std::vector<real> theta(d);
    t.site = 0;
    for (int j = 0; j < d; ++j) theta[j] = t.gaussian(real(0.0), real(prior_sigma));
    for (int i = 0; i < N; ++i) {
      real z = real(0.0);
      const double* xi = &X[(size_t)i * d];
      for (int j = 0; j < d; ++j) z = z + xi[j] * theta[j];
      real p = 1.0 / (1.0 + tmath::exp(-z));
      t.flip_observe(y[i] != 0, p);
    }
    for (int j = 0; j < d; ++j) t.returnValue.push_back(theta[j]);
  }
  QSO_PROGRAM_RUNNERS
};

// F4: mu ~ gaussian(0,5); logtau ~ gaussian(0,1); theta_j ~ gaussian(mu, exp(logtau));
//     gaussian.observe(y_j, theta_j, sigma_j).  Components: [mu, logtau, theta_0..theta_{J-1}].
struct EightSchools : Program {
  int J = 0;
  std::vector<double> y, sigma;
  template <class real> void body(Trace<real>& t) {
    t.site = 0;
    real mu = t.gaussian(real(0.0), real(5.0));
    t.site = 1;
    real logtau = t.gaussian(real(0.0), real(1.0));
    t.site = 2;   // theta_j: one call site inside the loop
    real tau = tmath::exp(logtau);
    t.returnValue.push_back(mu);
    t.returnValue.push_back(logtau);
    for (int j = 0; j < J; ++j) {
      real th = t.gaussian(mu, tau);
      t.gaussian_observe(real(y[j]), th, real(sigma[j]));
      t.returnValue.push_back(th);
    }
  }
  QSO_PROGRAM_RUNNERS
};

// F5: v ~ gaussian(0,3); x_i ~ gaussian(0, exp(v/2)), i<d-1.  Components: [v, x_0..x_{d-2}].
struct Funnel : Program {
  int d = 0;
  template <class real> void body(Trace<real>& t) {
    t.site = 0;
    real v = t.gaussian(real(0.0), real(3.0));
    t.site = 1;   // x_i: one call site inside the loop
    real s = tmath::exp(v / 2.0);
    t.returnValue.push_back(v);
    for (int i = 0; i < d - 1; ++i) t.returnValue.push_back(t.gaussian(real(0.0), s));
  }
  QSO_PROGRAM_RUNNERS
};

}  // namespace qso
