// The reference's test shape (test/testsuite.t:92-141: expectedValueTest) against the C++ mirror: mean of 10 iid
// gaussian(0.1, 0.5) under HMC numSteps=10 (testsuite.t:719-730).  Needs a GPU to run; compiled and linked by the CPU suite.
#include <cstdio>
#include "qsb200.hpp"
int main() {
  std::vector<int32_t> kind(10, QSB_ERP_GAUSSIAN);
  std::vector<double> mu(10, 0.1), sigma(10, 0.5);
  qs::Program prog = qs::Program::Indep(kind.data(), mu.data(), sigma.data(), 10, true, 10.0);
  qs::HMCParams hp; hp.numSteps = 10;
  qs::MCMCParams mp; mp.numsamps = 150; mp.lag = 20; mp.numchains = 256; mp.seed = 1;
  try {
    auto run = qs::infer(prog, [](const qs::SampleVector& s) { double m = 0; for (double e : qs::Expectation(s)) m += e; return m / s.numchains; },
                         qs::MCMC(qs::HMCKernel(hp), mp));
    double est = run();
    std::printf("estimate %.4f (truth 0.1)\n", est);
    if (!(std::fabs(est - 0.1) <= 0.07)) return 1;
    // a parameter that reads an earlier choice:  mu = gaussian(1, 0.5); x = gaussian(0.5 + 2 mu, 1); return mu + x  ->  3.5
    std::vector<int32_t> k2(2, QSB_ERP_GAUSSIAN), pref = {-1, -1, 0, -1};
    std::vector<double> p0 = {1.0, 0.5}, p1 = {0.5, 1.0}, pcoef = {0.0, 0.0, 2.0, 0.0};
    qs::Program hier = qs::Program::Indep(k2.data(), p0.data(), p1.data(), 2, true, 1.0).LatentParams(pref.data(), pcoef.data());
    auto run2 = qs::infer(hier, [](const qs::SampleVector& s) { double m = 0; for (double e : qs::Expectation(s)) m += e; return m / s.numchains; },
                          qs::MCMC(qs::HMCKernel(hp), mp));
    double est2 = run2();
    std::printf("estimate %.4f (truth 3.5)\n", est2);
    return std::fabs(est2 - 3.5) <= 0.15 ? 0 : 1;
  } catch (const qs::Error& e) { std::fprintf(stderr, "%s\n", e.what()); return 2; }
}
