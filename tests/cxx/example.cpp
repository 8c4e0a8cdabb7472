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
    return std::fabs(est - 0.1) <= 0.07 ? 0 : 1;
  } catch (const qs::Error& e) { std::fprintf(stderr, "%s\n", e.what()); return 2; }
}
