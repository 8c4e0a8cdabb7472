// Streaming statistics of kept draws (QSB_FLAG_MOMENTS), shared by the transition kernels and the diagnostics reduction.
//
// Per (chain, return-value dimension) running statistics of the kept draws, so that split-R-hat, pooled moments and an
// effective sample size need no stored draws (config c4: one kept sweep of all chains is 1 GB per GPU).  Seven doubles:
// Welford mean / M2 of the first half of the run (kept-sample index < half) and of the second half (split-R-hat), the running
// sum of the current batch and Welford mean / M2 of the completed batch means (batch size `batch`: the batch-means estimate
// of the asymptotic variance, hence the ESS).  slot(stat) points at this (chain, dim)'s entry of statistic `stat`; the
// arrays are laid out like the kept samples ([stat][dim][chain] thread-per-chain, [stat][chain][dim] otherwise).
#pragma once
#include <cstdint>

namespace qsb {

static constexpr int MOM_STATS = 7;
template <class Slot>
__device__ __forceinline__ void mom_update(Slot slot, uint64_t sidx, uint64_t half, uint32_t batch, double v) {
  const int h = sidx >= half ? 2 : 0;
  const double n = (double)(sidx >= half ? sidx - half + 1 : sidx + 1);
  double* pm = slot(h); double* p2 = slot(h + 1);
  const double mean = *pm, dlt = v - mean, nm = mean + dlt / n;
  *pm = nm; *p2 = *p2 + dlt * (v - nm);
  double* ps = slot(4);
  const double bs = *ps + v;
  if ((sidx + 1) % batch == 0) {
    const double bm = bs / (double)batch, nb = (double)((sidx + 1) / batch);
    double* qm = slot(5); double* q2 = slot(6);
    const double m0 = *qm, d2 = bm - m0, m1 = m0 + d2 / nb;
    *qm = m1; *q2 = *q2 + d2 * (bm - m1);
    *ps = 0.0;
  } else *ps = bs;
}

}  // namespace qsb
