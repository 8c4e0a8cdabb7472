// Instantiation helper: one EngineVariant per (family, lanes per chain, elements per lane).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include "engine.cuh"

namespace qsb {

template <int FAM, int G, int NPL> struct VariantImpl {
  static constexpr int TPB = G == 1 ? 128 : 256;   // (the one-off kernels; any G > 1 runs them at 256)
  static unsigned grid(const EngineParams& P) { return (unsigned)(((size_t)P.C * G + TPB - 1) / TPB); }
  template <bool HAS_HMC, bool LEX = false, bool DRIFT_ONLY = false> static void advance_t(const EngineParams& P, void* st) {
    constexpr int TPBA = (G < 32 || HAS_HMC) ? 128 : 256;   // must equal advance_kernel's launch bounds
    constexpr size_t SMEM = ((size_t)Chain<FAM, G, NPL>::NB * TPBA + (FAM == FAM_GAUSS_PREC ? G * NPL * G * NPL : 0)) * sizeof(double);   // Gaussian staging columns (+ S)
    static bool attr[64] = {};   // the attribute is per device: a process may drive several (qsb_run_config.device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (SMEM > 48 * 1024 && !attr[dev & 63]) {
      cudaFuncSetAttribute(advance_kernel<FAM, G, NPL, HAS_HMC, LEX, DRIFT_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      attr[dev & 63] = true;
    }
    // Thread per chain with less than one wave of blocks (config c2: 65 536 chains = 512 blocks of 128 on 148 SMs x 4 slots): the
    // launch takes as long as the SM with the most warps, 4 blocks = 16 warps where the mean is 13.8.  Smaller blocks spread the
    // same warps evenly (14 on the fullest SM with blocks of 64); every shared-memory offset in the kernel is relative to blockDim.
    int tpb = TPBA;
    if (G == 1 && !LEX) {
      int nsm = 148;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
      size_t best = ~(size_t)0;
      for (int b : {128, 64, 32}) {
        const size_t blocks = ((size_t)P.C + b - 1) / b, worst = ((blocks + nsm - 1) / nsm) * (size_t)(b / 32);
        if (worst < best && blocks <= (size_t)nsm * 16) { best = worst; tpb = b; }
      }
    }
    const size_t smem = SMEM - (size_t)Chain<FAM, G, NPL>::NB * (TPBA - tpb) * sizeof(double);
    advance_kernel<FAM, G, NPL, HAS_HMC, LEX, DRIFT_ONLY><<<(unsigned)(((size_t)P.C * G + tpb - 1) / tpb), tpb, smem, (cudaStream_t)st>>>(P);
  }
  static void advance(const EngineParams& P, void* st) { advance_t<true>(P, st); }
  static void advance_nohmc(const EngineParams& P, void* st) { advance_t<false>(P, st); }
  static void advance_lex(const EngineParams& P, void* st) { advance_t<true, true>(P, st); }
  static void advance_drift(const EngineParams& P, void* st) { advance_t<false, false, true>(P, st); }
  static void init(const EngineParams& P, void* st) { init_kernel<FAM, G, NPL><<<grid(P), TPB, 0, (cudaStream_t)st>>>(P); }
  static void rescore(const EngineParams& P, void* st) { rescore_kernel<FAM, G, NPL><<<grid(P), TPB, 0, (cudaStream_t)st>>>(P); }
  static void logp_grad(const EngineParams& P, void* st) { logp_grad_kernel<FAM, G, NPL><<<grid(P), TPB, 0, (cudaStream_t)st>>>(P); }
  // warp-specialised kernel: persistent grid of CTAs of 4 trajectory warps + 4 * NPROD sampler warps
  template <bool HAS_HMC, bool DRIFT_ONLY> static void advance_ws_t(const EngineParams& P, void* st) {
    static const int nprod = getenv("QSB_WS_NPROD") ? atoi(getenv("QSB_WS_NPROD")) : 2;   // sampler warpgroups per trajectory warpgroup
    if (nprod == 1) advance_ws_n<HAS_HMC, DRIFT_ONLY, 1>(P, st); else advance_ws_n<HAS_HMC, DRIFT_ONLY, 2>(P, st);
  }
  template <bool HAS_HMC, bool DRIFT_ONLY, int NPROD> static void advance_ws_n(const EngineParams& P, void* st) {
    if constexpr (G == 32) {
      constexpr int THREADS = 128 * (1 + NPROD);
      constexpr size_t SMEM = (size_t)4 * (HAS_HMC ? 3 : 2) * NPL * 32 * sizeof(double) + 16 * sizeof(uint64_t);
      auto kern = advance_ws_kernel<FAM, NPL, HAS_HMC, DRIFT_ONLY, NPROD>;
      static int ctas_per_sm[64] = {};   // per device
      int dev = 0, nsm = 148;
      cudaGetDevice(&dev);
      if (!ctas_per_sm[dev & 63]) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, THREADS, SMEM);
        ctas_per_sm[dev & 63] = n > 0 ? n : 1;
      }
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
      const unsigned want = (unsigned)((P.C + 3) / 4);
      const unsigned grid = std::min<unsigned>(want, (unsigned)(nsm * ctas_per_sm[dev & 63]));
      kern<<<grid, THREADS, SMEM, (cudaStream_t)st>>>(P);
    }
  }
  static void advance_ws(const EngineParams& P, void* st) { advance_ws_t<true, false>(P, st); }
  static void advance_nohmc_ws(const EngineParams& P, void* st) { advance_ws_t<false, false>(P, st); }
  static void advance_drift_ws(const EngineParams& P, void* st) { advance_ws_t<false, true>(P, st); }
  static void debug_leap(const EngineParams& P, const DebugLeap& D, void* st) {
    constexpr size_t SMEM = FAM == FAM_GAUSS_PREC ? ((size_t)Chain<FAM, G, NPL>::NB * TPB + G * NPL * G * NPL) * sizeof(double) : 0;
    debug_leap_kernel<FAM, G, NPL><<<grid(P), TPB, SMEM, (cudaStream_t)st>>>(P, D);
  }
  static EngineVariant make() { return EngineVariant{FAM, G, NPL, &advance, &advance_nohmc, &advance_lex, &advance_drift, &init, &rescore, &logp_grad, &debug_leap,
                                                   G == 32 ? &advance_ws : nullptr, G == 32 ? &advance_nohmc_ws : nullptr, G == 32 ? &advance_drift_ws : nullptr}; }
};

}  // namespace qsb
