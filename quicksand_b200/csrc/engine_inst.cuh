// Instantiation helper: one EngineVariant per (family, lanes per chain, elements per lane).
#pragma once
#include <cuda_runtime.h>
#include "engine.cuh"

namespace qsb {

template <int FAM, int G, int NPL> struct VariantImpl {
  static constexpr int TPB = G == 1 ? 128 : 256;
  static unsigned grid(const EngineParams& P) { return (unsigned)(((size_t)P.C * G + TPB - 1) / TPB); }
  static void advance(const EngineParams& P, void* st) { advance_kernel<FAM, G, NPL><<<grid(P), TPB, 0, (cudaStream_t)st>>>(P); }
  static void init(const EngineParams& P, void* st) { init_kernel<FAM, G, NPL><<<grid(P), TPB, 0, (cudaStream_t)st>>>(P); }
  static void rescore(const EngineParams& P, void* st) { rescore_kernel<FAM, G, NPL><<<grid(P), TPB, 0, (cudaStream_t)st>>>(P); }
  static void logp_grad(const EngineParams& P, void* st) { logp_grad_kernel<FAM, G, NPL><<<grid(P), TPB, 0, (cudaStream_t)st>>>(P); }
  static EngineVariant make() { return EngineVariant{FAM, G, NPL, &advance, &init, &rescore, &logp_grad}; }
};

}  // namespace qsb
