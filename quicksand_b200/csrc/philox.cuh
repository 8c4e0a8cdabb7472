// Counter-based random stream of the engine: Philox4x32-10 in place of the reference's single
// uniform source  R.random = rand()/(RAND_MAX+1.0)  (qs/lib/random.t:5,11).
//
// Stream contract:
//   key     = (seed_lo32, seed_hi32)
//   counter = (block, slot, iter, chain)       block = 0,1,2,... inside the sub-stream
//   (w0,w1,w2,w3) = Philox4x32-10(counter, key)
//   uniform(2*block) = ((w1<<32|w0) >> 11) * 2^-53,  uniform(2*block+1) = ((w3<<32|w2) >> 11) * 2^-53
// A sub-stream (chain, iter, slot) belongs to one sampler call site of one MCMC iteration:
//   iter  = MCMC iteration (qs/mcmc.t:57-63), QSB_ITER_INIT for the forward-sampling initialisation
//           (qs/trace.t:612-623; slot = ERP index), QSB_ITER_SEARCH+probe for the probe-th momentum
//           refresh of searchForInitialStepSize (qs/hmc.t:345-391)
//   slot  = HMC: momentum i -> i, accept -> n;  Drift: proposal i -> i, accept -> n;
//           HARM: direction i -> i, step length -> n, accept -> n+1;  mixture choice -> QSB_SLOT_MIXTURE
// Because `chain` is the GLOBAL chain id, a chain's numbers do not depend on how chains are spread
// over threads, warps, CTAs or GPUs; because each component has its own slot, the n Gaussians of an
// iteration can be drawn by n different lanes.
#pragma once
#include <cstdint>

namespace qsb {

static constexpr uint32_t QSB_ITER_INIT = 0xFFFFFFFFu;
static constexpr uint32_t QSB_ITER_SEARCH = 0xF0000000u;
static constexpr uint32_t QSB_SLOT_MIXTURE = 0xFFFFFFFFu;
// TraceMHKernel (qs/mcmc.t:134-193): which choice, the proposal's sampler (a sequential sub-stream), the accept test
static constexpr uint32_t QSB_SLOT_MH_INDEX = 0xFFFFFFFEu, QSB_SLOT_MH_PROPOSAL = 0xFFFFFFFDu, QSB_SLOT_MH_ACCEPT = 0xFFFFFFFCu;

struct RngKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t (&w)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

__host__ __device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * 0x1.0p-53;
}

// Sequential reader of one sub-stream: random() returns uniform(0), uniform(1), ...
struct SubStream {
  RngKey key;
  uint32_t chain, iter, slot, k;
  uint32_t w2, w3;
  __device__ __forceinline__ SubStream(RngKey key_, uint32_t chain_, uint32_t iter_, uint32_t slot_)
      : key(key_), chain(chain_), iter(iter_), slot(slot_), k(0), w2(0), w3(0) {}
  __device__ __forceinline__ double random() {
    double u;
    if ((k & 1u) == 0u) {
      uint32_t w[4];
      philox4x32_10(k >> 1, slot, iter, chain, key.k0, key.k1, w);
      w2 = w[2]; w3 = w[3];
      u = u53(w[0], w[1]);
    } else {
      u = u53(w2, w3);
    }
    ++k;
    return u;
  }
};

// One uniform: the first number of sub-stream (chain, iter, slot).
__device__ __forceinline__ double uniform_at(RngKey key, uint32_t chain, uint32_t iter, uint32_t slot) {
  uint32_t w[4];
  philox4x32_10(0u, slot, iter, chain, key.k0, key.k1, w);
  return u53(w[0], w[1]);
}

// gaussian.sample(0,1) of qs/distrib.t:64-74 (Leva's ratio-of-uniforms) reading sub-stream
// (chain, iter, slot) from its start: attempt a consumes exactly Philox block a (u = 1 - uniform(2a),
// v from uniform(2a+1)).  Returns v/u; the caller forms mu + sigma*v/u in the reference's order.
// Products and sums of the acceptance test are issued un-contracted (no FMA) so the accept/reject of
// every attempt, and therefore the draw, is bit-identical to a CPU evaluation of the same expressions.
// not inlined: called once per component per iteration; inlining the 10-round Philox at every unrolled component
// multiplied the kernels' code size past the instruction cache
static __device__ __noinline__ void leva_vu(RngKey key, uint32_t chain, uint32_t iter, uint32_t slot, double& v_out, double& u_out) {
  double u, v;
  for (uint32_t blk = 0;; ++blk) {
    uint32_t w[4];
    philox4x32_10(blk, slot, iter, chain, key.k0, key.k1, w);
    u = __dsub_rn(1.0, u53(w[0], w[1]));
    v = __dmul_rn(1.7156, __dsub_rn(u53(w[2], w[3]), 0.5));
    double x = __dsub_rn(u, 0.449871);
    double y = __dadd_rn(fabs(v), 0.386595);
    double q = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, __dsub_rn(__dmul_rn(0.196, y), __dmul_rn(0.25472, x))));
    if (!(q >= 0.27597)) break;
    if (q > 0.27846) continue;
    double lhs = __dmul_rn(v, v);
    double rhs = __dmul_rn(__dmul_rn(__dmul_rn(-4.0, u), u), log(u));
    if (!(lhs > rhs)) break;
  }
  v_out = v; u_out = u;
}
// One attempt (= one Philox block) of the same loop: true when the pair (v, u) is accepted.
static __device__ __noinline__ bool leva_attempt(RngKey key, uint32_t chain, uint32_t iter, uint32_t slot, uint32_t blk, double& v_out, double& u_out) {
  uint32_t w[4];
  philox4x32_10(blk, slot, iter, chain, key.k0, key.k1, w);
  const double u = __dsub_rn(1.0, u53(w[0], w[1]));
  const double v = __dmul_rn(1.7156, __dsub_rn(u53(w[2], w[3]), 0.5));
  v_out = v; u_out = u;
  const double x = __dsub_rn(u, 0.449871);
  const double y = __dadd_rn(fabs(v), 0.386595);
  const double q = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, __dsub_rn(__dmul_rn(0.196, y), __dmul_rn(0.25472, x))));
  if (!(q >= 0.27597)) return true;
  if (q > 0.27846) return false;
  return !(__dmul_rn(v, v) > __dmul_rn(__dmul_rn(__dmul_rn(-4.0, u), u), log(u)));
}
__device__ __forceinline__ double gaussian_std(RngKey key, uint32_t chain, uint32_t iter, uint32_t slot) {
  double v, u;
  leva_vu(key, chain, iter, slot, v, u);
  return v / u;
}
// gaussian.sample(mu, sigma) = mu + sigma*v/u, evaluated left to right like the reference (distrib.t:73)
__device__ __forceinline__ double gaussian_sample(RngKey key, uint32_t chain, uint32_t iter, uint32_t slot, double mu, double sigma) {
  double v, u;
  leva_vu(key, chain, iter, slot, v, u);
  return __dadd_rn(mu, __dmul_rn(sigma, v) / u);
}

// The same sampler on a sequential sub-stream (forward sampling of gamma/beta draws several
// Gaussians and uniforms from one sub-stream: qs/distrib.t:109-128, 150-153).
__device__ __forceinline__ double gaussian_sample_seq(SubStream& s, double mu, double sigma) {
  double u, v;
  while (true) {
    u = __dsub_rn(1.0, s.random());
    v = __dmul_rn(1.7156, __dsub_rn(s.random(), 0.5));
    double x = __dsub_rn(u, 0.449871);
    double y = __dadd_rn(fabs(v), 0.386595);
    double q = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, __dsub_rn(__dmul_rn(0.196, y), __dmul_rn(0.25472, x))));
    if (!(q >= 0.27597)) break;
    if (q > 0.27846) continue;
    if (!(__dmul_rn(v, v) > __dmul_rn(__dmul_rn(__dmul_rn(-4.0, u), u), log(u)))) break;
  }
  return __dadd_rn(mu, __dmul_rn(sigma, v) / u);
}

}  // namespace qsb
