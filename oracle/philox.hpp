// ORACLE (test infrastructure, never shipped, never on the product path).
//
// The injected uniform source that stands in for the reference's
//   R.random = rand()/(RAND_MAX+1.0)                    (qs/lib/random.t:5,11)
// Every sampler in the reference draws from that single function; here the
// same call, random(), returns the next uniform of a Philox4x32-10 sub-stream
// addressed by (seed, chain, iter, slot).  The CUDA engine implements the same
// contract (quicksand_b200/csrc/philox.cuh), so CPU and GPU chains see the same
// numbers regardless of how chains or dimensions are spread over threads.
//
// Stream contract (shared by oracle and device):
//   key     = (seed_lo32, seed_hi32)
//   counter = (block, slot, iter, chain)          block = 0,1,2,... within the sub-stream
//   words   = Philox4x32-10(counter, key) = (w0,w1,w2,w3)
//   uniform(2*block)   = ((w1<<32 | w0) >> 11) * 2^-53
//   uniform(2*block+1) = ((w3<<32 | w2) >> 11) * 2^-53        both in [0,1)
// One Philox block is exactly the (u, v) pair of one attempt of the Gaussian
// ratio-of-uniforms loop (qs/distrib.t:66-72).
//
// Slot / iter assignment (who calls random() where):
//   iter = MCMC iteration i of doMCMC (qs/mcmc.t:57-63), or a phase code:
//     ITER_INIT            forward-sampling initialisation (qs/trace.t:612-623), slot = ERP index
//     ITER_SEARCH + probe  probe-th momentum refresh of searchForInitialStepSize (qs/hmc.t:345-391)
//   slot within an iteration with n real components:
//     HMC   : momentum i -> i ; accept uniform -> n
//     Drift : proposal i -> i ; accept uniform -> n
//     HARM  : direction i -> i ; step length uniform -> n ; accept uniform -> n+1
//     Mixture kernel choice (qs/mcmc.t:276) -> SLOT_MIXTURE
#pragma once
#include <cstdint>

namespace qso {

static const uint32_t ITER_INIT = 0xFFFFFFFFu;
static const uint32_t ITER_SEARCH = 0xF0000000u;
static const uint32_t SLOT_MIXTURE = 0xFFFFFFFFu;
// TraceMHKernel (qs/mcmc.t:134-193): which choice, the proposal's sampler (a sequential sub-stream), the accept test
static const uint32_t SLOT_MH_INDEX = 0xFFFFFFFEu, SLOT_MH_PROPOSAL = 0xFFFFFFFDu, SLOT_MH_ACCEPT = 0xFFFFFFFCu;

inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline double u53(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 11) * 0x1.0p-53;
}

// The process-global generator state, like libc's rand() state in the reference.
struct Rng {
  uint64_t seed = 0;
  uint32_t chain = 0, iter = 0, slot = 0;
  uint32_t k = 0;        // index of the next uniform inside the (chain, iter, slot) sub-stream
  uint64_t ndraws = 0;   // statistics only
  void at(uint32_t iter_, uint32_t slot_) { iter = iter_; slot = slot_; k = 0; }
  double random() {
    uint32_t ctr[4] = {k >> 1, slot, iter, chain};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[4];
    philox4x32_10(ctr, key, w);
    double u = (k & 1) ? u53(w[2], w[3]) : u53(w[0], w[1]);
    ++k; ++ndraws;
    return u;
  }
};

inline Rng& rng() { static thread_local Rng g; return g; }
inline double random_() { return rng().random(); }   // R.random (qs/lib/random.t:11)

}  // namespace qso
