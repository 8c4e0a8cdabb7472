// Bernoulli-logit GLM under HMC (config c3): theta_j ~ gaussian(0, s), flip.observe(y_i, 1/(1+exp(-x_i.theta))).
//
// What the reference does per gradient (qs/hmc.t:325-342): re-run the whole program on tape-AD numbers --
// N*d multiply-adds each allocating a node -- then sweep the tape backwards (qs/lib/ad.t:679-708).
// Here one gradient of ALL chains is a pair of dense fp64 contractions on the tensor cores (DMMA,
// mma.sync.m8n8k4.f64) fused around the sigmoid:
//     Z = X Theta  (N x C),   R = y - 1/(1+exp(-Z)),   G = X^T R  (d x C),   lp = sum log(y ? p : 1-p)
// Z and R never leave registers.  Orientation: chains are the M dimension of both MMAs
// (Z^T = Theta^T X^T, G^T = R^T X), so the accumulator fragment of the first MMA *is* the A fragment
// of the second (thread (g,t) holds chain g, two observation rows) -- no shuffles, no shared-memory
// round trip.  The k-index <-> dimension map of the first MMA and the k-index <-> row map of the
// second are chosen so that every thread only ever touches the 2*NJ dimensions it owns
// (dims 8J+2t, 8J+2t+1 of chain g) and all shared-memory fragment loads are bank-conflict free.
//
// X (N x d, row stride S) is streamed from L2 through shared memory with TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx) into a ring of 2..4 stages.  There is no producer warp: the warp whose release of a stage is the last one
// pending re-arms it.  The default build runs 12 warps x 8 chains per CTA (96 chains read the same staged rows, so a row
// block is fetched once per 96 chains) with the Theta of a warp's 8 chains in registers (168 registers); QSB_GLM_CW = 16 /
// QSB_GLM_THETA_SMEM select the measured-slower variants with the chain tile's Theta in shared memory.
// Work decomposition: by default the (chain tile, row block) units are cut into one cost-balanced run per SM (grid = 148,
// one wave for any chain count, host-computed run table); QSB_FLAG_INVARIANT selects the fixed grid (chain tiles) x (row
// splits) whose grouping of row blocks into partial sums does not depend on the chain tile, so that a chain's draws are
// bit-identical for any partition of the chains.  Partial gradients are summed in a fixed order by the vector kernel.
//
// The leapfrog updates, momenta, Hamiltonians, accept/reject, dual averaging and sample collection
// are the warp-per-chain vector kernels below (qs/hmc.t:134-186, 289-323, 394-402).
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/qsb200.h"
#include "distrib.cuh"
#include "glm_hmc.cuh"
#include "moments.cuh"
#include "mbar.cuh"
#include "philox.cuh"

namespace qsb {

static constexpr int MB = 32;          // observation rows per pipeline stage
static constexpr int NT = MB / 8;      // 8-row n-tiles per stage
static constexpr int TAB_OFF = 2 * 4 + 2;       // behind the ring: full[STAGES], empty[STAGES] mbarriers and the stream-K cursor ...
static constexpr int STAGES = 4;        // deepest ring; a sampler uses 2..4 stages (GradParams::stages), whatever fits 227 KB
static constexpr int TAB_N = 256;                // entries of the table 2^(j/256) of the sigmoid's exponential (2 KB) ...
static constexpr int BAR_DOUBLES = TAB_OFF + TAB_N;   // ... which sits behind them
// 2^(j/256), j = 0..255, correctly rounded (generated with 200-bit arithmetic)
__constant__ double c_exp2_tab[TAB_N] = {
    0x1.0000000000000p+0, 0x1.00b1afa5abcbfp+0, 0x1.0163da9fb3335p+0, 0x1.02168143b0281p+0, 0x1.02c9a3e778061p+0, 0x1.037d42e11bbccp+0,
    0x1.04315e86e7f85p+0, 0x1.04e5f72f654b1p+0, 0x1.059b0d3158574p+0, 0x1.0650a0e3c1f89p+0, 0x1.0706b29ddf6dep+0, 0x1.07bd42b72a836p+0,
    0x1.0874518759bc8p+0, 0x1.092bdf66607e0p+0, 0x1.09e3ecac6f383p+0, 0x1.0a9c79b1f3919p+0, 0x1.0b5586cf9890fp+0, 0x1.0c0f145e46c85p+0,
    0x1.0cc922b7247f7p+0, 0x1.0d83b23395decp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.0efa55fdfa9c5p+0, 0x1.0fb66affed31bp+0, 0x1.1073028d7233ep+0,
    0x1.11301d0125b51p+0, 0x1.11edbab5e2ab6p+0, 0x1.12abdc06c31ccp+0, 0x1.136a814f204abp+0, 0x1.1429aaea92de0p+0, 0x1.14e95934f312ep+0,
    0x1.15a98c8a58e51p+0, 0x1.166a45471c3c2p+0, 0x1.172b83c7d517bp+0, 0x1.17ed48695bbc0p+0, 0x1.18af9388c8deap+0, 0x1.1972658375d2fp+0,
    0x1.1a35beb6fcb75p+0, 0x1.1af99f8138a1cp+0, 0x1.1bbe084045cd4p+0, 0x1.1c82f95281c6bp+0, 0x1.1d4873168b9aap+0, 0x1.1e0e75eb44027p+0,
    0x1.1ed5022fcd91dp+0, 0x1.1f9c18438ce4dp+0, 0x1.2063b88628cd6p+0, 0x1.212be3578a819p+0, 0x1.21f49917ddc96p+0, 0x1.22bdda27912d1p+0,
    0x1.2387a6e756238p+0, 0x1.2451ffb82140ap+0, 0x1.251ce4fb2a63fp+0, 0x1.25e85711ece75p+0, 0x1.26b4565e27cddp+0, 0x1.2780e341ddf29p+0,
    0x1.284dfe1f56381p+0, 0x1.291ba7591bb70p+0, 0x1.29e9df51fdee1p+0, 0x1.2ab8a66d10f13p+0, 0x1.2b87fd0dad990p+0, 0x1.2c57e39771b2fp+0,
    0x1.2d285a6e4030bp+0, 0x1.2df961f641589p+0, 0x1.2ecafa93e2f56p+0, 0x1.2f9d24abd886bp+0, 0x1.306fe0a31b715p+0, 0x1.31432edeeb2fdp+0,
    0x1.32170fc4cd831p+0, 0x1.32eb83ba8ea32p+0, 0x1.33c08b26416ffp+0, 0x1.3496266e3fa2dp+0, 0x1.356c55f929ff1p+0, 0x1.36431a2de883bp+0,
    0x1.371a7373aa9cbp+0, 0x1.37f26231e754ap+0, 0x1.38cae6d05d866p+0, 0x1.39a401b7140efp+0, 0x1.3a7db34e59ff7p+0, 0x1.3b57fbfec6cf4p+0,
    0x1.3c32dc313a8e5p+0, 0x1.3d0e544ede173p+0, 0x1.3dea64c123422p+0, 0x1.3ec70df1c5175p+0, 0x1.3fa4504ac801cp+0, 0x1.40822c367a024p+0,
    0x1.4160a21f72e2ap+0, 0x1.423fb2709468ap+0, 0x1.431f5d950a897p+0, 0x1.43ffa3f84b9d4p+0, 0x1.44e086061892dp+0, 0x1.45c2042a7d232p+0,
    0x1.46a41ed1d0057p+0, 0x1.4786d668b3237p+0, 0x1.486a2b5c13cd0p+0, 0x1.494e1e192aed2p+0, 0x1.4a32af0d7d3dep+0, 0x1.4b17dea6db7d7p+0,
    0x1.4bfdad5362a27p+0, 0x1.4ce41b817c114p+0, 0x1.4dcb299fddd0dp+0, 0x1.4eb2d81d8abffp+0, 0x1.4f9b2769d2ca7p+0, 0x1.508417f4531eep+0,
    0x1.516daa2cf6642p+0, 0x1.5257de83f4eefp+0, 0x1.5342b569d4f82p+0, 0x1.542e2f4f6ad27p+0, 0x1.551a4ca5d920fp+0, 0x1.56070dde910d2p+0,
    0x1.56f4736b527dap+0, 0x1.57e27dbe2c4cfp+0, 0x1.58d12d497c7fdp+0, 0x1.59c0827ff07ccp+0, 0x1.5ab07dd485429p+0, 0x1.5ba11fba87a03p+0,
    0x1.5c9268a5946b7p+0, 0x1.5d84590998b93p+0, 0x1.5e76f15ad2148p+0, 0x1.5f6a320dceb71p+0, 0x1.605e1b976dc09p+0, 0x1.6152ae6cdf6f4p+0,
    0x1.6247eb03a5585p+0, 0x1.633dd1d1929fdp+0, 0x1.6434634ccc320p+0, 0x1.652b9febc8fb7p+0, 0x1.6623882552225p+0, 0x1.671c1c70833f6p+0,
    0x1.68155d44ca973p+0, 0x1.690f4b19e9538p+0, 0x1.6a09e667f3bcdp+0, 0x1.6b052fa75173ep+0, 0x1.6c012750bdabfp+0, 0x1.6cfdcddd47645p+0,
    0x1.6dfb23c651a2fp+0, 0x1.6ef9298593ae5p+0, 0x1.6ff7df9519484p+0, 0x1.70f7466f42e87p+0, 0x1.71f75e8ec5f74p+0, 0x1.72f8286ead08ap+0,
    0x1.73f9a48a58174p+0, 0x1.74fbd35d7cbfdp+0, 0x1.75feb564267c9p+0, 0x1.77024b1ab6e09p+0, 0x1.780694fde5d3fp+0, 0x1.790b938ac1cf6p+0,
    0x1.7a11473eb0187p+0, 0x1.7b17b0976cfdbp+0, 0x1.7c1ed0130c132p+0, 0x1.7d26a62ff86f0p+0, 0x1.7e2f336cf4e62p+0, 0x1.7f3878491c491p+0,
    0x1.80427543e1a12p+0, 0x1.814d2add106d9p+0, 0x1.82589994cce13p+0, 0x1.8364c1eb941f7p+0, 0x1.8471a4623c7adp+0, 0x1.857f4179f5b21p+0,
    0x1.868d99b4492edp+0, 0x1.879cad931a436p+0, 0x1.88ac7d98a6699p+0, 0x1.89bd0a478580fp+0, 0x1.8ace5422aa0dbp+0, 0x1.8be05bad61778p+0,
    0x1.8cf3216b5448cp+0, 0x1.8e06a5e0866d9p+0, 0x1.8f1ae99157736p+0, 0x1.902fed0282c8ap+0, 0x1.9145b0b91ffc6p+0, 0x1.925c353aa2fe2p+0,
    0x1.93737b0cdc5e5p+0, 0x1.948b82b5f98e5p+0, 0x1.95a44cbc8520fp+0, 0x1.96bdd9a7670b3p+0, 0x1.97d829fde4e50p+0, 0x1.98f33e47a22a2p+0,
    0x1.9a0f170ca07bap+0, 0x1.9b2bb4d53fe0dp+0, 0x1.9c49182a3f090p+0, 0x1.9d674194bb8d5p+0, 0x1.9e86319e32323p+0, 0x1.9fa5e8d07f29ep+0,
    0x1.a0c667b5de565p+0, 0x1.a1e7aed8eb8bbp+0, 0x1.a309bec4a2d33p+0, 0x1.a42c980460ad8p+0, 0x1.a5503b23e255dp+0, 0x1.a674a8af46052p+0,
    0x1.a799e1330b358p+0, 0x1.a8bfe53c12e59p+0, 0x1.a9e6b5579fdbfp+0, 0x1.ab0e521356ebap+0, 0x1.ac36bbfd3f37ap+0, 0x1.ad5ff3a3c2774p+0,
    0x1.ae89f995ad3adp+0, 0x1.afb4ce622f2ffp+0, 0x1.b0e07298db666p+0, 0x1.b20ce6c9a8952p+0, 0x1.b33a2b84f15fbp+0, 0x1.b468415b749b1p+0,
    0x1.b59728de5593ap+0, 0x1.b6c6e29f1c52ap+0, 0x1.b7f76f2fb5e47p+0, 0x1.b928cf22749e4p+0, 0x1.ba5b030a1064ap+0, 0x1.bb8e0b79a6f1fp+0,
    0x1.bcc1e904bc1d2p+0, 0x1.bdf69c3f3a207p+0, 0x1.bf2c25bd71e09p+0, 0x1.c06286141b33dp+0, 0x1.c199bdd85529cp+0, 0x1.c2d1cd9fa652cp+0,
    0x1.c40ab5fffd07ap+0, 0x1.c544778fafb22p+0, 0x1.c67f12e57d14bp+0, 0x1.c7ba88988c933p+0, 0x1.c8f6d9406e7b5p+0, 0x1.ca3405751c4dbp+0,
    0x1.cb720dcef9069p+0, 0x1.ccb0f2e6d1675p+0, 0x1.cdf0b555dc3fap+0, 0x1.cf3155b5bab74p+0, 0x1.d072d4a07897cp+0, 0x1.d1b532b08c968p+0,
    0x1.d2f87080d89f2p+0, 0x1.d43c8eacaa1d6p+0, 0x1.d5818dcfba487p+0, 0x1.d6c76e862e6d3p+0, 0x1.d80e316c98398p+0, 0x1.d955d71ff6075p+0,
    0x1.da9e603db3285p+0, 0x1.dbe7cd63a8315p+0, 0x1.dd321f301b460p+0, 0x1.de7d5641c0658p+0, 0x1.dfc97337b9b5fp+0, 0x1.e11676b197d17p+0,
    0x1.e264614f5a129p+0, 0x1.e3b333b16ee12p+0, 0x1.e502ee78b3ff6p+0, 0x1.e653924676d76p+0, 0x1.e7a51fbc74c83p+0, 0x1.e8f7977cdb740p+0,
    0x1.ea4afa2a490dap+0, 0x1.eb9f4867cca6ep+0, 0x1.ecf482d8e67f1p+0, 0x1.ee4aaa2188510p+0, 0x1.efa1bee615a27p+0, 0x1.f0f9c1cb6412ap+0,
    0x1.f252b376bba97p+0, 0x1.f3ac948dd7274p+0, 0x1.f50765b6e4540p+0, 0x1.f6632798844f8p+0, 0x1.f7bfdad9cbe14p+0, 0x1.f91d802243c89p+0,
    0x1.fa7c1819e90d8p+0, 0x1.fbdba3692d514p+0, 0x1.fd3c22b8f71f1p+0, 0x1.fe9d96b2a23d9p+0};
#ifndef QSB_GLM_CW
#define QSB_GLM_CW 12
#endif
// warps per CTA.  12 (default): Theta of the chain tile in registers (3 warps per SM sub-partition at 168 registers).
// 16: Theta in shared memory (the ring leaves 124 KB unused), 4 warps per sub-partition at 128 registers -- 1 % slower in
// a same-box A/B (1.183 vs 1.171 ms per launch on c3), as are 12 warps with Theta in shared memory (1.181 ms); 10 and 14
// warps leave sub-partitions unevenly loaded (1.41 / 1.35 ms).  Kept selectable for other shapes.
static constexpr int CW = QSB_GLM_CW;
#ifndef QSB_GLM_THETA_SMEM
#define QSB_GLM_THETA_SMEM (QSB_GLM_CW == 16)
#endif
static constexpr bool THETA_SMEM = QSB_GLM_THETA_SMEM;
static constexpr int CB = CW * 8;      // chains per CTA
static constexpr int GRAD_THREADS = CW * 32;
// Column-permuted X for config c3's shape (the TRIM instantiation: 13 dimension tiles, d <= 100).  Logical dimension 8J + n sits
// in physical column n * PERM_PJ + J, so the dimensions a thread needs from consecutive tiles J, J+1 are adjacent in shared
// memory: the B fragments of BOTH contractions come in as LDS.128 (two tiles per load) -- the second contraction had one
// LDS.64 per DMMA (104 per row block) in the natural layout.  Row stride PERM_S = 114 doubles and the identity row map make
// both load patterns conflict-free per quarter warp (checked exhaustively: tools in DESIGN.md 5.2).  0 = natural layout.
#ifndef QSB_GLM_PERM
#define QSB_GLM_PERM 1
#endif
static constexpr bool PERMX = QSB_GLM_PERM != 0;
static constexpr int PERM_PJ = 14, PERM_S = 114;

struct GradParams {
  const double* Xp; const double* yv;
  int S, d, nj, nblocks, ksplit, ntiles;
  int sk_U;                // > 0: balanced ("stream-K") decomposition, see glm_grad_kernel (the value is the number of CTAs);
                           // 0: grid = chain tiles x row splits
  const int* sk_start;     // [sk_U + 1] first row-block unit (tile-major) of every CTA's run
  const int* sk_first;     // [ntiles] first CTA whose run touches the tile (partial-sum slot = CTA - that)
  uint32_t C; int Dp;
  const double* theta;     // [C][Dp]
  double* partG;           // [ksplit][C][Dp]
  double* partLp;          // [ksplit][C]
  int want_lp;
  int stages;              // depth of the X ring in shared memory (2..STAGES)
  int want_grad;           // 0: only the first contraction and the log-likelihood (Drift / HARM score a proposal, no gradient)
};

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): every kernel of a transition is launched with programmatic stream serialization, lets
// its successor start launching at once (pdl_trigger) and touches nothing its predecessor produces before pdl_wait, which
// returns when the predecessor grid has completed and its writes are visible.  The ~2-3 us launch gap per kernel boundary
// (41 boundaries per transition) and the gradient kernel's prologue then overlap the tail of the previous kernel.  Without
// the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  // volatile: keeps the issue order written below (independent accumulators back to back), nothing else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// smem: per stage [MB*S + 8 doubles X (8 = zeroed over-read pad)] [MB doubles y]; then 2*STAGES mbarriers
__host__ __device__ inline size_t stage_doubles(int S) { return (size_t)MB * S + 8 + MB; }
// row stride of the Theta tile in shared memory: == 8 mod 16 doubles, so that the LDS.128 of a quarter warp (two chains x
// four dimension pairs) hits 32 distinct banks
__host__ __device__ inline int theta_stride(int Dp) { return ((Dp / 8) & 1) ? Dp : Dp + 8; }

// exp(-z) and 1/(1+exp(-z)) without branches: the library exp()/division carry slow-path branches that cost
// issue slots and reconvergence stalls inside the MMA pipeline (ncu: profiles/r01_c3_grad_v1).  Accurate to
// ~2 ulp (range reduction with a split ln2, degree-13 polynomial, exponent scaling, reciprocal by MUFU seed
// + two Newton steps); saturates exactly like the naive form for |z| > 708.
// -z clamped to [-708, 708] on the high word alone (|z| >= 708 <=> (hi & 0x7fffffff) >= 0x40862000: the same values as
// fmin(fmax(-z, -708), 708), which compiles to two DSETP on the FP64 pipe plus six select / logic instructions per element)
__device__ __forceinline__ double neg_clamp708(double z) {
#ifdef QSB_GLM_FPCLAMP      // A/B: the previous form
  return fmin(fmax(-z, -708.0), 708.0);
#else
  const int hi = __double2hiint(z);
  const bool big = (unsigned)(hi & 0x7fffffff) >= 0x40862000u;
  const int nhi = hi ^ (int)0x80000000;                                 // high word of -z
  return __hiloint2double(big ? ((nhi & (int)0x80000000) | 0x40862000) : nhi, big ? 0 : __double2loint(z));
#endif
}
// exp(t) = 2^k 2^(j/256) e^r with n = 256 k + j = round(256 t / ln 2) and |r| <= ln2/512: a degree-4 series (truncation
// r^5/120 < 4e-17) and one table value from shared memory replace the degree-13 series on |r| <= ln2/2 -- 8 instead of 17
// FP64-datapath operations up to 1 + e^t (the kernel is bound by the datapath the DMMAs and these DFMAs share), ~1.5 ulp on p as
// before (numpy model against 200-bit arithmetic).  n as a double comes from I2F (XU pipe) instead of subtracting the
// rounding shift again.  Measured: profiles/README.md (r02x, r02y).
__device__ __forceinline__ double sigmoid_fast(double z, const double* tab) {
  double t = neg_clamp708(z);
  const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52: round-to-nearest-integer by addition
  const int n = __double2loint(fma(t, 369.3299304675746, SHIFT));
  const double nd = (double)n;
  double r = fma(nd, -0x1.62e42fee00000p-9, t);
  r = fma(nd, -7.453964567463233e-13, r);
  double q = 4.1666666666666664e-02;           // 1/4!
  q = fma(q, r, 1.6666666666666666e-01);       // 1/3!
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const double tj = tab[n & (TAB_N - 1)];
  const double ts = __hiloint2double(__double2hiint(tj) + ((n >> 8) << 20), __double2loint(tj));   // 2^k 2^(j/256): stays normal (|k| <= 1022)
  double d = fma(q, ts, 1.0);                  // 1 + e^t in one rounding
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double err = fma(-d, y, 1.0);          // one cubic step: y (1 + e + e^2), 3 x the seed's ~23 bits
  err = fma(err, err, err);
  y = fma(y, err, y);
  return y;
}

// The same computation cut into 14 single-latency steps on a (zr, q, k) state, so that the steps of several
// elements can be issued round-robin between the tensor-core instructions of the next row block: no step waits on
// its predecessor and the MMA stream of the warp never stalls behind a long dependent DFMA chain
// (ncu profiles/r01_c3_grad_v2: stall_wait 33 % of samples before this change).
struct SigState { double zr, q; int k; };
template <int STEP> __device__ __forceinline__ void sigmoid_step(SigState& s, double yv, const double* tab) {
  constexpr double SHIFT = 6755399441055744.0;
  static_assert(TAB_N == 256, "the constants below are for 256 table entries");
  if (STEP == 0) s.zr = neg_clamp708(s.zr);
  else if (STEP == 1) s.q = fma(s.zr, 369.3299304675746, SHIFT);
  else if (STEP == 2) { s.k = __double2loint(s.q); s.q = (double)s.k; }
  else if (STEP == 3) s.zr = fma(s.q, -0x1.62e42fee00000p-9, s.zr);
  else if (STEP == 4) { s.zr = fma(s.q, -7.453964567463233e-13, s.zr); s.q = 4.1666666666666664e-02; }
  else if (STEP == 5) s.q = fma(s.q, s.zr, 1.6666666666666666e-01);
  else if (STEP == 6) s.q = fma(s.q, s.zr, 0.5);
  else if (STEP == 7) s.q = fma(s.q, s.zr, 1.0);
  else if (STEP == 8) s.q = fma(s.q, s.zr, 1.0);
  else if (STEP == 9) s.zr = tab[s.k & (TAB_N - 1)];                    // (zr is free from here to the reciprocal)
  else if (STEP == 10) s.q = fma(s.q, __hiloint2double(__double2hiint(s.zr) + ((s.k >> 8) << 20), __double2loint(s.zr)), 1.0);
  else if (STEP == 11) { asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s.zr) : "d"(s.q)); }
  else if (STEP == 12) { double err = fma(-s.q, s.zr, 1.0); err = fma(err, err, err); s.zr = fma(s.zr, err, s.zr); }   // cubic step
  else if (STEP == 13) s.zr = yv - s.zr;
}
static constexpr int SIG_STEPS = 14;
template <int S0, int S1> __device__ __forceinline__ void sigmoid_steps(SigState (&st)[4], const double (&yv)[4], const double* tab) {
  if constexpr (S0 < S1 && S0 < SIG_STEPS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sigmoid_step<S0>(st[i], yv[i], tab);
    sigmoid_steps<S0 + 1, S1>(st, yv, tab);
  }
}

// FULL: every one of the NJ dimension tiles is live (nj == NJ), so the tile loops carry no predicates.
// LP: also accumulate the log-likelihood (last leapfrog step of a trajectory, initialisation, step-size probes).
// TRIM: the last dimension tile holds at most 4 real dimensions (d = 100: dims 96..99 of 96..103).  In the first
// contraction the k <-> dimension map of that tile becomes k-step 0 = dims 8(NJ-1) + tig and its all-padding second
// k-step is not issued (100 instead of 104 DMMAs per row block); the second contraction is unchanged.
// SK: the balanced decomposition below (GradParams::sk_U > 0) instead of the (chain tile, row split) grid.
// COH: Theta was written by OTHER SMs earlier in this same launch (glm_traj_kernel: the leapfrog update between two gradient
// evaluations runs inside the kernel): read it past the non-coherent L1.
template <int NJ, bool FULL, bool LP, bool TRIM, bool SK, bool COH>
__device__ __forceinline__ void glm_grad_body(const GradParams& P, bool& pdl_pending) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  const size_t SD = stage_doubles(P.S);
  const int NST = P.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + SD * NST);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work decomposition.  sk_U == 0: CTA = (chain tile, row split), grid = ntiles x ksplit.  sk_U > 0: the ntiles x nblocks
  // row-block units, tile-major, are cut into sk_U runs of equal COST, one run per CTA (grid = number of SMs: one full
  // wave for any chain count; a unit of the partially filled last tile costs what its live warps cost -- glm_layout()); a
  // run that crosses a tile boundary is worked off as two (or more) segments, and a tile's segments land in consecutive
  // partial-sum slots (slot = CTA index - index of the first CTA that touches the tile).
  int& sk_su = *reinterpret_cast<int*>(empty + STAGES);    // first unit of the CTA's next segment (kept out of the register file)
  if (SK) {
    if (threadIdx.x == 0) sk_su = P.sk_start[blockIdx.x];   // (ntiles x nblocks < 2^31: checked at creation)
    __syncthreads();
  }
  for (;;) {
  int tile, split, b0, b1;
  if (SK) {
    const int su = sk_su, su_end = P.sk_start[blockIdx.x + 1];
    tile = su / P.nblocks; b0 = su - tile * P.nblocks;
    b1 = min(P.nblocks, b0 + (su_end - su));
    split = (int)blockIdx.x - P.sk_first[tile];
  } else {
    tile = blockIdx.x % P.ntiles; split = blockIdx.x / P.ntiles;
    b0 = (int)(((long)split * P.nblocks) / P.ksplit); b1 = (int)(((long)(split + 1) * P.nblocks) / P.ksplit);
  }

  // warps whose eight chains all lie past the last chain (the tail of the last tile) do no work at all: the stage-release
  // barriers count the live warps only, so a 64-chain tile costs its SM sub-partitions 2/3 of a full tile's time
  const int nvw = min(CW, (int)((P.C - (uint32_t)tile * CB + 7u) / 8u));
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nvw); }
    mbar_fence_init();
  }
  if ((int)threadIdx.x < NST * 8) sm[SD * (threadIdx.x >> 3) + (size_t)MB * P.S + (threadIdx.x & 7)] = 0.0;   // over-read pads
  for (int j = threadIdx.x; j < TAB_N; j += GRAD_THREADS) sm[SD * NST + TAB_OFF + j] = c_exp2_tab[j];   // 2^(j/256) for the sigmoid
  if (THETA_SMEM) {   // the chain tile's Theta, [CB][theta_stride]; rows of chains past the end are zero
    if (pdl_pending) { pdl_wait(); pdl_pending = false; }
    double* sT = sm + SD * NST + BAR_DOUBLES;
    const int Dp = P.Dp, stride = theta_stride(Dp);
    for (int idx = threadIdx.x; idx < CB * Dp; idx += GRAD_THREADS) {
      const int ch = idx / Dp, j = idx - ch * Dp;
      const uint32_t gch = (uint32_t)tile * CB + ch;
      sT[(size_t)ch * stride + j] = gch < P.C ? P.theta[(size_t)gch * Dp + j] : 0.0;
    }
  }
  __syncthreads();

  // ---- no dedicated producer warp (it would take a third of one sub-partition's register file): thread 0 starts
  // the ring; afterwards the warp whose release of a stage is the LAST one pending re-arms it with the next row
  // block (TMA bulk copies of whole row blocks: contiguous in HBM because Xp is stored with stride S)
  const uint32_t bytesX = (uint32_t)(MB * P.S * sizeof(double)), bytesY = (uint32_t)(MB * sizeof(double));
  auto issue_load = [&](int b, int st) {      // st = (b - b0) % NST, carried by the callers (no runtime division in the loop)
    mbar_expect_tx(&full[st], bytesX + bytesY);
    double* dst = sm + SD * st;
    bulk_g2s(dst, P.Xp + (size_t)b * MB * P.S, bytesX, &full[st]);
    bulk_g2s(dst + (size_t)MB * P.S + 8, P.yv + (size_t)b * MB, bytesY, &full[st]);
  };
  if (threadIdx.x == 0)
    for (int b = b0; b < b1 && b < b0 + NST; ++b) issue_load(b, b - b0);

  // X and y are model data, not produced by the predecessor kernel: the ring is already filling while we wait for Theta
  if (pdl_pending) { pdl_wait(); pdl_pending = false; }
  if (warp < nvw) {
  // ---- consumers: warp w owns chains tile*CB + 8w .. +7; thread (gid, tig) owns dims 8J+2tig+{0,1} of chain gid
  const int gid = lane >> 2, tig = lane & 3;
  const double* tab = sm + SD * NST + TAB_OFF;
  const uint32_t chain = (uint32_t)tile * CB + warp * 8 + gid;
  const bool valid = chain < P.C;
  const int S = P.S;
  const int nj = FULL ? NJ : P.nj;
  // row permutation inside an 8-row tile (see header): rho = [0,2,1,3,6,4,7,5]; identity with the column-permuted X
  constexpr bool PERM = TRIM && PERMX;
  const uint32_t RHO = PERM ? 0x76543210u : 0x57463120u;
  const int rho_g = (RHO >> (4 * gid)) & 7, rho_t0 = (RHO >> (8 * tig)) & 7, rho_t1 = (RHO >> (8 * tig + 4)) & 7;

  double th[THETA_SMEM ? 1 : NJ][2], G[NJ][2];
  const int STH = theta_stride(P.Dp);
  const double* sTh = sm + SD * NST + BAR_DOUBLES + (size_t)(warp * 8 + gid) * STH + 2 * tig;   // THETA_SMEM: this thread's Theta row
#pragma unroll
  for (int J = 0; J < NJ; ++J) {
    G[J][0] = G[J][1] = 0.0;
    if (!THETA_SMEM) {
      th[J % (THETA_SMEM ? 1 : NJ)][0] = th[J % (THETA_SMEM ? 1 : NJ)][1] = 0.0;
      if ((FULL || J < nj) && valid) {
        const double2* tp = reinterpret_cast<const double2*>(P.theta + (size_t)chain * P.Dp + 8 * J + 2 * tig);
        double2 v = COH ? __ldcg(tp) : *tp;
        th[J % (THETA_SMEM ? 1 : NJ)][0] = v.x; th[J % (THETA_SMEM ? 1 : NJ)][1] = v.y;
      }
    }
  }
  double thL = 0.0;                                     // TRIM: theta[8(NJ-1) + tig]
  if (TRIM && valid) thL = COH ? __ldcg(P.theta + (size_t)chain * P.Dp + 8 * (NJ - 1) + tig) : P.theta[(size_t)chain * P.Dp + 8 * (NJ - 1) + tig];
  double lpacc = 0.0;
  const int offA = rho_g * S + 2 * tig;                 // phase 1 B fragment: X[rho(gid)][8J + 2tig .. +1]
  const int offL = PERM ? rho_g * S + tig * PERM_PJ + (NJ - 1) : rho_g * S + 8 * (NJ - 1) + tig;      // TRIM: X[rho(gid)][8(NJ-1) + tig]
  const int offB0 = rho_t0 * S + (PERM ? gid * PERM_PJ : gid), offB1 = rho_t1 * S + (PERM ? gid * PERM_PJ : gid);   // phase 2 B fragment: X[rho(2tig+e)][8J + gid]
  const int offP = rho_g * S + 2 * tig * PERM_PJ;       // PERM, phase 1: X[rho(gid)][(2tig + s) PJ + J]
  constexpr bool want_lp = LP;

  // Software pipeline over row blocks: while the sigmoid epilogue of block b runs on the FP64 ALU path, the first
  // contraction (Z^T = Theta^T X^T) of block b+1 is already feeding the tensor pipe; then the second contraction
  // (G^T += R^T X) of block b.  acc holds Z of the block whose epilogue is due, nxt the one being accumulated.
  double acc[NT][2], nxt[NT][2];
  auto phase1_tile = [&](const double* Xs, int J, double (&a)[NT][2]) {
    if (TRIM && J == NJ - 1) {
      double xl[NT];
#pragma unroll
      for (int T = 0; T < NT; ++T) xl[T] = Xs[T * 8 * S + offL];
#pragma unroll
      for (int T = 0; T < NT; ++T) dmma(a[T], thL, xl[T]);
      return;
    }
    double2 xv[NT];
#pragma unroll
    for (int T = 0; T < NT; ++T) xv[T] = *reinterpret_cast<const double2*>(Xs + T * 8 * S + offA + 8 * J);
    double t0, t1;
    if (THETA_SMEM) { const double2 tv = *reinterpret_cast<const double2*>(sTh + 8 * J); t0 = tv.x; t1 = tv.y; }
    else { t0 = th[J % (THETA_SMEM ? 1 : NJ)][0]; t1 = th[J % (THETA_SMEM ? 1 : NJ)][1]; }
#pragma unroll
    for (int T = 0; T < NT; ++T) dmma(a[T], t0, xv[T].x);   // NT independent accumulators back to back,
#pragma unroll
    for (int T = 0; T < NT; ++T) dmma(a[T], t1, xv[T].y);   // then the dependent second k-step of each
  };
  // PERM: half of a tile pair.  slot k = 2 Jp + s: k-step s of tiles J = 2 Jp and J + 1 from one LDS.128 per n-tile
  // (4 loads, 8 DMMAs: the granularity of phase1_tile, so the lock-step sigmoid schedule below is unchanged)
  auto phase1_half = [&](const double* Xs, int k, double (&a)[NT][2]) {
    const int J = k & ~1, sidx = k & 1;
    double2 xv[NT];
#pragma unroll
    for (int T = 0; T < NT; ++T) xv[T] = *reinterpret_cast<const double2*>(Xs + T * 8 * S + offP + sidx * PERM_PJ + J);
    double t0, t1;
    if (THETA_SMEM) {      // (conflict-free as LDS.128 of the dimension pair; the k-step picks its half)
      const double2 v0 = *reinterpret_cast<const double2*>(sTh + 8 * J), v1 = *reinterpret_cast<const double2*>(sTh + 8 * (J + 1));
      t0 = sidx ? v0.y : v0.x; t1 = sidx ? v1.y : v1.x;
    } else { t0 = th[J % (THETA_SMEM ? 1 : NJ)][sidx]; t1 = th[(J + 1) % (THETA_SMEM ? 1 : NJ)][sidx]; }
#pragma unroll
    for (int T = 0; T < NT; ++T) dmma(a[T], t0, xv[T].x);
#pragma unroll
    for (int T = 0; T < NT; ++T) dmma(a[T], t1, xv[T].y);
  };
  auto phase1_slot = [&](const double* Xs, int k, double (&a)[NT][2]) {
    if constexpr (PERM) phase1_half(Xs, k, a); else phase1_tile(Xs, k, a);
  };
  // Log-likelihood of a row block: sum_i log w_i with w_i = (y_i ? p_i : 1 - p_i) (distrib.t:17-25 on the naive sigmoid) is
  // taken as ONE log of the product of the block's eight w_i per thread -- seven library logs fewer per block, and the
  // product's rounding (8 ulp) is smaller than that of a sum of eight logs.  Factors below 1e-30 (a hopelessly misfit
  // observation) are added as their own log so that the product cannot underflow.
  double wprod = 1.0;
  auto lik_factor = [&](double w) {
    if (w < 1e-30) lpacc += log(w); else wprod *= w;
  };
  auto lik_flush = [&]() { lpacc += log(wprod); wprod = 1.0; };
  auto epilogue_elem = [&](const double* ys, int T, int e) {
    const double y = ys[T * 8 + (e ? rho_t1 : rho_t0)];
    if (want_lp) {
      const double pr = 1.0 / (1.0 + exp(-acc[T][e]));
      if (y <= 1.0) lik_factor(y != 0.0 ? pr : 1.0 - pr);
      acc[T][e] = y - pr;
    } else {
      acc[T][e] = y - sigmoid_fast(acc[T][e], tab);
    }
  };

  if (b0 < b1) {
    mbar_wait(&full[0], 0u);
#pragma unroll
    for (int T = 0; T < NT; ++T) acc[T][0] = acc[T][1] = 0.0;
    if constexpr (PERM) {
#pragma unroll
      for (int k = 0; k < NJ - 1; ++k) phase1_half(sm, k, acc);
      phase1_tile(sm, NJ - 1, acc);
    } else {
#pragma unroll
      for (int J = 0; J < NJ; ++J) if (FULL || J < nj) phase1_tile(sm, J, acc);
    }
  }
  // ring position of the row block being worked on: stage st = (b - b0) % NST and the parity ph = ((b - b0) / NST) & 1 of its
  // use, advanced by increment-and-wrap (NST is a run-time value: `%` and `/` by it cost ~55 integer / MUFU instructions per
  // row block and sat in front of the wait for the next stage)
  int st = 0; uint32_t ph = 0u;
  for (int b = b0; b < b1; ++b) {
    const double* Xs = sm + SD * st;
    const double* ys = Xs + (size_t)MB * S + 8;
    const bool has_next = b + 1 < b1;
    int st2 = st + 1; uint32_t ph2 = ph;
    if (st2 == NST) { st2 = 0; ph2 ^= 1u; }
#ifdef QSB_GLM_RINGDIV       // A/B: the previous form
    { const int it2 = b - b0 + 1; st2 = it2 % NST; ph2 = (uint32_t)((it2 / NST) & 1); }
#endif
    if (has_next) {
      mbar_wait(&full[st2], ph2);
      const double* Xn = sm + SD * st2;
#pragma unroll
      for (int T = 0; T < NT; ++T) nxt[T][0] = nxt[T][1] = 0.0;
      if constexpr (NT == 4 && NJ >= 12) {
        // lock-step sigmoid: elements (T=0,1) advance 4 steps per dimension tile J = 0..5, elements (T=2,3) during J = 6..11
        SigState sa[4], sb[4];
        double ya[4], yb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          sa[i].zr = acc[i >> 1][i & 1]; sa[i].q = 0.0; sa[i].k = 0; ya[i] = ys[(i >> 1) * 8 + ((i & 1) ? rho_t1 : rho_t0)];
          sb[i].zr = acc[2 + (i >> 1)][i & 1]; sb[i].q = 0.0; sb[i].k = 0; yb[i] = ys[(2 + (i >> 1)) * 8 + ((i & 1) ? rho_t1 : rho_t0)];
        }
        if (FULL || 0 < nj) phase1_slot(Xn, 0, nxt);  sigmoid_steps<0, 3>(sa, ya, tab);
        if (FULL || 1 < nj) phase1_slot(Xn, 1, nxt);  sigmoid_steps<3, 5>(sa, ya, tab);
        if (FULL || 2 < nj) phase1_slot(Xn, 2, nxt);  sigmoid_steps<5, 7>(sa, ya, tab);
        if (FULL || 3 < nj) phase1_slot(Xn, 3, nxt);  sigmoid_steps<7, 10>(sa, ya, tab);
        if (FULL || 4 < nj) phase1_slot(Xn, 4, nxt);  sigmoid_steps<10, 12>(sa, ya, tab);
        if (FULL || 5 < nj) phase1_slot(Xn, 5, nxt);  sigmoid_steps<12, 14>(sa, ya, tab);
        if (FULL || 6 < nj) phase1_slot(Xn, 6, nxt);  sigmoid_steps<0, 3>(sb, yb, tab);
        if (FULL || 7 < nj) phase1_slot(Xn, 7, nxt);  sigmoid_steps<3, 5>(sb, yb, tab);
        if (FULL || 8 < nj) phase1_slot(Xn, 8, nxt);  sigmoid_steps<5, 7>(sb, yb, tab);
        if (FULL || 9 < nj) phase1_slot(Xn, 9, nxt);  sigmoid_steps<7, 10>(sb, yb, tab);
        if (FULL || 10 < nj) phase1_slot(Xn, 10, nxt); sigmoid_steps<10, 12>(sb, yb, tab);
        if (FULL || 11 < nj) phase1_slot(Xn, 11, nxt); sigmoid_steps<12, 14>(sb, yb, tab);
#pragma unroll
        for (int J = 12; J < NJ; ++J) if (FULL || J < nj) phase1_tile(Xn, J, nxt);
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i >> 1][i & 1] = sa[i].zr; acc[2 + (i >> 1)][i & 1] = sb[i].zr; }
        if (want_lp) {
          // from the residual r = y - p of the lock-step sigmoid: y = 1: p = 1 - r; y = 0: 1 - p = 1 + r; i.e. w = 1 - |r|
          // (equal to the naive form up to the rounding of 1 - p).  Padding rows carry y = 2.
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ya[i] <= 1.0) lik_factor(1.0 - fabs(sa[i].zr));
            if (yb[i] <= 1.0) lik_factor(1.0 - fabs(sb[i].zr));
          }
          lik_flush();
        }
      } else {
#pragma unroll
        for (int J = 0; J < NJ; ++J) {
          if (FULL || J < nj) phase1_tile(Xn, J, nxt);
          if (J < 2 * NT) epilogue_elem(ys, J >> 1, J & 1);      // NJ >= 2*NT interleaves all; otherwise the rest follows
        }
#pragma unroll
        for (int q = NJ; q < 2 * NT; ++q) epilogue_elem(ys, q >> 1, q & 1);
        if (want_lp) lik_flush();
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2 * NT; ++q) epilogue_elem(ys, q >> 1, q & 1);
      if (want_lp) lik_flush();
    }
    // phase 2: G^T += R^T X
    if (P.want_grad)
#pragma unroll
    for (int T = 0; T < NT; ++T) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double* xb = Xs + T * 8 * S + (e ? offB1 : offB0);
        if constexpr (PERM) {
#pragma unroll
          for (int J = 0; J < NJ; J += 2) {
            const double2 v = *reinterpret_cast<const double2*>(xb + J);
            dmma(G[J], acc[T][e], v.x);
            if (J + 1 < NJ) dmma(G[J + 1], acc[T][e], v.y);
          }
        } else {
#pragma unroll
        for (int J = 0; J < NJ; ++J)
          if (FULL || J < nj) dmma(G[J], acc[T][e], xb[8 * J]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      if (mbar_arrive_last(&empty[st])) {           // every warp is done with this stage: refill it
        mbar_wait(&empty[st], ph);                  // acquire the other warps' releases (already complete)
        if (b + NST < b1) issue_load(b + NST, st);
      }
    }
#pragma unroll
    for (int T = 0; T < NT; ++T) { acc[T][0] = nxt[T][0]; acc[T][1] = nxt[T][1]; }
    st = st2; ph = ph2;
  }

  if (valid && P.want_grad) {
    double* out = P.partG + ((size_t)split * P.C + chain) * P.Dp + 2 * tig;
#pragma unroll
    for (int J = 0; J < NJ; ++J)
      if (FULL || J < nj) *reinterpret_cast<double2*>(out + 8 * J) = make_double2(G[J][0], G[J][1]);
  }
  if (want_lp) {
    lpacc += __shfl_xor_sync(0xffffffffu, lpacc, 1);
    lpacc += __shfl_xor_sync(0xffffffffu, lpacc, 2);
    if (valid && tig == 0) P.partLp[(size_t)split * P.C + chain] = lpacc;
  }
  }   // warp < nvw
  if (!SK) break;
  __syncthreads();                       // next segment of this CTA's run: the ring starts over
  const int su_next = tile * P.nblocks + b1;
  if (su_next >= P.sk_start[blockIdx.x + 1]) break;
  if (threadIdx.x == 0) {
    sk_su = su_next;
    for (int s = 0; s < NST; ++s) { mbar_inval(&full[s]); mbar_inval(&empty[s]); }
  }
  __syncthreads();
  }
}

template <int NJ, bool FULL, bool LP, bool TRIM = false, bool SK = false>
__global__ void __launch_bounds__(GRAD_THREADS, 1) glm_grad_kernel(const __grid_constant__ GradParams P) {
  pdl_trigger();
  bool pdl_pending = true;      // Theta (and the partial-sum buffers this kernel overwrites) belong to the predecessor until pdl_wait
  glm_grad_body<NJ, FULL, LP, TRIM, SK, false>(P, pdl_pending);
}

// ---------------------------------------------------------------------------------------------
// Vector kernels: one warp per chain, lanes over dimensions.
enum { V_INIT_SAMPLE = 0, V_FINISH_EVAL = 1, V_TRANS_FIRST = 2, V_TRANS_MID = 3, V_TRANS_LAST = 4, V_PROBE_FIRST = 5, V_PROBE_LAST = 6,
       V_MH_PROPOSE = 7, V_MH_ACCEPT = 8, V_DEBUG_FIRST = 9, V_DEBUG_LAST = 10 };

struct VecParams {
  RngKey key; uint32_t chain_begin, C; int d, Dp, ksplit;
  int sk_U, sk_nblocks, sk_CB;   // balanced decomposition of the gradient kernel (sk_U > 0): partial-sum slots per chain tile vary
  const int* sk_np;              // [ntiles] number of partial sums of every chain tile
  double prior_sigma;
  double *x, *g, *xs, *gs, *p;                 // [C][Dp]
  const double* partG; const double* partLp;
  double *lp, *H0, *eps, *prevlp;
  double *da_gbar, *da_xbar, *da_x0, *da_lastx, *da_gamma; uint32_t* da_k;
  int *direction, *done; int* n_active;
  unsigned long long *made, *acc;
  int* status;
  unsigned long long* health;   // [0] transitions with a non-finite accept threshold, [1] HMC transitions with |dH| > 1000
  uint32_t iter, L, step, probe;
  double T, invT;   // trace temperature of this iteration (AnnealingKernel, mcmc.t:297-319; trace.t:737-741)
  int kind;         // 1 HMC, 2 Drift, 3 HARM
  double *dr_mean, *dr_m2, *dr_s0, *dr_mult; uint32_t* dr_n;   // Drift: RunningVar per component [C][Dp] + per-chain n, scale, shrink factor
  double* ha_scale;                                            // HARM: [C]
  int adapting;
  double stepSize0;
  // samples / record
  double *samples, *samp_lp; uint32_t SC; uint64_t sidx; int keep;
  double* mom; uint64_t mom_half; uint32_t mom_batch;   // QSB_FLAG_MOMENTS (engine.cuh: mom_update), layout [stat][chain][dim]
  double *rec_state, *rec_dh, *rec_step, *rec_leap; int* rec_accept; uint64_t rec_iters, rec_row;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// number of partial sums the stream-K gradient kernel left for chain c (its tile's segments)
__device__ __forceinline__ int n_partials_sk(const VecParams& P, uint32_t c) { return P.sk_np[c / (uint32_t)P.sk_CB]; }
// gradient and (optionally) log-density of the chain at xs from the partial sums of the GEMM kernel
template <bool COH = false>
__device__ __forceinline__ double finish_grad(const VecParams& P, uint32_t c, int i, double th) {
  double gl = 0.0;
  if (!P.sk_U) {
    for (int r = 0; r < P.ksplit; ++r) gl += P.partG[((size_t)r * P.C + c) * P.Dp + i];
  } else {
    const int np = n_partials_sk(P, c);
    if (COH) { for (int r = 0; r < np; ++r) gl += __ldcg(P.partG + ((size_t)r * P.C + c) * P.Dp + i); }   // written by other SMs in this launch
    else for (int r = 0; r < np; ++r) gl += P.partG[((size_t)r * P.C + c) * P.Dp + i];
  }
  const double g = gaussian_dlp_dx(th, 0.0, P.prior_sigma) + gl;
  return P.T != 1.0 ? g * P.invT : g;
}
__device__ __forceinline__ double finish_lp(const VecParams& P, uint32_t c, int lane, const double* pos) {
  double lpp = 0.0;
  for (int i = lane; i < P.d; i += 32) lpp += gaussian_lp(pos[i], 0.0, P.prior_sigma);
  lpp = warp_sum(lpp);
  double ll = 0.0;
  if (!P.sk_U) {
    for (int r = 0; r < P.ksplit; ++r) ll += P.partLp[(size_t)r * P.C + c];
  } else {
    const int np = n_partials_sk(P, c);
    for (int r = 0; r < np; ++r) ll += P.partLp[(size_t)r * P.C + c];
  }
  return lpp + ll;   // at temperature 1; callers divide
}

// finish leapfrog step `step` of chain c (second half-kick), start the next one (hmc.t:307-323); one warp
template <bool COH>
__device__ __forceinline__ void trans_mid_chain(const VecParams& P, uint32_t c, int lane, uint32_t step) {
  const int d = P.d;
  const size_t base = (size_t)c * P.Dp;
  double* xs = P.xs + base; double* p = P.p + base;
  const double eps = P.eps[c];
  const double he = 0.5 * eps;
  for (int i = lane; i < d; i += 32) {
    double pos = xs[i];
    double gn = finish_grad<COH>(P, c, i, pos);
    double m = p[i] + he * gn;
    if (P.rec_leap) P.rec_leap[(((size_t)c * P.rec_iters + P.rec_row) * P.L + step) * d + i] = pos;
    m = m + he * gn;
    p[i] = m; xs[i] = pos + eps * m * 1.0;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) glm_vec_kernel(const __grid_constant__ VecParams P) {
  const uint32_t c = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  pdl_trigger();
  pdl_wait();
  if (c >= P.C) return;
  const int lane = threadIdx.x & 31;
  const uint32_t gc = P.chain_begin + c;
  const int d = P.d;
  const size_t base = (size_t)c * P.Dp;
  double* x = P.x + base; double* g = P.g + base; double* xs = P.xs + base; double* gs = P.gs + base; double* p = P.p + base;

  if (MODE == V_INIT_SAMPLE) {
    // forward sample theta_j ~ gaussian(0, sigma) in program order (trace.t:612-623); lp follows from the first evaluation
    for (int i = lane; i < P.Dp; i += 32) {
      double v = i < d ? gaussian_sample(P.key, gc, QSB_ITER_INIT, (uint32_t)i, 0.0, P.prior_sigma) : 0.0;
      x[i] = v; xs[i] = v; g[i] = 0.0; gs[i] = 0.0; p[i] = 0.0;
    }
    return;
  }
  if (MODE == V_FINISH_EVAL) {
    // lp and gradient at x (= xs) after an evaluation outside a trajectory: initial trace logprob and
    // the gradient HMCKernel:checkForChanges computes on its first call (hmc.t:254)
    for (int i = lane; i < d; i += 32) g[i] = finish_grad(P, c, i, x[i]);
    double lp = finish_lp(P, c, lane, x);
    // currTrace.logprob comes from the initialisation at temperature 1 (trace.t:592-624); the dual trace is first
    // evaluated under the first iteration's temperature (hmc.t:218, 254)
    if (lane == 0) { P.lp[c] = lp; P.prevlp[c] = P.T != 1.0 ? lp / P.T : lp; }
    return;
  }
  if (MODE == V_MH_PROPOSE) {
    // DriftKernel:next / HARMKernel:next proposals (drift.t:140-142; harm.t:67-79), reference operation order
    if (P.kind == 2) {
      const uint32_t n_rv = P.dr_n[c];
      const double s0 = P.dr_s0[c], mult = P.dr_mult[c];
      for (int i = lane; i < d; i += 32) {
        double sc = s0;
        if (n_rv > 100) { sc = sqrt(P.dr_m2[base + i] / (double)(n_rv - 1)); if (mult != 1.0) sc = sc * mult; }   // drift.t:288-298
        xs[i] = gaussian_sample(P.key, gc, P.iter, (uint32_t)i, x[i], sc);
      }
    } else {
      double zz[4] = {0.0, 0.0, 0.0, 0.0};
      double nrm = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) { const int i = lane + 32 * q; if (i < d) { zz[q] = gaussian_std(P.key, gc, P.iter, (uint32_t)i); nrm += zz[q] * zz[q]; } }
      nrm = sqrt(warp_sum(nrm));
      const double step = uniform_at(P.key, gc, P.iter, (uint32_t)d) * P.ha_scale[c];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const int i = lane + 32 * q; if (i < d) xs[i] = x[i] + (step * (zz[q] / nrm)); }
    }
    if (lane == 0) P.made[c] += 1;
    return;
  }
  if (MODE == V_MH_ACCEPT) {
    // accept / reject against the cached log-density, adaptation, sample (drift.t:155-167, 274-300; harm.t:90-108)
    double lpn = finish_lp(P, c, lane, xs);
    if (P.T != 1.0) lpn = lpn / P.T;
    const double thresh = lpn - P.lp[c];
    if (lane == 0 && !(fabs(thresh) <= 1.7976931348623157e308)) atomicAdd(&P.health[0], 1ull);
    const double u = uniform_at(P.key, gc, P.iter, (uint32_t)d + (P.kind == 3 ? 1u : 0u));
    const bool accept = log(u) < thresh;
    double step_rec = 0.0;
    __syncwarp();
    if (accept) for (int i = lane; i < d; i += 32) x[i] = xs[i];
    const unsigned long long acc = P.acc[c] + (accept ? 1ull : 0ull), made = P.made[c];
    if (P.kind == 2) {
      const uint32_t n_rv = P.dr_n[c];
      const double s0 = P.dr_s0[c];
      step_rec = s0;
      if (P.adapting) {
        const double ratio = (double)acc / (double)made;
        uint32_t n_new = n_rv;
        if (ratio >= 0.05 && acc > 5) {   // RunningVar:update with the current state (drift.t:26-37)
          n_new = n_rv + 1;
          for (int i = lane; i < d; i += 32) {
            const double cur = accept ? xs[i] : x[i];
            if (n_new == 1) { P.dr_mean[base + i] = cur; P.dr_m2[base + i] = 0.0; }
            else {
              const double mean = P.dr_mean[base + i];
              const double newmean = mean + (cur - mean) / (double)n_new;
              P.dr_m2[base + i] = P.dr_m2[base + i] + (cur - mean) * (cur - newmean);
              P.dr_mean[base + i] = newmean;
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          const bool shrink = ratio < 0.05 && acc > 5;
          P.dr_n[c] = n_new;
          if (n_new > 100) P.dr_mult[c] = shrink ? 0.99 : 1.0;
          else if (shrink) P.dr_s0[c] = s0 * 0.99;
        }
      }
    } else {
      const double hscale = P.ha_scale[c];
      step_rec = hscale;
      double t = 0.234;
      if (d < 5) { double tt = (d - 1) / 4.0; t = (1.0 - tt) * 0.44 + tt * 0.234; }
      double ns = hscale;
      if (P.adapting) {
        if (accept) ns = hscale + (hscale / (t * (1.0 - t))) * (1.0 - t) / (double)(P.iter + 1u);
        else ns = fabs(hscale - (hscale / (t * (1.0 - t))) * t / (double)(P.iter + 1u));
      }
      if (lane == 0) P.ha_scale[c] = ns;
    }
    const double lp_keep = accept ? lpn : P.lp[c];
    __syncwarp();
    if (lane == 0) { P.acc[c] = acc; if (accept) P.lp[c] = lpn; }
    if (P.rec_state) {
      size_t r = (size_t)c * P.rec_iters + P.rec_row;
      for (int i = lane; i < d; i += 32) P.rec_state[r * d + i] = x[i];
      if (lane == 0) { P.rec_accept[r] = accept ? 1 : 0; P.rec_dh[r] = thresh; P.rec_step[r] = step_rec; }
    }
    if (P.keep && c < P.SC) {
      if (P.mom) {
        const size_t plane = (size_t)P.SC * d;
        for (int i = lane; i < d; i += 32) {
          double* b0 = P.mom + (size_t)c * d + i;
          mom_update([&](int st) { return b0 + (size_t)st * plane; }, P.sidx, P.mom_half, P.mom_batch, x[i]);
        }
      } else {
        for (int i = lane; i < d; i += 32) P.samples[((size_t)P.sidx * P.SC + c) * d + i] = x[i];
        if (lane == 0) P.samp_lp[(size_t)P.sidx * P.SC + c] = lp_keep * P.T;   // mcmc.t:69
      }
    }
    return;
  }
  const double eps = P.eps[c];
  const double he = 0.5 * eps;
  if (MODE == V_DEBUG_FIRST) {   // qsb_debug_leapfrog: first half-kick and drift of one step from injected (x, p, g, eps)
    for (int i = lane; i < d; i += 32) {
      double m = p[i] + he * g[i];
      double pos = x[i] + eps * m * 1.0;
      p[i] = m; xs[i] = pos;
    }
    return;
  }
  if (MODE == V_DEBUG_LAST) {    // ... gradient at the new position, second half-kick; lp when the step is a trajectory's last
    for (int i = lane; i < d; i += 32) {
      double gn = finish_grad(P, c, i, xs[i]);
      gs[i] = gn; p[i] = p[i] + he * gn;
    }
    if (P.keep) { double lp = finish_lp(P, c, lane, xs); if (lane == 0) P.lp[c] = lp; }
    return;
  }
  if (MODE == V_TRANS_FIRST || MODE == V_PROBE_FIRST) {
    if (MODE == V_PROBE_FIRST && P.done[c]) return;
    // sampleMomenta, initial Hamiltonian, first half-kick and drift (hmc.t:140-149, 307-315)
    const uint32_t it = MODE == V_PROBE_FIRST ? QSB_ITER_SEARCH + P.probe : P.iter;
    double kin = 0.0;
    for (int i = lane; i < d; i += 32) {
      double m = gaussian_std(P.key, gc, it, (uint32_t)i) * 1.0;
      kin += m * m * 1.0;
      m = m + he * g[i];
      double pos = x[i] + eps * m * 1.0;
      p[i] = m; xs[i] = pos;
    }
    if (MODE == V_TRANS_FIRST) {
      kin = -0.5 * warp_sum(kin);
      if (lane == 0) { P.H0[c] = kin + P.lp[c]; P.made[c] += 1; }
    }
    return;
  }
  if (MODE == V_TRANS_MID) { trans_mid_chain<false>(P, c, lane, P.step); return; }
  if (MODE == V_PROBE_LAST) {
    if (P.done[c]) return;
    // searchForInitialStepSize bookkeeping (hmc.t:345-391); the momentum half-kick is irrelevant to it
    double lp = finish_lp(P, c, lane, xs);
    if (P.T != 1.0) lp = lp / P.T;
    if (lane == 0) {
      const double LOG_HALF = log(0.5);
      double H = lp - P.prevlp[c];
      P.prevlp[c] = lp;
      double e2 = eps;
      int done = 0;
      if (P.probe == 0) P.direction[c] = H > LOG_HALF ? 1 : -1;
      else {
        int dir = P.direction[c];
        if (!(lp == lp)) { atomicOr(P.status, 1); done = 1; }
        else if (dir == 1 && H <= LOG_HALF) done = 1;
        else if (dir == -1 && H >= LOG_HALF) done = 1;
        else if (dir == 1) e2 = eps * 2.0;
        else e2 = eps * 0.5;
        if (!done && e2 > 1e300) { atomicOr(P.status, 2); done = 1; }
        if (!done && e2 == 0) { atomicOr(P.status, 4); done = 1; }
      }
      P.eps[c] = e2;
      if (done) {
        P.done[c] = 1;
        // adapter:reinit(stepSize, adaptRate) (hmc.t:261)
        P.da_k[c] = 0; P.da_x0[c] = e2; P.da_lastx[c] = e2; P.da_gbar[c] = 0.0; P.da_xbar[c] = 0.0; P.da_gamma[c] = 0.05;
      } else atomicAdd(P.n_active, 1);
    }
    return;
  }
  // ---- V_TRANS_LAST: last half-kick, final Hamiltonian, adaptation, accept/reject, sample (hmc.t:150-186)
  double kin = 0.0;
  for (int i = lane; i < d; i += 32) {
    double pos = xs[i];
    double gn = finish_grad(P, c, i, pos);
    double m = p[i] + he * gn;
    if (P.rec_leap) P.rec_leap[(((size_t)c * P.rec_iters + P.rec_row) * P.L + P.step) * d + i] = pos;
    gs[i] = gn; p[i] = m;
    kin += m * m * 1.0;
  }
  double newlp = finish_lp(P, c, lane, xs);
  if (P.T != 1.0) newlp = newlp / P.T;
  const double H_new = -0.5 * warp_sum(kin) + newlp;
  const double dH = H_new - P.H0[c];
  if (lane == 0) { if (!(fabs(dH) <= 1.7976931348623157e308)) atomicAdd(&P.health[0], 1ull); else if (fabs(dH) > 1000.0) atomicAdd(&P.health[1], 1ull); }
  double eps_new = eps;
  if (P.adapting) {   // adaptStepSize + DualAverage:update (hmc.t:38-48, 394-402)
    double EdH = exp(dH);
    if (EdH > 1.0) EdH = 1.0;
    if (!(EdH == EdH)) EdH = 0.0;
    const double target = P.L == 1 ? 0.574 : 0.65;
    double gr = target - EdH;
    uint32_t k = P.da_k[c] + 1;
    double avgeta = 1.0 / (double)(k + 10);
    double muk = 0.5 * sqrt((double)k) / P.da_gamma[c];
    double gbar = avgeta * gr + (1 - avgeta) * P.da_gbar[c];
    double lastx = P.da_x0[c] - muk * gbar;
    // DualAverage.xbar (hmc.t:44-46) is written by the reference but never read: not computed
    __syncwarp();
    if (lane == 0) { P.da_k[c] = k; P.da_gbar[c] = gbar; P.da_lastx[c] = lastx; }
    eps_new = exp(lastx);
  }
  const double u = uniform_at(P.key, gc, P.iter, (uint32_t)d);
  const bool accept = log(u) < dH;
  __syncwarp();
  if (accept) {
    for (int i = lane; i < d; i += 32) { x[i] = xs[i]; g[i] = gs[i]; }
    if (lane == 0) { P.acc[c] += 1; P.lp[c] = newlp; }
  }
  if (lane == 0) P.eps[c] = eps_new;
  __syncwarp();
  if (P.rec_state) {
    size_t r = (size_t)c * P.rec_iters + P.rec_row;
    for (int i = lane; i < d; i += 32) P.rec_state[r * d + i] = x[i];
    if (lane == 0) { P.rec_accept[r] = accept ? 1 : 0; P.rec_dh[r] = dH; P.rec_step[r] = eps; }
  }
  if (P.keep && c < P.SC) {
    if (P.mom) {
      const size_t plane = (size_t)P.SC * d;
      for (int i = lane; i < d; i += 32) {
        double* b0 = P.mom + (size_t)c * d + i;
        mom_update([&](int st) { return b0 + (size_t)st * plane; }, P.sidx, P.mom_half, P.mom_batch, x[i]);
      }
    } else {
      for (int i = lane; i < d; i += 32) P.samples[((size_t)P.sidx * P.SC + c) * d + i] = x[i];
      if (lane == 0) P.samp_lp[(size_t)P.sidx * P.SC + c] = (accept ? newlp : P.lp[c]) * P.T;   // mcmc.t:69
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The interior of a trajectory in ONE launch: `nsteps` gradient evaluations, each followed by the leapfrog update that
// produces the next Theta (V_TRANS_MID), with grid-wide barriers in place of the 2 x nsteps kernel boundaries.  Balanced
// decomposition only (one CTA per SM, all co-resident: cooperative launch).  The idea was to save the ~13 us per leapfrog step
// that sit outside the gradient kernel at small chain counts; measured, the two grid barriers per step cost MORE than the two
// kernel boundaries they replace once those overlap through programmatic dependent launch (see glm_sampler_create): an
// experiment kept behind QSB_FUSED_TRAJ=1, not the default path.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
  }
  __syncthreads();
}
template <int NJ, bool FULL, bool TRIM>
__global__ void __launch_bounds__(GRAD_THREADS, 1) glm_traj_kernel(const __grid_constant__ GradParams P, const __grid_constant__ VecParams V,
                                                                   int nsteps, unsigned step0, unsigned* counter) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + stage_doubles(P.S) * P.stages);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned target = 0;
  bool no_pdl = false;
  for (int s = 0; s < nsteps; ++s) {
    glm_grad_body<NJ, FULL, false, TRIM, true, true>(P, no_pdl);
    __syncthreads();
    if (threadIdx.x == 0) for (int q = 0; q < P.stages; ++q) { mbar_inval(&full[q]); mbar_inval(&empty[q]); }
    grid_barrier(counter, target);             // every partial sum of this evaluation is in L2
    for (uint32_t c = blockIdx.x * CW + warp; c < V.C; c += gridDim.x * CW) trans_mid_chain<true>(V, c, lane, step0 + (unsigned)s);
    grid_barrier(counter, target);             // the next Theta is complete
  }
}

// ---------------------------------------------------------------------------------------------
struct GlmSampler {
  GlmModel* m = nullptr;
  cudaStream_t stream = nullptr;
  VecParams V;
  GradParams Gp;
  std::vector<void*> bufs;
  uint32_t L = 1; bool adapting = true; double stepSize0 = 1.0;
  int kind = 1; double scale0 = 1.0;
  uint32_t flags = 0;
  bool searched = false;
  uint64_t iter = 0, burnin = 0, lag = 1, max_samples = 0, rec_iters = 0;
  uint64_t iter_base = 0;   // iterations of earlier runs on this chain state (the Philox iteration word keeps counting across begins)
  uint64_t grad_evals = 0, leapfrogs = 0, launches = 0;
  bool advanced = false;   // ev0 / ev1 recorded at least once
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int nj_inst = 0;
  size_t smem = 0;
  // QSB_FLAG_TIME_KERNEL: event pairs around the gradient-contraction launches
  std::vector<cudaEvent_t> tev;
  size_t tev_used = 0;
  double t_accum = 0.0; uint64_t t_launches = 0;
  std::vector<double> temps;   // AnnealingKernel schedule (host: the launches of an iteration carry its temperature)
  unsigned* traj_counter = nullptr;   // grid-barrier counter of glm_traj_kernel
  bool fused_traj = false;            // the interior of a trajectory runs as one cooperative launch
};

static int pick_ksplit(int ntiles, int nblocks, int nsm) {
  if (const char* env = getenv("QSB_GLM_KSPLIT")) { int k = atoi(env); if (k >= 1 && k <= nblocks) return k; }   // tuning override
  int best = 1; double beste = 0.0;
  const int kmax = std::max(1, nblocks / 12);
  for (int k = 1; k <= kmax; ++k) {
    long ctas = (long)ntiles * k;
    long waves = (ctas + nsm - 1) / nsm;
    double eff = (double)ctas / (double)(waves * nsm);
    // per-CTA prologue/epilogue costs roughly one row block of time
    eff *= (double)(nblocks / k) / (double)(nblocks / k + 0.75);
    if (eff > beste * 1.005) { beste = eff; best = k; }
  }
  return best;
}

static bool pdl_enabled() { static const bool on = !getenv("QSB_NO_PDL"); return on; }
template <class... KArgs, class... Args>
static cudaError_t launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <int NJ, bool FULL, bool LP, bool TRIM = false, bool SK = false> static cudaError_t launch_grad_tf(const GradParams& P, size_t smem, cudaStream_t st) {
  if constexpr (!SK) { if (P.sk_U) return launch_grad_tf<NJ, FULL, LP, TRIM, true>(P, smem, st); }
  static bool attr_set[64] = {};   // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(glm_grad_kernel<NJ, FULL, LP, TRIM, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  const int grid = P.sk_U ? P.sk_U : P.ntiles * P.ksplit;
  return launch_ex(glm_grad_kernel<NJ, FULL, LP, TRIM, SK>, (unsigned)grid, GRAD_THREADS, smem, st, pdl_enabled(), P);
}
template <int NJ> static cudaError_t launch_grad_t(const GradParams& P, size_t smem, cudaStream_t st) {
  if constexpr (NJ == 13) {   // config c3's shape (d = 100): last tile half empty
    if (P.nj == NJ && P.d <= 8 * (NJ - 1) + 4)
      return P.want_lp ? launch_grad_tf<NJ, true, true, true>(P, smem, st) : launch_grad_tf<NJ, true, false, true>(P, smem, st);
  }
  if (P.nj == NJ) return P.want_lp ? launch_grad_tf<NJ, true, true>(P, smem, st) : launch_grad_tf<NJ, true, false>(P, smem, st);
  return P.want_lp ? launch_grad_tf<NJ, false, true>(P, smem, st) : launch_grad_tf<NJ, false, false>(P, smem, st);
}
static cudaError_t launch_grad(int nj_inst, const GradParams& P, size_t smem, cudaStream_t st) {
  switch (nj_inst) {
    case 2: return launch_grad_t<2>(P, smem, st);
    case 4: return launch_grad_t<4>(P, smem, st);
    case 13: return launch_grad_t<13>(P, smem, st);
    case 16: return launch_grad_t<16>(P, smem, st);
  }
  return cudaErrorInvalidValue;
}
template <int NJ> static cudaError_t launch_traj_t(const GradParams& P, const VecParams& V, int nsteps, unsigned* counter, size_t smem, cudaStream_t st) {
  auto go = [&](auto kern) -> cudaError_t {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
      attr_set[dev & 63] = true;
    }
    unsigned step0 = 0;
    void* args[] = {(void*)&P, (void*)&V, (void*)&nsteps, (void*)&step0, (void*)&counter};
    return cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)P.sk_U), dim3(GRAD_THREADS), args, smem, st);
  };
  if constexpr (NJ == 13) { if (P.nj == NJ && P.d <= 8 * (NJ - 1) + 4) return go(glm_traj_kernel<NJ, true, true>); }
  if (P.nj == NJ) return go(glm_traj_kernel<NJ, true, false>);
  return go(glm_traj_kernel<NJ, false, false>);
}
static cudaError_t launch_traj(int nj_inst, const GradParams& P, const VecParams& V, int nsteps, unsigned* counter, size_t smem, cudaStream_t st) {
  switch (nj_inst) {
    case 2: return launch_traj_t<2>(P, V, nsteps, counter, smem, st);
    case 4: return launch_traj_t<4>(P, V, nsteps, counter, smem, st);
    case 13: return launch_traj_t<13>(P, V, nsteps, counter, smem, st);
    case 16: return launch_traj_t<16>(P, V, nsteps, counter, smem, st);
  }
  return cudaErrorInvalidValue;
}
static int pick_nj_inst(int nj) { for (int c : {2, 4, 13, 16}) if (nj <= c) return c; return 0; }

// Balanced decomposition of one gradient launch (see glm_grad_kernel): the ntiles x nblocks row-block units, tile-major, are
// cut into one run per CTA such that every run costs the same.  The last tile may hold fewer chains than a full one; its
// idle warps are skipped, but a row block then does not cost in proportion to the live warps: measured on B200 (c3 shape,
// r02b) a row block takes 6.3 us with 12 live warps, 4.85 us with 8 and 3.56 us with 4 -- about 2.2 us of per-block latency
// that only thread-level parallelism hides -- hence the cost table 100 / 78 / 57 per warps-per-sub-partition 3 / 2 / 1.
// A cut one row block away from a tile boundary snaps to it (a single-block segment is all prologue).
// Output: start[ncta+1], first[ntiles] (first CTA touching a tile), np[ntiles] (partial sums of a tile = CTAs touching it);
// returns ncta.
int glm_unit_cost(int live_warps) { const int per_sp = (live_warps + 3) / 4; return per_sp >= 3 ? 100 : (per_sp == 2 ? 78 : 57); }
int glm_layout(int ntiles, int nblocks, uint32_t C, int nsm, std::vector<int>& start, std::vector<int>& first, std::vector<int>& np) {
  const long total = (long)ntiles * nblocks;
  const int WF = glm_unit_cost(CW);
  const int live_last = std::min<int>(CW, (int)((C - (uint32_t)(ntiles - 1) * CB + 7u) / 8u));
  const int w_last = glm_unit_cost(live_last);
  const long Ufull = (long)(ntiles - 1) * nblocks;
  const double Cfull = (double)Ufull * WF, T = Cfull + (double)nblocks * w_last;
  int ncta = (int)std::min<long>(nsm, std::max<long>(1, total / 12));
  std::vector<long> cut(ncta + 1, 0);
  for (int i = 1; i < ncta; ++i) {
    const double target = T * i / ncta;
    long u = target <= Cfull ? llround(target / WF) : Ufull + llround((target - Cfull) / w_last);
    const long r = u % nblocks;
    if (r == 1) u -= 1; else if (nblocks - r == 1) u += 1;     // no single-block segments
    cut[i] = std::min(std::max(u, cut[i - 1]), total);
  }
  cut[ncta] = total;
  start.clear();
  for (int i = 0; i < ncta; ++i) if (cut[i + 1] > cut[i]) start.push_back((int)cut[i]);   // CTAs with an empty run are dropped
  ncta = (int)start.size();
  start.push_back((int)total);
  first.assign(ntiles, 0); np.assign(ntiles, 0);
  for (int t = 0, i = 0; t < ntiles; ++t) {
    const long lo = (long)t * nblocks, hi = lo + nblocks;
    while (start[i + 1] <= lo) ++i;           // first run that reaches into the tile
    int j = i;
    while (j + 1 < ncta && start[j + 1] < hi) ++j;
    first[t] = i; np[t] = j - i + 1;
  }
  return ncta;
}

int glm_chains_per_tile() { return CB; }
int glm_rows_per_block() { return MB; }

int glm_model_create(GlmModel* m, const double* X, const uint8_t* y, int N, int d, double prior_sigma, std::string& why) {
  if (d > 128) { why = "dim <= 128 supported by the tensor-core GLM path"; return 1; }
  m->N = N; m->d = d; m->prior_sigma = prior_sigma;
  m->Dp = 8 * ((d + 7) / 8);
  int S = d; while (S % 8 != 4) ++S;
  m->perm = PERMX && m->Dp / 8 == 13 && d <= 8 * 12 + 4;     // the shape launch_grad_t runs with the TRIM instantiation
  if (m->perm) S = PERM_S;
  m->S = S;
  m->Npad = MB * ((N + MB - 1) / MB);
  std::vector<double> Xp((size_t)m->Npad * S, 0.0), yv(m->Npad, 2.0);
  for (int i = 0; i < N; ++i) {
    if (m->perm) { for (int j = 0; j < d; ++j) Xp[(size_t)i * S + (j & 7) * PERM_PJ + (j >> 3)] = X[(size_t)i * d + j]; }
    else memcpy(&Xp[(size_t)i * S], X + (size_t)i * d, sizeof(double) * d);
    yv[i] = y[i] ? 1.0 : 0.0;
  }
  if (cudaMalloc((void**)&m->Xp, Xp.size() * 8) != cudaSuccess || cudaMalloc((void**)&m->yv, yv.size() * 8) != cudaSuccess) { why = "cudaMalloc failed"; return 2; }
  cudaMemcpy(m->Xp, Xp.data(), Xp.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(m->yv, yv.data(), yv.size() * 8, cudaMemcpyHostToDevice);
  return 0;
}
// column-permuted copy of freshly uploaded rows (glm_model_update): logical dimension j of row i -> column (j % 8) PJ + j / 8
__global__ void permute_x_kernel(const double* __restrict__ Xraw, double* __restrict__ Xp, int N, int d, int S) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * d) return;
  const int i = (int)(idx / d), j = (int)(idx - (size_t)i * d);
  Xp[(size_t)i * S + (j & 7) * PERM_PJ + (j >> 3)] = Xraw[idx];
}
__global__ void ybin_to_double_kernel(const uint8_t* yb, double* yv, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) yv[i] = yb[i] ? 1.0 : 0.0;
}

// fully asynchronous on `stream` when X and y are pinned: two H2D copies and a conversion kernel, no host sync
int glm_model_update(GlmModel* m, const double* X, const uint8_t* y, void* stream, std::string& why) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (m->perm) {
    if (!m->Xraw && cudaMalloc((void**)&m->Xraw, (size_t)m->N * m->d * 8) != cudaSuccess) { why = "cudaMalloc failed"; return 2; }
    e = cudaMemcpyAsync(m->Xraw, X, (size_t)m->N * m->d * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      const size_t n = (size_t)m->N * m->d;
      permute_x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->Xraw, m->Xp, m->N, m->d, m->S);
      e = cudaGetLastError();
    }
  }
  else if (m->S == m->d) e = cudaMemcpyAsync(m->Xp, X, (size_t)m->N * m->d * 8, cudaMemcpyHostToDevice, st);
  else e = cudaMemcpy2DAsync(m->Xp, (size_t)m->S * 8, X, (size_t)m->d * 8, (size_t)m->d * 8, m->N, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { why = cudaGetErrorString(e); return 100 + (int)e; }
  if (!m->ybin_dev && cudaMalloc((void**)&m->ybin_dev, m->N) != cudaSuccess) { why = "cudaMalloc failed"; return 2; }
  e = cudaMemcpyAsync(m->ybin_dev, y, (size_t)m->N, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { why = cudaGetErrorString(e); return 100 + (int)e; }
  ybin_to_double_kernel<<<(m->N + 255) / 256, 256, 0, st>>>(m->ybin_dev, m->yv, m->N);
  e = cudaGetLastError();
  if (e != cudaSuccess) { why = cudaGetErrorString(e); return 100 + (int)e; }
  return 0;
}
void glm_model_destroy(GlmModel* m) {
  if (m->Xp) cudaFree(m->Xp);
  if (m->Xraw) cudaFree(m->Xraw);
  m->Xraw = nullptr;
  if (m->yv) cudaFree(m->yv);
  if (m->ybin_dev) cudaFree(m->ybin_dev);
  m->Xp = m->yv = nullptr; m->ybin_dev = nullptr;
}

template <class T> static bool dalloc(GlmSampler* s, T** p, size_t n) {
  *p = nullptr;
  if (cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return false;
  cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
  s->bufs.push_back(*p);
  return true;
}

static void glm_free_run_buffers(GlmSampler* s) {
  VecParams& V = s->V;
  for (double** p : {&V.samples, &V.samp_lp, &V.rec_state, &V.rec_dh, &V.rec_step, &V.rec_leap, &V.mom}) if (*p) { cudaFree(*p); *p = nullptr; }
  if (V.rec_accept) { cudaFree(V.rec_accept); V.rec_accept = nullptr; }
  s->rec_iters = 0; V.rec_iters = 0;
}

GlmSampler* glm_sampler_create(GlmModel* m, int kind, double stepSize, uint32_t numSteps, double scale, bool adapt, uint64_t seed,
                               uint32_t chain_begin, uint32_t numchains, uint32_t sample_chains, uint32_t flags, void* stream, std::string& why) {
  GlmSampler* s = new GlmSampler();
  s->m = m; s->stream = (cudaStream_t)stream; s->L = numSteps; s->adapting = adapt; s->stepSize0 = stepSize; s->flags = flags;
  s->kind = kind; s->scale0 = scale;
  memset(&s->V, 0, sizeof(s->V)); memset(&s->Gp, 0, sizeof(s->Gp));
  VecParams& V = s->V; GradParams& G = s->Gp;
  const size_t C = numchains, Dp = m->Dp;
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  G.Xp = m->Xp; G.yv = m->yv; G.S = m->S; G.d = m->d; G.nj = m->Dp / 8; G.nblocks = m->Npad / MB;
  G.ntiles = (int)((C + CB - 1) / CB); G.ksplit = pick_ksplit(G.ntiles, G.nblocks, nsm);
  G.C = numchains; G.Dp = m->Dp;
  G.sk_U = 0;
  // Default: the balanced decomposition (one equal-cost run of row blocks per SM).  QSB_FLAG_INVARIANT: the (chain tile x
  // 12 row splits) grid, whose grouping of a chain's row blocks into partial sums depends on nothing but the data shape, so
  // that a chain's draws are bit-identical for ANY partition of the chains over samplers / GPUs.
  const char* env_sk = getenv("QSB_GLM_STREAMK");   // tuning override (0 / 1)
  const bool invariant = env_sk ? atoi(env_sk) == 0 : (flags & QSB_FLAG_INVARIANT) != 0;
  int* sk_tab = nullptr;
  std::vector<int> sk_start, sk_first, sk_np;
  if (invariant) {
    if (!getenv("QSB_GLM_KSPLIT")) G.ksplit = std::max(1, std::min(12, G.nblocks / 12));
  } else if ((long)G.ntiles * G.nblocks < (1L << 30)) {
    G.sk_U = glm_layout(G.ntiles, G.nblocks, numchains, nsm, sk_start, sk_first, sk_np);
    G.ksplit = *std::max_element(sk_np.begin(), sk_np.end());   // most partial-sum slots a tile has
  }
  s->nj_inst = pick_nj_inst(G.nj);
  G.stages = STAGES;
  auto smem_for = [&](int nst) { return stage_doubles(m->S) * nst * 8 + BAR_DOUBLES * 8 + (THETA_SMEM ? (size_t)CB * theta_stride(m->Dp) * 8 : 0); };
  while (G.stages > 2 && smem_for(G.stages) > 227 * 1024) --G.stages;      // d = 128: the Theta tile leaves room for a 2-deep ring
  s->smem = smem_for(G.stages);
  if (s->smem > 227 * 1024) { why = "shared memory of the gradient kernel exceeds 227 KB for this dimension"; delete s; return nullptr; }
  V.key = RngKey{(uint32_t)seed, (uint32_t)(seed >> 32)}; V.chain_begin = chain_begin; V.C = numchains; V.d = m->d; V.Dp = m->Dp;
  V.T = 1.0; V.invT = 1.0;
  V.sk_U = G.sk_U; V.sk_nblocks = G.nblocks; V.sk_CB = CB;
  if (G.sk_U) {
    std::vector<int> tab(sk_start);
    tab.insert(tab.end(), sk_first.begin(), sk_first.end());
    tab.insert(tab.end(), sk_np.begin(), sk_np.end());
    if (!dalloc(s, &sk_tab, tab.size()) || cudaMemcpy(sk_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
      why = "cudaMalloc failed"; glm_sampler_destroy(s); return nullptr;
    }
    G.sk_start = sk_tab; G.sk_first = sk_tab + sk_start.size(); V.sk_np = sk_tab + sk_start.size() + sk_first.size();
  }
  V.ksplit = G.ksplit; V.prior_sigma = m->prior_sigma; V.L = numSteps; V.adapting = adapt ? 1 : 0; V.stepSize0 = stepSize; V.SC = sample_chains;
  bool ok = dalloc(s, &V.x, C * Dp) && dalloc(s, &V.g, C * Dp) && dalloc(s, &V.xs, C * Dp) && dalloc(s, &V.gs, C * Dp) && dalloc(s, &V.p, C * Dp);
  double *partG = nullptr, *partLp = nullptr;
  ok = ok && dalloc(s, &partG, (size_t)G.ksplit * C * Dp) && dalloc(s, &partLp, (size_t)G.ksplit * C);
  ok = ok && dalloc(s, &V.lp, C) && dalloc(s, &V.H0, C) && dalloc(s, &V.eps, C) && dalloc(s, &V.prevlp, C);
  ok = ok && dalloc(s, &V.da_gbar, C) && dalloc(s, &V.da_xbar, C) && dalloc(s, &V.da_x0, C) && dalloc(s, &V.da_lastx, C) && dalloc(s, &V.da_gamma, C) && dalloc(s, &V.da_k, C);
  ok = ok && dalloc(s, &V.health, 2);
  ok = ok && dalloc(s, &V.direction, C) && dalloc(s, &V.done, C) && dalloc(s, &V.n_active, 1) && dalloc(s, &V.made, C) && dalloc(s, &V.acc, C) && dalloc(s, &V.status, 1);
  V.kind = kind;
  if (kind == 2) ok = ok && dalloc(s, &V.dr_mean, C * Dp) && dalloc(s, &V.dr_m2, C * Dp) && dalloc(s, &V.dr_s0, C) && dalloc(s, &V.dr_mult, C) && dalloc(s, &V.dr_n, C);
  if (kind == 3) ok = ok && dalloc(s, &V.ha_scale, C);
  ok = ok && cudaEventCreate(&s->ev0) == cudaSuccess && cudaEventCreate(&s->ev1) == cudaSuccess;
  if (!ok || !s->nj_inst) { why = !s->nj_inst ? "unsupported dim" : "cudaMalloc failed"; glm_sampler_destroy(s); return nullptr; }
  G.partG = partG; G.partLp = partLp; V.partG = partG; V.partLp = partLp;
  G.theta = V.xs;
  // QSB_FUSED_TRAJ=1: the interior of a trajectory as ONE cooperative launch (glm_traj_kernel).  Measured SLOWER than the
  // separate launches with programmatic dependent launch (r02e, same box: 6.81 vs 7.01 M leapfrog steps/s at 8 192 chains,
  // 6.13 vs 6.28 M at 1 024): two grid-wide barriers per step cost more than two PDL-overlapped kernel boundaries.  Kept
  // selectable, off by default.
  if (G.sk_U && kind == 1 && numSteps >= 2 && getenv("QSB_FUSED_TRAJ") && atoi(getenv("QSB_FUSED_TRAJ")) != 0) {
    int coop = 0, occ = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (coop && G.sk_U <= nsm && dalloc(s, &s->traj_counter, 1)) s->fused_traj = true;
    (void)occ;
  }
  return s;
}

void glm_sampler_destroy(GlmSampler* s) {
  if (!s) return;
  cudaStreamSynchronize(s->stream);
  glm_free_run_buffers(s);
  for (void* p : s->bufs) cudaFree(p);
  for (cudaEvent_t e : s->tev) cudaEventDestroy(e);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
}

#define GCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { why = std::string(#call) + ": " + cudaGetErrorString(e_); return 100 + (int)e_; } } while (0)

static void glm_collect_times(GlmSampler* s) {
  if (!s->tev_used) return;
  cudaStreamSynchronize(s->stream);
  for (size_t i = 0; i + 1 < s->tev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->tev[i], s->tev[i + 1]) == cudaSuccess) { s->t_accum += ms * 1e-3; ++s->t_launches; }
    else (void)cudaGetLastError();
  }
  s->tev_used = 0;
}

template <int MODE> static cudaError_t launch_vec(GlmSampler* s, const VecParams& V) {
  ++s->launches;
  return launch_ex(glm_vec_kernel<MODE>, (unsigned)(((size_t)V.C * 32 + 255) / 256), 256u, 0, s->stream, pdl_enabled(), V);
}
static cudaError_t run_grad(GlmSampler* s, const double* theta, int want_lp, int want_grad = 1) {
  GradParams G = s->Gp;
  G.theta = theta; G.want_lp = want_lp; G.want_grad = want_grad;
  ++s->launches; ++s->grad_evals;
  if (!(s->flags & QSB_FLAG_TIME_KERNEL)) return launch_grad(s->nj_inst, G, s->smem, s->stream);
  if (s->tev_used + 2 > s->tev.size()) {
    if (s->tev.size() >= 8192) { glm_collect_times(s); }
    else for (int i = 0; i < 512; ++i) { cudaEvent_t e; cudaEventCreate(&e); s->tev.push_back(e); }
  }
  cudaEventRecord(s->tev[s->tev_used], s->stream);
  cudaError_t r = launch_grad(s->nj_inst, G, s->smem, s->stream);
  cudaEventRecord(s->tev[s->tev_used + 1], s->stream);
  s->tev_used += 2;
  return r;
}

static int reset_chain_state(GlmSampler* s, std::string& why) {
  VecParams& V = s->V;
  const size_t C = V.C;
  std::vector<double> e(C, s->stepSize0);
  GCU(cudaMemcpyAsync(V.eps, e.data(), 8 * C, cudaMemcpyHostToDevice, s->stream));
  GCU(cudaStreamSynchronize(s->stream));
  for (double* p : {V.da_gbar, V.da_xbar, V.da_x0, V.da_lastx, V.da_gamma}) GCU(cudaMemsetAsync(p, 0, 8 * C, s->stream));
  GCU(cudaMemsetAsync(V.da_k, 0, 4 * C, s->stream));
  GCU(cudaMemsetAsync(V.done, 0, 4 * C, s->stream)); GCU(cudaMemsetAsync(V.direction, 0, 4 * C, s->stream));
  GCU(cudaMemsetAsync(V.made, 0, 8 * C, s->stream)); GCU(cudaMemsetAsync(V.acc, 0, 8 * C, s->stream));
  GCU(cudaMemsetAsync(V.status, 0, 4, s->stream));
  GCU(cudaMemsetAsync(V.health, 0, 16, s->stream));
  if (s->kind == 2) {
    std::vector<double> sc(C, s->scale0), one(C, 1.0);
    GCU(cudaMemcpyAsync(V.dr_s0, sc.data(), 8 * C, cudaMemcpyHostToDevice, s->stream));
    GCU(cudaMemcpyAsync(V.dr_mult, one.data(), 8 * C, cudaMemcpyHostToDevice, s->stream));
    GCU(cudaStreamSynchronize(s->stream));
    GCU(cudaMemsetAsync(V.dr_n, 0, 4 * C, s->stream));
    GCU(cudaMemsetAsync(V.dr_mean, 0, 8 * C * V.Dp, s->stream)); GCU(cudaMemsetAsync(V.dr_m2, 0, 8 * C * V.Dp, s->stream));
  }
  if (s->kind == 3) {
    std::vector<double> sc(C, s->scale0);
    GCU(cudaMemcpyAsync(V.ha_scale, sc.data(), 8 * C, cudaMemcpyHostToDevice, s->stream));
    GCU(cudaStreamSynchronize(s->stream));
  }
  s->searched = false; s->iter = 0; s->iter_base = 0; s->grad_evals = 0; s->leapfrogs = 0;
  return 0;
}

// lp and gradient at x: the trace's initial logprob (trace.t:612-623) and HMC's first gradient (hmc.t:254)
static int eval_at_x(GlmSampler* s, std::string& why) {
  GCU(run_grad(s, s->V.x, 1));
  GCU(launch_vec<V_FINISH_EVAL>(s, s->V));
  return 0;
}

int glm_sampler_init(GlmSampler* s, std::string& why) {
  if (int rc = reset_chain_state(s, why)) return rc;
  GCU(launch_vec<V_INIT_SAMPLE>(s, s->V));
  return eval_at_x(s, why);
}

int glm_sampler_set_state(GlmSampler* s, const double* x, std::string& why) {
  if (int rc = reset_chain_state(s, why)) return rc;
  const VecParams& V = s->V;
  std::vector<double> h((size_t)V.C * V.Dp, 0.0);
  for (size_t c = 0; c < V.C; ++c) memcpy(&h[c * V.Dp], x + c * V.d, 8 * V.d);
  GCU(cudaMemcpyAsync(V.x, h.data(), 8 * h.size(), cudaMemcpyHostToDevice, s->stream));
  GCU(cudaMemcpyAsync(V.xs, h.data(), 8 * h.size(), cudaMemcpyHostToDevice, s->stream));
  GCU(cudaStreamSynchronize(s->stream));
  return eval_at_x(s, why);
}

int glm_sampler_get_state(GlmSampler* s, double* x, std::string& why) {
  const VecParams& V = s->V;
  std::vector<double> h((size_t)V.C * V.Dp);
  GCU(cudaMemcpyAsync(h.data(), V.x, 8 * h.size(), cudaMemcpyDeviceToHost, s->stream));
  GCU(cudaStreamSynchronize(s->stream));
  for (size_t c = 0; c < V.C; ++c) memcpy(x + c * V.d, &h[c * V.Dp], 8 * V.d);
  return 0;
}

int glm_sampler_begin(GlmSampler* s, uint64_t max_samples, uint64_t burnin, uint64_t lag, uint64_t numiters, std::string& why) {
  VecParams& V = s->V;
  glm_free_run_buffers(s);
  s->iter_base += s->iter;
  s->iter = 0; s->burnin = burnin; s->lag = lag; s->max_samples = max_samples;
  if (max_samples > 0 && (s->flags & QSB_FLAG_MOMENTS)) {
    const size_t bytes = 8 * (size_t)MOM_STATS * V.SC * V.d;
    GCU(cudaMalloc((void**)&V.mom, bytes));
    GCU(cudaMemsetAsync(V.mom, 0, bytes, s->stream));
    V.mom_half = max_samples / 2; V.mom_batch = (uint32_t)std::max<uint64_t>(1, (uint64_t)std::floor(std::sqrt((double)max_samples)));
  } else if (max_samples > 0) {
    GCU(cudaMalloc((void**)&V.samples, 8 * max_samples * V.SC * V.d));
    GCU(cudaMalloc((void**)&V.samp_lp, 8 * max_samples * V.SC));
  }
  if (s->flags & QSB_FLAG_RECORD) {
    if (numiters == 0) { why = "QSB_FLAG_RECORD needs numiters"; return 2; }
    const size_t C = V.C, d = V.d;
    s->rec_iters = numiters; V.rec_iters = numiters;
    GCU(cudaMalloc((void**)&V.rec_state, 8 * C * numiters * d));
    GCU(cudaMalloc((void**)&V.rec_dh, 8 * C * numiters));
    GCU(cudaMalloc((void**)&V.rec_step, 8 * C * numiters));
    GCU(cudaMalloc((void**)&V.rec_accept, 4 * C * numiters));
    if (s->flags & QSB_FLAG_RECORD_LEAP) {
      GCU(cudaMalloc((void**)&V.rec_leap, 8 * C * numiters * s->L * d));
      GCU(cudaMemset(V.rec_leap, 0, 8 * C * numiters * s->L * d));
    }
  }
  return 0;
}

// searchForInitialStepSize (hmc.t:345-391): every probe is one leapfrog step of all chains still searching;
// the host only loops until no chain is left (one 4-byte read per probe; initialisation only)
static double temp_at(const GlmSampler* s, uint64_t it) {
  if (s->temps.empty()) return 1.0;
  return s->temps[it < s->temps.size() ? it : s->temps.size() - 1];
}

int glm_sampler_set_temperatures(GlmSampler* s, const double* temps, uint64_t n, std::string&) {
  s->temps.clear();
  if (temps && n) s->temps.assign(temps, temps + n);
  return 0;
}

static int step_size_search(GlmSampler* s, std::string& why) {
  VecParams V = s->V;
  V.T = temp_at(s, 0); V.invT = 1.0 / V.T;
  for (uint32_t probe = 0; probe < 4096; ++probe) {
    V.probe = probe;
    GCU(cudaMemsetAsync(V.n_active, 0, 4, s->stream));
    GCU(launch_vec<V_PROBE_FIRST>(s, V));
    GCU(run_grad(s, V.xs, 1));
    GCU(launch_vec<V_PROBE_LAST>(s, V));
    int n_active = 0;
    GCU(cudaMemcpyAsync(&n_active, V.n_active, 4, cudaMemcpyDeviceToHost, s->stream));
    GCU(cudaStreamSynchronize(s->stream));
    if (n_active == 0) break;
  }
  int st = 0;
  GCU(cudaMemcpy(&st, V.status, 4, cudaMemcpyDeviceToHost));
  if (st & 1) { why = "HMC step size search failed because logprob became NaN"; return 20; }
  if (st & 2) { why = "HMC step size search diverged to infinity (Is your probability distribution flat?)"; return 21; }
  if (st & 4) { why = "HMC step size search collapsed to zero (Is your probability distribution discontinuous?)"; return 22; }
  return 0;
}

int glm_sampler_advance(GlmSampler* s, uint64_t iters, std::string& why) {
  if (s->rec_iters && s->iter + iters > s->rec_iters) { why = "record buffer holds numiters iterations only"; return 2; }
  if (s->kind != 1) {   // Drift / HARM: propose, score (first contraction + log-likelihood only), accept
    GCU(cudaEventRecord(s->ev0, s->stream));
    for (uint64_t k = 0; k < iters; ++k) {
      const uint64_t it = s->iter + k;
      VecParams V = s->V;
      V.iter = (uint32_t)(s->iter_base + it); V.rec_row = it;
      V.T = temp_at(s, it); V.invT = 1.0 / V.T;
      GCU(launch_vec<V_MH_PROPOSE>(s, V));
      GCU(run_grad(s, V.xs, 1, 0));
      --s->grad_evals;   // a log-density evaluation, not a gradient
      V.keep = (it >= s->burnin && it % s->lag == 0 && (V.samples || V.mom)) ? 1 : 0;
      V.sidx = it / s->lag - (s->burnin + s->lag - 1) / s->lag;
      if (V.sidx >= s->max_samples) V.keep = 0;
      GCU(launch_vec<V_MH_ACCEPT>(s, V));
    }
    GCU(cudaEventRecord(s->ev1, s->stream));
    s->iter += iters; s->advanced = true;
    return 0;
  }
  if (!s->searched) {
    if (temp_at(s, s->iter) != 1.0) {   // first dual-trace evaluation under the first temperature: gradient and search baseline
      VecParams V0 = s->V;
      V0.T = temp_at(s, s->iter); V0.invT = 1.0 / V0.T;
      GCU(launch_vec<V_FINISH_EVAL>(s, V0));
    }
    ++s->grad_evals;   // the reference's first traceUpdate (hmc.t:254); its value was computed with the initial logprob
    if (s->adapting) if (int rc = step_size_search(s, why)) return rc;
    s->searched = true;
  }
  GCU(cudaEventRecord(s->ev0, s->stream));
  for (uint64_t k = 0; k < iters; ++k) {
    const uint64_t it = s->iter + k;
    VecParams V = s->V;
    V.iter = (uint32_t)(s->iter_base + it); V.rec_row = it;
    V.T = temp_at(s, it); V.invT = 1.0 / V.T;
    GCU(launch_vec<V_TRANS_FIRST>(s, V));
    uint32_t st_begin = 0;
    if (s->fused_traj && !(s->flags & QSB_FLAG_TIME_KERNEL)) {
      // leapfrog steps 0 .. L-2 (gradient + update) as one cooperative launch; the last gradient (which also evaluates the
      // log-likelihood) and the accept step follow as before.  With QSB_FLAG_TIME_KERNEL the launches stay separate so that
      // the event pairs bracket the gradient contraction alone.
      GradParams Gp = s->Gp;
      Gp.theta = V.xs; Gp.want_lp = 0; Gp.want_grad = 1;
      GCU(cudaMemsetAsync(s->traj_counter, 0, sizeof(unsigned), s->stream));
      GCU(launch_traj(s->nj_inst, Gp, V, (int)s->L - 1, s->traj_counter, s->smem, s->stream));
      ++s->launches; s->grad_evals += s->L - 1;
      st_begin = s->L - 1;
    }
    for (uint32_t st = st_begin; st < s->L; ++st) {
      const bool last = st + 1 == s->L;
      V.step = st;
      GCU(run_grad(s, V.xs, last ? 1 : 0));
      if (!last) GCU(launch_vec<V_TRANS_MID>(s, V));
      else {
        V.keep = (it >= s->burnin && it % s->lag == 0 && (V.samples || V.mom)) ? 1 : 0;
        V.sidx = it / s->lag - (s->burnin + s->lag - 1) / s->lag;
        if (V.sidx >= s->max_samples) V.keep = 0;
        GCU(launch_vec<V_TRANS_LAST>(s, V));
      }
    }
    s->leapfrogs += s->L;
  }
  GCU(cudaEventRecord(s->ev1, s->stream));
  s->iter += iters; s->advanced = true;
  return 0;
}

void glm_sampler_moments_view(GlmSampler* s, const double** mom, uint64_t* half, uint32_t* batch) { *mom = s->V.mom; *half = s->V.mom_half; *batch = s->V.mom_batch; }
void glm_sampler_sample_view(GlmSampler* s, const double** samples, const double** samp_lp, int* G, uint32_t* SC, int* retdim) {
  *samples = s->V.samples; *samp_lp = s->V.samp_lp; *G = 32; *SC = s->V.SC; *retdim = s->V.d;
}

int glm_sampler_stats(GlmSampler* s, uint64_t* made, uint64_t* acc, uint64_t* grad_evals, uint64_t* leapfrogs, uint64_t* launches,
                      double* seconds, std::string& why) {
  const VecParams& V = s->V;
  std::vector<unsigned long long> m(V.C), a(V.C);
  GCU(cudaStreamSynchronize(s->stream));
  GCU(cudaMemcpy(m.data(), V.made, 8 * V.C, cudaMemcpyDeviceToHost));
  GCU(cudaMemcpy(a.data(), V.acc, 8 * V.C, cudaMemcpyDeviceToHost));
  *made = 0; *acc = 0;
  for (size_t c = 0; c < V.C; ++c) { *made += m[c]; *acc += a[c]; }
  *grad_evals = s->grad_evals * V.C; *leapfrogs = s->leapfrogs * V.C; *launches = s->launches;
  float ms = 0.f;
  *seconds = 0.0;
  if (s->advanced) { if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) *seconds = ms * 1e-3; else (void)cudaGetLastError(); }
  return 0;
}

void glm_sampler_health(GlmSampler* s, uint64_t* nonfinite, uint64_t* divergent) {
  unsigned long long h[2] = {0, 0};
  cudaStreamSynchronize(s->stream);
  if (cudaMemcpy(h, s->V.health, 16, cudaMemcpyDeviceToHost) != cudaSuccess) (void)cudaGetLastError();
  *nonfinite = h[0]; *divergent = h[1];
}

int glm_sampler_fetch_record(GlmSampler* s, double* state, int32_t* accept, double* dh, double* step, double* leap, std::string& why) {
  const VecParams& V = s->V;
  if (!V.rec_state) { why = "sampler was not created with QSB_FLAG_RECORD"; return 2; }
  GCU(cudaStreamSynchronize(s->stream));
  const size_t C = V.C, d = V.d, T = s->rec_iters;
  if (state) GCU(cudaMemcpy(state, V.rec_state, 8 * C * T * d, cudaMemcpyDeviceToHost));
  if (accept) GCU(cudaMemcpy(accept, V.rec_accept, 4 * C * T, cudaMemcpyDeviceToHost));
  if (dh) GCU(cudaMemcpy(dh, V.rec_dh, 8 * C * T, cudaMemcpyDeviceToHost));
  if (step) GCU(cudaMemcpy(step, V.rec_step, 8 * C * T, cudaMemcpyDeviceToHost));
  if (leap) { if (!V.rec_leap) { why = "leapfrog positions need QSB_FLAG_RECORD_LEAP"; return 2; } GCU(cudaMemcpy(leap, V.rec_leap, 8 * C * T * s->L * d, cudaMemcpyDeviceToHost)); }
  return 0;
}

void glm_sampler_set_timing(GlmSampler* s, bool on) {
  if (on) s->flags |= QSB_FLAG_TIME_KERNEL; else s->flags &= ~(uint32_t)QSB_FLAG_TIME_KERNEL;
}

int glm_sampler_kernel_time(GlmSampler* s, double* seconds, uint64_t* launches, std::string& why) {
  glm_collect_times(s);
  *seconds = s->t_accum; *launches = s->t_launches;
  s->t_accum = 0.0; s->t_launches = 0;
  return 0;
}

// Test hook: one HMCKernel:step of n injected states through the product's gradient kernel and vector arithmetic.
int glm_debug_leapfrog(GlmModel* m, int n, const double* x, const double* p, const double* g, const double* eps, int last,
                       double* x_out, double* p_out, double* g_out, double* lp_out, std::string& why) {
  GlmSampler* s = glm_sampler_create(m, 1, 1.0, 1, 1.0, false, 0, 0, (uint32_t)n, (uint32_t)n, 0, nullptr, why);
  if (!s) return 5;
  int rc = glm_sampler_set_state(s, x, why);   // also evaluates the gradient at x (used when g == nullptr)
  auto run = [&]() -> int {
    VecParams V = s->V;
    const size_t Dp = V.Dp, d = V.d;
    std::vector<double> h((size_t)n * Dp, 0.0);
    auto up = [&](const double* src, double* dst) -> cudaError_t {
      for (int c = 0; c < n; ++c) memcpy(&h[c * Dp], src + c * d, 8 * d);
      return cudaMemcpy(dst, h.data(), 8 * h.size(), cudaMemcpyHostToDevice);
    };
    GCU(up(p, V.p));
    if (g) GCU(up(g, V.g));
    GCU(cudaMemcpy(V.eps, eps, 8 * (size_t)n, cudaMemcpyHostToDevice));
    GCU(launch_vec<V_DEBUG_FIRST>(s, V));
    GCU(run_grad(s, V.xs, last ? 1 : 0));
    V.keep = last ? 1 : 0;
    GCU(launch_vec<V_DEBUG_LAST>(s, V));
    GCU(cudaStreamSynchronize(s->stream));
    auto down = [&](const double* src, double* dst) -> cudaError_t {
      cudaError_t e = cudaMemcpy(h.data(), src, 8 * h.size(), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) for (int c = 0; c < n; ++c) memcpy(dst + c * d, &h[c * Dp], 8 * d);
      return e;
    };
    if (x_out) GCU(down(V.xs, x_out));
    if (p_out) GCU(down(V.p, p_out));
    if (g_out) GCU(down(V.gs, g_out));
    if (lp_out) GCU(cudaMemcpy(lp_out, V.lp, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    return 0;
  };
  if (!rc) rc = run();
  glm_sampler_destroy(s);
  return rc;
}

int glm_logp_grad(GlmModel* m, const double* x, int n, double* lp_dual, double* grad, double* lp_float, std::string& why) {
  GlmSampler* s = glm_sampler_create(m, 1, 1.0, 1, 1.0, false, 0, 0, (uint32_t)n, (uint32_t)n, 0, nullptr, why);
  if (!s) return 5;
  int rc = glm_sampler_set_state(s, x, why);
  if (!rc) {
    const VecParams& V = s->V;
    std::vector<double> g((size_t)n * V.Dp), lp(n);
    cudaError_t e = cudaMemcpy(g.data(), V.g, 8 * g.size(), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(lp.data(), V.lp, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { why = cudaGetErrorString(e); rc = 100 + (int)e; }
    else {
      for (int c = 0; c < n; ++c) {
        if (grad) memcpy(grad + (size_t)c * V.d, &g[(size_t)c * V.Dp], 8 * V.d);
        if (lp_dual) lp_dual[c] = lp[c];
        if (lp_float) lp_float[c] = lp[c];
      }
    }
  }
  glm_sampler_destroy(s);
  return rc;
}

}  // namespace qsb
