// Chain-parallel MCMC engine: the fused per-chain transition kernels that replace the reference's
// doMCMC loop (qs/mcmc.t:50-90) with its HMC (qs/hmc.t), drift (qs/drift.t), HARM (qs/harm.t) and
// mixture (qs/mcmc.t:199-291) kernels, for many independent chains at once.
//
// Mapping.  A chain is owned by a *group* of G lanes: G = 1 (thread per chain, small n: the whole
// state of a transition lives in registers) or G = 32 (warp per chain, n up to 1024: lane l owns
// components l, l+32, ...; reductions are warp shuffles; all global accesses are coalesced over
// the components of one chain).  The same source serves both: lane-local arrays of NPL elements,
// element e <-> component e*G + lane.
//
// Data layout in HBM (flat structure-of-arrays, fp64; replaces the per-chain trace of
// qs/trace.t:523-540 and the kernels' S.Vector members, hmc.t:82-101, drift.t:87-98, harm.t:30-41):
//   x, g          unconstrained positions / cached gradient:  G=1: [n][C] (chain-minor),  G=32: [C][n]
//   lp            [C] currTrace.logprob
//   eps, da_*     [C] HMC step size + DualAverage state        (hmc.t:14-48)
//   dr_mean,dr_m2 like x: RunningVar per component; dr_n, dr_s0, dr_mult [C]   (drift.t:15-50, 274-300)
//   ha_scale      [C]                                           (harm.t:30-41)
//   made, acc     [nsub][C] KernelPropStats                     (mcmc.t:116-125)
// The drift kernel's per-component `scale` is not stored: it is recomputed from (dr_m2, dr_n) with
// the operations of RunningVar:getStdDev, which is what drift.t:288-298 assigns each iteration.
#pragma once
#include <cstdint>
#include "distrib.cuh"
#include "mbar.cuh"
#include "moments.cuh"
#include "philox.cuh"

namespace qsb {

enum { FAM_INDEP = 1, FAM_GAUSS_PREC = 2, FAM_LOGISTIC = 3, FAM_EIGHT_SCHOOLS = 4, FAM_FUNNEL = 5 };
enum { K_HMC = 1, K_DRIFT = 2, K_HARM = 3, K_TRACEMH = 4 };
enum { FLAG_HMC_INIT = 1, FLAG_GRAD_VALID = 2 };
enum { ERR_NAN_SEARCH = 1, ERR_SEARCH_INF = 2, ERR_SEARCH_ZERO = 4 };
static constexpr int MAX_SUB = 8;

struct ModelDev {
  int family, dim, nobs, ret_mode;
  double ret_div, prior_sigma;
  const int* erp_kind; const double* p0; const double* p1;   // INDEP
  // INDEP: optional Gaussian factor on the value of a scalar choice, qs.factor(gaussian.logprob(value, fmu, fsigma)) ==
  // qs.gaussian.observe(value, fmu, fsigma) (trace.t:1099-1107; erp.t:687-696; testsuite.t:394-442); fsigma <= 0: none
  const double* fmu; const double* fsigma;
  // INDEP: optional hard constraint after scalar choice i, qs.condition(value_i > cond_lo[i] and value_i < cond_hi[i])
  // (trace.t:794-796, 1103-1107); nullptr: the program has none
  const double* cond_lo; const double* cond_hi;
  // INDEP, dirichlet components (thread per chain; warp per chain through the scratch evaluator): choice length K, sum of the later alphas of the choice,
  // and the ERP index of every component (= the Philox slot of its forward sample; equals the component index without vector choices)
  const int* erp_len; const double* erp_asum; const int* erp_slot;
  int has_vector;
  // INDEP, latent parameters (thread per chain; warp per chain through the scratch evaluator): parameter q of choice i = f(p_q[i] + lat_b[2i+q] * value of
  // choice lat_ref[2i+q]) with f = identity / exp (lat_tf); lat_ref < 0: constant; lat_ref == nullptr: none in the program
  const int* lat_ref; const double* lat_b; const int* lat_tf;
  // DriftKernel{lexicalScaleSharing} (drift.t:101-115, 208-235): scale adapter of every component.  INDEP: lex_map;
  // GAUSS_PREC: one; EIGHT_SCHOOLS: mu, logtau, theta-loop; FUNNEL: v, x-loop
  const int* lex_map; int n_adapters;
  // TraceMHKernel: the random choices of the program (INDEP: first component of every choice; elsewhere one per component)
  const int* erp_first; int n_choices;
  // EIGHT_SCHOOLS observations, precomputed once per model with the same device operations the
  // per-term logprob would use (bit-identical): cy = 1.8378770664093453 + 2*log(sigma), s2 = sigma*sigma
  const double* y; const double* cy; const double* s2; const double* is2;   // is2 = 1/s2
};

struct EngineParams {
  ModelDev m;
  RngKey key;
  uint32_t chain_begin, C;
  int d, retdim;
  int nsub;
  int kind[MAX_SUB], doAdapt[MAX_SUB];
  double stepSize0[MAX_SUB], scale0[MAX_SUB], weight[MAX_SUB];
  uint32_t numSteps[MAX_SUB];
  double Sc[16 * 16];               // GAUSS_PREC: A + A^T padded to stride NPL, read from the constant bank
  // chain state
  double *x, *g, *lp;
  uint8_t* flags;
  double *eps, *da_gbar, *da_xbar, *da_x0, *da_lastx, *da_gamma;
  double *hmc_dualT, *hmc_gT;       // [C] dualTrace.temperature (hmc.t:266) and the temperature the cached gradient was computed under
  uint32_t* da_k;
  double *dr_mean, *dr_m2, *dr_s0, *dr_mult;
  uint32_t* dr_n;
  int dr_lexical;                   // lexicalScaleSharing: ScaleAdapter {scale, RunningVar} per adapter and chain, [adapter][C]
  double *ls_scale, *ls_mean, *ls_m2;
  uint32_t* ls_n;
  double* ha_scale;
  unsigned long long *made, *acc;   // [nsub][C]
  unsigned long long* counters;     // [0] gradient evaluations, [1] leapfrog steps (inside :next)
  int* status;                      // sticky error bits
  // run window.  Iterations are numbered continuously over the life of the chain state (the Philox iteration word and
  // HARM's 1/(iter+1) gain never restart, so a second run on the same sampler does not replay its random numbers);
  // it_origin = number of the first iteration of the current run: burn-in, lag, sample slots and the temperature
  // schedule count from it.
  uint64_t it_begin, n_iters, burnin, lag, it_origin;
  // AnnealingKernel (mcmc.t:297-319): temperature of iteration i = temps[min(i, ntemps-1)]; nullptr = 1
  const double* temps; uint64_t ntemps;
  // kept samples
  double *samples, *samp_lp;
  uint32_t sample_chains;
  uint64_t max_samples;
  // QSB_FLAG_MOMENTS: streaming statistics of the kept draws instead of the draws (see mom_update); nullptr = off
  double* mom; uint64_t mom_half; uint32_t mom_batch;
  // per-iteration record (parity tests)
  double *rec_state, *rec_dh, *rec_step, *rec_leap;
  int *rec_accept, *rec_kidx;
  uint64_t rec_iters;
  // INDEP programs with vector (dirichlet) choices or latent parameters in the warp-per-chain mapping: VS_ARRAYS arrays of
  // vs_dpad doubles per warp of the grid (see indep_general_warp); nullptr otherwise
  double* vscratch; int vs_dpad;
};
static constexpr int VS_ARRAYS = 6;   // x, g, and the evaluator's four work arrays

// ---------------------------------------------------------------------------------------------
// G lanes per chain, G a power of two in 1..32: a warp holds 32/G chains side by side.  Reductions are xor butterflies over
// the low log2(G) lane bits, broadcasts are shuffles of width G; both name only the G lanes of the chain in their mask, so
// the chains of a warp may sit in different branches (mixture kernels, step-size search loops of different length).
template <int G> struct Grp {
  static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "lanes per chain: a power of two up to 32");
  static __device__ __forceinline__ int lane() { return G == 1 ? 0 : (int)(threadIdx.x & (unsigned)(G - 1)); }
  static __device__ __forceinline__ unsigned mask() {
    if (G == 32) return 0xffffffffu;
    return ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
  }
  static __device__ __forceinline__ double sum(double v) {
    if (G > 1) {
      const unsigned m = mask();
#pragma unroll
      for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    }
    return v;
  }
  static __device__ __forceinline__ double bcast(double v, int src) {
    if (G > 1) return __shfl_sync(mask(), v, src, G);
    return v;
  }
};

template <int G> __device__ __forceinline__ size_t vidx(const EngineParams& P, int i, uint32_t c) {
  return G == 1 ? (size_t)i * P.C + c : (size_t)c * P.d + i;
}

#define QSB_FOR_E _Pragma("unroll") for (int e = 0; e < NPL; ++e)
// GAUSS_PREC: where the transition kernels read S = A + A^T from: 1 = the copy staged in shared memory (broadcast LDS),
// 0 = the kernel parameter itself (constant-bank operands of the DFMAs, no load instructions)
#ifndef QSB_GP_SMEM
#define QSB_GP_SMEM 1
#endif

// One ERP site of a dual trace: prior logprob of the constrained value, log-Jacobian, and
// d(logprob + logJ)/dx  (erp.t:623-640).  Not inlined: the INDEP family is not a throughput path
// and the four-way kind switch would otherwise be replicated at every unrolled component.
static __device__ __noinline__ void erp_eval(int kind, double x, double p0, double p1, double fmu, double fsigma, double& lpy, double& lj, double& g, double& yout) {
  double dydx, dlpdy, dlj;
  double y = erp_fwd(kind, x, p0, p1, &dydx);
  yout = y;
  lpy = erp_prior_lp(kind, y, p0, p1, &dlpdy);
  if (fsigma > 0.0) {   // the factor statement that follows the choice: likelihood term on the constrained value
    lpy = lpy + gaussian_lp(y, fmu, fsigma);
    dlpdy = dlpdy + gaussian_dlp_dx(y, fmu, fsigma);
  }
  lj = erp_logjac(kind, x, p0, p1, &dlj);
  g = dlpdy * dydx + dlj;
}
// Parameter q of choice i of a program with latent parameters, given the value yj of the choice it depends on.
static __device__ __noinline__ void latent_param(const ModelDev& m, int i, int q, double yj, double& p, double& dpdy) {
  const double b = m.lat_b[2 * i + q];
  const double a = (q ? m.p1[i] : m.p0[i]) + b * yj;
  if (m.lat_tf[2 * i + q]) { p = exp(a); dpdy = b * p; } else { p = a; dpdy = b; }
}
static __device__ __noinline__ double erp_fwd_noinline(int kind, double x, double p0, double p1) { return erp_fwd(kind, x, p0, p1, nullptr); }
static __device__ __noinline__ double dirichlet_component_noinline(SimplexCarry& c, int k, int K, double x, double alpha, double asum_after, double& g) {
  return dirichlet_component(c, k, K, x, alpha, asum_after, &g);
}

// INDEP program with latent parameters, one thread per chain, rolled loops over local arrays (not a throughput path).
// Forward in program order: value, resolved parameters, log-density terms and the adjoint each term sends to the values
// its parameters were computed from; then g_i = (dlp_i/dy_i + adjoint_i) dy_i/dx_i + dlogJ_i/dx_i  -- the reverse sweep of
// the reference's tape (ad.t) over  sum_i [lp_i(y_i; p(y_parents)) + logJ_i(x_i)]  (erp.t:623-640).
// work: 4 * dim doubles (values, value adjoints, dlp/dy, dy/dx) -- a local array in the thread-per-chain mapping, this warp's
// slice of EngineParams::vscratch in the warp-per-chain mapping (where lane 0 runs this function for its chain).
static __device__ __noinline__ double indep_latent_eval(const ModelDev& m, const double* x, double* g, double* lp_float, bool grad, double* cond_bad, double* work) {
  const int d = m.dim;
  double bad = 0.0;
  double* yv = work; double* yb = work + d; double* A = work + 2 * d; double* B = work + 3 * d;
  double lpd = 0.0, lpf = 0.0;
  SimplexCarry sc; sc.stick = 1.0; sc.lp = 0.0; sc.lj = 0.0;
  for (int i = 0; i < d; ++i) { yb[i] = 0.0; yv[i] = 0.0; }
  for (int i = 0; i < d; ++i) {
    const int kind = m.erp_kind[i];
    if (kind == ERP_DIRICHLET) {
      double gi;
      const int k = (int)m.p1[i], K = m.erp_len[i];
      dirichlet_component_noinline(sc, k, K, x[i], m.p0[i], m.erp_asum[i], gi);
      if (k == K - 1) { lpf += sc.lp; lpd += sc.lp + sc.lj; }
      A[i] = 0.0; B[i] = 0.0; g[i] = gi;
      continue;
    }
    double p0 = m.p0[i], p1 = m.p1[i], dp0 = 0.0, dp1 = 0.0;
    const int r0 = m.lat_ref ? m.lat_ref[2 * i] : -1, r1 = m.lat_ref ? m.lat_ref[2 * i + 1] : -1;
    if (r0 >= 0) latent_param(m, i, 0, yv[r0], p0, dp0);
    if (r1 >= 0) latent_param(m, i, 1, yv[r1], p1, dp1);
    double dydx, dlpdy, dlj;
    const double y = erp_fwd(kind, x[i], p0, p1, &dydx);
    yv[i] = y;
    if (m.cond_lo && !(y > m.cond_lo[i] && y < m.cond_hi[i])) bad += 1.0;
    double lpy = erp_prior_lp(kind, y, p0, p1, &dlpdy);
    if (grad && (r0 >= 0 || r1 >= 0)) {
      double d0, d1;
      erp_dlp_dparams(kind, y, p0, p1, d0, d1);
      if (r0 >= 0) yb[r0] += d0 * dp0;
      if (r1 >= 0) yb[r1] += d1 * dp1;
    }
    if (m.fsigma && m.fsigma[i] > 0.0) { lpy = lpy + gaussian_lp(y, m.fmu[i], m.fsigma[i]); dlpdy = dlpdy + gaussian_dlp_dx(y, m.fmu[i], m.fsigma[i]); }
    const double lj = erp_logjac(kind, x[i], p0, p1, &dlj);
    lpf += lpy; lpd += lpy + lj;
    A[i] = dlpdy; B[i] = dydx; g[i] = dlj;
  }
  if (grad) for (int i = 0; i < d; ++i) g[i] = (A[i] + yb[i]) * B[i] + g[i];
  *lp_float = lpf;
  if (cond_bad) *cond_bad = bad;
  return lpd;
}

// ---------------------------------------------------------------------------------------------
// Model log-densities.  Returns the dual-trace logprob (with log-Jacobians: what HMC's traceUpdate
// sees, hmc.t:325-342 + erp.t:623-640); lp_float is the float-trace logprob (no Jacobian: what
// Drift/HARM and the very first HMC Hamiltonian see, erp.t:361).  Inactive elements (component
// index >= dim) carry x = 0 and must not contribute.
// SSM: GAUSS_PREC reads A + A^T from the copy advance_kernel staged in shared memory behind the Gaussian columns
// (the pointer is formed here from the shared array itself so that the loads compile to LDS, not generic LD).
template <int FAM, int G, int NPL, bool GRAD, bool LP = true, bool SSM = false>
__device__ __forceinline__ double model_eval(const EngineParams& P, int lane, const double (&x)[NPL], double (&g)[NPL], double& lp_float,
                                             double* cond_bad = nullptr) {
  // cond_bad (INDEP): receives the number of qs.condition statements the evaluated point violates (0 = conditionsSatisfied)
  const ModelDev& m = P.m;
  const int d = m.dim;
  if (cond_bad) *cond_bad = 0.0;
  if (FAM == FAM_INDEP) {
    if (G == 1 && NPL <= 32 && m.lat_ref) {   // latent parameters: the general (slow) evaluator on local copies
      double xl[NPL], gl[NPL], work[4 * NPL];
      QSB_FOR_E { xl[e] = x[e]; gl[e] = 0.0; }
      const double lpd = indep_latent_eval(m, xl, gl, &lp_float, GRAD, cond_bad, work);
      if (GRAD) { QSB_FOR_E g[e] = e < d ? gl[e] : 0.0; }
      return lpd;
    }
    if (G == 32 && P.vscratch) {
      // Vector choices / latent parameters with a warp per chain (n up to 32 NPL): the stick-breaking recurrence and the
      // parameter links run through the program in order, so lane 0 evaluates the chain with the same sequential evaluator
      // the thread-per-chain mapping uses (same arithmetic, same order: the same numbers), on this warp's scratch slice;
      // the lanes hand their components in and take their gradient components out.  Not a throughput path.
      double* w = P.vscratch + (size_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5) * VS_ARRAYS * P.vs_dpad;
      double* xs = w; double* gsx = w + P.vs_dpad;
      QSB_FOR_E { const int i = e * G + lane; if (i < d) xs[i] = x[e]; }
      __syncwarp();
      double lpd = 0.0, lpf = 0.0, bad = 0.0;
      if (lane == 0) lpd = indep_latent_eval(m, xs, gsx, &lpf, GRAD, &bad, w + 2 * P.vs_dpad);
      __syncwarp();
      lpd = __shfl_sync(0xffffffffu, lpd, 0); lpf = __shfl_sync(0xffffffffu, lpf, 0); bad = __shfl_sync(0xffffffffu, bad, 0);
      if (GRAD) { QSB_FOR_E { const int i = e * G + lane; g[e] = i < d ? gsx[i] : 0.0; } }
      __syncwarp();
      lp_float = lpf;
      if (cond_bad) *cond_bad = bad;
      return lpd;
    }
    double lpd = 0.0, lpf = 0.0, bad = 0.0;
    SimplexCarry sc; sc.stick = 1.0; sc.lp = 0.0; sc.lj = 0.0;
    QSB_FOR_E {
      int i = e * G + lane;
      if (i < d) {
        const int kind = m.erp_kind[i];
        if (G == 1 && kind == ERP_DIRICHLET) {   // one component of a vector choice; the recurrence is carried in sc
          double gi;
          const int k = (int)m.p1[i], K = m.erp_len[i];
          dirichlet_component_noinline(sc, k, K, x[e], m.p0[i], m.erp_asum[i], gi);
          if (k == K - 1) { lpf += sc.lp; lpd += sc.lp + sc.lj; }
          if (GRAD) g[e] = gi;
        } else {
          double lpy, lj, gi, y;
          erp_eval(kind, x[e], m.p0[i], m.p1[i], m.fsigma ? m.fmu[i] : 0.0, m.fsigma ? m.fsigma[i] : 0.0, lpy, lj, gi, y);
          if (m.cond_lo && !(y > m.cond_lo[i] && y < m.cond_hi[i])) bad += 1.0;
          lpf += lpy;
          lpd += lpy + lj;
          if (GRAD) g[e] = gi;
        }
      } else if (GRAD) g[e] = 0.0;
    }
    if (cond_bad && m.cond_lo) *cond_bad = Grp<G>::sum(bad);
    lp_float = Grp<G>::sum(lpf);
    return Grp<G>::sum(lpd);
  } else if (FAM == FAM_GAUSS_PREC) {
    // log p = sum_i gaussian_lp(x_i;0,1) + factor(-0.5 x'Ax),  x'Ax = 0.5 x'(A+A')x ;  G == 1 only
    double lp = 0.0, q = 0.0;
    extern __shared__ double qsb_zbuf[];
    const double* Ssm = qsb_zbuf + (size_t)(2 * ((NPL + 1) / 2)) * blockDim.x;
    if (G > 1) {
      // G lanes per chain: lane l owns rows l, l+G, ..; x_j is broadcast from its owner and every lane advances its own rows'
      // sums in ascending j (the row sums equal the thread-per-chain form's bit for bit; q and lp are reduced over the lanes)
      constexpr int DP = G * NPL;                    // padded dimension = row stride of the staged S
      double sx[NPL];
      QSB_FOR_E sx[e] = 0.0;
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        const double xj = Grp<G>::bcast(x[j / G], j % G);
        if (j < d) {
          QSB_FOR_E {
            const int i = e * G + lane;
            if (SSM) sx[e] += Ssm[j * DP + i] * xj;            // S symmetric: column j read as the contiguous row j
            else sx[e] += P.Sc[(j * DP + i) & 255] * xj;
          }
        }
      }
      QSB_FOR_E {
        if (e * G + lane < d) {
          if (LP) { q += x[e] * sx[e]; lp += gaussian_lp(x[e], 0.0, 1.0); }
          if (GRAD) g[e] = -x[e] - 0.5 * sx[e];
        } else if (GRAD) g[e] = 0.0;
      }
      if (LP) { q = Grp<G>::sum(q); lp = Grp<G>::sum(lp); lp += -0.5 * (0.5 * q); }
      lp_float = lp;
      return lp;
    }
    if (d == NPL) {   // every element live (config c2: n = 10 on the NPL = 10 variant): no per-element predicates
      // column-major sweep: the NPL row sums advance together (NPL independent DFMA chains instead of one chain of NPL
      // dependent DFMAs per row); S is symmetric, so column j is read as the contiguous row j.  Each row still adds its
      // terms in ascending j: bit-identical to the row-major form.
      double sx[NPL];
      QSB_FOR_E sx[e] = 0.0;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        QSB_FOR_E {
          if (SSM) sx[e] += Ssm[j * NPL + e] * x[j];
          else sx[e] += P.Sc[(j * NPL + e) & 255] * x[j];
        }
      }
      QSB_FOR_E {
        if (LP) { q += x[e] * sx[e]; lp += gaussian_lp(x[e], 0.0, 1.0); }
        if (GRAD) g[e] = -x[e] - 0.5 * sx[e];
      }
    } else {
      QSB_FOR_E {
        if (e < d) {
          double sx = 0.0;
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            if (j < d) {
              if (SSM) sx += Ssm[e * NPL + j] * x[j];                 // shared memory (advance kernel)
              else sx += P.Sc[(e * NPL + j) & 255] * x[j];            // kernel parameter (one-off kernels)
            }
          }
          if (LP) { q += x[e] * sx; lp += gaussian_lp(x[e], 0.0, 1.0); }
          if (GRAD) g[e] = -x[e] - 0.5 * sx;
        } else if (GRAD) g[e] = 0.0;
      }
    }
    if (LP) lp += -0.5 * (0.5 * q);
    lp_float = lp;
    return lp;
  } else if (FAM == FAM_EIGHT_SCHOOLS) {
    // components: 0 mu, 1 logtau, 2+j theta_j
    double x0 = 0.0, x1 = 0.0;
    if (G == 1) { x0 = x[0]; x1 = NPL > 1 ? x[NPL > 1 ? 1 : 0] : 0.0; }
    else { x0 = Grp<G>::bcast(x[0], 0); x1 = Grp<G>::bcast(x[0], 1); }
    const double mu = x0, lt = x1;
    const double tau = exp(lt);
    const double tau2 = tau * tau;
    const double c_th = LP ? 1.8378770664093453 + 2.0 * log(tau) : 0.0;
    const double inv_tau2 = 1.0 / tau2;
    double lp = 0.0, gmu = 0.0, glt = 0.0;
    QSB_FOR_E {
      int i = e * G + lane;
      if (i >= 2 && i < d) {
        int j = i - 2;
        double th = x[e];
        double r = th - mu;
        double ry = m.y[j] - th;
        if (LP) {   // evaluations whose log-density is used: the reference's expressions, division for division
          double s2 = m.s2[j];
          double r2t = r * r / tau2;
          lp += -.5 * (c_th + r2t);
          lp += -.5 * (m.cy[j] + ry * ry / s2);
          if (GRAD) {
            double rt = r / tau2;
            g[e] = -rt + ry / s2;
            gmu += rt;
            glt += -1.0 + r2t;
          }
        } else if (GRAD) {   // inside a trajectory only the gradient is needed: reciprocals instead of divisions (<= 1 ulp apart)
          double rt = r * inv_tau2;
          g[e] = -rt + ry * m.is2[j];
          gmu += rt;
          glt += -1.0 + r * rt;
        }
      } else if (GRAD) g[e] = 0.0;
    }
    if (LP) {
      lp = Grp<G>::sum(lp);
      lp += gaussian_lp(mu, 0.0, 5.0) + gaussian_lp(lt, 0.0, 1.0);
    }
    if (GRAD) {
      gmu = Grp<G>::sum(gmu);
      glt = Grp<G>::sum(glt);
      QSB_FOR_E {
        int i = e * G + lane;
        if (i == 0) g[e] = -mu / 25.0 + gmu;
        if (i == 1) g[e] = -lt + glt;
      }
    }
    lp_float = lp;
    return lp;
  } else {  // FAM_FUNNEL: component 0 = v, i >= 1: x_i ~ gaussian(0, exp(v/2))
    const double v = Grp<G>::bcast(x[0], 0);
    const double s = exp(v / 2.0);
    const double s2 = s * s;
    const double c_x = LP ? 1.8378770664093453 + 2.0 * log(s) : 0.0;
    const double inv_s2 = 1.0 / s2;
    double lp = 0.0, gv = 0.0;
    QSB_FOR_E {
      int i = e * G + lane;
      if (i >= 1 && i < d) {
        double xi = x[e];
        if (LP) {
          double q = xi * xi / s2;
          lp += -.5 * (c_x + q);
          if (GRAD) { g[e] = -xi / s2; gv += -0.5 + 0.5 * q; }
        } else if (GRAD) {   // inside a trajectory: one reciprocal per evaluation instead of two divisions per component
          double t = xi * inv_s2;
          g[e] = -t; gv += -0.5 + 0.5 * (xi * t);
        }
      } else if (GRAD) g[e] = 0.0;
    }
    if (LP) {
      lp = Grp<G>::sum(lp);
      lp += gaussian_lp(v, 0.0, 3.0);
    }
    if (GRAD) {
      gv = Grp<G>::sum(gv);
      QSB_FOR_E { if (e * G + lane == 0) g[e] = -v / 9.0 + gv; }
    }
    lp_float = lp;
    return lp;
  }
}

// ---------------------------------------------------------------------------------------------
// Gradient of the hierarchical families from chain-level scalars.  For FUNNEL and EIGHT_SCHOOLS every component's
// gradient is a function of that component, two reduced scalars and (EIGHT_SCHOOLS) its observation, so the
// warp-per-chain HMC trajectory keeps positions and momenta in registers and NO gradient vector: 2*NPL instead of
// 3*NPL doubles per lane, which is what lets n = 1000 run without register spills.  Operation for operation the
// in-trajectory (GRAD && !LP) branch of model_eval.
template <int FAM> struct GradScalars {
  double a, b, g0, g1;   // FUNNEL: a = 1/s^2, g0 = d/dv.  EIGHT_SCHOOLS: a = mu, b = 1/tau^2, g0 = d/dmu, g1 = d/dlogtau
  template <int G, int NPL> __device__ __forceinline__ void reduce(const EngineParams& P, int lane, const double (&x)[NPL]) {
    const int d = P.m.dim;
    if (FAM == FAM_FUNNEL) {
      const double v = Grp<G>::bcast(x[0], 0);
      const double s = exp(v / 2.0);
      const double s2 = s * s;
      a = 1.0 / s2; b = 0.0; g1 = 0.0;
      // sum_i (-0.5 + 0.5 x_i^2/s^2) = 0.5 * sum_i x_i (x_i/s^2) - 0.5 (d-1): no per-element predicate (elements past the
      // last component hold x = 0; component 0 is masked once) and four independent accumulators per lane
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      QSB_FOR_E {
        double xe = x[e];
        if (e == 0) xe = lane == 0 ? 0.0 : xe;
        acc[e & 3] = fma(xe, xe * a, acc[e & 3]);
      }
      const double su = Grp<G>::sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
      g0 = -v / 9.0 + (0.5 * su - 0.5 * (double)(d - 1));
    } else {
      const double mu = Grp<G>::bcast(x[0], 0), lt = Grp<G>::bcast(x[0], 1);
      const double tau = exp(lt);
      const double tau2 = tau * tau;
      a = mu; b = 1.0 / tau2;
      double gmu = 0.0, glt = 0.0;
      QSB_FOR_E { int i = e * G + lane; if (i >= 2 && i < d) { const double r = x[e] - mu; const double rt = r * b; gmu += rt; glt += -1.0 + r * rt; } }
      g0 = -mu / 25.0 + Grp<G>::sum(gmu);
      g1 = -lt + Grp<G>::sum(glt);
    }
  }
  // gradient of component i at value xi
  __device__ __forceinline__ double elem(const EngineParams& P, int i, double xi) const {
    const int d = P.m.dim;
    if (FAM == FAM_FUNNEL) return i == 0 ? g0 : -(xi * a);   // elements past the last component hold x = 0: gradient -0
    if (i >= d) return 0.0;
    if (i == 0) return g0;
    if (i == 1) return g1;
    const double rt = (xi - a) * b;
    return -rt + (P.m.y[i - 2] - xi) * P.m.is2[i - 2];
  }
};

// ---------------------------------------------------------------------------------------------
// The float-trace log-density as a stream over components (what Drift / HARM score: erp.t:361, no Jacobian), for the
// warp-per-chain kernels: a proposal is generated, scored and parked in HBM component by component, so the kernel holds
// no per-lane vectors in registers and runs at 3 blocks per SM instead of 1.  The per-lane accumulation order equals
// model_eval's (components ascending, then the warp butterfly), so both forms give bit-identical sums.
template <int FAM> struct ModelStream {
  double g0, g1, a, b;   // component 0 and 1 of the proposal and two derived scalars (a is a reciprocal)
  mutable double bad = 0.0;   // INDEP: qs.condition statements violated by the components this lane scored
  struct Obs { double y, cy, is2; };   // per-component observation data, fetched ahead of use
  __device__ __forceinline__ void prep(const EngineParams& P, double x0, double x1) {
    g0 = x0; g1 = x1; a = 0.0; b = 0.0;
    if (FAM == FAM_EIGHT_SCHOOLS) { const double tau = exp(x1); a = 1.0 / (tau * tau); b = 1.8378770664093453 + 2.0 * log(tau); }
    if (FAM == FAM_FUNNEL) { const double s = exp(x0 / 2.0); a = 1.0 / (s * s); b = 1.8378770664093453 + 2.0 * log(s); }
  }
  __device__ __forceinline__ void fetch(const EngineParams& P, int i, int d, Obs& o) const {
    o.y = 0.0; o.cy = 0.0; o.is2 = 0.0;
    if (FAM == FAM_EIGHT_SCHOOLS && i >= 2 && i < d) { o.y = P.m.y[i - 2]; o.cy = P.m.cy[i - 2]; o.is2 = P.m.is2[i - 2]; }
  }
  // adds component i's terms to lp, in model_eval's order.  Quotients by chain-level or per-observation constants are
  // products with their reciprocals here (<= 1 ulp per term from the reference's divisions).
  __device__ __forceinline__ void term(const EngineParams& P, int i, double xi, const Obs& o, double& lp) const {
    if (FAM == FAM_INDEP) {
      double lpy, lj, gi, y;
      erp_eval(P.m.erp_kind[i], xi, P.m.p0[i], P.m.p1[i], P.m.fsigma ? P.m.fmu[i] : 0.0, P.m.fsigma ? P.m.fsigma[i] : 0.0, lpy, lj, gi, y);
      if (P.m.cond_lo && !(y > P.m.cond_lo[i] && y < P.m.cond_hi[i])) bad += 1.0;
      lp += lpy;
    } else if (FAM == FAM_EIGHT_SCHOOLS) {
      if (i >= 2) {
        const double r = xi - g0, ry = o.y - xi;
        lp += -.5 * (b + r * r * a);
        lp += -.5 * (o.cy + ry * ry * o.is2);
      }
    } else if (FAM == FAM_FUNNEL) {
      if (i >= 1) lp += -.5 * (b + xi * xi * a);
    }
  }
  __device__ __forceinline__ double finish(double lp_sum) const {
    if (FAM == FAM_EIGHT_SCHOOLS) return lp_sum + (gaussian_lp(g0, 0.0, 5.0) + gaussian_lp(g1, 0.0, 1.0));
    if (FAM == FAM_FUNNEL) return lp_sum + gaussian_lp(g0, 0.0, 3.0);
    return lp_sum;
  }
};

// ---------------------------------------------------------------------------------------------
// Standard-normal draws z_j = v/u (distrib.t:64-74 with mu = 0, sigma = 1) of the n components [32*row_begin, 32*row_begin + n)
// of ONE chain's iteration, by the 32 lanes of the chain's warp.  Component j reads sub-stream (chain, iter, slot = j) and
// accepts its first Leva pair that passes -- which pair that is does not depend on who evaluates it, so the lanes pull
// components from a common cursor: every trip of the loop makes 32 attempts for 32 different components and a lane whose
// attempt passed takes the next unassigned component.  Trips ~ 1.37 n / 32 + a short tail, instead of the maximum over lanes
// of a lane's own total (1.7 n / 32).  Results go to this warp's columns of the block's staging array: component
// 32*(row_begin + r) + l -> qsb_zbuf[r * rstride + (warp's first thread) + l], i.e. lane l later reads its own component of row r.
// One out-of-line copy per kernel (the 10 Philox rounds are inlined in it).
// LANE_CHAINS (thread per chain): the 32 lanes of the warp own 32 consecutive chains (gc = the chain of lane 0); work item
// j is component row_begin + (j >> 5) of chain gc + (j & 31) and lands in the same place of the staging array, so the warp
// shares the attempts of its 32 chains the same way.  Needs all 32 lanes of the warp at the call.
// QSB_FILL_U work items can be in flight per lane (U independent attempts per trip as straight-line code, the votes / cursor /
// loop branch paid once per U attempts).  Measured on B200 (gpurun r02k, same box, c5 / c4 / c2): U = 1 1.095 / 2.061 / 0.0402 ms
// per step, U = 2 1.124 / 2.090 / 0.0432, U = 4 1.364 / 2.356 / 0.0475 -- the interleaving buys nothing (a warp inside the fill
// already issues at the rate its IMAD.WIDE / FP64 mix allows; the kernels lack WARPS, not ILP) and the tail of the fill, where
// most slots idle but every trip still costs U attempts, grows with U.  U = 1 is the default.
#ifndef QSB_FILL_U
#define QSB_FILL_U 1
#endif
template <int G, int U = QSB_FILL_U>
static __device__ __forceinline__ void gaussian_fill_body(uint32_t k0, uint32_t k1, uint32_t gc, uint32_t iter, int row_begin, int n, double* zw, int rstride) {
  // work item j: chain q = j % CPW of the warp's CPW = 32/G chains (gc = the chain of lane 0), component row_begin*G + j / CPW;
  // it belongs to lane q*G + (component % G), row component / G - row_begin of the staging rows at zw.  n = number of work items.
  constexpr int CPW = 32 / G;
  const int lane = (int)(threadIdx.x & 31u);
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t comp0 = (uint32_t)row_begin * (uint32_t)G;
  int j[U], next = 32 * U;
  uint32_t blk[U];
  bool active[U];
#pragma unroll
  for (int t = 0; t < U; ++t) { j[t] = lane + 32 * t; blk[t] = 0; active[t] = j[t] < n; }
  for (;;) {
    bool any = false;
#pragma unroll
    for (int t = 0; t < U; ++t) any = any || active[t];
    if (!__any_sync(0xffffffffu, any)) break;
    double u[U], v[U], qq[U];
    uint32_t q[U], ci[U];
    // the U attempts, unconditionally (an idle slot recomputes its last item: no branch separates the chains)
#pragma unroll
    for (int t = 0; t < U; ++t) {
      uint32_t w[4];
      q[t] = (uint32_t)j[t] % (uint32_t)CPW; ci[t] = (uint32_t)j[t] / (uint32_t)CPW;     // chain in the warp, component - comp0
      philox4x32_10(blk[t], comp0 + ci[t], iter, gc + q[t], k0, k1, w);
      u[t] = __dsub_rn(1.0, u53(w[0], w[1]));
      v[t] = __dmul_rn(1.7156, __dsub_rn(u53(w[2], w[3]), 0.5));
      const double x = __dsub_rn(u[t], 0.449871);
      const double y = __dadd_rn(fabs(v[t]), 0.386595);
      qq[t] = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, __dsub_rn(__dmul_rn(0.196, y), __dmul_rn(0.25472, x))));
    }
    bool ok[U];
#pragma unroll
    for (int t = 0; t < U; ++t) {
      ok[t] = !(qq[t] >= 0.27597);
      if (!ok[t] && !(qq[t] > 0.27846)) {
        // The exact test v^2 > -4 u^2 log(u) (distrib.t:71) is reached by 0.85 % of the attempts, i.e. by some lane of a warp in
        // 23 % of the trips, and the double-precision log() then costs every lane ~90 instructions.  A single-precision log
        // decides it first: |__logf(u) - log(u)| < 1e-5 for every u in [2^-53, 1], so outside the band |v^2 - R_f| <= 4 u^2 * 2e-5
        // the decision equals the exact one; inside it (~1e-6 of the attempts) the exact expression is evaluated.
        const double u2x4 = __dmul_rn(__dmul_rn(4.0, u[t]), u[t]), vv = __dmul_rn(v[t], v[t]);
        const double rf = u2x4 * (double)(-__logf((float)u[t])), band = u2x4 * 2e-5;
        if (vv > rf + band) ok[t] = false;
        else if (vv < rf - band) ok[t] = true;
        else ok[t] = !(vv > __dmul_rn(__dmul_rn(__dmul_rn(-4.0, u[t]), u[t]), log(u[t])));
      }
      ok[t] = ok[t] && active[t];
    }
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const double z = v[t] / u[t];          // u in (0, 1]: finite for idle slots too
      if (ok[t]) zw[(ci[t] / (uint32_t)G) * rstride + q[t] * (uint32_t)G + (ci[t] % (uint32_t)G)] = z;
    }
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const unsigned m = __ballot_sync(0xffffffffu, ok[t]);
      if (ok[t]) { j[t] = next + __popc(m & lt); blk[t] = 0; active[t] = j[t] < n; }
      else if (active[t]) ++blk[t];
      next += __popc(m);
    }
  }
  __syncwarp();
}
// One out-of-line copy per kernel for the fused (every warp samples and integrates) kernels.
template <int G>
static __device__ __noinline__ void gaussian_fill_warp(uint32_t k0, uint32_t k1, uint32_t gc, uint32_t iter, int row_begin, int n, double* zw, int rstride) {
  gaussian_fill_body<G>(k0, k1, gc, iter, row_begin, n, zw, rstride);
}

// Hand-off between the sampler warps and the trajectory warp of one chain slot in the warp-specialised kernel
// (advance_ws_kernel): two staging buffers of NPL x 32 standard normals alternate, full[b] / empty[b] are mbarriers.
struct WsLink {
  double* zbuf;        // this slot's two buffers (2 x NPL x 32 doubles: row r of buffer b at zbuf + (b * NPL + r) * 32)
  double* zsearch;     // a third buffer for the trajectory warp's own draws (step-size search probes), or nullptr
  uint64_t* full; uint64_t* empty;
  uint32_t k;          // hand-offs consumed so far: buffer k & 1, phase (k >> 1) & 1
  bool held;           // the trajectory warp currently reads buffer k & 1
};

// ---------------------------------------------------------------------------------------------
// WS: the chain runs on a trajectory warp of the warp-specialised kernel -- the standard normals of an iteration arrive
// from the sampler warps through `link` instead of being drawn in place.
template <int FAM, int G, int NPL, bool WS = false> struct Chain {
  const EngineParams& P;
  const uint32_t c;       // local chain index
  const uint32_t gc;      // global chain id (Philox counter word)
  const int lane;
  const int d;
  WsLink* const link;
  mutable const double* zptr = nullptr;   // WS: this lane's column of the staging rows currently being read
  // gradient evaluations / leapfrog steps / transitions whose accept threshold was not finite (NaN or +-inf log-density: the
  // chain left the support or overflowed) / HMC transitions with |dH| > 1000 (divergent); of this group, flushed once per launch
  mutable unsigned long long cnt[4] = {0, 0, 0, 0};
  // currTrace.temperature of this iteration (trace.t:652-655): every log-density evaluated during the iteration is
  // divided by it (trace.t:737-741) and its gradient multiplied by 1/T (the adjoint of that division, ad.t:405-412);
  // values cached from earlier iterations (lp, gradient) keep the temperature they were computed under, as in the reference
  mutable double T = 1.0, invT = 1.0;
  // conditionsSatisfied of the trace evaluated last (trace.t:723, 794-796), as a count of violated qs.condition statements;
  // the accept tests read it: `conditionsSatisfied and log(random()) < ...` (hmc.t:164, drift.t:158, harm.t:93, mcmc.t:181)
  mutable double cond_bad = 0.0;
  __device__ __forceinline__ Chain(const EngineParams& P_, uint32_t c_, WsLink* link_ = nullptr)
      : P(P_), c(c_), gc(P_.chain_begin + c_), lane(Grp<G>::lane()), d(P_.d), link(link_) {}

  __device__ __forceinline__ void load(const double* base, double (&v)[NPL]) const {
    QSB_FOR_E { int i = e * G + lane; v[e] = i < d ? base[vidx<G>(P, i, c)] : 0.0; }
  }
  __device__ __forceinline__ void store(double* base, const double (&v)[NPL]) const {
    QSB_FOR_E { int i = e * G + lane; if (i < d) base[vidx<G>(P, i, c)] = v[e]; }
  }
  __device__ __forceinline__ bool leader() const { return lane == 0; }
  __device__ __forceinline__ void count(int which, unsigned long long n) const {
    if (leader()) cnt[which] += n;
  }
  // one atomic per warp and counter per launch (a per-chain atomic on one address serialised 9 % of the c2 kernel)
  __device__ __forceinline__ void flush_counts() const {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      unsigned long long v = cnt[w];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31u) == 0 && v) atomicAdd(&P.counters[w], v);
    }
  }

  // ---- Gaussian draws of one iteration.  Component i of the chain reads sub-stream (chain, iter, slot = i) and
  // accepts the first Leva pair that passes (distrib.t:66-72).  Doing that in lock step over the components of a
  // lane would make every component cost the *maximum* attempt count of the 32 lanes of the warp (E ~ 3.6 instead
  // of 1.37).  Instead every lane walks through ITS OWN components, one attempt per trip, and parks the accepted
  // pair in its private column of shared memory (column = thread, row = component slot: conflict-free); the
  // consumer then reads the rows with compile-time indices.  Warp trips ~ max over lanes of the lane's total attempts.
  static constexpr int HALF = (NPL + 1) / 2;
  static constexpr int NB = 2 * HALF;                 // shared-memory rows per thread
  __device__ __forceinline__ const double* zcol() const {
    if (WS) return zptr;
    extern __shared__ double qsb_zbuf[];
    return qsb_zbuf + threadIdx.x;
  }
  __device__ __forceinline__ double* zcol_w() const {
    extern __shared__ double qsb_zbuf[];
    return qsb_zbuf + threadIdx.x;
  }
  __device__ __forceinline__ int zstride() const { return WS ? 32 : (int)blockDim.x; }
  // WS: take / give back the staging buffer the sampler warps filled for this iteration
  __device__ __forceinline__ void ws_acquire() const {
    const uint32_t b = link->k & 1u;
    mbar_wait(&link->full[b], (link->k >> 1) & 1u);
    zptr = link->zbuf + (size_t)b * (NPL * 32) + (threadIdx.x & 31u);
    link->held = true;
  }
  __device__ __forceinline__ void ws_release() const {
    if (!link->held) return;
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) mbar_arrive(&link->empty[link->k & 1u]);
    ++link->k;
    link->held = false;
  }
  // VU = false: row r (component e_begin + r) <- v/u;  VU = true: row r <- v, row HALF + r <- u
  template <bool VU>
  __device__ __forceinline__ void gaussian_fill(uint32_t iter, int e_begin, int e_end) const {
    if constexpr (WS) {
      const int hi = e_end * 32 < d ? e_end * 32 : d;
      if (iter < QSB_ITER_SEARCH) { ws_acquire(); return; }      // drawn ahead of time by the sampler warps
      // the probes of the step-size search: drawn in place, into the slot's third buffer
      gaussian_fill_warp<32>(P.key.k0, P.key.k1, gc, iter, e_begin, hi - e_begin * 32, link->zsearch + e_begin * 32, 32);
      zptr = link->zsearch + (threadIdx.x & 31u);
      return;
    }
    if (G == 32 && !VU) {   // warp per chain: the lanes share the attempts of all components of these rows
      const int hi = e_end * 32 < d ? e_end * 32 : d;
      extern __shared__ double qsb_zbuf[];
      gaussian_fill_warp<32>(P.key.k0, P.key.k1, gc, iter, e_begin, hi - e_begin * 32, qsb_zbuf + (threadIdx.x & ~31u), (int)blockDim.x);
      return;
    }
    if (G < 32 && !VU && P.nsub == 1 && __activemask() == 0xffffffffu) {
      // several chains per warp, a single kernel (every lane of the warp is at this call) and a full warp: the 32/G chains
      // share their attempts.  Partial tail warps and mixtures (lanes in different sub-kernels) take the per-lane loop below.
      const int hi = e_end * G < d ? e_end * G : d;
      extern __shared__ double qsb_zbuf[];
      gaussian_fill_warp<G>(P.key.k0, P.key.k1, gc - (uint32_t)((threadIdx.x & 31u) / (unsigned)G), iter, e_begin, (hi - e_begin * G) * (32 / G),
                            qsb_zbuf + (threadIdx.x & ~31u), (int)blockDim.x);
      return;
    }
    double* z = zcol_w();
    const int stride = blockDim.x;
    int e = e_begin;
    uint32_t blk = 0;
    while (e < e_end && e * G + lane < d) {
      double v, u;
      if (leva_attempt(P.key, gc, iter, (uint32_t)(e * G + lane), blk, v, u)) {
        if (VU) { z[(e - e_begin) * stride] = v; z[(HALF + e - e_begin) * stride] = u; }
        else z[(e - e_begin) * stride] = v / u;
        ++e; blk = 0;
      } else ++blk;
    }
    if (G == 32) __syncwarp();
  }

  // ---- sampleMomenta + kinetic energy (hmc.t:289-305), invMasses == 1
  __device__ __forceinline__ void sample_momenta(uint32_t iter, double (&p)[NPL]) const {
    gaussian_fill<false>(iter, 0, NPL);
    const double* z = zcol();
    const int stride = zstride();
    QSB_FOR_E {
      int i = e * G + lane;
      p[e] = i < d ? z[e * stride] * 1.0 : 0.0;
    }
    if constexpr (WS) { if (iter < QSB_ITER_SEARCH) ws_release(); }   // the momenta are in registers: the sampler warps may refill
  }
  __device__ __forceinline__ double kinetic(const double (&p)[NPL]) const {
    double k = 0.0;
    QSB_FOR_E k += p[e] * p[e] * 1.0;
    return -0.5 * Grp<G>::sum(k);
  }
  // ---- HMCKernel:step (hmc.t:307-323): two separate half-kicks, reference operation order
  // LPW: also return the log-density at the new position.  Inside a trajectory only the last step's value is used
  // (hmc.t:150-152 computes and overwrites it every step), so the other steps skip its logs and divisions.
  template <bool LPW = true>
  __device__ __forceinline__ double leapfrog(double eps, double (&xs)[NPL], double (&gs)[NPL], double (&p)[NPL]) const {
    const double he = 0.5 * eps;
    QSB_FOR_E p[e] = p[e] + he * gs[e];
    QSB_FOR_E xs[e] = xs[e] + eps * p[e] * 1.0;
    double lpf;
    double lp = model_eval<FAM, G, NPL, true, LPW, (FAM == FAM_GAUSS_PREC && QSB_GP_SMEM)>(P, lane, xs, gs, lpf, &cond_bad);
    if (T != 1.0) { lp = lp / T; QSB_FOR_E gs[e] = gs[e] * invT; }
    QSB_FOR_E p[e] = p[e] + he * gs[e];
    return lp;
  }

  // ---- DualAverage:update + adaptStepSize (hmc.t:38-48, 394-402)
  __device__ __forceinline__ double adapt_step_size(double dH, double target) const {
    double EdH = exp(dH);
    if (EdH > 1.0) EdH = 1.0;
    if (!(EdH == EdH)) EdH = 0.0;
    double gr = target - EdH;
    uint32_t k = P.da_k[c] + 1;
    double avgeta = 1.0 / (double)(k + 10);
    double muk = 0.5 * sqrt((double)k) / P.da_gamma[c];
    double gbar = avgeta * gr + (1 - avgeta) * P.da_gbar[c];
    double lastx = P.da_x0[c] - muk * gbar;
    // DualAverage.xbar (hmc.t:44-46: xbar = k^-0.75 * lastx + (1 - k^-0.75) * xbar) is written by the reference but never
    // read: not computed here -- it would cost a pow() per transition for a value nothing observes
    if (G == 32) __syncwarp();   // every lane has read the chain's adapter state before the leader overwrites it
    if (leader()) { P.da_k[c] = k; P.da_gbar[c] = gbar; P.da_lastx[c] = lastx; }
    return exp(lastx);
  }

  // ---- searchForInitialStepSize (hmc.t:345-391).  x, g are the current position / gradient in HBM.
  __device__ __forceinline__ double search_step_size(double eps, double dual_lp) const {
    double xs[NPL], gs[NPL], p[NPL];   // own scratch: arrays passed by reference into a call would be forced into local memory
    const double LOG_HALF = log(0.5);
    double prevlp = dual_lp;
    int direction = 0;
    unsigned long long evals = 0;
    for (uint32_t probe = 0;; ++probe) {     // one leapfrog call site for the first probe and the doubling/halving loop
      sample_momenta(QSB_ITER_SEARCH + probe, p);
      load(P.x, xs); load(P.g, gs);
      const double lp = leapfrog(eps, xs, gs, p);
      ++evals;
      const double H = lp - prevlp;      // hmc.t:349-351, 357: dualTrace.logprob is the previous probe's value
      prevlp = lp;
      if (probe == 0) { direction = H > LOG_HALF ? 1 : -1; continue; }
      if (!(lp == lp)) { if (leader()) atomicOr(P.status, ERR_NAN_SEARCH); break; }
      if (direction == 1 && H <= LOG_HALF) break;
      else if (direction == -1 && H >= LOG_HALF) break;
      else if (direction == 1) eps = eps * 2.0;
      else eps = eps * 0.5;
      if (eps > 1e300) { if (leader()) atomicOr(P.status, ERR_SEARCH_INF); break; }
      if (eps == 0) { if (leader()) atomicOr(P.status, ERR_SEARCH_ZERO); break; }
    }
    count(0, evals);
    return eps;
  }

  // ---- HMCKernel:next (hmc.t:134-186)
  __device__ __forceinline__ bool hmc_next(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    double xs[NPL], gs[NPL], p[NPL];
    const bool adapting = P.doAdapt[k] != 0;
    const uint32_t L = P.numSteps[k];
    const double target = L == 1 ? 0.574 : 0.65;
    if (leader()) P.made[(size_t)k * P.C + c] += 1;
    uint8_t fl = P.flags[c];
    double eps = P.eps[c];
    // checkForChanges (hmc.t:188-287)
    if (!(fl & FLAG_GRAD_VALID)) {
      load(P.x, xs);
      double lpf;
      double dual_lp = model_eval<FAM, G, NPL, true, true, (FAM == FAM_GAUSS_PREC && QSB_GP_SMEM)>(P, lane, xs, gs, lpf);
      if (P.temps) {
        // hmc.t:254 runs before hmc.t:266: the re-gathered gradient is taken under the temperature the dual trace was
        // left with by the previous HMC call (first call: the copy of currTrace's, hmc.t:218 / trace.t:167-173)
        const double Tg = (fl & FLAG_HMC_INIT) ? P.hmc_dualT[c] : T;
        if (Tg != 1.0) { const double iTg = 1.0 / Tg; dual_lp = dual_lp / Tg; QSB_FOR_E gs[e] = gs[e] * iTg; }
      }
      store(P.g, gs);
      count(0, 1);
      if (!(fl & FLAG_HMC_INIT) && adapting) {
        if (G == 32) __syncwarp();
        eps = search_step_size(eps, dual_lp);
        if (leader()) {   // adapter:reinit(stepSize, adaptRate) (hmc.t:261)
          P.da_k[c] = 0; P.da_x0[c] = eps; P.da_lastx[c] = eps; P.da_gbar[c] = 0.0; P.da_xbar[c] = 0.0; P.da_gamma[c] = 0.05;
        }
        if (G == 32) __syncwarp();
      }
      fl |= FLAG_HMC_INIT | FLAG_GRAD_VALID;
    }
    if (P.temps && leader()) P.hmc_dualT[c] = T;   // hmc.t:266
    sample_momenta(iter, p);
    load(P.x, xs); load(P.g, gs);
    const double lp_cur = P.lp[c];
    const double H = kinetic(p) + lp_cur;
    double newlp = 0.0;
    rec_step = eps;
    for (uint32_t s = 0; s + 1 < L; ++s) {
      leapfrog<false>(eps, xs, gs, p);
      if (P.rec_leap) {
        size_t base = (((size_t)c * P.rec_iters + (iter - P.it_begin)) * L + s) * d;
        QSB_FOR_E { int i = e * G + lane; if (i < d) P.rec_leap[base + i] = xs[e]; }
      }
    }
    {
      newlp = leapfrog<true>(eps, xs, gs, p);
      if (P.rec_leap) {
        size_t base = (((size_t)c * P.rec_iters + (iter - P.it_begin)) * L + (L - 1)) * d;
        QSB_FOR_E { int i = e * G + lane; if (i < d) P.rec_leap[base + i] = xs[e]; }
      }
    }
    count(0, L); count(1, L);
    const double H_new = kinetic(p) + newlp;
    const double dH = H_new - H;
    rec_dh = dH;
    if (adapting) eps = adapt_step_size(dH, target);
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d);
    const bool accept = cond_bad == 0.0 && log(u) < dH;      // dualTrace.conditionsSatisfied after the last step (hmc.t:164)
    if (accept) {
      store(P.x, xs); store(P.g, gs);
      if (leader()) { P.acc[(size_t)k * P.C + c] += 1; P.lp[c] = newlp; }
    }
    if (leader()) { P.eps[c] = eps; P.flags[c] = fl; }
    return accept;
  }

  // ---- HMCKernel:next for the warp-per-chain hierarchical families: positions and momenta in registers, the gradient
  // recomputed from GradScalars at every kick (two kicks per evaluation inside a trajectory, hmc.t:307-323 kept as two
  // separate half-kicks).  The cached gradient of the reference (hmc.t:82-101) is a function of the cached position, so
  // it is not stored.
  static constexpr bool SCALAR_GRAD = G == 32 && (FAM == FAM_FUNNEL || FAM == FAM_EIGHT_SCHOOLS);
  __device__ __forceinline__ void kick(double he, const GradScalars<FAM>& gs, const double (&x)[NPL], double (&p)[NPL], bool twice, double f) const {
    QSB_FOR_E {
      double g = gs.elem(P, e * G + lane, x[e]);
      if (f != 1.0) g = g * f;
      p[e] = p[e] + he * g;
      if (twice) p[e] = p[e] + he * g;
    }
  }
  __device__ __forceinline__ double lp_only(const double (&x)[NPL]) const {
    double lpf, gd[NPL];
    const double lp = model_eval<FAM, G, NPL, false, true, false>(P, lane, x, gd, lpf);
    return T != 1.0 ? lp / T : lp;
  }
  __device__ __forceinline__ double search_step_size_warp(double eps, double dual_lp) const {
    double x[NPL], p[NPL];
    const double LOG_HALF = log(0.5);
    double prevlp = dual_lp;
    int direction = 0;
    unsigned long long evals = 0;
    GradScalars<FAM> gs;
    for (uint32_t probe = 0;; ++probe) {
      sample_momenta(QSB_ITER_SEARCH + probe, p);     // (before the positions are loaded: the fill runs with few live registers)
      load(P.x, x);
      gs.template reduce<G, NPL>(P, lane, x);
      kick(0.5 * eps, gs, x, p, false, invT);
      QSB_FOR_E x[e] = x[e] + eps * p[e] * 1.0;
      const double lp = lp_only(x);
      ++evals;
      const double H = lp - prevlp;      // hmc.t:349-351, 357: dualTrace.logprob is the previous probe's value
      prevlp = lp;
      if (probe == 0) { direction = H > LOG_HALF ? 1 : -1; continue; }
      if (!(lp == lp)) { if (leader()) atomicOr(P.status, ERR_NAN_SEARCH); break; }
      if (direction == 1 && H <= LOG_HALF) break;
      else if (direction == -1 && H >= LOG_HALF) break;
      else if (direction == 1) eps = eps * 2.0;
      else eps = eps * 0.5;
      if (eps > 1e300) { if (leader()) atomicOr(P.status, ERR_SEARCH_INF); break; }
      if (eps == 0) { if (leader()) atomicOr(P.status, ERR_SEARCH_ZERO); break; }
    }
    count(0, evals);
    return eps;
  }
  __device__ __forceinline__ bool hmc_next_warp(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    double x[NPL], p[NPL];
    const bool adapting = P.doAdapt[k] != 0;
    const uint32_t L = P.numSteps[k];
    const double target = L == 1 ? 0.574 : 0.65;
    if (leader()) P.made[(size_t)k * P.C + c] += 1;
    uint8_t fl = P.flags[c];
    double eps = P.eps[c];
    double Tg0 = T;     // temperature carried by the reference's cached gradient (first half-kick of this call)
    if (P.temps) {
      if (fl & FLAG_GRAD_VALID) Tg0 = P.hmc_gT[c];
      else if (fl & FLAG_HMC_INIT) Tg0 = P.hmc_dualT[c];   // re-gather under the dual trace's stale temperature (hmc.t:254 before :266)
    }
    const double f0 = Tg0 != 1.0 ? 1.0 / Tg0 : 1.0;
    if (!(fl & FLAG_GRAD_VALID)) {   // checkForChanges (hmc.t:188-287): the reference re-evaluates the gradient here
      count(0, 1);
      if (!(fl & FLAG_HMC_INIT) && adapting) {
        load(P.x, x);
        const double dual_lp = lp_only(x);
        __syncwarp();
        eps = search_step_size_warp(eps, dual_lp);
        if (leader()) {   // adapter:reinit(stepSize, adaptRate) (hmc.t:261)
          P.da_k[c] = 0; P.da_x0[c] = eps; P.da_lastx[c] = eps; P.da_gbar[c] = 0.0; P.da_xbar[c] = 0.0; P.da_gamma[c] = 0.05;
        }
        __syncwarp();
      }
      fl |= FLAG_HMC_INIT | FLAG_GRAD_VALID;
    }
    sample_momenta(iter, p);
    load(P.x, x);
    const double H = kinetic(p) + P.lp[c];
    rec_step = eps;
    const double he = 0.5 * eps;
    GradScalars<FAM> gs;
    gs.template reduce<G, NPL>(P, lane, x);
    kick(he, gs, x, p, false, f0);
    for (uint32_t s = 0; s < L; ++s) {
      QSB_FOR_E x[e] = x[e] + eps * p[e] * 1.0;
      gs.template reduce<G, NPL>(P, lane, x);
      kick(he, gs, x, p, s + 1 < L, invT);
      if (P.rec_leap) {
        size_t base = (((size_t)c * P.rec_iters + (iter - P.it_begin)) * L + s) * d;
        QSB_FOR_E { int i = e * G + lane; if (i < d) P.rec_leap[base + i] = x[e]; }
      }
    }
    const double newlp = lp_only(x);
    count(0, L); count(1, L);
    const double dH = (kinetic(p) + newlp) - H;
    rec_dh = dH;
    if (adapting) eps = adapt_step_size(dH, target);
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d);
    const bool accept = log(u) < dH;
    if (accept) {
      store(P.x, x);
      if (leader()) { P.acc[(size_t)k * P.C + c] += 1; P.lp[c] = newlp; }
    }
    if (leader()) {
      P.eps[c] = eps; P.flags[c] = fl;
      if (P.temps) { P.hmc_dualT[c] = T; P.hmc_gT[c] = accept ? T : Tg0; }
    }
    return accept;
  }

  // ---- DriftKernel:next + adapt (drift.t:133-172, 274-300), lexicalScaleSharing = false
  __device__ __forceinline__ bool drift_next(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    double x[NPL], xp[NPL];
    const bool adapting = P.doAdapt[k] != 0;
    const size_t kc = (size_t)k * P.C + c;
    unsigned long long made = P.made[kc] + 1, acc = P.acc[kc];
    const uint32_t n_rv = P.dr_n[c];
    const double s0 = P.dr_s0[c], mult = P.dr_mult[c];
    load(P.x, x);
    // proposal x'_i = gaussian.sample(x_i, scale_i) = x_i + scale_i*v/u (distrib.t:73), in two passes of HALF components
    {
      const double* z = zcol();
      const int stride = zstride();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        gaussian_fill<true>(iter, h * HALF, h * HALF + HALF < NPL ? h * HALF + HALF : NPL);
#pragma unroll
        for (int r = 0; r < HALF; ++r) {
          const int e = h * HALF + r;
          if (e < NPL) {
            int i = e * G + lane;
            if (i < d) {
              double sc = s0;
              if (n_rv > 100) {   // scale = varEst:getStdDev() [* 0.99]  (drift.t:288-298)
                sc = sqrt(P.dr_m2[vidx<G>(P, i, c)] / (double)(n_rv - 1));
                if (mult != 1.0) sc = sc * mult;
              }
              xp[e] = __dadd_rn(x[e], __dmul_rn(sc, z[r * stride]) / z[(HALF + r) * stride]);
            } else xp[e] = 0.0;
          }
        }
        if (G == 32) __syncwarp();
      }
    }
    rec_step = s0;
    double lpf, gdummy[NPL];
    model_eval<FAM, G, NPL, false, true, (FAM == FAM_GAUSS_PREC && QSB_GP_SMEM)>(P, lane, xp, gdummy, lpf, &cond_bad);
    if (T != 1.0) lpf = lpf / T;
    const double lp_cur = P.lp[c];
    const double thresh = lpf - lp_cur;
    rec_dh = thresh;
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d);
    const bool accept = cond_bad == 0.0 && log(u) < thresh;   // `conditionsSatisfied and log(random()) < ...`
    if (accept) {
      acc += 1;
      store(P.x, xp);
      QSB_FOR_E x[e] = xp[e];
      if (leader()) { P.lp[c] = lpf; P.flags[c] = P.flags[c] & ~FLAG_GRAD_VALID; }
    }
    if (adapting) {
      const double ratio = (double)acc / (double)made;
      uint32_t n_new = n_rv;
      if (ratio >= 0.05 && acc > 5) {   // RunningVar:update with the current state (drift.t:26-37)
        n_new = n_rv + 1;
        QSB_FOR_E {
          int i = e * G + lane;
          if (i < d) {
            size_t ix = vidx<G>(P, i, c);
            if (n_new == 1) { P.dr_mean[ix] = x[e]; P.dr_m2[ix] = 0.0; }
            else {
              double mean = P.dr_mean[ix];
              double newmean = mean + (x[e] - mean) / (double)n_new;
              P.dr_m2[ix] = P.dr_m2[ix] + (x[e] - mean) * (x[e] - newmean);
              P.dr_mean[ix] = newmean;
            }
          }
        }
      }
      if (leader()) {
        const bool shrink = ratio < 0.05 && acc > 5;
        P.dr_n[c] = n_new;
        if (n_new > 100) P.dr_mult[c] = shrink ? 0.99 : 1.0;
        else if (shrink) P.dr_s0[c] = s0 * 0.99;
      }
    }
    if (leader()) { P.made[kc] = made; P.acc[kc] = acc; }
    return accept;
  }

  // ---- DriftKernel{lexicalScaleSharing = true} (drift.t:101-115, 208-235, 274-300): one ScaleAdapter per call site (and index
  // inside a vector choice); every component of the site feeds the same RunningVar and is proposed with the same scale.
  __device__ __forceinline__ int adapter_of(int i) const {
    if (FAM == FAM_INDEP) return P.m.lex_map[i];
    if (FAM == FAM_EIGHT_SCHOOLS) return i < 2 ? i : 2;
    if (FAM == FAM_FUNNEL) return i == 0 ? 0 : 1;
    return 0;
  }
  // scale update of every adapter after the variance updates (drift.t:286-298); lane/thread-serial over adapters
  __device__ __forceinline__ void lex_rescale(int a, bool shrink) const {
    const size_t ix = (size_t)a * P.C + c;
    const uint32_t n = P.ls_n[ix];
    double sc = P.ls_scale[ix];
    if (n > 100) sc = sqrt(P.ls_m2[ix] / (double)(n - 1));
    if (shrink) sc = sc * 0.99;
    P.ls_scale[ix] = sc;
  }
  // thread per chain: the reference's component order, RunningVar:update operation for operation
  __device__ __forceinline__ bool drift_next_lex(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    double x[NPL], xp[NPL];
    const bool adapting = P.doAdapt[k] != 0;
    const size_t kc = (size_t)k * P.C + c;
    unsigned long long made = P.made[kc] + 1, acc = P.acc[kc];
    load(P.x, x);
    {
      const double* z = zcol();
      const int stride = zstride();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        gaussian_fill<true>(iter, h * HALF, h * HALF + HALF < NPL ? h * HALF + HALF : NPL);
#pragma unroll
        for (int r = 0; r < HALF; ++r) {
          const int e = h * HALF + r;
          if (e < NPL) {
            int i = e * G + lane;
            if (i < d) {
              const double sc = P.ls_scale[(size_t)adapter_of(i) * P.C + c];
              xp[e] = __dadd_rn(x[e], __dmul_rn(sc, z[r * stride]) / z[(HALF + r) * stride]);
            } else xp[e] = 0.0;
          }
        }
        if (G == 32) __syncwarp();
      }
    }
    rec_step = P.ls_scale[c];
    double lpf, gdummy[NPL];
    model_eval<FAM, G, NPL, false, true, (FAM == FAM_GAUSS_PREC && QSB_GP_SMEM)>(P, lane, xp, gdummy, lpf, &cond_bad);
    if (T != 1.0) lpf = lpf / T;
    const double thresh = lpf - P.lp[c];
    rec_dh = thresh;
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d);
    const bool accept = cond_bad == 0.0 && log(u) < thresh;   // `conditionsSatisfied and log(random()) < ...`
    if (accept) {
      acc += 1;
      store(P.x, xp);
      QSB_FOR_E x[e] = xp[e];
      P.lp[c] = lpf; P.flags[c] = P.flags[c] & ~FLAG_GRAD_VALID;
    }
    if (adapting) {
      const double ratio = (double)acc / (double)made;
      if (ratio >= 0.05 && acc > 5) {
        QSB_FOR_E {
          const int i = e * G + lane;
          if (i < d) {   // RunningVar:update (drift.t:26-37)
            const size_t ix = (size_t)adapter_of(i) * P.C + c;
            const uint32_t n = P.ls_n[ix] + 1;
            P.ls_n[ix] = n;
            if (n == 1) { P.ls_mean[ix] = x[e]; P.ls_m2[ix] = 0.0; }
            else {
              const double mean = P.ls_mean[ix];
              const double newmean = mean + (x[e] - mean) / (double)n;
              P.ls_m2[ix] = P.ls_m2[ix] + (x[e] - mean) * (x[e] - newmean);
              P.ls_mean[ix] = newmean;
            }
          }
        }
      }
      const bool shrink = ratio < 0.05 && acc > 5;
      for (int a = 0; a < P.m.n_adapters; ++a) lex_rescale(a, shrink);
    }
    P.made[kc] = made; P.acc[kc] = acc;
    return accept;
  }
  // warp per chain: the components of a site are folded into its RunningVar as one batch -- with S1 = sum(x - mean),
  // S2 = sum (x - mean)^2 over the m components of the site: n' = n + m, mean' = mean + S1/n', M2' = M2 + S2 - S1^2/n'
  // (the closed form of m successive RunningVar:update calls; equal up to rounding)
  __device__ __forceinline__ bool drift_next_stream_lex(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    const bool adapting = P.doAdapt[k] != 0;
    const size_t kc = (size_t)k * P.C + c;
    unsigned long long made = P.made[kc] + 1, acc = P.acc[kc];
    const double* z = zcol();
    const int stride = zstride();
    const int ne = (d + G - 1) / G;
    const size_t base = (size_t)c * d;
    double* __restrict__ xw = P.x + base;
    const int A = P.m.n_adapters;
    constexpr bool FEW = FAM != FAM_INDEP;   // at most 3 adapters, chosen by the component index alone
    double s3[3] = {0.0, 0.0, 0.0};
    if (FEW) { for (int a = 0; a < 3; ++a) if (a < A) s3[a] = P.ls_scale[(size_t)a * P.C + c]; }
    auto scale_of = [&](int i) -> double {
      if (FEW) { const int a = adapter_of(i); return a == 0 ? s3[0] : (a == 1 ? s3[1] : s3[2]); }
      return P.ls_scale[(size_t)adapter_of(i) * P.C + c];
    };
    gaussian_fill<false>(iter, 0, ne);
    ModelStream<FAM> ms;
    double lp = 0.0;
    for (int e0 = 0; e0 < ne; e0 += RB) {
      double xv[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) { const int i = (e0 + r) * G + lane; xv[r] = i < d ? xw[i] : 0.0; }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int e = e0 + r;
        if (e < ne) {
          const int i = e * G + lane;
          double xpi = 0.0;
          if (i < d) xpi = __dadd_rn(xv[r], __dmul_rn(scale_of(i), z[e * stride]));
          if (e == 0) ms.prep(P, Grp<G>::bcast(xpi, 0), Grp<G>::bcast(xpi, 1));
          if (i < d) { typename ModelStream<FAM>::Obs ob; ms.fetch(P, i, d, ob); ms.term(P, i, xpi, ob, lp); }
        }
      }
    }
    rec_step = FEW ? s3[0] : P.ls_scale[c];
    double lpf = ms.finish(Grp<G>::sum(lp));
    if (FAM == FAM_INDEP && P.m.cond_lo) cond_bad = Grp<G>::sum(ms.bad); else cond_bad = 0.0;
    if (T != 1.0) lpf = lpf / T;
    const double thresh = lpf - P.lp[c];
    rec_dh = thresh;
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d);
    const bool accept = cond_bad == 0.0 && log(u) < thresh;   // `conditionsSatisfied and log(random()) < ...`
    if (accept) acc += 1;
    const double ratio = (double)acc / (double)made;
    const bool upd = adapting && ratio >= 0.05 && acc > 5;
    if (accept) {
      for (int e0 = 0; e0 < ne; e0 += RB) {
        double xv[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) { const int i = (e0 + r) * G + lane; xv[r] = i < d ? xw[i] : 0.0; }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int e = e0 + r; const int i = e * G + lane;
          if (e < ne && i < d) xw[i] = __dadd_rn(xv[r], __dmul_rn(scale_of(i), z[e * stride]));
        }
      }
      __syncwarp();
    }
    if (upd) {
      if (FEW) {
        double mean3[3] = {0.0, 0.0, 0.0}, S1[3] = {0.0, 0.0, 0.0}, S2[3] = {0.0, 0.0, 0.0};
        int cnt[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) if (a < A) mean3[a] = P.ls_mean[(size_t)a * P.C + c];
        for (int e0 = 0; e0 < ne; e0 += RB) {
          double xv[RB];
#pragma unroll
          for (int r = 0; r < RB; ++r) { const int i = (e0 + r) * G + lane; xv[r] = i < d ? xw[i] : 0.0; }
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const int e = e0 + r; const int i = e * G + lane;
            if (e < ne && i < d) {
              const int a = adapter_of(i);
#pragma unroll
              for (int q = 0; q < 3; ++q) if (a == q) { const double dv = xv[r] - mean3[q]; S1[q] += dv; S2[q] += dv * dv; cnt[q] += 1; }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          S1[q] = Grp<G>::sum(S1[q]); S2[q] = Grp<G>::sum(S2[q]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) cnt[q] += __shfl_xor_sync(0xffffffffu, cnt[q], o);
          if (lane == q && q < A && cnt[q] > 0) {
            const size_t ix = (size_t)q * P.C + c;
            const uint32_t n2 = P.ls_n[ix] + (uint32_t)cnt[q];
            const double inv = 1.0 / (double)n2;
            P.ls_n[ix] = n2;
            P.ls_mean[ix] = mean3[q] + S1[q] * inv;
            P.ls_m2[ix] = P.ls_m2[ix] + (S2[q] - S1[q] * S1[q] * inv);
          }
        }
      } else {
        for (int a = 0; a < A; ++a) {
          const size_t ix = (size_t)a * P.C + c;
          const double mean = P.ls_mean[ix];
          double S1 = 0.0, S2 = 0.0; int cnt = 0;
          for (int e = 0; e < ne; ++e) {
            const int i = e * G + lane;
            if (i < d && adapter_of(i) == a) { const double dv = xw[i] - mean; S1 += dv; S2 += dv * dv; cnt += 1; }
          }
          S1 = Grp<G>::sum(S1); S2 = Grp<G>::sum(S2);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
          if (leader() && cnt > 0) {
            const uint32_t n2 = P.ls_n[ix] + (uint32_t)cnt;
            const double inv = 1.0 / (double)n2;
            P.ls_n[ix] = n2;
            P.ls_mean[ix] = mean + S1 * inv;
            P.ls_m2[ix] = P.ls_m2[ix] + (S2 - S1 * S1 * inv);
          }
        }
      }
      __syncwarp();
    }
    if (adapting && lane < A) lex_rescale(lane, ratio < 0.05 && acc > 5);
    __syncwarp();
    if (leader()) {
      if (accept) { P.lp[c] = lpf; P.flags[c] = P.flags[c] & ~FLAG_GRAD_VALID; }
      P.made[kc] = made; P.acc[kc] = acc;
    }
    return accept;
  }

  // ---- TraceMHKernel:next with {doStruct=false} (mcmc.t:134-193): one random choice is re-proposed -- a gaussian ERP
  // drifts around its current value with its own sigma (erpdefs.t:46-57), every other ERP resamples from its prior
  // (erp.t:18-38) -- the whole program is rescored (float trace: no Jacobian), accept on (lp' - lp) + rvs - fwd.
  // Every lane of a chain's group computes the (chain-uniform) proposal itself.
  __device__ __forceinline__ void choice_params(int j, const double (&x)[NPL], int& kind, double& p0, double& p1) const {
    kind = ERP_GAUSSIAN; p0 = 0.0; p1 = 1.0;
    if (FAM == FAM_INDEP) {
      kind = P.m.erp_kind[j]; p0 = P.m.p0[j]; p1 = P.m.p1[j];
      if (P.m.lat_ref) {   // current parameters from the current values of the choices they depend on
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          const int r = P.m.lat_ref[2 * j + q];
          if (r >= 0) {
            const double yr = erp_fwd_noinline(P.m.erp_kind[r], component(x, r), P.m.p0[r], P.m.p1[r]);
            double dp;
            latent_param(P.m, j, q, yr, q ? p1 : p0, dp);
          }
        }
      }
    } else if (FAM == FAM_EIGHT_SCHOOLS) {
      if (j == 0) p1 = 5.0;
      else if (j >= 2) { p0 = Grp<G>::bcast(x[0], 0); p1 = exp(G == 1 ? x[NPL > 1 ? 1 : 0] : Grp<G>::bcast(x[0], 1)); }
    } else if (FAM == FAM_FUNNEL) {
      if (j == 0) p1 = 3.0; else p1 = exp(Grp<G>::bcast(x[0], 0) / 2.0);
    }
  }
  __device__ __forceinline__ double component(const double (&x)[NPL], int j) const {   // x_j, known to every lane of the group
    double v = 0.0;
    QSB_FOR_E { if (e * G + lane == j) v = x[e]; }
    return Grp<G>::sum(v);
  }
  __device__ __forceinline__ bool tracemh_next(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    double x[NPL];
    const size_t kc = (size_t)k * P.C + c;
    load(P.x, x);
    const int nch = P.m.n_choices;
    const int r = (int)(uniform_at(P.key, gc, iter, QSB_SLOT_MH_INDEX) * (double)nch);   // random()*numchoices, truncated (mcmc.t:171-172)
    rec_step = (double)r;
    const int j0 = FAM == FAM_INDEP ? P.m.erp_first[r] : r;
    SubStream s(P.key, gc, iter, QSB_SLOT_MH_PROPOSAL);
    double fwdlp, rvslp;
    int kind; double p0, p1;
    choice_params(j0, x, kind, p0, p1);
    if (FAM == FAM_INDEP && kind == ERP_DIRICHLET) {   // (every lane of the chain's group computes the chain-uniform proposal)
      // vector choice: K gammas from the proposal sub-stream, normalised (distrib.t:339-350); current value by stick
      // breaking, new unconstrained coordinates by its inverse (erp.t:205-233)
      const int K = P.m.erp_len[j0], Km1 = K - 1;
      double t[32], cur[32], alpha[32];
      double ssum = 0.0, asum = 0.0, stick = 1.0;
      for (int q = 0; q < K; ++q) {
        alpha[q] = P.m.p0[j0 + q];
        const double xq = component(x, j0 + q);
        if (q < Km1) { const double z = logistic_(xq - log((double)(Km1 - q))); cur[q] = stick * z; stick = stick - cur[q]; } else cur[q] = stick;
        t[q] = gamma_sample_seq(s, alpha[q], 1.0); ssum = ssum + t[q]; asum = asum + alpha[q];
      }
      for (int q = 0; q < K; ++q) t[q] = t[q] / ssum;
      fwdlp = log_gamma5(asum); rvslp = fwdlp;
      for (int q = 0; q < K; ++q) {
        fwdlp = fwdlp + (alpha[q] - 1.0) * log(t[q]); fwdlp = fwdlp - log_gamma5(alpha[q]);
        rvslp = rvslp + (alpha[q] - 1.0) * log(cur[q]); rvslp = rvslp - log_gamma5(alpha[q]);
      }
      double sticklen = t[Km1];
      QSB_FOR_E { if (e * G + lane == j0 + Km1) x[e] = 0.0; }
      for (int q = Km1 - 1; q >= 0; --q) {
        sticklen = sticklen + t[q];
        const double xn = invlogistic_(t[q] / sticklen) + log((double)(Km1 - q));
        QSB_FOR_E { if (e * G + lane == j0 + q) x[e] = xn; }
      }
    } else {
      const double cur_x = component(x, j0);
      const double cur_v = erp_fwd_noinline(kind, cur_x, p0, p1);
      double new_v, new_x;
      if (kind == ERP_GAUSSIAN) {
        new_v = gaussian_sample_seq(s, cur_v, p1);
        fwdlp = gaussian_lp(new_v, cur_v, p1);
        rvslp = gaussian_lp(cur_v, new_v, p1);
        new_x = new_v;
      } else {
        new_v = erp_sample_seq(kind, s, p0, p1);
        fwdlp = erp_prior_lp(kind, new_v, p0, p1, nullptr);
        rvslp = erp_prior_lp(kind, cur_v, p0, p1, nullptr);
        new_x = erp_inv(kind, new_v, p0, p1);
      }
      QSB_FOR_E { if (e * G + lane == j0) x[e] = new_x; }
    }
    double lpf, gdummy[NPL];
    model_eval<FAM, G, NPL, false, true, false>(P, lane, x, gdummy, lpf, &cond_bad);
    if (T != 1.0) lpf = lpf / T;
    const double thresh = (lpf - P.lp[c]) + rvslp - fwdlp;
    rec_dh = thresh;
    const double u = uniform_at(P.key, gc, iter, QSB_SLOT_MH_ACCEPT);
    const bool accept = cond_bad == 0.0 && log(u) < thresh;   // `conditionsSatisfied and log(random()) < ...`
    if (accept) store(P.x, x);
    if (G == 32) __syncwarp();
    if (leader()) {
      P.made[kc] += 1;
      if (accept) { P.acc[kc] += 1; P.lp[c] = lpf; P.flags[c] = P.flags[c] & ~FLAG_GRAD_VALID; }
    }
    return accept;
  }

  // ---- HARMKernel:next (harm.t:60-113)
  __device__ __forceinline__ bool harm_next(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    double x[NPL], xp[NPL];
    const bool adapting = P.doAdapt[k] != 0;
    const size_t kc = (size_t)k * P.C + c;
    load(P.x, x);
    double nrm = 0.0;
    gaussian_fill<false>(iter, 0, NPL);
    {
      const double* z = zcol();
      const int stride = zstride();
      QSB_FOR_E {
        int i = e * G + lane;
        xp[e] = i < d ? z[e * stride] : 0.0;
        nrm += xp[e] * xp[e];
      }
    }
    nrm = sqrt(Grp<G>::sum(nrm));
    const double hscale = P.ha_scale[c];
    rec_step = hscale;
    const double step = uniform_at(P.key, gc, iter, (uint32_t)d) * hscale;
    QSB_FOR_E {
      int i = e * G + lane;
      xp[e] = i < d ? x[e] + (step * (xp[e] / nrm)) : 0.0;
    }
    double lpf, gdummy[NPL];
    model_eval<FAM, G, NPL, false, true, (FAM == FAM_GAUSS_PREC && QSB_GP_SMEM)>(P, lane, xp, gdummy, lpf, &cond_bad);
    if (T != 1.0) lpf = lpf / T;
    double t = 0.234;
    if (d < 5) { double tt = (d - 1) / 4.0; t = (1.0 - tt) * 0.44 + tt * 0.234; }
    const double thresh = lpf - P.lp[c];
    rec_dh = thresh;
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d + 1u);
    const bool accept = cond_bad == 0.0 && log(u) < thresh;   // `conditionsSatisfied and log(random()) < ...`
    double ns = hscale;
    if (accept) {
      store(P.x, xp);
      if (adapting) ns = hscale + (hscale / (t * (1.0 - t))) * (1.0 - t) / (double)(iter + 1u);
    } else {
      if (adapting) ns = fabs(hscale - (hscale / (t * (1.0 - t))) * t / (double)(iter + 1u));
    }
    if (leader()) {
      P.made[kc] += 1;
      if (accept) { P.acc[kc] += 1; P.lp[c] = lpf; P.flags[c] = P.flags[c] & ~FLAG_GRAD_VALID; }
      P.ha_scale[c] = ns;
    }
    return accept;
  }

  // ---- warp-per-chain streaming forms of the two Metropolis kernels: no per-lane vectors in registers, the proposal is
  // never written to HBM (on acceptance it is recomputed from the staged Gaussians with the same operations), global
  // loads are issued RB rows ahead of their use.  Against the register forms above (and the reference's expression
  // order) three quotients become products with a reciprocal -- sigma*(v/u) for (sigma*v)/u, m2*(1/(n-1)), (x-mean)*(1/n),
  // z*(1/norm) -- each within 1 ulp.
  static constexpr int RB = 4;
  // Addressing: every array of the chain is walked through a per-lane pointer advanced once per RB rows, so that an
  // element is reached with an immediate offset (row r of the chunk = +32*r doubles) and no per-element index arithmetic;
  // TPB (= blockDim.x, the row stride of the Gaussian staging array) is a compile-time constant for the same reason.
  template <int TPB>
  __device__ __forceinline__ bool drift_next_stream(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    const bool adapting = P.doAdapt[k] != 0;
    const size_t kc = (size_t)k * P.C + c;
    unsigned long long made = P.made[kc] + 1, acc = P.acc[kc];
    const uint32_t n_rv = P.dr_n[c];
    const double s0 = P.dr_s0[c], mult = P.dr_mult[c];
    const int ne = (d + G - 1) / G;                     // component rows of this chain
    const size_t base = (size_t)c * d + lane;
    const bool use_var = n_rv > 100;                    // scale = varEst:getStdDev() [* 0.99]  (drift.t:288-298)
    const double inv_nm1 = use_var ? 1.0 / (double)(n_rv - 1) : 0.0;
    gaussian_fill<false>(iter, 0, ne);
    ModelStream<FAM> ms;
    double lp = 0.0;
    {
      // x and m2 of rows e0+RB.. are requested while rows e0.. are scored (double buffer in registers); the observation
      // data are model constants shared by every chain (L1 hits) and are read at their use
      const double* __restrict__ xl = P.x + base;
      const double* __restrict__ ml = P.dr_m2 + base;
      const double* zl = zcol();
      int i0 = lane;                                   // component index of row e0 for this lane
      double xn[RB], mn[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const bool ok = i0 + 32 * r < d;
        xn[r] = ok ? xl[32 * r] : 0.0;
        mn[r] = (ok && use_var) ? ml[32 * r] : 0.0;
      }
      for (int e0 = 0; e0 < ne; e0 += RB) {
        double xv[RB], mv[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) { xv[r] = xn[r]; mv[r] = mn[r]; }
        xl += 32 * RB; ml += 32 * RB;
        if (e0 + RB < ne) {
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const bool ok = i0 + 32 * (RB + r) < d;
            xn[r] = ok ? xl[32 * r] : 0.0;
            mn[r] = (ok && use_var) ? ml[32 * r] : 0.0;
          }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int i = i0 + 32 * r;
          double xpi = 0.0;
          if (i < d) {
            double sc = s0;
            if (use_var) { sc = sqrt(mv[r] * inv_nm1); if (mult != 1.0) sc = sc * mult; }
            xpi = __dadd_rn(xv[r], __dmul_rn(sc, zl[r * TPB]));
          }
          if (e0 == 0 && r == 0) ms.prep(P, Grp<G>::bcast(xpi, 0), Grp<G>::bcast(xpi, 1));
          if (i < d) { typename ModelStream<FAM>::Obs ob; ms.fetch(P, i, d, ob); ms.term(P, i, xpi, ob, lp); }
        }
        i0 += 32 * RB; zl += RB * TPB;
      }
    }
    rec_step = s0;
    double lpf = ms.finish(Grp<G>::sum(lp));
    if (FAM == FAM_INDEP && P.m.cond_lo) cond_bad = Grp<G>::sum(ms.bad); else cond_bad = 0.0;
    if (T != 1.0) lpf = lpf / T;
    const double thresh = lpf - P.lp[c];
    rec_dh = thresh;
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d);
    const bool accept = cond_bad == 0.0 && log(u) < thresh;   // `conditionsSatisfied and log(random()) < ...`
    if (accept) acc += 1;
    const double ratio = (double)acc / (double)made;
    const bool upd = adapting && ratio >= 0.05 && acc > 5;   // RunningVar:update with the current state (drift.t:26-37, 279-284)
    const uint32_t n_new = upd ? n_rv + 1 : n_rv;
    if (accept || upd) {
      const double inv_n = 1.0 / (double)n_new;
      double* __restrict__ xl = P.x + base;
      double* __restrict__ ml = P.dr_m2 + base;
      double* __restrict__ al = P.dr_mean + base;
      const double* zl = zcol();
      const bool need_m2 = use_var || (upd && n_new != 1), need_mean = upd && n_new != 1;
      int i0 = lane;
      for (int e0 = 0; e0 < ne; e0 += RB) {
        double xv[RB], mv[RB], av[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const bool ok = i0 + 32 * r < d;
          xv[r] = ok ? xl[32 * r] : 0.0;
          mv[r] = (ok && need_m2) ? ml[32 * r] : 0.0;
          av[r] = (ok && need_mean) ? al[32 * r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          if (i0 + 32 * r < d) {
            double cur = xv[r];
            if (accept) {   // the accepted proposal, recomputed exactly as it was scored
              double sc = s0;
              if (use_var) { sc = sqrt(mv[r] * inv_nm1); if (mult != 1.0) sc = sc * mult; }
              cur = __dadd_rn(xv[r], __dmul_rn(sc, zl[r * TPB]));
              xl[32 * r] = cur;
            }
            if (upd) {
              if (n_new == 1) { al[32 * r] = cur; ml[32 * r] = 0.0; }
              else {
                const double mean = av[r];
                const double newmean = mean + (cur - mean) * inv_n;
                ml[32 * r] = mv[r] + (cur - mean) * (cur - newmean);
                al[32 * r] = newmean;
              }
            }
          }
        }
        i0 += 32 * RB; xl += 32 * RB; ml += 32 * RB; al += 32 * RB; zl += RB * TPB;
      }
    }
    __syncwarp();
    if (leader()) {
      if (accept) { P.lp[c] = lpf; P.flags[c] = P.flags[c] & ~FLAG_GRAD_VALID; }
      if (adapting) {
        const bool shrink = ratio < 0.05 && acc > 5;
        P.dr_n[c] = n_new;
        if (n_new > 100) P.dr_mult[c] = shrink ? 0.99 : 1.0;
        else if (shrink) P.dr_s0[c] = s0 * 0.99;
      }
      P.made[kc] = made; P.acc[kc] = acc;
    }
    return accept;
  }

  __device__ __forceinline__ bool harm_next_stream(int k, uint32_t iter, double& rec_dh, double& rec_step) const {
    const bool adapting = P.doAdapt[k] != 0;
    const size_t kc = (size_t)k * P.C + c;
    const int stride = zstride();
    const int ne = (d + G - 1) / G;
    const size_t base = (size_t)c * d;
    double* __restrict__ xw = P.x + base;
    gaussian_fill<false>(iter, 0, ne);
    const double* z = zcol();       // (after the fill: the warp-specialised kernel hands the buffer over in it)
    double nrm = 0.0;
    for (int e = 0; e < ne; ++e) { if (e * G + lane < d) { const double zi = z[e * stride]; nrm += zi * zi; } }
    nrm = sqrt(Grp<G>::sum(nrm));
    const double hscale = P.ha_scale[c];
    rec_step = hscale;
    const double step = uniform_at(P.key, gc, iter, (uint32_t)d) * hscale;
    const double sn = step * (1.0 / nrm);      // x' = x + step*(z/norm) (harm.t:78-79) as x + (step/norm)*z
    ModelStream<FAM> ms;
    double lp = 0.0;
    for (int e0 = 0; e0 < ne; e0 += RB) {
      double xv[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) { const int i = (e0 + r) * G + lane; xv[r] = i < d ? xw[i] : 0.0; }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int e = e0 + r;
        if (e < ne) {
          const int i = e * G + lane;
          double xpi = 0.0;
          if (i < d) xpi = fma(sn, z[e * stride], xv[r]);
          if (e == 0) ms.prep(P, Grp<G>::bcast(xpi, 0), Grp<G>::bcast(xpi, 1));
          if (i < d) { typename ModelStream<FAM>::Obs ob; ms.fetch(P, i, d, ob); ms.term(P, i, xpi, ob, lp); }
        }
      }
    }
    double lpf = ms.finish(Grp<G>::sum(lp));
    if (FAM == FAM_INDEP && P.m.cond_lo) cond_bad = Grp<G>::sum(ms.bad); else cond_bad = 0.0;
    if (T != 1.0) lpf = lpf / T;
    double t = 0.234;
    if (d < 5) { double tt = (d - 1) / 4.0; t = (1.0 - tt) * 0.44 + tt * 0.234; }
    const double thresh = lpf - P.lp[c];
    rec_dh = thresh;
    const double u = uniform_at(P.key, gc, iter, (uint32_t)d + 1u);
    const bool accept = cond_bad == 0.0 && log(u) < thresh;   // `conditionsSatisfied and log(random()) < ...`
    double ns = hscale;
    if (accept) {
      for (int e0 = 0; e0 < ne; e0 += RB) {
        double xv[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) { const int i = (e0 + r) * G + lane; xv[r] = i < d ? xw[i] : 0.0; }
#pragma unroll
        for (int r = 0; r < RB; ++r) { const int e = e0 + r; const int i = e * G + lane; if (e < ne && i < d) xw[i] = fma(sn, z[e * stride], xv[r]); }
      }
      if (adapting) ns = hscale + (hscale / (t * (1.0 - t))) * (1.0 - t) / (double)(iter + 1u);
    } else {
      if (adapting) ns = fabs(hscale - (hscale / (t * (1.0 - t))) * t / (double)(iter + 1u));
    }
    __syncwarp();
    if (leader()) {
      P.made[kc] += 1;
      if (accept) { P.acc[kc] += 1; P.lp[c] = lpf; P.flags[c] = P.flags[c] & ~FLAG_GRAD_VALID; }
      P.ha_scale[c] = ns;
    }
    return accept;
  }

  // ---- one iteration of doMCMC's loop body (mcmc.t:57-72)
  template <bool HAS_HMC, bool LEX, bool DRIFT_ONLY>
  __device__ __forceinline__ void iterate(uint64_t it) const {
    const uint32_t iter = (uint32_t)it;
    const uint64_t lit = it - P.it_origin;       // iteration of this run (mcmc.t:57: i = 0 .. iters-1)
    if (P.temps) { T = P.temps[lit < P.ntemps ? lit : P.ntemps - 1]; invT = 1.0 / T; }   // AnnealingKernel:next (mcmc.t:308-311)
    int k = 0;
    if (P.nsub > 1) {   // MixtureKernel:next (mcmc.t:275-287) + categorical_array.sample (distrib.t:281-292)
      double sum = 0.0;
      for (int j = 0; j < P.nsub; ++j) sum = sum + P.weight[j];
      double xr = uniform_at(P.key, gc, iter, QSB_SLOT_MIXTURE) * sum;
      double pa = 0.0;
      int result = 0;
      do { pa = pa + P.weight[result]; result = result + 1; } while (!(pa > xr || result == P.nsub));
      k = result - 1;
    }
    double rdh = 0.0, rstep = 0.0;
    bool accept;
    const int kind = P.kind[k];
    constexpr bool STREAM = G == 32 && FAM != FAM_GAUSS_PREC;
    if (HAS_HMC && kind == K_HMC) {
      if constexpr (SCALAR_GRAD) accept = hmc_next_warp(k, iter, rdh, rstep);
      else accept = hmc_next(k, iter, rdh, rstep);
    }
    else if (LEX && kind == K_TRACEMH) accept = tracemh_next(k, iter, rdh, rstep);
    // the streaming Drift / HARM scorers take a program term by term; INDEP programs with vector choices or latent parameters
    // (vscratch set) are scored by the general evaluator through the non-streaming forms
    else if (DRIFT_ONLY || kind == K_DRIFT) {
      if (LEX && P.dr_lexical) {
        if constexpr (STREAM && FAM != FAM_INDEP) accept = drift_next_stream_lex(k, iter, rdh, rstep);
        else if constexpr (STREAM) accept = P.vscratch ? drift_next_lex(k, iter, rdh, rstep) : drift_next_stream_lex(k, iter, rdh, rstep);
        else accept = drift_next_lex(k, iter, rdh, rstep);
      } else {
        constexpr int ZS = WS ? 32 : ((G == 1 || HAS_HMC) ? 128 : 256);
        if constexpr (STREAM && FAM != FAM_INDEP) accept = drift_next_stream<ZS>(k, iter, rdh, rstep);
        else if constexpr (STREAM) accept = P.vscratch ? drift_next(k, iter, rdh, rstep) : drift_next_stream<ZS>(k, iter, rdh, rstep);
        else accept = drift_next(k, iter, rdh, rstep);
      }
    }
    else if constexpr (STREAM && FAM != FAM_INDEP) accept = harm_next_stream(k, iter, rdh, rstep);
    else if constexpr (STREAM) accept = P.vscratch ? harm_next(k, iter, rdh, rstep) : harm_next_stream(k, iter, rdh, rstep);
    else accept = harm_next(k, iter, rdh, rstep);
    if constexpr (WS) ws_release();
    // health counters: a non-finite accept threshold, an HMC energy error beyond 1000
    if (!(fabs(rdh) <= 1.7976931348623157e308)) count(2, 1);
    else if (HAS_HMC && kind == K_HMC && fabs(rdh) > 1000.0) count(3, 1);
    if (G == 32) __syncwarp();
    const bool keep = lit >= P.burnin && (lit % P.lag) == 0 && c < P.sample_chains && (P.samples != nullptr || P.mom != nullptr);
    if (keep || P.rec_state) {
      double x[NPL];
      load(P.x, x);
      if (P.rec_state) {
        size_t r = (size_t)c * P.rec_iters + (it - P.it_begin);
        QSB_FOR_E { int i = e * G + lane; if (i < d) P.rec_state[r * d + i] = x[e]; }
        if (leader()) {
          P.rec_accept[r] = accept ? 1 : 0; P.rec_dh[r] = rdh; P.rec_step[r] = rstep;
          if (P.rec_kidx) P.rec_kidx[r] = k;
        }
      }
      if (keep) {
        const uint64_t sidx = lit / P.lag - (P.burnin + P.lag - 1) / P.lag;
        if (sidx < P.max_samples) {
          const uint32_t SC = P.sample_chains;
          // where kept value j of this chain goes: the sample array, or (QSB_FLAG_MOMENTS) the running statistics
          auto put = [&](int j, double v) {
            if (P.mom) {
              const size_t plane = (size_t)SC * P.retdim;
              double* base = P.mom + (G == 1 ? (size_t)j * SC + c : (size_t)c * P.retdim + j);
              mom_update([&](int st) { return base + (size_t)st * plane; }, sidx, P.mom_half, P.mom_batch, v);
            } else P.samples[G == 1 ? ((size_t)sidx * P.retdim + j) * SC + c : ((size_t)sidx * SC + c) * P.retdim + j] = v;
          };
          if (FAM == FAM_INDEP && G == 32 && P.vscratch && P.m.has_vector) {
            // the values of a dirichlet choice come out of the stick-breaking recurrence in component order: lane 0 walks
            // the chain's components through the scratch slice, the lanes then store (or sum) their own
            double* w = P.vscratch + (size_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5) * VS_ARRAYS * P.vs_dpad;
            QSB_FOR_E { const int i = e * G + lane; if (i < d) w[i] = x[e]; }
            __syncwarp();
            if (lane == 0) {
              SimplexCarry sc; sc.stick = 1.0; sc.lp = 0.0; sc.lj = 0.0;
              double* ys = w + P.vs_dpad;
              for (int i = 0; i < d; ++i) {
                double gdummy;
                if (P.m.erp_kind[i] == ERP_DIRICHLET) ys[i] = dirichlet_component_noinline(sc, (int)P.m.p1[i], P.m.erp_len[i], w[i], P.m.p0[i], P.m.erp_asum[i], gdummy);
                else ys[i] = erp_fwd_noinline(P.m.erp_kind[i], w[i], P.m.p0[i], P.m.p1[i]);
              }
            }
            __syncwarp();
            if (P.m.ret_mode == 0) {
              if (lane == 0) { double sum = 0.0; for (int i = 0; i < d; ++i) sum += w[P.vs_dpad + i]; put(0, sum / P.m.ret_div); }
            } else {
              QSB_FOR_E { const int i = e * G + lane; if (i < d) put(i, w[P.vs_dpad + i]); }
            }
            __syncwarp();
          } else if (FAM == FAM_INDEP) {
            double sum = 0.0;
            SimplexCarry sc; sc.stick = 1.0; sc.lp = 0.0; sc.lj = 0.0;
            QSB_FOR_E {
              int i = e * G + lane;
              if (i < d) {
                double y, gdummy;
                if (G == 1 && P.m.erp_kind[i] == ERP_DIRICHLET)
                  y = dirichlet_component_noinline(sc, (int)P.m.p1[i], P.m.erp_len[i], x[e], P.m.p0[i], P.m.erp_asum[i], gdummy);
                else y = erp_fwd_noinline(P.m.erp_kind[i], x[e], P.m.p0[i], P.m.p1[i]);
                if (P.m.ret_mode == 0) sum += y;
                else put(i, y);
              }
            }
            if (P.m.ret_mode == 0) {
              sum = Grp<G>::sum(sum);
              if (leader()) put(0, sum / P.m.ret_div);
            }
          } else {
            QSB_FOR_E {
              int i = e * G + lane;
              if (i < d) put(i, x[e]);
            }
          }
          if (leader() && P.samp_lp) P.samp_lp[(size_t)sidx * SC + c] = P.lp[c] * T;   // mcmc.t:69
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Kernels
// HAS_HMC = false: the kernel list holds no HMCKernel; the trajectory code (and its register vectors) is not compiled
// in, which lets the warp-per-chain Drift / HARM kernels run at 3 blocks per SM.
// LEX: the general kernel -- everything above plus DriftKernel{lexicalScaleSharing} and TraceMHKernel; always built with
// HAS_HMC = true.  The specialised kernels do not carry that code, so their register allocation is not bent by it.
// DRIFT_ONLY: a single DriftKernel (config c4): no HARM code in the kernel, so the register allocator serves one path.
template <int FAM, int G, int NPL, bool HAS_HMC, bool LEX = false, bool DRIFT_ONLY = false>
__global__ void __launch_bounds__(G < 32 || HAS_HMC ? 128 : 256, G == 1 ? 4 : (G == 32 ? 3 : 5))
advance_kernel(const __grid_constant__ EngineParams P) {
  const uint32_t grp = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G);
  if (FAM == FAM_GAUSS_PREC) {
    extern __shared__ double qsb_zbuf[];
    double* sS = qsb_zbuf + (size_t)Chain<FAM, G, NPL>::NB * blockDim.x;   // after the Gaussian staging columns
    for (int t = threadIdx.x; t < G * NPL * G * NPL; t += blockDim.x) sS[t] = P.Sc[t & 255];
    __syncthreads();
  }
  const bool active = grp < P.C;               // whole warps: C*G and the block size are multiples of 32 or the tail warp idles
  Chain<FAM, G, NPL> ch(P, active ? grp : 0);
  if (active)
    for (uint64_t it = P.it_begin; it < P.it_begin + P.n_iters; ++it) ch.template iterate<HAS_HMC, LEX, DRIFT_ONLY>(it);
  ch.flush_counts();
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised form of advance_kernel for the warp-per-chain mapping (G = 32): the Gaussian sampling (Philox + Leva:
// half of the fused kernels' instructions, ~40 registers) and the transition itself (the trajectory state: 2 x NPL doubles per
// lane) run on DIFFERENT warps of one CTA with different register budgets (setmaxnreg) and meet in shared memory.
//   warpgroup 0   (warps 0..3)          4 trajectory warps, one chain slot each
//   warpgroups 1..NPROD (4 NPROD warps)  sampler warps: warp j of warpgroup 1+p fills rows [p RP, (p+1) RP) of slot j's buffer
// Every iteration of every chain needs exactly the standard normals z_i(chain, iteration), i < n, whatever kernel the
// mixture picks (HMC momenta, HARM direction, drift proposal: slots 0..n-1 of the iteration's sub-streams), so the sampler
// warps run ahead through the same (chain, iteration) sequence as their trajectory warp, two staging buffers per slot
// alternating under a full / empty mbarrier pair.  Only the probes of the step-size search (once per chain) are drawn by the
// trajectory warp itself.  The kernel is persistent: gridDim = CTAs that fit the device, slot s works off chains s, s + S, ...
static constexpr int WS_CONSUMER_REGS = 160, WS_PRODUCER_REGS = 40;
template <int FAM, int NPL, bool HAS_HMC, bool DRIFT_ONLY, int NPROD>
__global__ void __launch_bounds__(128 * (1 + NPROD), 2)
advance_ws_kernel(const __grid_constant__ EngineParams P) {
  extern __shared__ double qsb_zbuf[];
  constexpr int BUF = NPL * 32;                      // doubles per staging buffer
  constexpr int SLOT = (HAS_HMC ? 3 : 2) * BUF;      // two hand-off buffers (+ the trajectory warp's own)
  uint64_t* bars = reinterpret_cast<uint64_t*>(qsb_zbuf + 4 * SLOT);    // per slot: full[2], empty[2]
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
  const int slot = warp & 3, role = warp >> 2;       // role 0: trajectory; 1..NPROD: sampler
  if (threadIdx.x < 4) {
    uint64_t* b = bars + 4 * threadIdx.x;
    mbar_init(&b[0], NPROD); mbar_init(&b[1], NPROD); mbar_init(&b[2], 1); mbar_init(&b[3], 1);
    mbar_fence_init();
  }
  __syncthreads();
  double* zslot = qsb_zbuf + (size_t)slot * SLOT;
  uint64_t* full = bars + 4 * slot;
  uint64_t* empty = full + 2;
  const uint32_t nslots = gridDim.x * 4u, s0 = blockIdx.x * 4u + (uint32_t)slot;
  if (role == 0) {
    reg_alloc<WS_CONSUMER_REGS>();
    WsLink link{zslot, HAS_HMC ? zslot + 2 * BUF : nullptr, full, empty, 0u, false};
    unsigned long long tot[4] = {0ull, 0ull, 0ull, 0ull};
    for (uint32_t c = s0; c < P.C; c += nslots) {
      Chain<FAM, 32, NPL, true> ch(P, c, &link);
      for (uint64_t it = P.it_begin; it < P.it_begin + P.n_iters; ++it) ch.template iterate<HAS_HMC, false, DRIFT_ONLY>(it);
      for (int w = 0; w < 4; ++w) tot[w] += ch.cnt[w];
    }
    if (lane == 0) for (int w = 0; w < 4; ++w) if (tot[w]) atomicAdd(&P.counters[w], tot[w]);
  } else {
    reg_dealloc<WS_PRODUCER_REGS>();
    constexpr int RP = (NPL + NPROD - 1) / NPROD;    // rows per sampler warp
    const int row0 = (role - 1) * RP;
    const int d = P.d;
    const int n = min(d, (row0 + RP) * 32) - row0 * 32;   // components this warp draws (may be <= 0 past the last row)
    uint32_t k = 0;
    for (uint32_t c = s0; c < P.C; c += nslots) {
      const uint32_t gc = P.chain_begin + c;
      for (uint64_t it = P.it_begin; it < P.it_begin + P.n_iters; ++it, ++k) {
        const uint32_t b = k & 1u;
        mbar_wait(&empty[b], ((k >> 1) & 1u) ^ 1u);      // (a fresh barrier passes the wait on the preceding phase)
        if (n > 0) gaussian_fill_body<32>(P.key.k0, P.key.k1, gc, (uint32_t)it, row0, n, zslot + (size_t)b * BUF + row0 * 32, 32);
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[b]);
      }
    }
  }
}

// The forward sample of a whole INDEP chain in program order by ONE lane (warp-per-chain mapping with vector choices or latent
// parameters: a later choice reads the sampled values of earlier ones, a dirichlet choice is one draw of K components).  The
// statements are those of the thread-per-chain branch of init_kernel below, over arrays in the warp's scratch slice.
static __device__ __noinline__ void indep_init_seq(const EngineParams& P, uint32_t gc, uint32_t it_init, double* xs, double* yv, double& lp_out, double& bad_out) {
  const ModelDev& m = P.m;
  const int d = m.dim;
  double lp = 0.0, bad = 0.0;
  for (int i = 0; i < d; ++i) xs[i] = 0.0;
  for (int i = 0; i < d; ++i) {
    const int kind = m.erp_kind[i];
    double p0 = m.p0[i], p1 = m.p1[i];
    if (m.lat_ref) {
      double dp;
      const int r0 = m.lat_ref[2 * i], r1 = m.lat_ref[2 * i + 1];
      if (r0 >= 0) latent_param(m, i, 0, yv[r0], p0, dp);
      if (r1 >= 0) latent_param(m, i, 1, yv[r1], p1, dp);
    }
    const uint32_t slot = m.erp_slot ? (uint32_t)m.erp_slot[i] : (uint32_t)i;
    if (kind == ERP_DIRICHLET) {
      if ((int)p1 == 0) {
        const int K = m.erp_len[i], Km1 = K - 1;
        SubStream s(P.key, gc, it_init, slot);
        double t[32];
        double ssum = 0.0, asum = 0.0;
        for (int k = 0; k < K; ++k) { t[k] = gamma_sample_seq(s, m.p0[i + k], 1.0); ssum = ssum + t[k]; asum = asum + m.p0[i + k]; }
        for (int k = 0; k < K; ++k) t[k] = t[k] / ssum;
        double lpv = log_gamma5(asum);
        for (int k = 0; k < K; ++k) { const double a = m.p0[i + k]; lpv = lpv + (a - 1.0) * log(t[k]); lpv = lpv - log_gamma5(a); }
        lp += lpv;
        double sticklen = t[Km1];
        for (int k = Km1 - 1; k >= 0; --k) {
          sticklen = sticklen + t[k];
          const double z = t[k] / sticklen;
          xs[i + k] = invlogistic_(z) + log((double)(Km1 - k));
        }
      }
      yv[i] = 0.0;
    } else {
      SubStream s(P.key, gc, it_init, slot);
      const double y = erp_sample_seq(kind, s, p0, p1);
      lp += erp_prior_lp(kind, y, p0, p1, nullptr);
      if (m.fsigma && m.fsigma[i] > 0.0) lp += gaussian_lp(y, m.fmu[i], m.fsigma[i]);
      xs[i] = erp_inv(kind, y, p0, p1);
      if (m.cond_lo && !(y > m.cond_lo[i] && y < m.cond_hi[i])) bad += 1.0;
      yv[i] = y;
    }
  }
  lp_out = lp; bad_out = bad;
}

// Forward-sampling initialisation (trace.t:592-624: update(true) creates every choice by sampling it,
// erp.t:567-572) followed by the kernels' first gather of unconstrained components (erp.t:519-522).
// The ERPs of a chain are sampled in program order; ERP i reads sub-stream (chain, ITER_INIT, i).
template <int FAM, int G, int NPL>
__global__ void __launch_bounds__(G == 1 ? 128 : 256) init_kernel(const __grid_constant__ EngineParams P) {
  const uint32_t grp = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G);
  if (grp >= P.C) return;
  Chain<FAM, G, NPL> ch(P, grp);
  const int lane = ch.lane, d = P.d;
  double x[NPL];
  double lp0 = 0.0;
  if (FAM == FAM_INDEP && G == 32 && P.vscratch) {
    double* w = P.vscratch + (size_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5) * VS_ARRAYS * P.vs_dpad;
    double lp = 0.0;
    for (uint32_t attempt = 0;; ++attempt) {     // rejection initialisation as below
      double bad = 0.0;
      if (lane == 0) indep_init_seq(P, ch.gc, QSB_ITER_INIT - attempt, w, w + 2 * P.vs_dpad, lp, bad);
      __syncwarp();
      bad = __shfl_sync(0xffffffffu, bad, 0);
      if (!P.m.cond_lo || bad == 0.0 || attempt >= 100000u) break;
    }
    lp0 = __shfl_sync(0xffffffffu, lp, 0);
    QSB_FOR_E { const int i = e * G + lane; x[e] = i < d ? w[i] : 0.0; }
    __syncwarp();
  } else if (FAM == FAM_INDEP) {
    double lp = 0.0;
    double yv[G == 1 ? NPL : 1];   // sampled values (programs with latent parameters: later choices read them)
    // rejection initialisation (trace.t:612-623): re-run the program from scratch until every qs.condition holds; attempt a
    // reads the sub-streams of iteration code QSB_ITER_INIT - a
    for (uint32_t attempt = 0;; ++attempt) {
    const uint32_t it_init = QSB_ITER_INIT - attempt;
    double bad = 0.0;
    lp = 0.0;
    QSB_FOR_E x[e] = 0.0;
    QSB_FOR_E {
      int i = e * G + lane;
      if (i < d) {
        int kind = P.m.erp_kind[i];
        double p0 = P.m.p0[i], p1 = P.m.p1[i];
        if (G == 1 && P.m.lat_ref) {
          double dp;
          const int r0 = P.m.lat_ref[2 * i], r1 = P.m.lat_ref[2 * i + 1];
          if (r0 >= 0) latent_param(P.m, i, 0, yv[r0 < NPL ? r0 : 0], p0, dp);
          if (r1 >= 0) latent_param(P.m, i, 1, yv[r1 < NPL ? r1 : 0], p1, dp);
        }
        const uint32_t slot = P.m.erp_slot ? (uint32_t)P.m.erp_slot[i] : (uint32_t)i;   // ERP index (trace.t:948-994 creates choices in program order)
        if (G == 1 && kind == ERP_DIRICHLET) {
          if ((int)p1 == 0) {
            // dirichlet.sample (distrib.t:339-350): K gammas from the choice's sub-stream, normalised; logprob of the
            // sampled vector (distrib.t:352-362); inverse stick breaking (erp.t:219-233), x[K-1] = 0
            const int K = P.m.erp_len[i], Km1 = K - 1;
            SubStream s(P.key, ch.gc, it_init, slot);
            double t[32];
            double ssum = 0.0, asum = 0.0;
            for (int k = 0; k < K; ++k) { t[k] = gamma_sample_seq(s, P.m.p0[i + k], 1.0); ssum = ssum + t[k]; asum = asum + P.m.p0[i + k]; }
            for (int k = 0; k < K; ++k) t[k] = t[k] / ssum;
            double lpv = log_gamma5(asum);
            for (int k = 0; k < K; ++k) { const double a = P.m.p0[i + k]; lpv = lpv + (a - 1.0) * log(t[k]); lpv = lpv - log_gamma5(a); }
            lp += lpv;
            double sticklen = t[Km1];
            for (int k = Km1 - 1; k >= 0; --k) {
              sticklen = sticklen + t[k];
              const double z = t[k] / sticklen;
              x[e + k < NPL ? e + k : NPL - 1] = invlogistic_(z) + log((double)(Km1 - k));
            }
          }
        } else {
          SubStream s(P.key, ch.gc, it_init, slot);
          double y = erp_sample_seq(kind, s, p0, p1);
          lp += erp_prior_lp(kind, y, p0, p1, nullptr);   // float trace: logprob of the sampled value itself
          if (P.m.fsigma && P.m.fsigma[i] > 0.0) lp += gaussian_lp(y, P.m.fmu[i], P.m.fsigma[i]);   // and the factor that follows it
          x[e] = erp_inv(kind, y, p0, p1);
          if (P.m.cond_lo && !(y > P.m.cond_lo[i] && y < P.m.cond_hi[i])) bad += 1.0;
          if (G == 1) yv[e] = y;
        }
      }
    }
    if (!P.m.cond_lo || Grp<G>::sum(bad) == 0.0 || attempt >= 100000u) break;
    }
    lp0 = Grp<G>::sum(lp);
  } else {
    if (FAM == FAM_GAUSS_PREC) {
      QSB_FOR_E { int i = e * G + lane; x[e] = i < d ? gaussian_sample(P.key, ch.gc, QSB_ITER_INIT, (uint32_t)i, 0.0, 1.0) : 0.0; }
    } else if (FAM == FAM_EIGHT_SCHOOLS) {
      double mu = gaussian_sample(P.key, ch.gc, QSB_ITER_INIT, 0u, 0.0, 5.0);
      double lt = gaussian_sample(P.key, ch.gc, QSB_ITER_INIT, 1u, 0.0, 1.0);
      double tau = exp(lt);
      QSB_FOR_E {
        int i = e * G + lane;
        x[e] = 0.0;
        if (i == 0) x[e] = mu;
        else if (i == 1) x[e] = lt;
        else if (i < d) x[e] = gaussian_sample(P.key, ch.gc, QSB_ITER_INIT, (uint32_t)i, mu, tau);
      }
    } else {
      double v = gaussian_sample(P.key, ch.gc, QSB_ITER_INIT, 0u, 0.0, 3.0);
      double s = exp(v / 2.0);
      QSB_FOR_E {
        int i = e * G + lane;
        x[e] = 0.0;
        if (i == 0) x[e] = v;
        else if (i < d) x[e] = gaussian_sample(P.key, ch.gc, QSB_ITER_INIT, (uint32_t)i, 0.0, s);
      }
    }
    double g[NPL], lpf;
    model_eval<FAM, G, NPL, false>(P, lane, x, g, lpf);
    lp0 = lpf;
  }
  ch.store(P.x, x);
  if (ch.leader()) P.lp[grp] = lp0;
}

// lp of an explicitly set state (qsb_sampler_set_state)
template <int FAM, int G, int NPL>
__global__ void __launch_bounds__(G == 1 ? 128 : 256) rescore_kernel(const __grid_constant__ EngineParams P) {
  const uint32_t grp = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G);
  if (grp >= P.C) return;
  Chain<FAM, G, NPL> ch(P, grp);
  double x[NPL], g[NPL], lpf;
  ch.load(P.x, x);
  model_eval<FAM, G, NPL, false>(P, ch.lane, x, g, lpf);
  if (ch.leader()) P.lp[grp] = lpf;
}

// test hook: lp_dual, gradient and lp_float at the states in P.x; gradient to P.g, lp_dual to P.lp, lp_float to P.eps
template <int FAM, int G, int NPL>
__global__ void __launch_bounds__(G == 1 ? 128 : 256) logp_grad_kernel(const __grid_constant__ EngineParams P) {
  const uint32_t grp = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G);
  if (grp >= P.C) return;
  Chain<FAM, G, NPL> ch(P, grp);
  double x[NPL], g[NPL], lpf;
  ch.load(P.x, x);
  double lpd = model_eval<FAM, G, NPL, true>(P, ch.lane, x, g, lpf);
  ch.store(P.g, g);
  if (ch.leader()) { P.lp[grp] = lpd; P.eps[grp] = lpf; }
}

// test hook (qsb_debug_leapfrog): exactly one HMCKernel:step (hmc.t:307-323) from an injected (x, p, eps[, g]) with the
// device functions the transition kernels use -- Chain::leapfrog for the register form, GradScalars + kick for the
// scalar-gradient warp form.  `last` selects the variant a trajectory runs on its last step (log-density evaluated with
// the reference's expressions) or on its interior steps (gradient only).  Arrays are laid out like P.x.
struct DebugLeap {
  const double* p; const double* g; const double* eps;   // g == nullptr: gradient recomputed at x
  double *x_out, *p_out, *g_out, *lp_out;
  int last;
};
template <int FAM, int G, int NPL>
__global__ void __launch_bounds__(G == 1 ? 128 : 256) debug_leap_kernel(const __grid_constant__ EngineParams P, const DebugLeap D) {
  const uint32_t grp = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G);
  if (FAM == FAM_GAUSS_PREC) {
    extern __shared__ double qsb_zbuf[];
    double* sS = qsb_zbuf + (size_t)Chain<FAM, G, NPL>::NB * blockDim.x;
    for (int t = threadIdx.x; t < G * NPL * G * NPL; t += blockDim.x) sS[t] = P.Sc[t & 255];
    __syncthreads();
  }
  if (grp >= P.C) return;
  Chain<FAM, G, NPL> ch(P, grp);
  const int lane = ch.lane;
  double x[NPL], p[NPL], g[NPL];
  ch.load(P.x, x); ch.load(D.p, p);
  const double eps = D.eps[grp];
  double lp = 0.0;
  if constexpr (Chain<FAM, G, NPL>::SCALAR_GRAD) {
    const double he = 0.5 * eps;
    GradScalars<FAM> gs;
    gs.template reduce<G, NPL>(P, lane, x);
    ch.kick(he, gs, x, p, false, 1.0);
    QSB_FOR_E x[e] = x[e] + eps * p[e] * 1.0;
    gs.template reduce<G, NPL>(P, lane, x);
    ch.kick(he, gs, x, p, false, 1.0);
    QSB_FOR_E g[e] = gs.elem(P, e * G + lane, x[e]);
    if (D.last) lp = ch.lp_only(x);
  } else {
    if (D.g) ch.load(D.g, g);
    else { double lpf; model_eval<FAM, G, NPL, true, true, (FAM == FAM_GAUSS_PREC && QSB_GP_SMEM)>(P, lane, x, g, lpf); }
    if (D.last) lp = ch.template leapfrog<true>(eps, x, g, p);
    else ch.template leapfrog<false>(eps, x, g, p);
  }
  ch.store(D.x_out, x); ch.store(D.p_out, p); ch.store(D.g_out, g);
  if (ch.leader()) D.lp_out[grp] = lp;
}

// host-side launcher table entry
struct EngineVariant {
  int family, G, npl;
  void (*advance)(const EngineParams&, void* stream);
  void (*advance_nohmc)(const EngineParams&, void* stream);
  void (*advance_lex)(const EngineParams&, void* stream);
  void (*advance_drift)(const EngineParams&, void* stream);
  void (*init)(const EngineParams&, void* stream);
  void (*rescore)(const EngineParams&, void* stream);
  void (*logp_grad)(const EngineParams&, void* stream);
  void (*debug_leap)(const EngineParams&, const DebugLeap&, void* stream);
  // warp-specialised forms (G == 32 only, otherwise nullptr): same semantics as advance / advance_nohmc / advance_drift
  void (*advance_ws)(const EngineParams&, void* stream);
  void (*advance_nohmc_ws)(const EngineParams&, void* stream);
  void (*advance_drift_ws)(const EngineParams&, void* stream);
};

}  // namespace qsb
