// C-ABI of libqsb200.so (include/qsb200.h): host-side sampler objects over the chain-parallel
// engine (engine.cuh) and the GLM tensor-core path (glm_hmc.cu).  Plain CUDA runtime; no torch.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/qsb200.h"
#include "engine.cuh"
#include "glm_hmc.cuh"

namespace qsb {
extern const EngineVariant g_variants_f1[]; extern const int g_nvariants_f1;
extern const EngineVariant g_variants_f1w[]; extern const int g_nvariants_f1w;
extern const EngineVariant g_variants_f2[]; extern const int g_nvariants_f2;
extern const EngineVariant g_variants_f4[]; extern const int g_nvariants_f4;
extern const EngineVariant g_variants_f5[]; extern const int g_nvariants_f5;
}  // namespace qsb

using namespace qsb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(100 + (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct DevArena {
  std::vector<void*> ptrs;
  size_t bytes = 0;
  template <class T> cudaError_t alloc(T** p, size_t n, bool zero = true) {
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(*p);
    bytes += n * sizeof(T);
    if (zero) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
  }
  void release() { for (void* p : ptrs) cudaFree(p); ptrs.clear(); }
};

struct qsb_model {
  qsb_model_desc desc;
  int device = 0;
  int retdim = 0;
  std::vector<double> A;      // GAUSS_PREC host copy
  ModelDev dev;
  DevArena arena;
  GlmModel glm;               // LOGISTIC
  double* sigma_dev = nullptr; // EIGHT_SCHOOLS
};

struct qsb_sampler {
  qsb_model* m = nullptr;
  qsb_run_config cfg;
  std::vector<qsb_kernel_spec> specs;
  EngineParams P;
  const EngineVariant* var = nullptr;
  GlmSampler* glm = nullptr;
  DevArena arena;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint64_t iter = 0, burnin = 0, lag = 1, max_samples = 0, numiters = 0;
  uint64_t iter_base = 0;   // iterations of earlier runs on this chain state: the Philox iteration word keeps counting across qsb_sampler_begin
  uint64_t launches = 0;
  double seconds = 0.0;
  bool inited = false;
  bool advanced = false;    // ev0 / ev1 have been recorded at least once
  uint64_t rec_iters = 0;
  double t_accum = 0.0; uint64_t t_launches = 0;   // QSB_FLAG_TIME_KERNEL
  std::vector<cudaEvent_t> tev; size_t tev_used = 0;   // event pairs around the timed launches, folded without stalling the stream
  double* stage = nullptr; size_t stage_bytes = 0;  // device staging of fetch_samples (kept between calls)
  double* temps_dev = nullptr;                      // AnnealingKernel schedule
  // qsb_sampler_fetch_samples_async: two staging buffers alternate; the gather runs on the compute stream, the D2H copy
  // on a copy stream, so the copy of one step's draws overlaps the next step's transition
  cudaStream_t copy_stream = nullptr;
  double* astage[2] = {nullptr, nullptr}; size_t astage_bytes[2] = {0, 0};
  cudaEvent_t gathered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
  int aflip = 0;
};

// ---------------------------------------------------------------------------------------------
__global__ void precompute_obs_kernel(const double* sigma, double* cy, double* s2, double* is2, int n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) { double s = sigma[j]; cy[j] = 1.8378770664093453 + 2.0 * log(s); s2[j] = s * s; is2[j] = 1.0 / (s * s); }
}

// device sample layout -> AoS Sample elements { value[retdim], logprob, loglikelihood } (infer.t:13-18)
__global__ void gather_samples_kernel(const double* samples, const double* samp_lp, int G, uint32_t SC, int retdim,
                                      uint32_t c0, uint32_t nc, uint64_t s0, uint64_t ns, double* out, size_t stride_d, int with_lp) {
  size_t total = (size_t)nc * ns * (retdim + (with_lp ? 2 : 0));
  int w = retdim + (with_lp ? 2 : 0);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int j = (int)(t % w);
    size_t r = t / w;
    uint64_t k = s0 + r % ns;
    uint32_t c = c0 + (uint32_t)(r / ns);
    double v;
    if (j < retdim) v = samples[G == 1 ? ((size_t)k * retdim + j) * SC + c : ((size_t)k * SC + c) * retdim + j];
    else if (j == retdim) v = samp_lp[(size_t)k * SC + c];
    else v = 0.0;   // loglikelihood = 0 (mcmc.t:70)
    out[r * stride_d + j] = v;
  }
}

// per-(chain, dim) statistics of kept draws -> per-dim sums (see qsb_sampler_diagnostics in qsb200.h)
__global__ void diagnostics_kernel(const double* samples, int G, uint32_t SC, int retdim, uint64_t n, int maxlag, double* out) {
  size_t total = (size_t)SC * retdim;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t c; int j;
  if (G == 1) { c = (uint32_t)(t % SC); j = (int)(t / SC); } else { j = (int)(t % retdim); c = (uint32_t)(t / retdim); }
  auto at = [&](uint64_t k) { return samples[G == 1 ? ((size_t)k * retdim + j) * SC + c : ((size_t)k * SC + c) * retdim + j]; };
  const uint64_t h = n / 2;
  double* o = out + (size_t)j * 12;
  double sum_all = 0.0, sq_all = 0.0;
  for (int half = 0; half < 2; ++half) {
    if (h < 2) break;
    uint64_t b = half * h;
    double s = 0.0;
    for (uint64_t k = b; k < b + h; ++k) s += at(k);
    double mean = s / (double)h, m2 = 0.0;
    for (uint64_t k = b; k < b + h; ++k) { double dlt = at(k) - mean; m2 += dlt * dlt; }
    atomicAdd(&o[2], mean); atomicAdd(&o[3], mean * mean); atomicAdd(&o[4], m2 / (double)(h - 1));
  }
  for (uint64_t k = 0; k < n; ++k) { double v = at(k); sum_all += v; sq_all += v * v; }
  atomicAdd(&o[5], sum_all); atomicAdd(&o[6], sq_all);
  {
    double mean = sum_all / (double)n, m2 = 0.0;
    for (uint64_t k = 0; k < n; ++k) { double dlt = at(k) - mean; m2 += dlt * dlt; }
    atomicAdd(&o[9], mean); atomicAdd(&o[10], mean * mean); atomicAdd(&o[11], m2 / (double)(n - 1));
  }
  if (maxlag > 0 && n > 1) {
    double mean = sum_all / (double)n;
    double* ac = out + (size_t)retdim * 12 + (size_t)j * (maxlag + 1);
    for (int lagv = 0; lagv <= maxlag && (uint64_t)lagv < n; ++lagv) {
      double s = 0.0;
      for (uint64_t k = 0; k + lagv < n; ++k) s += (at(k) - mean) * (at(k + lagv) - mean);
      atomicAdd(&ac[lagv], s);
    }
  }
}
// The same per-dimension sums from the running statistics of QSB_FLAG_MOMENTS (moments.cuh) instead of stored draws: the two
// halves are the kept-sample ranges [0, half) and [half, n); a chain's moments are their Chan combination.  Behind the 12 sums
// per dimension: [0] sum over chains of the sample variance of the completed batch means, [1] the batch size, [2] the number
// of completed batches (per chain): asymptotic variance ~ batch * var(batch means), ESS = chains * n * s^2 / that.
__global__ void diagnostics_moments_kernel(const double* mom, int layout1, uint32_t SC, int retdim, uint64_t n, uint64_t half, uint32_t batch, double* out) {
  size_t total = (size_t)SC * retdim;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t c; int j;
  if (layout1) { c = (uint32_t)(t % SC); j = (int)(t / SC); } else { j = (int)(t % retdim); c = (uint32_t)(t / retdim); }
  const size_t plane = (size_t)SC * retdim;
  const double* b = mom + (layout1 ? (size_t)j * SC + c : (size_t)c * retdim + j);
  const double n0 = (double)(n < half ? n : half), n1 = (double)(n > half ? n - half : 0);
  const double m0 = b[0], q0 = b[plane], m1 = b[2 * plane], q1 = b[3 * plane];
  double* o = out + (size_t)j * 12;
  if (n0 >= 2.0 && n1 >= 2.0) {
    atomicAdd(&o[2], m0 + m1); atomicAdd(&o[3], m0 * m0 + m1 * m1); atomicAdd(&o[4], q0 / (n0 - 1.0) + q1 / (n1 - 1.0));
  }
  const double nn = n0 + n1;
  const double mean = (n0 * m0 + n1 * m1) / nn;
  const double M2 = q0 + q1 + (n1 > 0.0 ? (m1 - m0) * (m1 - m0) * n0 * n1 / nn : 0.0);
  atomicAdd(&o[5], nn * mean); atomicAdd(&o[6], M2 + nn * mean * mean);
  atomicAdd(&o[9], mean); atomicAdd(&o[10], mean * mean); atomicAdd(&o[11], M2 / (nn - 1.0));
  const uint64_t nb = n / batch;
  if (nb >= 2) atomicAdd(&out[(size_t)retdim * 12 + (size_t)j * 3], b[6 * plane] / (double)(nb - 1));
  if (c == 0) { out[(size_t)retdim * 12 + (size_t)j * 3 + 1] = (double)batch; out[(size_t)retdim * 12 + (size_t)j * 3 + 2] = (double)nb; }
}
__global__ void diagnostics_header_kernel(double* out, int retdim, double mh, double nh, double M, double N) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < retdim) { double* o = out + (size_t)j * 12; o[0] = mh; o[1] = nh; o[7] = M; o[8] = N; }
}

// ---- queries (infer.t:146-264).  View of kept samples: value(k, c, j), logprob(k, c), loglikelihood(k, c)
struct SampleView {
  const double* samples; const double* samp_lp; const double* samp_ll;   // samp_ll == nullptr: loglikelihood = 0 (mcmc.t:70)
  int layout;        // 1: [k][j][c]   32: [k][c][j]   0: AoS host vector staged on the device, element stride `stride` doubles
  uint32_t SC; int retdim; uint64_t n; size_t stride;
  __device__ __forceinline__ double value(uint64_t k, uint32_t c, int j) const {
    if (layout == 1) return samples[((size_t)k * retdim + j) * SC + c];
    if (layout == 32) return samples[((size_t)k * SC + c) * retdim + j];
    return samples[((size_t)c * n + k) * stride + j];
  }
  __device__ __forceinline__ double logprob(uint64_t k, uint32_t c) const {
    return layout ? samp_lp[(size_t)k * SC + c] : samples[((size_t)c * n + k) * stride + retdim];
  }
  __device__ __forceinline__ double loglik(uint64_t k, uint32_t c) const {
    return layout ? 0.0 : samples[((size_t)c * n + k) * stride + retdim + 1];
  }
};

// Expectation: thread per (chain, dim), samples summed in the reference's order (infer.t:154-165)
__global__ void expectation_mean_kernel(SampleView V, uint32_t c0, uint32_t nc, double* mean) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= (size_t)nc * V.retdim) return;
  const uint32_t c = c0 + (uint32_t)(t / V.retdim); const int j = (int)(t % V.retdim);
  double m = V.value(0, c, j), totalw = 0.0;
  for (uint64_t i = 1; i < V.n; ++i) {
    const double w = exp(V.loglik(i, c));
    totalw = totalw + w;
    m = m + w * V.value(i, c, j);
  }
  mean[t] = m / totalw;
}
// variance: warp per chain; per sample the inner product diff.diff over dims (lanes), accumulated in sample order (infer.t:169-176)
__global__ void expectation_var_kernel(SampleView V, uint32_t c0, uint32_t nc, const double* mean, double* var) {
  const uint32_t w = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= nc) return;
  const uint32_t c = c0 + w;
  double v = 0.0;
  for (uint64_t i = 0; i < V.n; ++i) {
    double dd = 0.0;
    for (int j = lane; j < V.retdim; j += 32) { const double diff = V.value(i, c, j) - mean[(size_t)w * V.retdim + j]; dd = dd + diff * diff; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
    v = v + dd;
  }
  if (lane == 0) var[w] = v / (double)V.n;
}
// MAP: thread per chain, first strict maximum (infer.t:195-201)
__global__ void map_kernel(SampleView V, uint32_t c0, uint32_t nc, double* value, double* logprob) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nc) return;
  const uint32_t c = c0 + w;
  uint64_t best = 0; double bl = V.logprob(0, c);
  for (uint64_t i = 1; i < V.n; ++i) { const double l = V.logprob(i, c); if (l > bl) { bl = l; best = i; } }
  for (int j = 0; j < V.retdim; ++j) value[(size_t)w * V.retdim + j] = V.value(best, c, j);
  if (logprob) logprob[w] = bl;
}
// Autocorrelation: block per chain, thread per lag t; the sum over i and the inner product over dims are sequential
// like the reference's (infer.t:221-241).  mean: [nc][retdim] (per chain) or [retdim] (shared, per_chain = 0)
__global__ void autocorrelation_kernel(SampleView V, uint32_t c0, const double* mean, const double* var, double shared_var, int per_chain, double* ac) {
  const uint32_t w = blockIdx.x;
  const uint32_t c = c0 + w;
  const double* m = mean + (per_chain ? (size_t)w * V.retdim : 0);
  const double v = per_chain ? var[w] : shared_var;
  for (uint64_t t = threadIdx.x; t < V.n; t += blockDim.x) {
    double act = 0.0;
    const uint64_t n = V.n - t;
    for (uint64_t i = 0; i < n; ++i) {
      double ip = 0.0;
      for (int j = 0; j < V.retdim; ++j) ip = ip + (V.value(i, c, j) - m[j]) * (V.value(i + t, c, j) - m[j]);
      act = act + ip;
    }
    if (n > 0) act = act / ((double)n * v);
    ac[(size_t)w * V.n + t] = act;
  }
}

__global__ void uniforms_kernel(RngKey key, uint32_t chain, uint32_t iter, uint32_t slot, int n, double* out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    SubStream s(key, chain, iter, slot);
    for (int i = 0; i < n; ++i) out[i] = s.random();
  }
}
__global__ void gaussians_kernel(RngKey key, uint32_t chain, uint32_t iter, uint32_t slot0, int n, double mu, double sigma, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = gaussian_sample(key, chain, iter, slot0 + (uint32_t)i, mu, sigma);
}
__global__ void logprob_kernel(int which, const double* a, int n, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (which == 8) {   // dirichlet (distrib.t:352-362) over the n tuples (theta_k, alpha_k, -): one value, in out[0]; others NaN
    double r = nan("");
    if (i == 0) {
      double sum = 0.0;
      for (int k = 0; k < n; ++k) sum = sum + a[3 * k + 1];
      r = log_gamma5(sum);
      for (int k = 0; k < n; ++k) { const double al = a[3 * k + 1]; r = r + (al - 1.0) * log(a[3 * k]); r = r - log_gamma5(al); }
    }
    out[i] = r;
    return;
  }
  const double* v = a + 3 * (size_t)i;
  double r;
  switch (which) {
    case 0: r = bernoulli_lp(v[0] != 0.0, v[1]); break;
    case 1: r = uniform_lp(v[0], v[1], v[2]); break;
    case 2: r = gaussian_lp(v[0], v[1], v[2]); break;
    case 3: r = gamma_lp(v[0], v[1], v[2]); break;
    case 4: r = beta_lp(v[0], v[1], v[2]); break;
    case 9: r = log_gamma5(v[0]); break;
    default: r = nan(""); break;
  }
  out[i] = r;
}

// ---------------------------------------------------------------------------------------------
static const EngineVariant* pick_variant(int family, int d, int mapping, std::string& why) {
  const EngineVariant* tab = nullptr; int n = 0;
  switch (family) {
    case QSB_FAM_INDEP: tab = g_variants_f1; n = g_nvariants_f1; break;
    case QSB_FAM_GAUSS_PREC: tab = g_variants_f2; n = g_nvariants_f2; break;
    case QSB_FAM_EIGHT_SCHOOLS: tab = g_variants_f4; n = g_nvariants_f4; break;
    case QSB_FAM_FUNNEL: tab = g_variants_f5; n = g_nvariants_f5; break;
    default: why = "family has no chain-parallel engine variant"; return nullptr;
  }
  const EngineVariant* best = nullptr;
  auto consider = [&](const EngineVariant* t, int cnt) {
    for (int i = 0; i < cnt; ++i) {
      const EngineVariant& v = t[i];
      if (mapping != 0 && v.G != mapping) continue;
      if ((long)v.G * v.npl < d) continue;
      if (!best || (long)v.G * v.npl < (long)best->G * best->npl) best = &v;
    }
  };
  consider(tab, n);
  if (family == QSB_FAM_INDEP) consider(g_variants_f1w, g_nvariants_f1w);
  if (!best) why = "no engine variant for dim=" + std::to_string(d) + " (family " + std::to_string(family) + ", mapping " + std::to_string(mapping) + ")";
  return best;
}

template <class T> static cudaError_t fill(T* p, size_t n, T v, cudaStream_t st) {
  std::vector<T> h(n, v);
  return cudaMemcpyAsync(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, st) == cudaSuccess ? cudaStreamSynchronize(st) : cudaGetLastError();
}

extern "C" {

int qsb_version(void) { return QSB_VERSION; }
const char* qsb_last_error(void) { return g_err.c_str(); }
#ifndef QSB_SOURCE_HASH
#define QSB_SOURCE_HASH "unknown"
#endif
const char* qsb_source_hash(void) { return QSB_SOURCE_HASH; }
// internal (multi.cu): not part of include/qsb200.h
void qsb_set_error_(const char* msg) { g_err = msg ? msg : ""; }
int qsb_sampler_retdim_(qsb_sampler* s) { return s ? s->m->retdim : -1; }
void* qsb_sampler_stream_(qsb_sampler* s) { return s ? (void*)s->stream : nullptr; }

int qsb_model_create(const qsb_model_desc* d, int32_t device, qsb_model** out) {
  if (!d || !out) return fail(1, "qsb_model_create: null argument");
  *out = nullptr;
  if (d->dim <= 0) return fail(2, "qsb_model_create: dim must be positive");
  if (device >= 0) CU(cudaSetDevice(device)); else CU(cudaGetDevice(&device));
  qsb_model* m = new qsb_model();
  m->desc = *d; m->device = device;
  memset(&m->dev, 0, sizeof(m->dev));
  m->dev.family = d->family; m->dev.dim = d->dim; m->dev.nobs = d->nobs; m->dev.ret_mode = d->ret_mode;
  m->dev.ret_div = d->ret_div; m->dev.prior_sigma = d->prior_sigma;
  m->retdim = d->dim;
  auto bail = [&](int code, const std::string& msg) { m->arena.release(); delete m; return fail(code, msg); };
  switch (d->family) {
    case QSB_FAM_INDEP: {
      if (!d->erp_kind || !d->erp_p0 || !d->erp_p1) return bail(3, "INDEP: erp_kind/erp_p0/erp_p1 required");
      for (int i = 0; i < d->dim; ++i) if (d->erp_kind[i] < 0 || d->erp_kind[i] > 5) return bail(3, "INDEP: unsupported ERP kind (continuous gaussian/gamma/beta/uniform/dirichlet/neggamma only)");
      // vector choices: K consecutive components of kind DIRICHLET with erp_p1 = 0, 1, .., K-1
      std::vector<int> len(d->dim, 1), slot(d->dim, 0);
      std::vector<double> asum(d->dim, 0.0);
      int nerp = 0;
      for (int i = 0; i < d->dim;) {
        if (d->erp_kind[i] != QSB_ERP_DIRICHLET) { slot[i] = nerp++; ++i; continue; }
        if (d->erp_p1[i] != 0.0) return bail(3, "INDEP: a dirichlet choice must start with erp_p1 = 0 (component index inside the choice)");
        int j = i + 1;
        while (j < d->dim && d->erp_kind[j] == QSB_ERP_DIRICHLET && d->erp_p1[j] == (double)(j - i)) ++j;
        const int K = j - i;
        if (K < 2 || K > 32) return bail(3, "INDEP: dirichlet choices of 2..32 components are supported");
        for (int q = i; q < j; ++q) {
          if (!(d->erp_p0[q] > 0.0)) return bail(3, "INDEP: dirichlet parameters must be positive");
          len[q] = K; slot[q] = nerp;
          for (int r = q + 1; r < j; ++r) asum[q] += d->erp_p0[r];
        }
        ++nerp; m->dev.has_vector = 1;
        i = j;
      }
      int *k, *dl, *ds; double *p0, *p1, *da;
      if (m->arena.alloc(&k, d->dim) || m->arena.alloc(&p0, d->dim) || m->arena.alloc(&p1, d->dim) ||
          m->arena.alloc(&dl, d->dim) || m->arena.alloc(&ds, d->dim) || m->arena.alloc(&da, d->dim)) return bail(4, "cudaMalloc failed");
      cudaMemcpy(dl, len.data(), sizeof(int) * d->dim, cudaMemcpyHostToDevice);
      cudaMemcpy(ds, slot.data(), sizeof(int) * d->dim, cudaMemcpyHostToDevice);
      cudaMemcpy(da, asum.data(), sizeof(double) * d->dim, cudaMemcpyHostToDevice);
      m->dev.erp_len = dl; m->dev.erp_slot = ds; m->dev.erp_asum = da;
      {
        // DriftKernel lexical scale sharing (drift.t:208-235): adapter of a component = (group of its choice's call site in
        // order of first appearance, index inside the choice), numbered in order of first appearance
        std::vector<int> amap(d->dim, 0);
        std::vector<std::pair<int, int>> seen;   // (lexid, k) -> adapter id = position
        for (int i = 0; i < d->dim; ++i) {
          const int lex = d->erp_lexid ? d->erp_lexid[i] : slot[i];
          const int kk = d->erp_kind[i] == QSB_ERP_DIRICHLET ? (int)d->erp_p1[i] : 0;
          int a = -1;
          for (size_t q = 0; q < seen.size(); ++q) if (seen[q].first == lex && seen[q].second == kk) { a = (int)q; break; }
          if (a < 0) { a = (int)seen.size(); seen.emplace_back(lex, kk); }
          amap[i] = a;
        }
        int* dm;
        if (m->arena.alloc(&dm, d->dim)) return bail(4, "cudaMalloc failed");
        cudaMemcpy(dm, amap.data(), sizeof(int) * d->dim, cudaMemcpyHostToDevice);
        m->dev.lex_map = dm; m->dev.n_adapters = (int)seen.size();
        // TraceMHKernel picks among the random choices (a vector choice is one): first component of each
        std::vector<int> first;
        for (int i = 0; i < d->dim; ++i) if (i == 0 || slot[i] != slot[i - 1]) first.push_back(i);
        int* df;
        if (m->arena.alloc(&df, first.size())) return bail(4, "cudaMalloc failed");
        cudaMemcpy(df, first.data(), sizeof(int) * first.size(), cudaMemcpyHostToDevice);
        m->dev.erp_first = df; m->dev.n_choices = (int)first.size();
      }
      cudaMemcpy(k, d->erp_kind, sizeof(int) * d->dim, cudaMemcpyHostToDevice);
      cudaMemcpy(p0, d->erp_p0, sizeof(double) * d->dim, cudaMemcpyHostToDevice);
      cudaMemcpy(p1, d->erp_p1, sizeof(double) * d->dim, cudaMemcpyHostToDevice);
      m->dev.erp_kind = k; m->dev.p0 = p0; m->dev.p1 = p1;
      if (d->erp_fsigma) {
        if (!d->erp_fmu) return bail(3, "INDEP: erp_fmu is required with erp_fsigma");
        for (int i = 0; i < d->dim; ++i) if (d->erp_fsigma[i] > 0.0 && d->erp_kind[i] == QSB_ERP_DIRICHLET) return bail(3, "INDEP: factors on the components of a dirichlet choice are not supported");
        double *fm, *fs;
        if (m->arena.alloc(&fm, d->dim) || m->arena.alloc(&fs, d->dim)) return bail(4, "cudaMalloc failed");
        cudaMemcpy(fm, d->erp_fmu, sizeof(double) * d->dim, cudaMemcpyHostToDevice);
        cudaMemcpy(fs, d->erp_fsigma, sizeof(double) * d->dim, cudaMemcpyHostToDevice);
        m->dev.fmu = fm; m->dev.fsigma = fs;
      }
      if (d->erp_cond_lo || d->erp_cond_hi) {
        if (!d->erp_cond_lo || !d->erp_cond_hi) return bail(3, "INDEP: erp_cond_lo and erp_cond_hi come together");
        bool any = false;
        for (int i = 0; i < d->dim; ++i) {
          const bool has = d->erp_cond_lo[i] > -INFINITY || d->erp_cond_hi[i] < INFINITY;
          if (has && d->erp_kind[i] == QSB_ERP_DIRICHLET) return bail(3, "INDEP: conditions on the components of a dirichlet choice are not supported");
          if (has && !(d->erp_cond_lo[i] < d->erp_cond_hi[i])) return bail(3, "INDEP: an empty condition interval can never be satisfied");
          any |= has;
        }
        if (any) {
          double *cl, *chh;
          if (m->arena.alloc(&cl, d->dim) || m->arena.alloc(&chh, d->dim)) return bail(4, "cudaMalloc failed");
          cudaMemcpy(cl, d->erp_cond_lo, sizeof(double) * d->dim, cudaMemcpyHostToDevice);
          cudaMemcpy(chh, d->erp_cond_hi, sizeof(double) * d->dim, cudaMemcpyHostToDevice);
          m->dev.cond_lo = cl; m->dev.cond_hi = chh;
        }
      }
      if (d->erp_pref) {
        if (!d->erp_pcoef) return bail(3, "INDEP: erp_pcoef is required with erp_pref");
        std::vector<int> tf(2 * (size_t)d->dim, 0);
        bool any = false;
        for (int i = 0; i < d->dim; ++i)
          for (int q = 0; q < 2; ++q) {
            const int r = d->erp_pref[2 * i + q];
            if (d->erp_ptf) tf[2 * i + q] = d->erp_ptf[2 * i + q] ? 1 : 0;
            if (r < 0) continue;
            any = true;
            if (r >= i) return bail(3, "INDEP: a latent parameter must refer to an earlier choice");
            if (d->erp_kind[r] == QSB_ERP_DIRICHLET) return bail(3, "INDEP: a latent parameter must refer to a scalar choice");
            if (d->erp_kind[i] == QSB_ERP_UNIFORM || d->erp_kind[i] == QSB_ERP_DIRICHLET)
              return bail(3, "INDEP: latent parameters are supported on gaussian, gamma and beta choices");
          }
        // (dim <= 32: thread per chain; beyond that the warp-per-chain mapping with the sequential evaluator on scratch memory)
        if (any) {
          int *lr, *lt; double* lb;
          if (m->arena.alloc(&lr, 2 * (size_t)d->dim) || m->arena.alloc(&lt, 2 * (size_t)d->dim) || m->arena.alloc(&lb, 2 * (size_t)d->dim)) return bail(4, "cudaMalloc failed");
          cudaMemcpy(lr, d->erp_pref, sizeof(int) * 2 * d->dim, cudaMemcpyHostToDevice);
          cudaMemcpy(lt, tf.data(), sizeof(int) * 2 * d->dim, cudaMemcpyHostToDevice);
          cudaMemcpy(lb, d->erp_pcoef, sizeof(double) * 2 * d->dim, cudaMemcpyHostToDevice);
          m->dev.lat_ref = lr; m->dev.lat_tf = lt; m->dev.lat_b = lb;
          m->dev.has_vector = 1;   // same consequence as a vector choice: thread-per-chain mapping only
        }
      }
      if (d->ret_mode == 0) m->retdim = 1;
      break;
    }
    case QSB_FAM_GAUSS_PREC: {
      if (!d->A) return bail(3, "GAUSS_PREC: A required");
      if (d->dim > 16) return bail(3, "GAUSS_PREC: dim <= 16 supported (thread-per-chain register path)");
      m->A.assign(d->A, d->A + (size_t)d->dim * d->dim);
      m->dev.n_adapters = 1; m->dev.n_choices = d->dim;
      break;
    }
    case QSB_FAM_LOGISTIC: {
      if (!d->X || !d->ybin || d->nobs <= 0) return bail(3, "LOGISTIC: X, ybin, nobs required");
      std::string why;
      if (glm_model_create(&m->glm, d->X, d->ybin, d->nobs, d->dim, d->prior_sigma, why)) return bail(5, "LOGISTIC: " + why);
      break;
    }
    case QSB_FAM_EIGHT_SCHOOLS: {
      if (!d->y || !d->sigma || d->nobs <= 0 || d->dim != d->nobs + 2) return bail(3, "EIGHT_SCHOOLS: y, sigma required and dim == nobs + 2");
      double *y, *sg, *cy, *s2, *is2;
      if (m->arena.alloc(&y, d->nobs) || m->arena.alloc(&sg, d->nobs) || m->arena.alloc(&cy, d->nobs) || m->arena.alloc(&s2, d->nobs) ||
          m->arena.alloc(&is2, d->nobs)) return bail(4, "cudaMalloc failed");
      cudaMemcpy(y, d->y, sizeof(double) * d->nobs, cudaMemcpyHostToDevice);
      cudaMemcpy(sg, d->sigma, sizeof(double) * d->nobs, cudaMemcpyHostToDevice);
      precompute_obs_kernel<<<(d->nobs + 127) / 128, 128>>>(sg, cy, s2, is2, d->nobs);
      if (cudaDeviceSynchronize() != cudaSuccess) return bail(6, "precompute_obs_kernel failed");
      m->dev.y = y; m->dev.cy = cy; m->dev.s2 = s2; m->dev.is2 = is2; m->sigma_dev = sg;
      m->dev.n_adapters = 3; m->dev.n_choices = d->dim;
      break;
    }
    case QSB_FAM_FUNNEL: m->dev.n_adapters = 2; m->dev.n_choices = d->dim; break;
    default: return bail(2, "qsb_model_create: unknown family");
  }
  *out = m;
  return 0;
}

int qsb_model_update_data(qsb_model* m, const qsb_model_desc* d, void* stream) {
  if (!m || !d) return fail(1, "null argument");
  if (d->family != m->desc.family || d->dim != m->desc.dim || d->nobs != m->desc.nobs) return fail(2, "qsb_model_update_data: shape mismatch");
  CU(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (d->family == QSB_FAM_LOGISTIC) {
    if (!d->X || !d->ybin) return fail(3, "LOGISTIC: X, ybin required");
    std::string why;
    if (glm_model_update(&m->glm, d->X, d->ybin, stream, why)) return fail(5, "LOGISTIC: " + why);
    return 0;
  }
  if (d->family == QSB_FAM_EIGHT_SCHOOLS) {
    if (!d->y || !d->sigma) return fail(3, "EIGHT_SCHOOLS: y, sigma required");
    CU(cudaMemcpyAsync((void*)m->dev.y, d->y, sizeof(double) * d->nobs, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync((void*)m->sigma_dev, d->sigma, sizeof(double) * d->nobs, cudaMemcpyHostToDevice, st));
    precompute_obs_kernel<<<(d->nobs + 127) / 128, 128, 0, st>>>(m->sigma_dev, (double*)m->dev.cy, (double*)m->dev.s2, (double*)m->dev.is2, d->nobs);
    CU(cudaGetLastError());
    return 0;
  }
  return fail(2, "qsb_model_update_data: family has no observation data");
}

int qsb_model_destroy(qsb_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  glm_model_destroy(&m->glm);
  m->arena.release();
  delete m;
  return 0;
}
int qsb_model_dim(const qsb_model* m) { return m ? m->desc.dim : -1; }
int qsb_model_retdim(const qsb_model* m) { return m ? m->retdim : -1; }

static void free_run_buffers(qsb_sampler* s) {
  EngineParams& P = s->P;
  for (double** p : {&P.samples, &P.samp_lp, &P.rec_state, &P.rec_dh, &P.rec_step, &P.rec_leap, &P.mom}) if (*p) { cudaFree(*p); *p = nullptr; }
  for (int** p : {&P.rec_accept, &P.rec_kidx}) if (*p) { cudaFree(*p); *p = nullptr; }
  s->rec_iters = 0; P.rec_iters = 0;
}

static int alloc_engine_state(qsb_sampler* s) {
  EngineParams& P = s->P;
  const size_t C = P.C, d = P.d, ns = P.nsub;
  DevArena& a = s->arena;
  bool has_hmc = false, has_drift = false, has_harm = false;
  for (auto& k : s->specs) { has_hmc |= k.kind == QSB_KERNEL_HMC; has_drift |= k.kind == QSB_KERNEL_DRIFT; has_harm |= k.kind == QSB_KERNEL_HARM; }
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  A(a.alloc(&P.x, C * d)); A(a.alloc(&P.lp, C)); A(a.alloc(&P.flags, C));
  if (P.vs_dpad > 0) A(a.alloc(&P.vscratch, ((C + 7) / 8) * 8 * (size_t)VS_ARRAYS * P.vs_dpad));   // one slice per warp of the grid (blocks of 8 warps)
  A(a.alloc(&P.eps, C));   // also the lp_float output of the logp_grad test hook
  A(a.alloc(&P.g, C * d));
  if (has_hmc) {
    A(a.alloc(&P.da_gbar, C)); A(a.alloc(&P.da_xbar, C)); A(a.alloc(&P.da_x0, C)); A(a.alloc(&P.da_lastx, C)); A(a.alloc(&P.da_gamma, C));
    A(a.alloc(&P.da_k, C));
    A(a.alloc(&P.hmc_dualT, C)); A(a.alloc(&P.hmc_gT, C));
  }
  if (has_drift && !P.dr_lexical) { A(a.alloc(&P.dr_mean, C * d)); A(a.alloc(&P.dr_m2, C * d)); A(a.alloc(&P.dr_s0, C)); A(a.alloc(&P.dr_mult, C)); A(a.alloc(&P.dr_n, C)); }
  if (has_drift && P.dr_lexical) {   // one ScaleAdapter per (call site, index inside the choice) and chain: [adapter][chain]
    const size_t na = (size_t)P.m.n_adapters;
    A(a.alloc(&P.ls_scale, na * C)); A(a.alloc(&P.ls_mean, na * C)); A(a.alloc(&P.ls_m2, na * C)); A(a.alloc(&P.ls_n, na * C));
  }
  if (has_harm) A(a.alloc(&P.ha_scale, C));
  A(a.alloc(&P.made, ns * C)); A(a.alloc(&P.acc, ns * C));
  A(a.alloc(&P.counters, 4)); A(a.alloc(&P.status, 1));
  if (e != cudaSuccess) return fail(4, std::string("cudaMalloc (chain state): ") + cudaGetErrorString(e));
  return 0;
}

static int reset_kernel_state(qsb_sampler* s) {
  EngineParams& P = s->P;
  const size_t C = P.C, d = P.d;
  cudaStream_t st = s->stream;
  CU(cudaMemsetAsync(P.flags, 0, C, st));
  CU(cudaMemsetAsync(P.made, 0, sizeof(unsigned long long) * P.nsub * C, st));
  CU(cudaMemsetAsync(P.acc, 0, sizeof(unsigned long long) * P.nsub * C, st));
  CU(cudaMemsetAsync(P.counters, 0, sizeof(unsigned long long) * 4, st));
  CU(cudaMemsetAsync(P.status, 0, sizeof(int), st));
  for (int k = 0; k < P.nsub; ++k) {
    if (P.kind[k] == K_HMC) {
      CU(fill(P.eps, C, P.stepSize0[k], st));
      CU(fill(P.hmc_dualT, C, 1.0, st)); CU(fill(P.hmc_gT, C, 1.0, st));
      CU(cudaMemsetAsync(P.da_gbar, 0, 8 * C, st)); CU(cudaMemsetAsync(P.da_xbar, 0, 8 * C, st)); CU(cudaMemsetAsync(P.da_x0, 0, 8 * C, st));
      CU(cudaMemsetAsync(P.da_lastx, 0, 8 * C, st)); CU(cudaMemsetAsync(P.da_gamma, 0, 8 * C, st)); CU(cudaMemsetAsync(P.da_k, 0, 4 * C, st));
    } else if (P.kind[k] == K_DRIFT && P.dr_lexical) {
      const size_t na = (size_t)P.m.n_adapters;
      CU(fill(P.ls_scale, na * C, P.scale0[k], st));
      CU(cudaMemsetAsync(P.ls_mean, 0, 8 * na * C, st)); CU(cudaMemsetAsync(P.ls_m2, 0, 8 * na * C, st)); CU(cudaMemsetAsync(P.ls_n, 0, 4 * na * C, st));
    } else if (P.kind[k] == K_DRIFT) {
      CU(fill(P.dr_s0, C, P.scale0[k], st)); CU(fill(P.dr_mult, C, 1.0, st));
      CU(cudaMemsetAsync(P.dr_n, 0, 4 * C, st)); CU(cudaMemsetAsync(P.dr_mean, 0, 8 * C * d, st)); CU(cudaMemsetAsync(P.dr_m2, 0, 8 * C * d, st));
    } else if (P.kind[k] == K_HARM) {
      CU(fill(P.ha_scale, C, P.scale0[k], st));
    }
  }
  return 0;
}

int qsb_sampler_create(qsb_model* m, const qsb_kernel_spec* specs, int32_t nspecs, const qsb_run_config* cfg, qsb_sampler** out) {
  if (!m || !specs || !cfg || !out) return fail(1, "qsb_sampler_create: null argument");
  *out = nullptr;
  if (nspecs < 1 || nspecs > MAX_SUB) return fail(2, "qsb_sampler_create: 1..8 kernels");
  if (cfg->numchains == 0) return fail(2, "qsb_sampler_create: numchains must be positive");
  int nh = 0, nd = 0, nr = 0, nm = 0;
  for (int i = 0; i < nspecs; ++i) {
    switch (specs[i].kind) {
      case QSB_KERNEL_HMC: ++nh; if (specs[i].numSteps < 1 || specs[i].numSteps > 0xFFFFFFFFull) return fail(2, "HMC: numSteps out of range"); break;
      case QSB_KERNEL_DRIFT: ++nd; break;
      case QSB_KERNEL_HARM: ++nr; break;
      case QSB_KERNEL_TRACEMH: ++nm; break;
      default: return fail(2, "qsb_sampler_create: unknown kernel kind");
    }
  }
  if (nh > 1 || nd > 1 || nr > 1 || nm > 1) return fail(2, "qsb_sampler_create: at most one kernel of each kind per mixture");
  CU(cudaSetDevice(m->device));
  qsb_sampler* s = new qsb_sampler();
  s->m = m; s->cfg = *cfg; s->specs.assign(specs, specs + nspecs);
  s->stream = (cudaStream_t)cfg->stream;
  if (s->cfg.sample_chains == 0 || s->cfg.sample_chains > cfg->numchains) s->cfg.sample_chains = cfg->numchains;
  auto bail = [&](int code, const std::string& msg) { s->arena.release(); if (s->glm) glm_sampler_destroy(s->glm); delete s; return fail(code, msg); };
  if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) return bail(4, "cudaEventCreate failed");
  if (m->desc.family == QSB_FAM_LOGISTIC) {
    if (nspecs != 1 || specs[0].kind == QSB_KERNEL_TRACEMH) return bail(2, "LOGISTIC: the tensor-core path runs a single kernel (HMC, Drift or HARM), not a mixture or TraceMH");
    if (specs[0].kind == QSB_KERNEL_DRIFT && specs[0].lexicalScaleSharing) return bail(2, "LOGISTIC: DriftKernel{lexicalScaleSharing} is not implemented on the tensor-core path");
    std::string why;
    s->glm = glm_sampler_create(&m->glm, specs[0].kind, specs[0].stepSize, (uint32_t)specs[0].numSteps, specs[0].scale, specs[0].doAdapt != 0,
                                cfg->seed, cfg->chain_begin, cfg->numchains, s->cfg.sample_chains, cfg->flags, (void*)s->stream, why);
    if (!s->glm) return bail(5, "LOGISTIC: " + why);
    *out = s;
    return 0;
  }
  EngineParams& P = s->P;
  memset(&P, 0, sizeof(P));
  P.m = m->dev;
  P.key = RngKey{(uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32)};
  P.chain_begin = cfg->chain_begin; P.C = cfg->numchains; P.d = m->desc.dim; P.retdim = m->retdim;
  P.nsub = nspecs;
  for (int i = 0; i < nspecs; ++i) {
    P.kind[i] = specs[i].kind; P.doAdapt[i] = specs[i].doAdapt; P.stepSize0[i] = specs[i].stepSize;
    P.numSteps[i] = (uint32_t)specs[i].numSteps; P.scale0[i] = specs[i].scale; P.weight[i] = specs[i].weight;
    if (specs[i].kind == QSB_KERNEL_DRIFT && specs[i].lexicalScaleSharing) P.dr_lexical = 1;
  }
  if (P.dr_lexical && m->dev.n_adapters > 32) return bail(2, "DriftKernel{lexicalScaleSharing}: at most 32 scale adapters (call site x index inside the choice)");
  P.sample_chains = s->cfg.sample_chains;
  P.lag = 1;
  std::string why;
  // vector (dirichlet) choices and latent parameters: thread per chain up to 32 components (every component of the chain in
  // one thread's registers), warp per chain with the general evaluator on scratch memory beyond that or when asked for
  int mapping = cfg->reserved;
  if (m->dev.has_vector) {
    if (mapping != 0 && mapping != 1 && mapping != 32) return bail(2, "dirichlet choices and latent parameters run in the thread-per-chain (<= 32 components) or the warp-per-chain mapping");
    if (mapping == 0) mapping = P.d <= 32 ? 1 : 32;
    // the per-adapter running variances of lexicalScaleSharing are updated component by component in program order by the
    // non-streaming drift kernel: one thread per chain
    if (P.dr_lexical && mapping == 32) return bail(2, "DriftKernel{lexicalScaleSharing} over dirichlet choices / latent parameters runs in the thread-per-chain mapping (<= 32 components)");
  }
  // c2-like shapes: four lanes per chain when the chains fill whole warps (8 per warp) and there are enough of them to matter
  if (mapping == 0 && m->desc.family == QSB_FAM_GAUSS_PREC && cfg->numchains % 8 == 0 && getenv("QSB_G4")) mapping = 4;   // (measured 2 x slower than thread per chain at c2's shape: not the default)
  if (mapping > 1 && mapping < 32 && cfg->numchains % (32 / mapping) != 0) return bail(2, "a mapping of " + std::to_string(mapping) + " lanes per chain needs a chain count that fills whole warps");
  s->var = pick_variant(m->desc.family, P.d, mapping, why);
  if (!s->var) return bail(2, why);
  if (m->dev.has_vector && s->var->G == 32) P.vs_dpad = 32 * s->var->npl;   // scratch of the general evaluator (alloc_engine_state)
  if (m->desc.family == QSB_FAM_GAUSS_PREC) {
    const int d = P.d, npl = s->var->npl * s->var->G;   // row stride of the staged S = padded dimension
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) P.Sc[i * npl + j] = m->A[(size_t)i * d + j] + m->A[(size_t)j * d + i];
  }
  if (int rc = alloc_engine_state(s)) { std::string e = g_err; return bail(rc, e); }
  if (int rc = reset_kernel_state(s)) { std::string e = g_err; return bail(rc, e); }
  *out = s;
  return 0;
}

int qsb_sampler_destroy(qsb_sampler* s) {
  if (!s) return 0;
  cudaSetDevice(s->m->device);
  cudaStreamSynchronize(s->stream);
  if (s->glm) glm_sampler_destroy(s->glm);
  else free_run_buffers(s);
  if (s->stage) cudaFree(s->stage);
  if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
  for (int k = 0; k < 2; ++k) { if (s->astage[k]) cudaFree(s->astage[k]); if (s->gathered[k]) cudaEventDestroy(s->gathered[k]); if (s->copied[k]) cudaEventDestroy(s->copied[k]); }
  if (s->temps_dev) cudaFree(s->temps_dev);
  for (cudaEvent_t e : s->tev) cudaEventDestroy(e);
  s->arena.release();
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
  return 0;
}

static int check_status(qsb_sampler* s) {
  int st = 0;
  CU(cudaMemcpyAsync(&st, s->P.status, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (st & ERR_NAN_SEARCH) return fail(20, "HMC step size search failed because logprob became NaN");                                  // hmc.t:361-365
  if (st & ERR_SEARCH_INF) return fail(21, "HMC step size search diverged to infinity (Is your probability distribution flat?)");        // hmc.t:381-384
  if (st & ERR_SEARCH_ZERO) return fail(22, "HMC step size search collapsed to zero (Is your probability distribution discontinuous?)"); // hmc.t:385-388
  return 0;
}

int qsb_sampler_init(qsb_sampler* s) {
  if (!s) return fail(1, "null sampler");
  CU(cudaSetDevice(s->m->device));
  if (s->glm) { std::string why; if (glm_sampler_init(s->glm, why)) return fail(5, why); s->inited = true; s->iter = 0; s->iter_base = 0; return 0; }
  if (int rc = reset_kernel_state(s)) return rc;
  s->var->init(s->P, s->stream); ++s->launches;
  CU(cudaGetLastError());
  s->inited = true; s->iter = 0; s->iter_base = 0;
  return 0;
}

int qsb_sampler_set_state(qsb_sampler* s, const double* x) {
  if (!s || !x) return fail(1, "null argument");
  CU(cudaSetDevice(s->m->device));
  if (s->glm) { std::string why; if (glm_sampler_set_state(s->glm, x, why)) return fail(5, why); s->inited = true; s->iter = 0; s->iter_base = 0; return 0; }
  EngineParams& P = s->P;
  const size_t C = P.C, d = P.d;
  std::vector<double> h(C * d);
  if (s->var->G == 1) { for (size_t c = 0; c < C; ++c) for (size_t i = 0; i < d; ++i) h[i * C + c] = x[c * d + i]; }
  else std::copy(x, x + C * d, h.begin());
  if (int rc = reset_kernel_state(s)) return rc;
  CU(cudaMemcpyAsync(P.x, h.data(), sizeof(double) * C * d, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->var->rescore(P, s->stream); ++s->launches;
  CU(cudaGetLastError());
  s->inited = true; s->iter = 0; s->iter_base = 0;
  return 0;
}

int qsb_sampler_get_state(qsb_sampler* s, double* x) {
  if (!s || !x) return fail(1, "null argument");
  CU(cudaSetDevice(s->m->device));
  if (s->glm) { std::string why; if (glm_sampler_get_state(s->glm, x, why)) return fail(5, why); return 0; }
  EngineParams& P = s->P;
  const size_t C = P.C, d = P.d;
  std::vector<double> h(C * d);
  CU(cudaMemcpyAsync(h.data(), P.x, sizeof(double) * C * d, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (s->var->G == 1) { for (size_t c = 0; c < C; ++c) for (size_t i = 0; i < d; ++i) x[c * d + i] = h[i * C + c]; }
  else std::copy(h.begin(), h.end(), x);
  return 0;
}

int qsb_sampler_begin(qsb_sampler* s, uint64_t max_samples, uint64_t burnin, uint64_t lag, uint64_t numiters) {
  if (!s) return fail(1, "null sampler");
  if (lag == 0) return fail(2, "lag must be >= 1");
  CU(cudaSetDevice(s->m->device));
  if (!s->inited) if (int rc = qsb_sampler_init(s)) return rc;
  s->iter_base += s->iter;   // a further run on the same chain state continues the iteration numbering (fresh random numbers)
  s->iter = 0; s->burnin = burnin; s->lag = lag; s->max_samples = max_samples; s->numiters = numiters;
  if (s->glm) { std::string why; if (glm_sampler_begin(s->glm, max_samples, burnin, lag, numiters, why)) return fail(5, why); return 0; }
  EngineParams& P = s->P;
  // (re)allocate the sample and record buffers of this run; the chain state persists
  free_run_buffers(s);
  P.max_samples = max_samples; P.burnin = burnin; P.lag = lag;
  if (max_samples > 0 && (s->cfg.flags & QSB_FLAG_MOMENTS)) {   // running statistics of the kept draws instead of the draws
    const size_t bytes = sizeof(double) * MOM_STATS * P.sample_chains * P.retdim;
    CU(cudaMalloc((void**)&P.mom, bytes));
    CU(cudaMemsetAsync(P.mom, 0, bytes, s->stream));
    P.mom_half = max_samples / 2;
    P.mom_batch = (uint32_t)std::max<uint64_t>(1, (uint64_t)std::floor(std::sqrt((double)max_samples)));
  } else if (max_samples > 0) {
    CU(cudaMalloc((void**)&P.samples, sizeof(double) * max_samples * P.sample_chains * P.retdim));
    CU(cudaMalloc((void**)&P.samp_lp, sizeof(double) * max_samples * P.sample_chains));
  }
  if (s->cfg.flags & QSB_FLAG_RECORD) {
    if (numiters == 0) return fail(2, "QSB_FLAG_RECORD needs numiters");
    const size_t C = P.C, d = P.d;
    s->rec_iters = numiters; P.rec_iters = numiters;
    CU(cudaMalloc((void**)&P.rec_state, sizeof(double) * C * numiters * d));
    CU(cudaMalloc((void**)&P.rec_dh, sizeof(double) * C * numiters));
    CU(cudaMalloc((void**)&P.rec_step, sizeof(double) * C * numiters));
    CU(cudaMalloc((void**)&P.rec_accept, sizeof(int) * C * numiters));
    CU(cudaMemset(P.rec_state, 0, sizeof(double) * C * numiters * d));
    if (P.nsub > 1) CU(cudaMalloc((void**)&P.rec_kidx, sizeof(int) * C * numiters));
    if ((s->cfg.flags & QSB_FLAG_RECORD_LEAP) && P.nsub == 1 && P.kind[0] == K_HMC) {
      CU(cudaMalloc((void**)&P.rec_leap, sizeof(double) * C * numiters * P.numSteps[0] * d));
      CU(cudaMemset(P.rec_leap, 0, sizeof(double) * C * numiters * P.numSteps[0] * d));
    }
  }
  return 0;
}

}  // extern "C"
static int fold_kernel_times(qsb_sampler* s) {
  if (!s->tev_used) return 0;
  CU(cudaStreamSynchronize(s->stream));
  for (size_t i = 0; i + 1 < s->tev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->tev[i], s->tev[i + 1]) == cudaSuccess) { s->t_accum += ms * 1e-3; ++s->t_launches; }
    else (void)cudaGetLastError();
  }
  s->tev_used = 0;
  return 0;
}
extern "C" {

int qsb_sampler_advance(qsb_sampler* s, uint64_t iters) {
  if (!s) return fail(1, "null sampler");
  if (iters == 0) return 0;
  CU(cudaSetDevice(s->m->device));
  if (s->iter_base + s->iter + iters >= QSB_ITER_SEARCH) return fail(2, "iteration counter would overflow the Philox iteration word");
  if (s->glm) {
    std::string why;
    if (glm_sampler_advance(s->glm, iters, why)) return fail(5, why);
    s->iter += iters;
    return 0;
  }
  EngineParams& P = s->P;
  if (s->rec_iters && s->iter + iters > s->rec_iters) return fail(2, "record buffer holds numiters iterations only");
  const bool timed = (s->cfg.flags & QSB_FLAG_TIME_KERNEL) != 0;
  if (timed && s->tev_used + 2 > s->tev.size()) {
    if (s->tev.size() >= 8192) { if (int rc = fold_kernel_times(s)) return rc; }
    else for (int i = 0; i < 512; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); s->tev.push_back(e); }
  }
  P.it_begin = s->iter_base + s->iter; P.n_iters = iters; P.it_origin = s->iter_base;
  CU(cudaEventRecord(s->ev0, s->stream));
  if (timed) CU(cudaEventRecord(s->tev[s->tev_used], s->stream));
  {
    // the kernel indexes record rows by (it - it_begin); shift the bases so rows are absolute iterations
    EngineParams R = P;
    if (R.rec_state) {
      const size_t d = R.d, off = s->iter;
      R.rec_state += off * d; R.rec_dh += off; R.rec_step += off; R.rec_accept += off;
      if (R.rec_kidx) R.rec_kidx += off;
      if (R.rec_leap) R.rec_leap += off * R.numSteps[0] * d;
    }
    bool has_hmc = false;
    for (int k = 0; k < R.nsub; ++k) has_hmc |= R.kind[k] == K_HMC;
    bool has_mh = false;
    for (int k = 0; k < R.nsub; ++k) has_mh |= R.kind[k] == K_TRACEMH;
    // QSB_WS=1: the warp-per-chain variants run warp-specialised (sampler warps + trajectory warps, engine.cuh:
    // advance_ws_kernel) instead of fused (every warp samples and integrates)
    static const bool ws = getenv("QSB_WS") && atoi(getenv("QSB_WS")) != 0;   // (not the default yet: see profiles/README.md)
    const bool use_ws = ws && s->var->advance_ws != nullptr && !R.vscratch;   // (the scratch slices are indexed by the fused kernels' warp number)
    if (R.dr_lexical || has_mh) s->var->advance_lex(R, s->stream);   // the general kernel
    else if (R.nsub == 1 && R.kind[0] == K_DRIFT) (use_ws ? s->var->advance_drift_ws : s->var->advance_drift)(R, s->stream);
    else if (has_hmc) (use_ws ? s->var->advance_ws : s->var->advance)(R, s->stream);
    else (use_ws ? s->var->advance_nohmc_ws : s->var->advance_nohmc)(R, s->stream);
  }
  ++s->launches;
  if (timed) { CU(cudaEventRecord(s->tev[s->tev_used + 1], s->stream)); s->tev_used += 2; }
  CU(cudaEventRecord(s->ev1, s->stream));
  CU(cudaGetLastError());
  s->iter += iters; s->advanced = true;
  return 0;
}

int qsb_sampler_set_temperatures(qsb_sampler* s, const double* temps, uint64_t n) {
  if (!s) return fail(1, "null sampler");
  CU(cudaSetDevice(s->m->device));
  for (uint64_t i = 0; temps && i < n; ++i) if (!(temps[i] > 0.0)) return fail(2, "qsb_sampler_set_temperatures: temperatures must be positive");
  if (s->glm) { std::string why; if (glm_sampler_set_temperatures(s->glm, temps, temps ? n : 0, why)) return fail(5, why); return 0; }
  CU(cudaStreamSynchronize(s->stream));
  if (s->temps_dev) { cudaFree(s->temps_dev); s->temps_dev = nullptr; }
  s->P.temps = nullptr; s->P.ntemps = 0;
  if (temps && n) {
    CU(cudaMalloc((void**)&s->temps_dev, 8 * n));
    CU(cudaMemcpy(s->temps_dev, temps, 8 * n, cudaMemcpyHostToDevice));
    s->P.temps = s->temps_dev; s->P.ntemps = n;
  }
  return 0;
}

int qsb_sampler_run(qsb_sampler* s, uint64_t nsamps, uint64_t burnin, uint64_t lag) {
  if (!s) return fail(1, "null sampler");
  if (lag == 0) return fail(2, "lag must be >= 1");
  const uint64_t iters = burnin + nsamps * lag;   // mcmc.t:53
  if (int rc = qsb_sampler_begin(s, nsamps + 1, burnin, lag, iters)) return rc;
  if (int rc = qsb_sampler_advance(s, iters)) return rc;
  CU(cudaStreamSynchronize(s->stream));
  if (!s->glm) if (int rc = check_status(s)) return rc;
  return 0;
}

uint64_t qsb_sampler_num_samples(const qsb_sampler* s) {
  if (!s) return 0;
  // iterations i in [0, iter) with i >= burnin and i % lag == 0
  if (s->iter == 0) return 0;
  const uint64_t last = (s->iter - 1) / s->lag;                  // largest multiple index below iter
  const uint64_t first = (s->burnin + s->lag - 1) / s->lag;      // first kept multiple index
  uint64_t n = last >= first ? last - first + 1 : 0;
  return std::min(n, s->max_samples);
}

static int fetch_common(qsb_sampler* s, uint32_t c0, uint32_t c1, uint64_t s0, uint64_t s1, void* dst, size_t stride_bytes, int with_lp) {
  if (!s || !dst) return fail(1, "null argument");
  CU(cudaSetDevice(s->m->device));
  const double *samples, *samp_lp; int G; uint32_t SC; int retdim;
  if (s->glm) glm_sampler_sample_view(s->glm, &samples, &samp_lp, &G, &SC, &retdim);
  else { samples = s->P.samples; samp_lp = s->P.samp_lp; G = s->var->G; SC = s->P.sample_chains; retdim = s->P.retdim; }
  const uint64_t have = qsb_sampler_num_samples(s);
  if (s->cfg.flags & QSB_FLAG_MOMENTS) return fail(2, "fetch: QSB_FLAG_MOMENTS keeps running statistics, no draws are stored");
  if (have > 0 && !samples) return fail(2, "fetch: the sampler holds no stored draws");
  if (c1 > SC || c0 > c1 || s1 > have || s0 > s1) return fail(2, "fetch: chain/sample range out of bounds (chains stored: " + std::to_string(SC) + ", samples kept: " + std::to_string(have) + ")");
  const int w = retdim + (with_lp ? 2 : 0);
  if (stride_bytes < (size_t)w * 8 || stride_bytes % 8) return fail(2, "fetch: stride must be a multiple of 8 and >= element size");
  const uint64_t ns = s1 - s0;
  if (ns == 0 || c1 == c0) return 0;
  const size_t stride_d = stride_bytes / 8;
  // staged through a device buffer, at most ~256 MiB at a time
  uint32_t chunk = (uint32_t)std::max<size_t>(1, (256u << 20) / (ns * stride_bytes));
  chunk = std::min<uint32_t>(chunk, c1 - c0);
  const size_t need = (size_t)chunk * ns * stride_bytes;
  if (s->stage_bytes < need) {
    if (s->stage) cudaFree(s->stage);
    s->stage = nullptr; s->stage_bytes = 0;
    CU(cudaMalloc((void**)&s->stage, need));
    s->stage_bytes = need;
  }
  double* stage = s->stage;
  int rc = 0;
  for (uint32_t c = c0; c < c1 && !rc; c += chunk) {
    uint32_t nc = std::min<uint32_t>(chunk, c1 - c);
    size_t total = (size_t)nc * ns * w;
    unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
    gather_samples_kernel<<<grid, 256, 0, s->stream>>>(samples, samp_lp, G, SC, retdim, c, nc, s0, ns, stage, stride_d, with_lp);
    ++s->launches;
    cudaError_t e = cudaMemcpyAsync((char*)dst + (size_t)(c - c0) * ns * stride_bytes, stage, (size_t)nc * ns * stride_bytes, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) rc = fail(100 + (int)e, std::string("fetch: ") + cudaGetErrorString(e));
  }
  return rc;
}

int qsb_sampler_fetch_samples(qsb_sampler* s, uint32_t c0, uint32_t c1, uint64_t s0, uint64_t s1, void* dst, size_t stride_bytes) {
  return fetch_common(s, c0, c1, s0, s1, dst, stride_bytes, 1);
}
int qsb_sampler_fetch_samples_async(qsb_sampler* s, uint32_t c0, uint32_t c1, uint64_t s0, uint64_t s1, void* dst, size_t stride_bytes) {
  if (!s || !dst) return fail(1, "null argument");
  CU(cudaSetDevice(s->m->device));
  const double *samples, *samp_lp; int G; uint32_t SC; int retdim;
  if (s->glm) glm_sampler_sample_view(s->glm, &samples, &samp_lp, &G, &SC, &retdim);
  else { samples = s->P.samples; samp_lp = s->P.samp_lp; G = s->var->G; SC = s->P.sample_chains; retdim = s->P.retdim; }
  const uint64_t have = qsb_sampler_num_samples(s);
  if (s->cfg.flags & QSB_FLAG_MOMENTS) return fail(2, "fetch: QSB_FLAG_MOMENTS keeps running statistics, no draws are stored");
  if (have > 0 && !samples) return fail(2, "fetch: the sampler holds no stored draws");
  if (c1 > SC || c0 > c1 || s1 > have || s0 > s1) return fail(2, "fetch: chain/sample range out of bounds (chains stored: " + std::to_string(SC) + ", samples kept: " + std::to_string(have) + ")");
  const int w = retdim + 2;
  if (stride_bytes < (size_t)w * 8 || stride_bytes % 8) return fail(2, "fetch: stride must be a multiple of 8 and >= element size");
  const uint64_t ns = s1 - s0; const uint32_t nc = c1 - c0;
  if (ns == 0 || nc == 0) return 0;
  const size_t need = (size_t)nc * ns * stride_bytes;
  if (need > ((size_t)4 << 30)) return fail(2, "fetch_samples_async: at most 4 GiB per call");
  if (!s->copy_stream) {
    CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) { CU(cudaEventCreateWithFlags(&s->gathered[k], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&s->copied[k], cudaEventDisableTiming)); }
  }
  const int k = s->aflip; s->aflip ^= 1;
  if (s->astage_bytes[k] < need) {
    CU(cudaStreamSynchronize(s->copy_stream));
    if (s->astage[k]) cudaFree(s->astage[k]);
    s->astage[k] = nullptr; s->astage_bytes[k] = 0;
    CU(cudaMalloc((void**)&s->astage[k], need));
    s->astage_bytes[k] = need;
  }
  CU(cudaStreamWaitEvent(s->stream, s->copied[k], 0));      // the copy that last read this staging buffer is done
  const size_t total = (size_t)nc * ns * w;
  const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
  gather_samples_kernel<<<grid, 256, 0, s->stream>>>(samples, samp_lp, G, SC, retdim, c0, nc, s0, ns, s->astage[k], stride_bytes / 8, 1);
  ++s->launches;
  CU(cudaGetLastError());
  CU(cudaEventRecord(s->gathered[k], s->stream));
  CU(cudaStreamWaitEvent(s->copy_stream, s->gathered[k], 0));
  CU(cudaMemcpyAsync(dst, s->astage[k], need, cudaMemcpyDeviceToHost, s->copy_stream));
  CU(cudaEventRecord(s->copied[k], s->copy_stream));
  return 0;
}
int qsb_sampler_fetch_sync(qsb_sampler* s) {
  if (!s) return fail(1, "null sampler");
  CU(cudaSetDevice(s->m->device));
  if (s->copy_stream) CU(cudaStreamSynchronize(s->copy_stream));
  return 0;
}
int qsb_sampler_fetch_values(qsb_sampler* s, uint32_t c0, uint32_t c1, uint64_t s0, uint64_t s1, double* dst) {
  if (!s) return fail(1, "null sampler");
  int retdim = s->glm ? s->m->retdim : s->P.retdim;
  return fetch_common(s, c0, c1, s0, s1, dst, (size_t)retdim * 8, 0);
}

int qsb_sampler_stats(qsb_sampler* s, qsb_stats* out) {
  if (!s || !out) return fail(1, "null argument");
  CU(cudaSetDevice(s->m->device));
  memset(out, 0, sizeof(*out));
  CU(cudaStreamSynchronize(s->stream));
  if (s->glm) {
    std::string why;
    if (glm_sampler_stats(s->glm, &out->props_made, &out->props_accepted, &out->grad_evals, &out->leapfrog_steps, &out->kernel_launches, &out->seconds, why)) return fail(5, why);
    glm_sampler_health(s->glm, &out->nonfinite, &out->divergent);
    out->sub_made[0] = out->props_made; out->sub_accepted[0] = out->props_accepted;
    out->iterations = s->iter;
    return 0;
  }
  EngineParams& P = s->P;
  const size_t C = P.C;
  std::vector<unsigned long long> made(P.nsub * C), acc(P.nsub * C);
  unsigned long long counters[4];
  CU(cudaMemcpy(made.data(), P.made, 8 * made.size(), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(acc.data(), P.acc, 8 * acc.size(), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(counters, P.counters, 32, cudaMemcpyDeviceToHost));
  for (int k = 0; k < P.nsub; ++k) for (size_t c = 0; c < C; ++c) {
    out->sub_made[k] += made[k * C + c]; out->sub_accepted[k] += acc[k * C + c];
  }
  for (int k = 0; k < P.nsub; ++k) { out->props_made += out->sub_made[k]; out->props_accepted += out->sub_accepted[k]; }
  out->grad_evals = counters[0]; out->leapfrog_steps = counters[1];
  out->nonfinite = counters[2]; out->divergent = counters[3];
  out->iterations = s->iter; out->kernel_launches = s->launches;
  float ms = 0.f;
  // (events that were never recorded make cudaEventElapsedTime fail AND leave that error pending for the next
  // cudaGetLastError of this thread: only ask once an advance has recorded them, and clear a failure)
  if (s->advanced) { if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) out->seconds = ms * 1e-3; else (void)cudaGetLastError(); }
  return 0;
}

int qsb_sampler_set_kernel_timing(qsb_sampler* s, int32_t on) {
  if (!s) return fail(1, "null sampler");
  CU(cudaSetDevice(s->m->device));
  if (on) s->cfg.flags |= QSB_FLAG_TIME_KERNEL; else s->cfg.flags &= ~(uint32_t)QSB_FLAG_TIME_KERNEL;
  if (s->glm) glm_sampler_set_timing(s->glm, on != 0);
  return 0;
}

int qsb_sampler_kernel_time(qsb_sampler* s, double* seconds, uint64_t* launches) {
  if (!s || !seconds || !launches) return fail(1, "null argument");
  CU(cudaSetDevice(s->m->device));
  if (s->glm) { std::string why; if (glm_sampler_kernel_time(s->glm, seconds, launches, why)) return fail(5, why); return 0; }
  if (int rc = fold_kernel_times(s)) return rc;
  *seconds = s->t_accum; *launches = s->t_launches;
  s->t_accum = 0.0; s->t_launches = 0;
  return 0;
}

int qsb_sampler_diagnostics(qsb_sampler* s, int32_t maxlag, double* dst, int32_t dst_is_device) {
  if (!s || !dst) return fail(1, "null argument");
  CU(cudaSetDevice(s->m->device));
  if (s->cfg.flags & QSB_FLAG_MOMENTS) {   // from the running statistics: 12 + 3 doubles per dimension, whatever maxlag says
    const double* mom = nullptr; uint64_t half = 0; uint32_t batch = 1; int G; uint32_t SC; int retdim;
    const double *sm_, *sl_;
    if (s->glm) { glm_sampler_sample_view(s->glm, &sm_, &sl_, &G, &SC, &retdim); glm_sampler_moments_view(s->glm, &mom, &half, &batch); }
    else { mom = s->P.mom; half = s->P.mom_half; batch = s->P.mom_batch; G = s->var->G; SC = s->P.sample_chains; retdim = s->P.retdim; }
    const uint64_t n = qsb_sampler_num_samples(s);
    const size_t len = (size_t)retdim * 15;
    double* out = dst;
    if (!dst_is_device) CU(cudaMalloc((void**)&out, len * 8));
    CU(cudaMemsetAsync(out, 0, len * 8, s->stream));
    if (n >= 2 && mom) {
      const size_t total = (size_t)SC * retdim;
      diagnostics_moments_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s->stream>>>(mom, G == 1 ? 1 : 0, SC, retdim, n, half, batch, out);
      const double nh = (n > half && half >= 2 && n - half >= 2) ? 0.5 * (double)n : 0.0;
      diagnostics_header_kernel<<<(retdim + 127) / 128, 128, 0, s->stream>>>(out, retdim, nh > 0.0 ? 2.0 * SC : 0.0, nh, (double)SC, (double)n);
      s->launches += 2;
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !dst_is_device) e = cudaMemcpyAsync(dst, out, len * 8, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (!dst_is_device) cudaFree(out);
    if (e != cudaSuccess) return fail(100 + (int)e, std::string("diagnostics: ") + cudaGetErrorString(e));
    return 0;
  }
  if (maxlag < 0) return fail(2, "maxlag must be >= 0");
  const double *samples, *samp_lp; int G; uint32_t SC; int retdim;
  if (s->glm) glm_sampler_sample_view(s->glm, &samples, &samp_lp, &G, &SC, &retdim);
  else { samples = s->P.samples; samp_lp = s->P.samp_lp; G = s->var->G; SC = s->P.sample_chains; retdim = s->P.retdim; }
  const uint64_t n = qsb_sampler_num_samples(s);
  const size_t len = (size_t)retdim * 12 + (maxlag > 0 ? (size_t)retdim * (maxlag + 1) : 0);
  double* out = dst;
  if (!dst_is_device) CU(cudaMalloc((void**)&out, len * 8));
  CU(cudaMemsetAsync(out, 0, len * 8, s->stream));
  if (n >= 2 && samples) {
    size_t total = (size_t)SC * retdim;
    diagnostics_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s->stream>>>(samples, G, SC, retdim, n, maxlag, out);
    diagnostics_header_kernel<<<(retdim + 127) / 128, 128, 0, s->stream>>>(out, retdim, (n / 2 >= 2) ? 2.0 * SC : 0.0, (double)(n / 2), (double)SC, (double)n);
    s->launches += 2;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && !dst_is_device) e = cudaMemcpyAsync(dst, out, len * 8, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  if (!dst_is_device) cudaFree(out);
  if (e != cudaSuccess) return fail(100 + (int)e, std::string("diagnostics: ") + cudaGetErrorString(e));
  return 0;
}

// ---- queries
}  // extern "C"

struct DevTmp {   // scoped device scratch
  std::vector<void*> p;
  template <class T> cudaError_t get(T** q, size_t n) { *q = nullptr; cudaError_t e = cudaMalloc((void**)q, std::max<size_t>(n, 1) * sizeof(T)); if (e == cudaSuccess) p.push_back(*q); return e; }
  ~DevTmp() { for (void* q : p) cudaFree(q); }
};

static int sampler_view(qsb_sampler* s, uint32_t c0, uint32_t c1, SampleView& V) {
  if (!s) return fail(1, "null sampler");
  CU(cudaSetDevice(s->m->device));
  int G; const double *samples, *samp_lp; uint32_t SC; int retdim;
  if (s->glm) glm_sampler_sample_view(s->glm, &samples, &samp_lp, &G, &SC, &retdim);
  else { samples = s->P.samples; samp_lp = s->P.samp_lp; G = s->var->G; SC = s->P.sample_chains; retdim = s->P.retdim; }
  const uint64_t n = qsb_sampler_num_samples(s);
  if (n == 0 || !samples) return fail(2, "query: the sampler holds no samples");   // S.assert(samples:size() > 0) (infer.t:153, 194)
  if (c1 > SC || c0 >= c1) return fail(2, "query: chain range out of bounds (chains stored: " + std::to_string(SC) + ")");
  V.samples = samples; V.samp_lp = samp_lp; V.samp_ll = nullptr; V.layout = G == 1 ? 1 : 32; V.SC = SC; V.retdim = retdim; V.n = n; V.stride = 0;
  return 0;
}
static int host_view(const void* samples, size_t stride_bytes, uint32_t nchains, uint64_t nsamps, int32_t retdim, DevTmp& tmp, SampleView& V) {
  if (!samples || nchains == 0 || retdim <= 0) return fail(1, "query: bad argument");
  if (nsamps == 0) return fail(2, "query: empty sample vector");
  if (stride_bytes < (size_t)(retdim + 2) * 8 || stride_bytes % 8) return fail(2, "query: stride must be a multiple of 8 and >= element size");
  double* d;
  const size_t bytes = (size_t)nchains * nsamps * stride_bytes;
  CU(tmp.get(&d, bytes / 8));
  CU(cudaMemcpy(d, samples, bytes, cudaMemcpyHostToDevice));
  V.samples = d; V.samp_lp = nullptr; V.samp_ll = nullptr; V.layout = 0; V.SC = nchains; V.retdim = retdim; V.n = nsamps; V.stride = stride_bytes / 8;
  return 0;
}
static int run_expectation(const SampleView& V, uint32_t c0, uint32_t nc, cudaStream_t st, double* mean, double* var, double* mean_dev_out, double* var_dev_out) {
  DevTmp tmp;
  double *dm = mean_dev_out, *dv = var_dev_out;
  if (!dm) CU(tmp.get(&dm, (size_t)nc * V.retdim));
  if (!dv && (var || var_dev_out)) CU(tmp.get(&dv, nc));
  const size_t total = (size_t)nc * V.retdim;
  expectation_mean_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(V, c0, nc, dm);
  if (dv) expectation_var_kernel<<<(unsigned)(((size_t)nc * 32 + 127) / 128), 128, 0, st>>>(V, c0, nc, dm, dv);
  CU(cudaGetLastError());
  if (mean) CU(cudaMemcpyAsync(mean, dm, total * 8, cudaMemcpyDeviceToHost, st));
  if (var) CU(cudaMemcpyAsync(var, dv, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}
static int run_map(const SampleView& V, uint32_t c0, uint32_t nc, cudaStream_t st, double* value, double* logprob) {
  if (!value) return fail(1, "query: null output");
  DevTmp tmp;
  double *dv, *dl;
  CU(tmp.get(&dv, (size_t)nc * V.retdim)); CU(tmp.get(&dl, nc));
  map_kernel<<<(nc + 127) / 128, 128, 0, st>>>(V, c0, nc, dv, dl);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(value, dv, (size_t)nc * V.retdim * 8, cudaMemcpyDeviceToHost, st));
  if (logprob) CU(cudaMemcpyAsync(logprob, dl, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}
static int run_autocorrelation(const SampleView& V, uint32_t c0, uint32_t nc, cudaStream_t st, const double* mean, double variance, double* ac) {
  if (!ac) return fail(1, "query: null output");
  DevTmp tmp;
  double *dm, *dv = nullptr, *dac;
  CU(tmp.get(&dac, (size_t)nc * V.n));
  if (mean) {
    CU(tmp.get(&dm, V.retdim));
    CU(cudaMemcpyAsync(dm, mean, (size_t)V.retdim * 8, cudaMemcpyHostToDevice, st));
  } else {   // noMeanAndVar (infer.t:244-247): each chain's own Expectation(true)
    CU(tmp.get(&dm, (size_t)nc * V.retdim)); CU(tmp.get(&dv, nc));
    if (int rc = run_expectation(V, c0, nc, st, nullptr, nullptr, dm, dv)) return rc;
  }
  autocorrelation_kernel<<<nc, 128, 0, st>>>(V, c0, dm, dv, variance, mean ? 0 : 1, dac);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ac, dac, (size_t)nc * V.n * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

extern "C" {

int qsb_query_expectation(qsb_sampler* s, uint32_t c0, uint32_t c1, double* mean, double* var) {
  SampleView V;
  if (int rc = sampler_view(s, c0, c1, V)) return rc;
  if (!mean) return fail(1, "query: null output");
  s->launches += var ? 2 : 1;
  return run_expectation(V, c0, c1 - c0, s->stream, mean, var, nullptr, nullptr);
}
int qsb_query_expectation_host(const void* samples, size_t stride_bytes, uint32_t nchains, uint64_t nsamps, int32_t retdim, double* mean, double* var) {
  DevTmp tmp; SampleView V;
  if (int rc = host_view(samples, stride_bytes, nchains, nsamps, retdim, tmp, V)) return rc;
  if (!mean) return fail(1, "query: null output");
  return run_expectation(V, 0, nchains, nullptr, mean, var, nullptr, nullptr);
}
int qsb_query_map(qsb_sampler* s, uint32_t c0, uint32_t c1, double* value, double* logprob) {
  SampleView V;
  if (int rc = sampler_view(s, c0, c1, V)) return rc;
  ++s->launches;
  return run_map(V, c0, c1 - c0, s->stream, value, logprob);
}
int qsb_query_map_host(const void* samples, size_t stride_bytes, uint32_t nchains, uint64_t nsamps, int32_t retdim, double* value, double* logprob) {
  DevTmp tmp; SampleView V;
  if (int rc = host_view(samples, stride_bytes, nchains, nsamps, retdim, tmp, V)) return rc;
  return run_map(V, 0, nchains, nullptr, value, logprob);
}
int qsb_query_autocorrelation(qsb_sampler* s, uint32_t c0, uint32_t c1, const double* mean, double variance, double* ac) {
  SampleView V;
  if (int rc = sampler_view(s, c0, c1, V)) return rc;
  s->launches += mean ? 1 : 3;
  return run_autocorrelation(V, c0, c1 - c0, s->stream, mean, variance, ac);
}
int qsb_query_autocorrelation_host(const void* samples, size_t stride_bytes, uint32_t nchains, uint64_t nsamps, int32_t retdim,
                                   const double* mean, double variance, double* ac) {
  DevTmp tmp; SampleView V;
  if (int rc = host_view(samples, stride_bytes, nchains, nsamps, retdim, tmp, V)) return rc;
  return run_autocorrelation(V, 0, nchains, nullptr, mean, variance, ac);
}

int qsb_sampler_fetch_record(qsb_sampler* s, double* state, int32_t* accept, double* dh, double* step, int32_t* kidx, double* leap) {
  if (!s) return fail(1, "null sampler");
  CU(cudaSetDevice(s->m->device));
  if (s->glm) { std::string why; if (glm_sampler_fetch_record(s->glm, state, accept, dh, step, leap, why)) return fail(5, why); return 0; }
  EngineParams& P = s->P;
  if (!P.rec_state) return fail(2, "sampler was not created with QSB_FLAG_RECORD");
  CU(cudaStreamSynchronize(s->stream));
  const size_t C = P.C, d = P.d, T = s->rec_iters;
  if (state) CU(cudaMemcpy(state, P.rec_state, 8 * C * T * d, cudaMemcpyDeviceToHost));
  if (accept) CU(cudaMemcpy(accept, P.rec_accept, 4 * C * T, cudaMemcpyDeviceToHost));
  if (dh) CU(cudaMemcpy(dh, P.rec_dh, 8 * C * T, cudaMemcpyDeviceToHost));
  if (step) CU(cudaMemcpy(step, P.rec_step, 8 * C * T, cudaMemcpyDeviceToHost));
  if (kidx) { if (!P.rec_kidx) return fail(2, "kernel index is recorded for mixtures only"); CU(cudaMemcpy(kidx, P.rec_kidx, 4 * C * T, cudaMemcpyDeviceToHost)); }
  if (leap) { if (!P.rec_leap) return fail(2, "leapfrog positions need QSB_FLAG_RECORD_LEAP and a single HMC kernel"); CU(cudaMemcpy(leap, P.rec_leap, 8 * C * T * P.numSteps[0] * d, cudaMemcpyDeviceToHost)); }
  return 0;
}

// ---- test hooks
int qsb_logp_grad(qsb_model* m, const double* x, int32_t n, double* lp_dual, double* grad, double* lp_float) {
  if (!m || !x || n <= 0) return fail(1, "bad argument");
  CU(cudaSetDevice(m->device));
  if (m->desc.family == QSB_FAM_LOGISTIC) {
    std::string why;
    if (glm_logp_grad(&m->glm, x, n, lp_dual, grad, lp_float, why)) return fail(5, why);
    return 0;
  }
  qsb_kernel_spec spec = {QSB_KERNEL_HMC, 0, 1.0, 1, 1.0, 1.0, 0, 0};
  qsb_run_config cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.numchains = (uint32_t)n; cfg.device = m->device;
  const char* env = getenv("QSB_TEST_MAPPING");
  if (env) cfg.reserved = atoi(env);
  qsb_sampler* s = nullptr;
  if (int rc = qsb_sampler_create(m, &spec, 1, &cfg, &s)) return rc;
  int rc = qsb_sampler_set_state(s, x);
  if (!rc) {
    s->var->logp_grad(s->P, s->stream);
    const size_t C = n, d = m->desc.dim;
    std::vector<double> g(C * d);
    cudaError_t e = cudaMemcpy(g.data(), s->P.g, 8 * C * d, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && lp_dual) e = cudaMemcpy(lp_dual, s->P.lp, 8 * C, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && lp_float) e = cudaMemcpy(lp_float, s->P.eps, 8 * C, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(100 + (int)e, cudaGetErrorString(e));
    else if (grad) {
      if (s->var->G == 1) { for (size_t c = 0; c < C; ++c) for (size_t i = 0; i < d; ++i) grad[c * d + i] = g[i * C + c]; }
      else std::copy(g.begin(), g.end(), grad);
    }
  }
  qsb_sampler_destroy(s);
  return rc;
}

int qsb_debug_leapfrog(qsb_model* m, int32_t n, const double* x, const double* p, const double* g, const double* eps, int32_t last,
                       double* x_out, double* p_out, double* g_out, double* lp_out) {
  if (!m || !x || !p || !eps || n <= 0) return fail(1, "bad argument");
  CU(cudaSetDevice(m->device));
  if (m->desc.family == QSB_FAM_LOGISTIC) {
    std::string why;
    if (glm_debug_leapfrog(&m->glm, n, x, p, g, eps, last, x_out, p_out, g_out, lp_out, why)) return fail(5, why);
    return 0;
  }
  qsb_kernel_spec spec = {QSB_KERNEL_HMC, 0, 1.0, 1, 1.0, 1.0, 0, 0};
  qsb_run_config cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.numchains = (uint32_t)n; cfg.device = m->device;
  const char* env = getenv("QSB_TEST_MAPPING");
  if (env) cfg.reserved = atoi(env);
  qsb_sampler* s = nullptr;
  if (int rc = qsb_sampler_create(m, &spec, 1, &cfg, &s)) return rc;
  int rc = qsb_sampler_set_state(s, x);
  if (!rc) {
    const size_t C = n, d = m->desc.dim;
    const bool tr = s->var->G == 1;   // [n][C] layout
    DevTmp tmp;
    double *dp = nullptr, *dg = nullptr, *de = nullptr, *ox = nullptr, *op = nullptr, *og = nullptr, *ol = nullptr;
    cudaError_t e = tmp.get(&dp, C * d);
    if (e == cudaSuccess && g) e = tmp.get(&dg, C * d);
    if (e == cudaSuccess) e = tmp.get(&de, C);
    if (e == cudaSuccess) e = tmp.get(&ox, C * d);
    if (e == cudaSuccess) e = tmp.get(&op, C * d);
    if (e == cudaSuccess) e = tmp.get(&og, C * d);
    if (e == cudaSuccess) e = tmp.get(&ol, C);
    std::vector<double> h(C * d);
    auto up = [&](const double* src, double* dst) {
      if (tr) { for (size_t c = 0; c < C; ++c) for (size_t i = 0; i < d; ++i) h[i * C + c] = src[c * d + i]; }
      else std::copy(src, src + C * d, h.begin());
      return cudaMemcpy(dst, h.data(), 8 * C * d, cudaMemcpyHostToDevice);
    };
    auto down = [&](const double* src, double* dst) {
      cudaError_t e2 = cudaMemcpy(h.data(), src, 8 * C * d, cudaMemcpyDeviceToHost);
      if (e2 != cudaSuccess) return e2;
      if (tr) { for (size_t c = 0; c < C; ++c) for (size_t i = 0; i < d; ++i) dst[c * d + i] = h[i * C + c]; }
      else std::copy(h.begin(), h.end(), dst);
      return e2;
    };
    if (e == cudaSuccess) e = up(p, dp);
    if (e == cudaSuccess && g) e = up(g, dg);
    if (e == cudaSuccess) e = cudaMemcpy(de, eps, 8 * C, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      DebugLeap D{dp, dg, de, ox, op, og, ol, last ? 1 : 0};
      s->var->debug_leap(s->P, D, s->stream);
      e = cudaStreamSynchronize(s->stream);
    }
    if (e == cudaSuccess && x_out) e = down(ox, x_out);
    if (e == cudaSuccess && p_out) e = down(op, p_out);
    if (e == cudaSuccess && g_out) e = down(og, g_out);
    if (e == cudaSuccess && lp_out) e = cudaMemcpy(lp_out, ol, 8 * C, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(100 + (int)e, cudaGetErrorString(e));
  }
  qsb_sampler_destroy(s);
  return rc;
}

int qsb_debug_glm_layout(uint32_t numchains, int32_t nobs, int32_t nsm, int32_t capacity, int32_t* start, int32_t* first, int32_t* np,
                         int32_t* ntiles, int32_t* nblocks, int32_t* ncta) {
  if (numchains == 0 || nobs <= 0 || nsm <= 0 || !ntiles || !nblocks || !ncta) return fail(1, "bad argument");
  const int CBt = glm_chains_per_tile(), MBr = glm_rows_per_block();
  const int nt = (int)((numchains + CBt - 1) / CBt), nb = (nobs + MBr - 1) / MBr;
  std::vector<int> st, fi, npv;
  const int nc = glm_layout(nt, nb, numchains, nsm, st, fi, npv);
  *ntiles = nt; *nblocks = nb; *ncta = nc;
  if (start) { if (capacity < nc + 1) return fail(2, "capacity"); std::copy(st.begin(), st.end(), start); }
  if (first) { if (capacity < nt) return fail(2, "capacity"); std::copy(fi.begin(), fi.end(), first); }
  if (np) { if (capacity < nt) return fail(2, "capacity"); std::copy(npv.begin(), npv.end(), np); }
  return 0;
}

int qsb_logprob(int32_t which, const double* args, int32_t n, double* out) {
  if (!args || !out || n <= 0) return fail(1, "bad argument");
  double *da = nullptr, *dout = nullptr;
  CU(cudaMalloc((void**)&da, 24 * (size_t)n)); CU(cudaMalloc((void**)&dout, 8 * (size_t)n));
  cudaMemcpy(da, args, 24 * (size_t)n, cudaMemcpyHostToDevice);
  logprob_kernel<<<(n + 127) / 128, 128>>>(which, da, n, dout);
  cudaError_t e = cudaMemcpy(out, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(dout);
  if (e != cudaSuccess) return fail(100 + (int)e, cudaGetErrorString(e));
  return 0;
}

int qsb_uniforms(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot, int32_t n, double* out) {
  if (!out || n <= 0) return fail(1, "bad argument");
  double* d = nullptr;
  CU(cudaMalloc((void**)&d, 8 * (size_t)n));
  uniforms_kernel<<<1, 1>>>(RngKey{(uint32_t)seed, (uint32_t)(seed >> 32)}, chain, iter, slot, n, d);
  cudaError_t e = cudaMemcpy(out, d, 8 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(100 + (int)e, cudaGetErrorString(e));
  return 0;
}

int qsb_gaussians(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot0, int32_t n, double mu, double sigma, double* out) {
  if (!out || n <= 0) return fail(1, "bad argument");
  double* d = nullptr;
  CU(cudaMalloc((void**)&d, 8 * (size_t)n));
  gaussians_kernel<<<(n + 127) / 128, 128>>>(RngKey{(uint32_t)seed, (uint32_t)(seed >> 32)}, chain, iter, slot0, n, mu, sigma, d);
  cudaError_t e = cudaMemcpy(out, d, 8 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(100 + (int)e, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
