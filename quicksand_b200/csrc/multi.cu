// Multi-GPU layer of the C ABI (include/qsb200.h: qsb_group_*, qsb_comm_*) and the in-process FP64 peak measurement.
//
// Chains are independent: a group is one qsb_sampler per device over a block partition of the global
// chain range, driven by ONE host thread like the reference's single-threaded host (qs/trace.t:452-454).  No data-path
// collective exists; the only exchange is the all-reduce(sum) of the diagnostics' sufficient statistics.  It runs on the
// library's own NCCL communicator (NVLink / NVSwitch).  NCCL is bound at run time (dlopen of libnccl.so.2) so that the
// single-device library keeps depending on the CUDA runtime alone -- a Terra process that never asks for several devices
// never loads it.  Shards that share a device (tests on a one-GPU box) are reduced by a device kernel instead.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/qsb200.h"

// capi.cu internals (not in the public header): the thread-local message behind qsb_last_error(), and two accessors
extern "C" void qsb_set_error_(const char* msg);
extern "C" int qsb_sampler_retdim_(qsb_sampler* s);
extern "C" void* qsb_sampler_stream_(qsb_sampler* s);

namespace {

int fail(int code, const std::string& msg) { qsb_set_error_(msg.c_str()); return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(100 + (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
  bool load() {
    if (h) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { why = std::string("NCCL is not loadable (") + dlerror() + ")"; return false; }
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) why = std::string("NCCL symbol missing: ") + n; return p; };
    GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) { h = nullptr; return false; }
    return true;
  }
};
Nccl& nccl() { static Nccl n; return n; }
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(300 + (int)r_, std::string(#call) + ": " + nccl().GetErrorString(r_)); } while (0)

__global__ void sum_into_kernel(double* dst, const double* src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// maxlag < 0: the QSB_FLAG_MOMENTS layout (12 + 3 doubles per dimension)
size_t diag_len(int retdim, int maxlag) { return maxlag < 0 ? (size_t)retdim * 15 : (size_t)retdim * 12 + (maxlag > 0 ? (size_t)retdim * (maxlag + 1) : 0); }

// The header entries [0] half-chains, [1] draws per half-chain, [7] chains, [8] draws per chain of every dimension are
// written identically by every shard: after a sum over shards [1] and [8] must be divided back.
void fix_header_after_sum(double* h, int retdim, int nshards, int maxlag) {
  for (int j = 0; j < retdim; ++j) { h[(size_t)j * 12 + 1] /= nshards; h[(size_t)j * 12 + 8] /= nshards; }
  if (maxlag < 0) for (int j = 0; j < retdim; ++j) { h[(size_t)retdim * 12 + 3 * j + 1] /= nshards; h[(size_t)retdim * 12 + 3 * j + 2] /= nshards; }
}

}  // namespace

struct qsb_comm {
  ncclComm_t comm = nullptr;
  int device = 0, rank = 0, nranks = 1;
};

struct qsb_group {
  std::vector<int> devices;
  std::vector<qsb_model*> models;
  std::vector<qsb_sampler*> samplers;
  std::vector<cudaStream_t> streams;
  std::vector<uint32_t> begin, count;       // chains of shard i: [begin, begin + count) in group-local numbering
  std::vector<uint32_t> stored;             // leading chains of shard i whose samples are kept
  std::vector<ncclComm_t> comms;            // empty: shards share devices, reduced by sum_into_kernel
  std::vector<double*> dbuf; size_t dbuf_len = 0;
  int retdim = 0;
};

extern "C" {

int32_t qsb_group_size(const qsb_group* g) { return g ? (int32_t)g->samplers.size() : 0; }
qsb_sampler* qsb_group_sampler(qsb_group* g, int32_t i) { return (g && i >= 0 && i < (int32_t)g->samplers.size()) ? g->samplers[i] : nullptr; }

int qsb_group_destroy(qsb_group* g) {
  if (!g) return 0;
  for (size_t i = 0; i < g->samplers.size(); ++i) {
    cudaSetDevice(g->devices[i]);
    if (g->samplers[i]) qsb_sampler_destroy(g->samplers[i]);
    if (i < g->models.size() && g->models[i]) qsb_model_destroy(g->models[i]);
    if (i < g->dbuf.size() && g->dbuf[i]) cudaFree(g->dbuf[i]);
  }
  for (ncclComm_t c : g->comms) if (c) nccl().CommDestroy(c);
  for (size_t i = 0; i < g->streams.size(); ++i) { cudaSetDevice(g->devices[i]); if (g->streams[i]) cudaStreamDestroy(g->streams[i]); }
  delete g;
  return 0;
}

int qsb_group_create(const qsb_model_desc* desc, const qsb_kernel_spec* specs, int32_t nspecs, const qsb_run_config* cfg,
                     const int32_t* devices, int32_t ndev, qsb_group** out) {
  if (!desc || !specs || !cfg || !devices || !out) return fail(1, "qsb_group_create: null argument");
  *out = nullptr;
  if (ndev < 1 || ndev > 64) return fail(2, "qsb_group_create: 1..64 devices");
  if (cfg->numchains < (uint32_t)ndev) return fail(2, "qsb_group_create: fewer chains than devices");
  int visible = 0;
  CU(cudaGetDeviceCount(&visible));
  for (int i = 0; i < ndev; ++i) if (devices[i] < 0 || devices[i] >= visible) return fail(2, "qsb_group_create: device ordinal out of range (visible devices: " + std::to_string(visible) + ")");
  qsb_group* g = new qsb_group();
  g->devices.assign(devices, devices + ndev);
  g->models.assign(ndev, nullptr); g->samplers.assign(ndev, nullptr); g->streams.assign(ndev, nullptr); g->dbuf.assign(ndev, nullptr);
  auto bail = [&](int rc) { std::string e = qsb_last_error(); qsb_group_destroy(g); return fail(rc, e); };
  // block partition: the first (numchains % ndev) shards own one chain more
  const uint32_t base = cfg->numchains / ndev, rem = cfg->numchains % ndev;
  const uint32_t keep_total = (cfg->sample_chains == 0 || cfg->sample_chains > cfg->numchains) ? cfg->numchains : cfg->sample_chains;
  uint32_t at = 0;
  for (int i = 0; i < ndev; ++i) {
    const uint32_t n = base + ((uint32_t)i < rem ? 1u : 0u);
    g->begin.push_back(at); g->count.push_back(n);
    // kept samples: all chains, or the leading ceil(sample_chains / ndev) chains of every shard (a spread subset of the chains)
    g->stored.push_back(keep_total == cfg->numchains ? n : std::min<uint32_t>(n, (keep_total + ndev - 1) / ndev));
    at += n;
  }
  for (int i = 0; i < ndev; ++i) {
    if (cudaSetDevice(devices[i]) != cudaSuccess) { qsb_set_error_("cudaSetDevice failed"); return bail(4); }
    if (cudaStreamCreateWithFlags(&g->streams[i], cudaStreamNonBlocking) != cudaSuccess) { qsb_set_error_("cudaStreamCreate failed"); return bail(4); }
    if (int rc = qsb_model_create(desc, devices[i], &g->models[i])) return bail(rc);
    qsb_run_config c = *cfg;
    c.chain_begin = cfg->chain_begin + g->begin[i]; c.numchains = g->count[i]; c.sample_chains = g->stored[i];
    c.device = devices[i]; c.stream = (void*)g->streams[i];
    if (int rc = qsb_sampler_create(g->models[i], specs, nspecs, &c, &g->samplers[i])) return bail(rc);
  }
  g->retdim = qsb_model_retdim(g->models[0]);
  bool distinct = true;
  for (int i = 0; i < ndev; ++i) for (int j = 0; j < i; ++j) if (devices[i] == devices[j]) distinct = false;
  if (ndev > 1 && distinct) {
    if (!nccl().load()) { qsb_set_error_(nccl().why.c_str()); return bail(30); }
    g->comms.assign(ndev, nullptr);
    ncclResult_t r = nccl().CommInitAll(g->comms.data(), ndev, g->devices.data());
    if (r != ncclSuccess) { qsb_set_error_((std::string("ncclCommInitAll: ") + nccl().GetErrorString(r)).c_str()); g->comms.clear(); return bail(300 + (int)r); }
  }
  *out = g;
  return 0;
}

#define FOR_SHARDS(expr) do { for (size_t i_ = 0; i_ < g->samplers.size(); ++i_) { qsb_sampler* s = g->samplers[i_]; (void)s; if (int rc_ = (expr)) return rc_; } } while (0)

int qsb_group_init(qsb_group* g) { if (!g) return fail(1, "null group"); FOR_SHARDS(qsb_sampler_init(s)); return 0; }
int qsb_group_begin(qsb_group* g, uint64_t max_samples, uint64_t burnin, uint64_t lag, uint64_t numiters) {
  if (!g) return fail(1, "null group");
  FOR_SHARDS(qsb_sampler_begin(s, max_samples, burnin, lag, numiters));
  return 0;
}
// every shard's launches are asynchronous on its own stream: the devices run concurrently, the host thread only enqueues
int qsb_group_advance(qsb_group* g, uint64_t iters) { if (!g) return fail(1, "null group"); FOR_SHARDS(qsb_sampler_advance(s, iters)); return 0; }
int qsb_group_sync(qsb_group* g) {
  if (!g) return fail(1, "null group");
  for (size_t i = 0; i < g->samplers.size(); ++i) { CU(cudaSetDevice(g->devices[i])); CU(cudaStreamSynchronize(g->streams[i])); }
  return 0;
}
int qsb_group_run(qsb_group* g, uint64_t nsamps, uint64_t burnin, uint64_t lag) {
  if (!g) return fail(1, "null group");
  if (lag == 0) return fail(2, "lag must be >= 1");
  const uint64_t iters = burnin + nsamps * lag;   // mcmc.t:53
  if (int rc = qsb_group_begin(g, nsamps + 1, burnin, lag, iters)) return rc;
  if (int rc = qsb_group_advance(g, iters)) return rc;
  if (int rc = qsb_group_sync(g)) return rc;
  qsb_stats st;
  FOR_SHARDS(qsb_sampler_stats(s, &st));   // surfaces sticky device-side errors of every shard
  return 0;
}

int qsb_group_fetch_samples(qsb_group* g, uint32_t c0, uint32_t c1, uint64_t s0, uint64_t s1, void* dst, size_t stride_bytes) {
  if (!g || !dst) return fail(1, "null argument");
  if (c0 > c1 || s0 > s1) return fail(2, "fetch: bad range");
  // group-local chain c lives in the shard whose block contains it; only a shard's leading `stored` chains hold samples
  const uint64_t ns = s1 - s0;
  for (size_t i = 0; i < g->samplers.size(); ++i) {
    const uint32_t b = g->begin[i], e = b + g->count[i];
    const uint32_t lo = std::max(c0, b), hi = std::min(c1, e);
    if (lo >= hi) continue;
    if (hi - b > g->stored[i]) return fail(2, "fetch: chain " + std::to_string(b + g->stored[i]) + " keeps no samples (sample_chains)");
    if (int rc = qsb_sampler_fetch_samples(g->samplers[i], lo - b, hi - b, s0, s1, (char*)dst + (size_t)(lo - c0) * ns * stride_bytes, stride_bytes)) return rc;
  }
  return 0;
}

int qsb_group_stats(qsb_group* g, qsb_stats* out) {
  if (!g || !out) return fail(1, "null argument");
  memset(out, 0, sizeof(*out));
  for (size_t i = 0; i < g->samplers.size(); ++i) {
    qsb_stats st;
    if (int rc = qsb_sampler_stats(g->samplers[i], &st)) return rc;
    out->props_made += st.props_made; out->props_accepted += st.props_accepted;
    for (int k = 0; k < 8; ++k) { out->sub_made[k] += st.sub_made[k]; out->sub_accepted[k] += st.sub_accepted[k]; }
    out->grad_evals += st.grad_evals; out->leapfrog_steps += st.leapfrog_steps; out->kernel_launches += st.kernel_launches;
    out->nonfinite += st.nonfinite; out->divergent += st.divergent;
    out->iterations = st.iterations;
    out->seconds = std::max(out->seconds, st.seconds);
  }
  return 0;
}

int qsb_group_diagnostics(qsb_group* g, int32_t maxlag, double* dst) {
  if (!g || !dst) return fail(1, "null argument");
  const size_t len = diag_len(g->retdim, maxlag);
  const int n = (int)g->samplers.size();
  if (g->dbuf_len < len) {
    for (int i = 0; i < n; ++i) { CU(cudaSetDevice(g->devices[i])); if (g->dbuf[i]) cudaFree(g->dbuf[i]); g->dbuf[i] = nullptr; CU(cudaMalloc((void**)&g->dbuf[i], len * 8)); }
    g->dbuf_len = len;
  }
  for (int i = 0; i < n; ++i) if (int rc = qsb_sampler_diagnostics(g->samplers[i], maxlag, g->dbuf[i], 1)) return rc;
  if (!g->comms.empty()) {
    NC(nccl().GroupStart());
    for (int i = 0; i < n; ++i) {
      ncclResult_t r = nccl().AllReduce(g->dbuf[i], g->dbuf[i], len, ncclDouble, ncclSum, g->comms[i], g->streams[i]);
      if (r != ncclSuccess) { nccl().GroupEnd(); return fail(300 + (int)r, std::string("ncclAllReduce: ") + nccl().GetErrorString(r)); }
    }
    NC(nccl().GroupEnd());
  } else if (n > 1) {   // shards share devices: fold everything into shard 0's buffer in shard order
    CU(cudaSetDevice(g->devices[0]));
    double* tmp = nullptr;
    CU(cudaMalloc((void**)&tmp, len * 8));
    for (int i = 1; i < n; ++i) {
      CU(cudaSetDevice(g->devices[i])); CU(cudaStreamSynchronize(g->streams[i]));
      CU(cudaSetDevice(g->devices[0]));
      cudaError_t e = cudaMemcpyPeerAsync(tmp, g->devices[0], g->dbuf[i], g->devices[i], len * 8, g->streams[0]);
      if (e == cudaSuccess) { sum_into_kernel<<<(unsigned)std::min<size_t>((len + 255) / 256, 1024), 256, 0, g->streams[0]>>>(g->dbuf[0], tmp, len); e = cudaGetLastError(); }
      if (e != cudaSuccess) { cudaFree(tmp); return fail(100 + (int)e, cudaGetErrorString(e)); }
    }
    cudaError_t e = cudaStreamSynchronize(g->streams[0]);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(100 + (int)e, cudaGetErrorString(e));
  }
  CU(cudaSetDevice(g->devices[0]));
  CU(cudaMemcpyAsync(dst, g->dbuf[0], len * 8, cudaMemcpyDeviceToHost, g->streams[0]));
  CU(cudaStreamSynchronize(g->streams[0]));
  if (n > 1) fix_header_after_sum(dst, g->retdim, n, maxlag);
  if (!g->comms.empty()) return qsb_group_sync(g);
  return 0;
}

// ---- one process per device
int qsb_comm_unique_id(uint8_t id[QSB_COMM_ID_BYTES]) {
  if (!id) return fail(1, "null argument");
  if (!nccl().load()) return fail(30, nccl().why);
  static_assert(sizeof(ncclUniqueId) == QSB_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  NC(nccl().GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return 0;
}
int qsb_comm_create(const uint8_t id[QSB_COMM_ID_BYTES], int32_t rank, int32_t nranks, int32_t device, qsb_comm** out) {
  if (!id || !out) return fail(1, "null argument");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(2, "qsb_comm_create: bad rank");
  if (!nccl().load()) return fail(30, nccl().why);
  if (device >= 0) CU(cudaSetDevice(device)); else CU(cudaGetDevice(&device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  qsb_comm* c = new qsb_comm();
  c->device = device; c->rank = rank; c->nranks = nranks;
  ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, u, rank);
  if (r != ncclSuccess) { delete c; return fail(300 + (int)r, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); }
  *out = c;
  return 0;
}
int qsb_comm_destroy(qsb_comm* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return 0;
}
int qsb_comm_allreduce_sum(qsb_comm* c, double* buf, size_t n, void* stream) {
  if (!c || !buf) return fail(1, "null argument");
  CU(cudaSetDevice(c->device));
  NC(nccl().AllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, (cudaStream_t)stream));
  return 0;
}
int qsb_sampler_diagnostics_allreduce(qsb_sampler* s, qsb_comm* c, int32_t maxlag, double* dst) {
  if (!s || !c || !dst) return fail(1, "null argument");
  CU(cudaSetDevice(c->device));
  // retdim from a zero-lag probe of the layout is not available here: ask the sampler through the public entry
  int retdim = qsb_sampler_retdim_(s);
  const size_t len = diag_len(retdim, maxlag);
  double* d = nullptr;
  CU(cudaMalloc((void**)&d, len * 8));
  int rc = qsb_sampler_diagnostics(s, maxlag, d, 1);
  cudaStream_t st = (cudaStream_t)qsb_sampler_stream_(s);
  if (!rc) rc = qsb_comm_allreduce_sum(c, d, len, (void*)st);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(dst, d, len * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(100 + (int)e, cudaGetErrorString(e));
  }
  cudaFree(d);
  if (!rc && c->nranks > 1) fix_header_after_sum(dst, retdim, c->nranks, maxlag);
  return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// FP64 peak of this device, measured in-process (roofline denominators: MEASURED_PEAKS.json has no fp64 figure).
namespace {
template <int ILP> __global__ void peak_dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}
template <int ILP> __global__ void peak_dmma_kernel(double* out, int iters) {
  double d[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { d[i][0] = i; d[i][1] = threadIdx.x; }
  const double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}
}  // namespace

extern "C" int qsb_measure_fp64_peak(int32_t which, double seconds, int32_t device, double* tflops) {
  if (!tflops || (which != 0 && which != 1)) return fail(1, "bad argument");
  if (device >= 0) CU(cudaSetDevice(device)); else CU(cudaGetDevice(&device));
  int nsm = 0;
  CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  double* out = nullptr;
  CU(cudaMalloc((void**)&out, 8));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  constexpr int ILP = 8, TPB = 256;
  const int grid = nsm * 8, iters = 20000;
  // flop per launch: DFMA 2 per thread-op; DMMA m8n8k4 = 8*8*4*2 = 512 per warp-op
  const double flop = which == 0 ? 2.0 * ILP * (double)iters * TPB * grid : 512.0 * ILP * (double)iters * (TPB / 32) * grid;
  auto launch = [&]() { if (which == 0) peak_dfma_kernel<ILP><<<grid, TPB>>>(out, iters, 0.999999, 1e-9); else peak_dmma_kernel<ILP><<<grid, TPB>>>(out, iters); };
  launch(); launch();
  cudaError_t e = cudaDeviceSynchronize();
  double best = 0.0, spent = 0.0;
  for (int rep = 0; e == cudaSuccess && rep < 200 && (rep < 3 || spent < seconds); ++rep) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess && ms > 0.f) { best = std::max(best, flop / (ms * 1e-3) / 1e12); spent += ms * 1e-3; }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  if (e != cudaSuccess) return fail(100 + (int)e, cudaGetErrorString(e));
  *tflops = best;
  return 0;
}
