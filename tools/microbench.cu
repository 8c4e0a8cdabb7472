// FP64 / DMMA / transcendental / Philox / HBM micro-benchmarks for B200 (sm_100a).
//
// MEASURED_PEAKS.json has no FP64 figure, and every roofline in this repo is an
// FP64 roofline, so the denominators are measured here.
// Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/microbench tools/microbench.cu
// Run  :  tools/microbench > gpurun_out/microbench.json
//
// Each kernel runs a dependent chain per accumulator with enough independent
// accumulators per thread to cover the pipe latency; the launch covers every SM
// with several resident CTAs. Timing: CUDA events, 3 warm-up launches, best of 5.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

// ---------------------------------------------------------------- DFMA
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
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

// ---------------------------------------------------------------- DMMA
__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
               "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dmma884(double* out, int iters) {
  double d[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { d[i][0] = i; d[i][1] = threadIdx.x; }
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
  if (s == 123.456) out[0] = s;
}

template <int ILP, int KK>
__global__ void k_dmma16(double* out, int iters) {
  double d[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { d[i][0] = i; d[i][1] = threadIdx.x; d[i][2] = 1; d[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - threadIdx.x * 1e-3 * i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KK == 4) { double a2[2] = {a[0], a[1]}; mma1684(d[i], a2, b[0]); }
      else if (KK == 8) { double a4[4] = {a[0], a[1], a[2], a[3]}; double b2[2] = {b[0], b[1]}; mma1688(d[i], a4, b2); }
      else { mma16816(d[i], a, b); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456) out[0] = s;
}

// Mixed: every warp issues DMMA (m16n8k8) and independent DFMA chains.
template <int ILP_M, int ILP_F>
__global__ void k_mixed(double* out, int iters, double fa, double fb) {
  double d[ILP_M][4];
#pragma unroll
  for (int i = 0; i < ILP_M; ++i) { d[i][0] = i; d[i][1] = threadIdx.x; d[i][2] = 1; d[i][3] = 2; }
  double a4[4], b2[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a4[i] = threadIdx.x * 1e-3 + i;
  b2[0] = 0.5; b2[1] = 0.25;
  double acc[ILP_F];
#pragma unroll
  for (int i = 0; i < ILP_F; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP_M; ++i) mma1688(d[i], a4, b2);
#pragma unroll
    for (int i = 0; i < ILP_F; ++i) acc[i] = fma(acc[i], fa, fb);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP_M; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
#pragma unroll
  for (int i = 0; i < ILP_F; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// Warp-specialised mix: even warps DMMA, odd warps DFMA.
__global__ void k_mixed_ws(double* out, int iters, double fa, double fb) {
  int w = threadIdx.x >> 5;
  double s = 0;
  if (w & 1) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], fa, fb);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], fa, fb);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
  } else {
    double d[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { d[i][0] = i; d[i][1] = threadIdx.x; d[i][2] = 1; d[i][3] = 2; }
    double a4[4] = {1, 2, 3, threadIdx.x * 1e-3}, b2[2] = {0.5, 0.25};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mma1688(d[i], a4, b2);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  }
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- transcendental rates
template <int OP>
__global__ void k_math(double* out, int iters, double seed) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = seed + (threadIdx.x + i * 37) * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (OP == 0) x[i] = exp(-x[i]) + 0.5;            // exp
      else if (OP == 1) x[i] = log(x[i]) + 2.5;        // log
      else if (OP == 2) x[i] = 1.0 / (1.0 + x[i]);     // div
      else if (OP == 3) x[i] = sqrt(x[i]) + 0.5;       // sqrt
      else if (OP == 4) x[i] = 1.0 / (1.0 + exp(-x[i])); // logistic
      else if (OP == 5) x[i] = log1p(exp(-x[i])) + 0.3;  // softplus(-x)
    }
  }
  double s = x[0] + x[1] + x[2] + x[3];
  if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void k_philox(uint32_t* out, int iters) {
  uint32_t acc = 0;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    uint32_t c[4] = {(uint32_t)it, tid, 7u, 9u};
    uint32_t k0 = 0x1234u, k1 = 0x5678u;
#pragma unroll
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    acc ^= c[0] ^ c[1] ^ c[2] ^ c[3];
  }
  if (acc == 0x12345u) out[0] = acc;
}

// ---------------------------------------------------------------- HBM copy
__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

// ---------------------------------------------------------------- DMMA layout check (m16n8k8 and m16n8k4, m16n8k16)
// A: 16 x K row-major, B: K x 8 (given as Bt[n][k]), C = A*B, one warp.
template <int KK>
__global__ void k_layout(const double* A, const double* Bt, double* C) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double d[4] = {0, 0, 0, 0};
  if (KK == 4) {
    double a[2] = {A[g * KK + t], A[(g + 8) * KK + t]};
    mma1684(d, a, Bt[g * KK + t]);
  } else if (KK == 8) {
    double a[4] = {A[g * KK + t], A[(g + 8) * KK + t], A[g * KK + t + 4], A[(g + 8) * KK + t + 4]};
    double b[2] = {Bt[g * KK + t], Bt[g * KK + t + 4]};
    mma1688(d, a, b);
  } else {
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[2 * i] = A[g * KK + t + 4 * i]; a[2 * i + 1] = A[(g + 8) * KK + t + 4 * i]; b[i] = Bt[g * KK + t + 4 * i]; }
    mma16816(d, a, b);
  }
  C[g * 8 + 2 * t] = d[0]; C[g * 8 + 2 * t + 1] = d[1];
  C[(g + 8) * 8 + 2 * t] = d[2]; C[(g + 8) * 8 + 2 * t + 1] = d[3];
}

template <int KK>
double layout_err() {
  std::vector<double> A(16 * KK), Bt(8 * KK), C(128), R(128, 0.0);
  for (int i = 0; i < 16 * KK; ++i) A[i] = sin(0.37 * i + 1.0);
  for (int i = 0; i < 8 * KK; ++i) Bt[i] = cos(0.11 * i + 0.3);
  for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) for (int k = 0; k < KK; ++k) R[m * 8 + n] += A[m * KK + k] * Bt[n * KK + k];
  double *dA, *dB, *dC;
  CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, Bt.size() * 8)); CK(cudaMalloc(&dC, 128 * 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bt.data(), Bt.size() * 8, cudaMemcpyHostToDevice));
  k_layout<KK><<<1, 32>>>(dA, dB, dC);
  CK(cudaMemcpy(C.data(), dC, 128 * 8, cudaMemcpyDeviceToHost));
  double e = 0; for (int i = 0; i < 128; ++i) e = fmax(e, fabs(C[i] - R[i]));
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return e;
}

// ---------------------------------------------------------------- harness
template <class F>
float best_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 1024));
  printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\",\n", prop.name, sms, prop.major, prop.minor);

  const int iters = 4096;
  // DFMA: threads x ILP x iters FMAs = 2 flops each
  for (int tpb : {256, 512, 1024}) {
    int blocks = sms * (2048 / tpb);
    float ms = best_ms([&] { k_dfma<8><<<blocks, tpb>>>(dout, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * blocks * tpb * 8.0 * iters;
    printf(" \"dfma_tflops_tpb%d\": %.3f,\n", tpb, fl / ms * 1e-9);
  }
  {
    int tpb = 256, blocks = sms * 4;
    float ms = best_ms([&] { k_dfma<16><<<blocks, tpb>>>(dout, iters, 1.0000001, 1e-9); });
    printf(" \"dfma_tflops_ilp16_occ1024\": %.3f,\n", 2.0 * blocks * tpb * 16.0 * iters / ms * 1e-9);
  }
  // DMMA
  for (int tpb : {256, 512}) {
    int blocks = sms * (1024 / tpb);
    double warps = (double)blocks * tpb / 32;
    float ms = best_ms([&] { k_dmma884<8><<<blocks, tpb>>>(dout, iters); });
    printf(" \"dmma_m8n8k4_tflops_tpb%d\": %.3f,\n", tpb, 2.0 * 8 * 8 * 4 * warps * 8 * iters / ms * 1e-9);
    ms = best_ms([&] { k_dmma16<4, 4><<<blocks, tpb>>>(dout, iters); });
    printf(" \"dmma_m16n8k4_tflops_tpb%d\": %.3f,\n", tpb, 2.0 * 16 * 8 * 4 * warps * 4 * iters / ms * 1e-9);
    ms = best_ms([&] { k_dmma16<4, 8><<<blocks, tpb>>>(dout, iters); });
    printf(" \"dmma_m16n8k8_tflops_tpb%d\": %.3f,\n", tpb, 2.0 * 16 * 8 * 8 * warps * 4 * iters / ms * 1e-9);
    ms = best_ms([&] { k_dmma16<4, 16><<<blocks, tpb>>>(dout, iters / 2); });
    printf(" \"dmma_m16n8k16_tflops_tpb%d\": %.3f,\n", tpb, 2.0 * 16 * 8 * 16 * warps * 4 * (iters / 2) / ms * 1e-9);
  }
  {
    int tpb = 256, blocks = sms * 2;
    double warps = (double)blocks * tpb / 32;
    float ms = best_ms([&] { k_dmma16<8, 8><<<blocks, tpb>>>(dout, iters); });
    printf(" \"dmma_m16n8k8_ilp8_tflops\": %.3f,\n", 2.0 * 16 * 8 * 8 * warps * 8 * iters / ms * 1e-9);
    ms = best_ms([&] { k_dmma16<1, 8><<<blocks, tpb>>>(dout, iters); });
    printf(" \"dmma_m16n8k8_ilp1_tflops\": %.3f,\n", 2.0 * 16 * 8 * 8 * warps * 1 * iters / ms * 1e-9);
    // one warp per SMSP, ILP 1: dependent-issue latency of one m16n8k8
    ms = best_ms([&] { k_dmma16<1, 8><<<sms, 128>>>(dout, iters); });
    printf(" \"dmma_m16n8k8_dep_latency_ns\": %.3f,\n", ms * 1e6 / iters);
    ms = best_ms([&] { k_dfma<1><<<sms, 128>>>(dout, iters, 1.0000001, 1e-9); });
    printf(" \"dfma_dep_latency_ns\": %.3f,\n", ms * 1e6 / iters);
  }
  // mixed
  {
    int tpb = 256, blocks = sms * 2;
    double warps = (double)blocks * tpb / 32, thr = (double)blocks * tpb;
    float ms = best_ms([&] { k_mixed<4, 8><<<blocks, tpb>>>(dout, iters, 1.0000001, 1e-9); });
    double fm = 2.0 * 16 * 8 * 8 * warps * 4 * iters, ff = 2.0 * thr * 8 * iters;
    printf(" \"mixed_same_warp_dmma_tflops\": %.3f, \"mixed_same_warp_dfma_tflops\": %.3f, \"mixed_same_warp_total_tflops\": %.3f,\n",
           fm / ms * 1e-9, ff / ms * 1e-9, (fm + ff) / ms * 1e-9);
    ms = best_ms([&] { k_mixed_ws<<<blocks, tpb>>>(dout, iters, 1.0000001, 1e-9); });
    fm = 2.0 * 16 * 8 * 8 * (warps / 2) * 4 * iters; ff = 2.0 * (thr / 2) * 16 * iters;
    printf(" \"mixed_ws_dmma_tflops\": %.3f, \"mixed_ws_dfma_tflops\": %.3f, \"mixed_ws_total_tflops\": %.3f,\n",
           fm / ms * 1e-9, ff / ms * 1e-9, (fm + ff) / ms * 1e-9);
  }
  // math
  {
    int tpb = 256, blocks = sms * 8, it2 = 1024;
    const char* names[6] = {"exp", "log", "div", "sqrt", "logistic", "softplus"};
    double n = (double)blocks * tpb * 4 * it2;
    float ms;
    ms = best_ms([&] { k_math<0><<<blocks, tpb>>>(dout, it2, 0.3); }); printf(" \"math_%s_gops\": %.3f,\n", names[0], n / ms * 1e-6);
    ms = best_ms([&] { k_math<1><<<blocks, tpb>>>(dout, it2, 1.3); }); printf(" \"math_%s_gops\": %.3f,\n", names[1], n / ms * 1e-6);
    ms = best_ms([&] { k_math<2><<<blocks, tpb>>>(dout, it2, 0.3); }); printf(" \"math_%s_gops\": %.3f,\n", names[2], n / ms * 1e-6);
    ms = best_ms([&] { k_math<3><<<blocks, tpb>>>(dout, it2, 0.3); }); printf(" \"math_%s_gops\": %.3f,\n", names[3], n / ms * 1e-6);
    ms = best_ms([&] { k_math<4><<<blocks, tpb>>>(dout, it2, 0.3); }); printf(" \"math_%s_gops\": %.3f,\n", names[4], n / ms * 1e-6);
    ms = best_ms([&] { k_math<5><<<blocks, tpb>>>(dout, it2, 0.3); }); printf(" \"math_%s_gops\": %.3f,\n", names[5], n / ms * 1e-6);
  }
  // philox
  {
    int tpb = 256, blocks = sms * 8, it2 = 2048;
    float ms = best_ms([&] { k_philox<<<blocks, tpb>>>((uint32_t*)dout, it2); });
    printf(" \"philox4x32_10_gblocks_per_s\": %.3f,\n", (double)blocks * tpb * it2 / ms * 1e-6);
  }
  // HBM copy
  {
    size_t n = (size_t)1 << 30;  // 1 GiB each way
    double2 *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n));
    CK(cudaMemset(a, 1, n)); CK(cudaMemset(b, 0, n));
    float ms = best_ms([&] { k_copy<<<sms * 8, 512>>>(a, b, n / 16); }, 10);
    printf(" \"hbm_copy_gbs\": %.1f,\n", 2.0 * n / ms * 1e-6);
    cudaFree(a); cudaFree(b);
  }
  printf(" \"layout_err_m16n8k4\": %.3e, \"layout_err_m16n8k8\": %.3e, \"layout_err_m16n8k16\": %.3e\n}\n",
         layout_err<4>(), layout_err<8>(), layout_err<16>());
  return 0;
}
