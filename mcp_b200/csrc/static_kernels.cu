// static_kernels.cu — the few kernels that do not depend on a problem description, compiled ahead of
// time by nvcc for sm_100a (the problem-specialised kernels go through NVRTC, see mcpb200.cpp).
//
//   fp64_fma_peak_kernel : chains of dependent-free DFMAs; measures the FP64 pipe's achievable TFLOP/s,
//                          the roofline denominator of the KKT factorisation (MEASURED_PEAKS.json has no
//                          FP64 entry; BASELINE.md §4 asks the builder to measure one).
//   flush_l2_kernel      : streams a buffer larger than the 126 MB L2 so timed iterations start cold.
#include <cuda_runtime.h>

#include <cstdio>

namespace {

constexpr int kChains = 8;
constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) fp64_fma_peak_kernel(double* out, double a, double b) {
  double acc[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

__global__ void flush_l2_kernel(double2* buf, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) buf[i] = make_double2(v, v);
}

double2* g_flush_buf[64] = {nullptr};
constexpr size_t kFlushBytes = 256ull << 20;

}  // namespace

#define TRY(call)                                                         \
  do {                                                                    \
    cudaError_t e_ = (call);                                              \
    if (e_ != cudaSuccess) {                                              \
      snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e_));     \
      cudaGetLastError();                                                 \
      return 1;                                                           \
    }                                                                     \
  } while (0)

extern "C" int mcpb200_static_fp64_peak(double* tflops_out, char* err, int errlen) {
  int dev = 0, sms = 0;
  TRY(cudaGetDevice(&dev));
  TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* out = nullptr;
  TRY(cudaMalloc(&out, 8));
  cudaEvent_t e0, e1;
  TRY(cudaEventCreate(&e0));
  TRY(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {  // first reps are warm-up
    TRY(cudaEventRecord(e0));
    fp64_fma_peak_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
    TRY(cudaEventRecord(e1));
    TRY(cudaEventSynchronize(e1));
    float ms = 0;
    TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * kChains * kIters * (double)blocks * threads;
    if (rep >= 2) best = best > flops / (ms * 1e-3) / 1e12 ? best : flops / (ms * 1e-3) / 1e12;
  }
  TRY(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops_out = best;
  return 0;
}

extern "C" int mcpb200_static_flush_l2(void* stream, char* err, int errlen) {
  int dev = 0;
  TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!g_flush_buf[dev]) TRY(cudaMalloc(&g_flush_buf[dev], kFlushBytes));
  flush_l2_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(g_flush_buf[dev], kFlushBytes / sizeof(double2), 1.0);
  TRY(cudaGetLastError());
  return 0;
}
