// Dependent-issue latencies on sm_100a that the LU kernels' critical paths are made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double seed, int nwarps_sync) {
  __shared__ double sm[1024];
  __shared__ unsigned smax;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double a = seed + lane * 1e-9, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  int slot = 0;
#define REC() if (threadIdx.x == 0) cyc[slot] = (t1 - t0); ++slot;
  // 0: dependent DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = fma(a, b, c);
  t1 = clock64(); REC();
  // 1: dependent DMUL
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = a * b;
  t1 = clock64(); REC();
  // 2: dependent DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = a + c;
  t1 = clock64(); REC();
  // 3: dependent full reciprocal 1.0/x
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) a = 1.0 / a;
  t1 = clock64(); REC();
  // 4: dependent shfl of a double
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
  t1 = clock64(); REC();
  // 5: dependent redux max (u32)
  unsigned u = (unsigned)__double2hiint(a) + lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) u = __reduce_max_sync(0xffffffffu, u) + lane;
  t1 = clock64(); REC();
  // 6: shared store -> syncwarp -> broadcast load round trip
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    if (lane == (i & 31)) sm[wid * 32 + (i & 7)] = a;
    __syncwarp();
    a += sm[wid * 32 + (i & 7)];
    __syncwarp();
  }
  t1 = clock64(); REC();
  // 7: __syncthreads with all warps of the CTA, nothing else
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64(); REC();
  // 8: one warp works (DFMA x4 dependent), the others wait at the barrier, then all do a shared load
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    if (wid == (i & 15)) {
      a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
      if (lane == 0) sm[512 + (i & 1)] = a;
    }
    __syncthreads();
    a += sm[512 + (i & 1)];
  }
  t1 = clock64(); REC();
  // 9: independent DFMA throughput, 8 accumulators per thread, all warps
  double r0 = a, r1 = a + 1, r2 = a + 2, r3 = a + 3, r4 = a + 4, r5 = a + 5, r6 = a + 6, r7 = a + 7;
  __syncthreads();
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    r0 = fma(r0, b, c); r1 = fma(r1, b, c); r2 = fma(r2, b, c); r3 = fma(r3, b, c);
    r4 = fma(r4, b, c); r5 = fma(r5, b, c); r6 = fma(r6, b, c); r7 = fma(r7, b, c);
  }
  t1 = clock64(); REC();
  // 10: atomicMax on shared + barrier
  if (threadIdx.x == 0) smax = 0;
  __syncthreads();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    if (lane == 0) atomicMax(&smax, u + i);
    __syncthreads();
    u += smax;
  }
  t1 = clock64(); REC();
  out[threadIdx.x] = a + u + r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
  const char* names[] = {"DFMA dep", "DMUL dep", "DADD dep", "1.0/x dep", "shfl.f64 dep", "redux.max dep", "STS->syncwarp->LDS", "__syncthreads only",
                         "1 warp 4xDFMA + bar + LDS", "8 indep DFMA/thread (per 8)", "atomicMax.shared + bar + LDS"};
  for (int threads : {32, 512}) {
    k<<<1, threads>>>(out, cyc, 1.25, threads / 32);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("threads=%d  (%s)\n", threads, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 11; ++i) printf("  %-32s %8.1f cycles/iter\n", names[i], (double)h[i] / N);
  }
  return 0;
}
