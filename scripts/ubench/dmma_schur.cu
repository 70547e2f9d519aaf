// Micro-benchmark: the Schur rank-k update of the dense QP kernel,
//   C (n×n) += Aᵀ diag(d) A,   A is ny×n  (n = ny = 100: /root/reference/benchmark/quadratic_program_benchmark.jl:7-11)
// once with register-tiled DFMA (the layout of dense kernel v3: 512 threads, lane ↔ rows {l, l+32, …}, warp ↔ columns
// {w, w+16, …}, a 4×7 tile per thread) and once with the FP64 tensor-core instruction mma.sync.aligned.m8n8k4.f64
// (16 warps, 8×8 output tiles, operands from the same shared-memory copy of A).  One CTA per instance, persistent over a
// batch; reports time per instance, FP64 flops/s and the fraction of the measured DFMA peak.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_schur dmma_schur.cu && ./dmma_schur
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int N = 100, NY = 100, NP = 104;       // NP: n padded to a multiple of 8
constexpr int LDA = NP + 1;                      // shared-memory row stride of A (odd: conflict-free column access)
constexpr int THREADS = 512;

__global__ void __launch_bounds__(THREADS, 1) schur_fma(const double* __restrict__ A, const double* __restrict__ d, double* __restrict__ C, int B, int reps) {
  extern __shared__ double sm[];
  double* As = sm;            // NY × LDA
  double* ds = sm + NY * LDA;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < NY * NP; i += THREADS) As[(i / NP) * LDA + i % NP] = (i % NP < N) ? A[(size_t)b * NY * N + (i / NP) * N + i % NP] : 0.0;
    for (int i = threadIdx.x; i < NY; i += THREADS) ds[i] = d[(size_t)b * NY + i];
    __syncthreads();
    double acc[4][7];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 7; ++c) acc[r][c] = 0.0;
    for (int rep = 0; rep < reps; ++rep)
#pragma unroll 2
      for (int k = 0; k < NY; ++k) {
        const double dk = ds[k];
        double ar[4], ac[7];
#pragma unroll
        for (int r = 0; r < 4; ++r) ar[r] = (lane + 32 * r < NP) ? As[k * LDA + lane + 32 * r] * dk : 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) ac[c] = (w + 16 * c < NP) ? As[k * LDA + w + 16 * c] : 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 7; ++c) acc[r][c] = fma(ar[r], ac[c], acc[r][c]);
      }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 7; ++c)
        if (lane + 32 * r < N && w + 16 * c < N) C[(size_t)b * N * N + (lane + 32 * r) * N + w + 16 * c] = acc[r][c];
  }
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 13×13 output tiles of 8×8; warp w owns tile rows {w} (w < 13) — 13 tiles each — so its A-operand fragment
// (rows 8w…8w+7 of (D A)ᵀ, 4 values of k) is loaded once per k-step and reused over the 13 column tiles.
__global__ void __launch_bounds__(THREADS, 1) schur_dmma(const double* __restrict__ A, const double* __restrict__ d, double* __restrict__ C, int B, int reps) {
  extern __shared__ double sm[];
  double* As = sm;
  double* ds = sm + NY * LDA;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;          // fragment coordinates: A[row g][k t], B[k t][col g], C[row g][cols 2t, 2t+1]
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < NY * NP; i += THREADS) As[(i / NP) * LDA + i % NP] = (i % NP < N) ? A[(size_t)b * NY * N + (i / NP) * N + i % NP] : 0.0;
    for (int i = threadIdx.x; i < NY; i += THREADS) ds[i] = d[(size_t)b * NY + i];
    __syncthreads();
    if (w < 13) {
      double c[13][2];
#pragma unroll
      for (int j = 0; j < 13; ++j) c[j][0] = c[j][1] = 0.0;
      for (int rep = 0; rep < reps; ++rep)
#pragma unroll 1
        for (int k0 = 0; k0 < NY; k0 += 4) {
          const double a = As[(k0 + t) * LDA + 8 * w + g] * ds[k0 + t];   // (D A)ᵀ[8w+g][k0+t]
#pragma unroll
          for (int j = 0; j < 13; ++j) dmma(c[j][0], c[j][1], a, As[(k0 + t) * LDA + 8 * j + g]);
        }
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int row = 8 * w + g, col = 8 * j + 2 * t;
        if (row < N && col < N) C[(size_t)b * N * N + row * N + col] = c[j][0];
        if (row < N && col + 1 < N) C[(size_t)b * N * N + row * N + col + 1] = c[j][1];
      }
    }
  }
}

__global__ void fma_peak(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int main() {
  const int B = 148 * 8, reps = 16;
  std::vector<double> hA((size_t)B * NY * N), hd((size_t)B * NY);
  srand(1);
  for (auto& v : hA) v = (rand() % 10 == 0) ? (rand() / (double)RAND_MAX - 0.5) : 0.0;
  for (auto& v : hd) v = 0.5 + rand() / (double)RAND_MAX;
  double *A, *d, *C1, *C2, *scratch;
  cudaMalloc(&A, hA.size() * 8); cudaMalloc(&d, hd.size() * 8);
  cudaMalloc(&C1, (size_t)B * N * N * 8); cudaMalloc(&C2, (size_t)B * N * N * 8); cudaMalloc(&scratch, 148 * 4 * 256 * 8);
  cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d, hd.data(), hd.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = (NY * LDA + NY) * 8;
  cudaFuncSetAttribute(schur_fma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(schur_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  // DFMA peak
  fma_peak<<<148 * 4, 256>>>(scratch, 1 << 16);
  cudaEventRecord(e0); fma_peak<<<148 * 4, 256>>>(scratch, 1 << 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double peak = 148.0 * 4 * 256 * 8 * 2 * (1 << 16) / (ms * 1e-3) / 1e12;
  printf("DFMA peak %.2f TFLOP/s\n", peak);
  const double flops = 2.0 * NP * NP * NY * reps * B;   // padded-tile flops actually executed
  const double useful = 2.0 * N * N * NY * reps * B;
  for (int v = 0; v < 2; ++v) {
    auto run = [&] { if (v == 0) schur_fma<<<148, THREADS, smem>>>(A, d, C1, B, reps); else schur_dmma<<<148, THREADS, smem>>>(A, d, C2, B, reps); };
    run(); cudaDeviceSynchronize();
    cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%s: %.3f ms for %d instances x %d accumulations: %.2f us per n=100 Schur product, %.2f TFLOP/s useful (%.1f %% of DFMA peak; %.2f TFLOP/s incl. padding)  err=%s\n",
           v == 0 ? "register-tiled DFMA (v3 layout)" : "mma.sync.m8n8k4.f64          ", ms, B, reps, ms * 1e3 / (B * reps), useful / (ms * 1e-3) / 1e12,
           100 * useful / (ms * 1e-3) / 1e12 / peak, flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  }
  std::vector<double> h1((size_t)N * N), h2((size_t)N * N);
  cudaMemcpy(h1.data(), C1, h1.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2.data(), C2, h2.size() * 8, cudaMemcpyDeviceToHost);
  double maxd = 0, maxv = 0;
  for (size_t i = 0; i < h1.size(); ++i) { maxd = fmax(maxd, fabs(h1[i] - h2[i])); maxv = fmax(maxv, fabs(h1[i])); }
  printf("max |C_fma - C_dmma| = %.3e (max |C| %.3e)\n", maxd, maxv);
  return 0;
}
