// Probe of the FP64 tensor-core instruction (mma.sync.m8n8k4.f64, SASS DMMA) on the B200:
//   (1) is D = A*B + C bit-identical to the chain fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))) (k ascending)?  or k descending? or neither?
//   (2) throughput: DMMAs per second per SM with 4 / 8 independent accumulators per warp, against DFMA with the same flops
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// one warp: A (8x4 row-major), B (4x8, element (k,n) at B[k*8+n]), C (8x8) -> D (8x8)
__global__ void k_check(const double* A, const double* B, const double* C, double* D) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  const double a = A[g * 4 + t];        // row g, col t
  const double b = B[t * 8 + g];        // row t (k), col g (n)
  double d0, d1;
  dmma(d0, d1, a, b, C[g * 8 + 2 * t], C[g * 8 + 2 * t + 1]);
  D[g * 8 + 2 * t] = d0;
  D[g * 8 + 2 * t + 1] = d1;
}

template <int NACC>
__global__ void k_dmma_rate(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-3, b = seed * 0.5 + threadIdx.x * 1e-4;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b, c[i][0], c[i][1]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma_rate(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-3, b = seed * 0.5 + threadIdx.x * 1e-4;
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(a, b, c[i]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double hA[32], hB[32], hC[64], hD[64];
  srand(7);
  int same_up = 0, same_down = 0, trials = 2000, other = 0;
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dC, sizeof hC); cudaMalloc(&dD, sizeof hD);
  for (int tr = 0; tr < trials; ++tr) {
    for (int i = 0; i < 32; ++i) { hA[i] = (rand() / (double)RAND_MAX - 0.5) * (1 + (rand() % 1000)); hB[i] = (rand() / (double)RAND_MAX - 0.5) * (1 + (rand() % 7)); }
    for (int i = 0; i < 64; ++i) hC[i] = (rand() / (double)RAND_MAX - 0.5) * 3;
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice); cudaMemcpy(dC, hC, sizeof hC, cudaMemcpyHostToDevice);
    k_check<<<1, 32>>>(dA, dB, dC, dD);
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    bool up = true, down = true;
    for (int m = 0; m < 8; ++m)
      for (int n = 0; n < 8; ++n) {
        double u = hC[m * 8 + n], d = hC[m * 8 + n];
        for (int k = 0; k < 4; ++k) u = __builtin_fma(hA[m * 4 + k], hB[k * 8 + n], u);
        for (int k = 3; k >= 0; --k) d = __builtin_fma(hA[m * 4 + k], hB[k * 8 + n], d);
        if (memcmp(&u, &hD[m * 8 + n], 8)) up = false;
        if (memcmp(&d, &hD[m * 8 + n], 8)) down = false;
      }
    same_up += up; same_down += down; other += (!up && !down);
  }
  printf("mma.m8n8k4.f64 vs FMA chain: k ascending identical in %d / %d trials, k descending in %d, neither in %d\n", same_up, trials, same_down, other);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = p.multiProcessorCount * 4;
  auto run = [&](auto kernel, const char* name, double flops_per_thread_iter) {
    kernel<<<blocks, 256>>>(out, 100, 1.0); cudaDeviceSynchronize();
    cudaEventRecord(e0); kernel<<<blocks, 256>>>(out, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = flops_per_thread_iter * iters * (double)blocks * 256;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, ms, fl / ms / 1e9);
  };
  // a DMMA is 8*8*4 FMA = 512 flop per warp = 16 flop per thread; a DFMA 2 flop per thread
  run(k_dmma_rate<4>, "DMMA, 4 accumulators/warp", 4 * 16.0);
  run(k_dmma_rate<8>, "DMMA, 8 accumulators/warp", 8 * 16.0);
  run(k_dfma_rate<8>, "DFMA, 8 accumulators/thread", 8 * 2.0);
  run(k_dfma_rate<16>, "DFMA, 16 accumulators/thread", 16 * 2.0);
  return 0;
}
