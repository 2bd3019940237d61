// store_pattern.cu -- what does the output write pattern of the interpolation kernel cost on its own?
// Writes n rows of 2304 bytes (C3: 12 modes x 12 complex) in several patterns and reports GB/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern store_pattern.cu && ./store_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <random>

constexpr int ROW = 2304;  // bytes
__device__ __forceinline__ void st16(void* p, double a, double b) { asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory"); }
__device__ __forceinline__ void st32(void* p, double a, double b) { asm volatile("st.global.v4.f64 [%0], {%1,%2,%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory"); }

// mode 0: a warp writes a row as contiguous 512-byte pieces (16 B per lane); 1: contiguous 1 KB pieces (32 B per lane)
// mode 2: lane owns 48-byte pieces, three 16-byte stores (stride 48); 3: lane owns 48-byte pieces, 32 + 16 split
template <int MODE>
__global__ void k_store(char* out, const uint32_t* rows, size_t n) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < n; r += nwarp) {
    // rows == nullptr: pseudo-random row from arithmetic (no load in front of the stores)
    const size_t ri = rows ? (size_t)rows[r] : (size_t)((r * 7919003ull) % n);
    char* row = out + ri * ROW;
    const double a = (double)r, b = (double)lane;
    if (MODE == 0) {
      for (int o = lane * 16; o < ROW; o += 512) st16(row + o, a, b);
    } else if (MODE == 1) {
      for (int o = lane * 32; o < ROW; o += 1024) st32(row + o, a, b);
    } else {
      for (int piece = lane; piece < ROW / 48; piece += 32) {
        char* p = row + piece * 48;
        if (MODE == 2) { st16(p, a, b); st16(p + 16, a, b); st16(p + 32, a, b); }
        else {
          const bool even = (piece & 1) == 0;  // ROW is a multiple of 32
          st32(even ? p : p + 16, a, b);
          st16(even ? p + 32 : p, a, b);
        }
      }
    }
  }
}

// mode 4: TMA bulk stores: one elected lane per warp issues a 2304-byte cp.async.bulk from shared memory per row
__global__ void k_store_tma(char* out, const uint32_t* rows, size_t n) {
  __shared__ __align__(128) char buf[8][ROW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = lane * 16; o < ROW; o += 512) *reinterpret_cast<double2*>(buf[w] + o) = make_double2(1.0, 2.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
  if (lane == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(buf[w]);
    for (size_t r = warp; r < n; r += nwarp) {
      char* row = out + (size_t)rows[r] * ROW;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(row), "r"(src), "r"(ROW) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
static void run_tma(const char* name, char* out, const uint32_t* rows, size_t n, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) k_store_tma<<<148 * ctas_per_sm, 256>>>(out, rows, n);
  cudaEventRecord(e0);
  for (int it = 0; it < 5; ++it) k_store_tma<<<148 * ctas_per_sm, 256>>>(out, rows, n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-44s %7.3f ms  %7.1f GB/s (%d CTAs/SM)\n", name, ms, (double)n * ROW / ms * 1e-6, ctas_per_sm);
}

// ctas_per_sm < 8: occupancy is limited with dynamic shared memory, like the interpolation kernel (2 CTAs of 256 threads)
// mode 5: like the interpolation kernel before the row-ownership fix: a row is written in two parts at different times by
// different warps (pieces 0..31 by one warp, pieces 32..47 -- together with the first 16 of the next row -- by another)
__global__ void k_store_split(char* out, size_t n, int delay) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
  // task t covers 32 consecutive 48-byte pieces of the concatenation of the (pseudo-randomly placed) rows
  const size_t n_task = n * 48 / 32;
  for (size_t t = warp; t < n_task; t += nwarp) {
    const size_t tt = (t + (size_t)delay * nwarp * ((t & 1) ? 1 : 0)) % n_task;  // odd tasks run `delay` sweeps later
    const size_t piece = tt * 32 + lane, r = piece / 48, k = piece - r * 48;
    char* p = out + ((r * 7919003ull) % n) * ROW + k * 48;
    const bool even = (k & 1) == 0;
    st32(even ? p : p + 16, (double)r, 1.0);
    st16(even ? p + 32 : p, (double)r, 1.0);
  }
}

template <int MODE>
static void run(const char* name, char* out, const uint32_t* rows, size_t n, int ctas_per_sm = 8) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t smem = ctas_per_sm >= 8 ? 0 : (size_t)(220 * 1024 / ctas_per_sm - 2048);
  cudaFuncSetAttribute(k_store<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = 148 * ctas_per_sm;
  for (int it = 0; it < 2; ++it) k_store<MODE><<<grid, 256, smem>>>(out, rows, n);
  cudaEventRecord(e0);
  for (int it = 0; it < 5; ++it) k_store<MODE><<<grid, 256, smem>>>(out, rows, n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("%-44s %7.3f ms  %7.1f GB/s (%d CTAs/SM)\n", name, ms, (double)n * ROW / ms * 1e-6, ctas_per_sm);
}

int main() {
  const size_t n = 10000000;
  char* out; uint32_t* rows;
  cudaMalloc(&out, n * ROW); cudaMalloc(&rows, n * 4);
  std::vector<uint32_t> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (uint32_t)i;
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemcpy(rows, h.data(), n * 4, cudaMemcpyHostToDevice);
    const char* tag = pass ? "random rows" : "sequential rows";
    printf("-- %s\n", tag);
    run<0>("contiguous 512 B per instruction", out, rows, n);
    run<1>("contiguous 1 KB per instruction (STG.256)", out, rows, n);
    run<2>("48-byte pieces, 3 x 16 B", out, rows, n);
    run<3>("48-byte pieces, 32 B + 16 B", out, rows, n);
    run<3>("48-byte pieces, 32 B + 16 B", out, rows, n, 4);
    run<3>("48-byte pieces, 32 B + 16 B", out, rows, n, 2);
    run<3>("48-byte pieces, 32 B + 16 B", out, rows, n, 1);
    run<0>("contiguous 512 B per instruction", out, rows, n, 2);
    if (pass) {
      printf("-- arithmetic random rows (no index load)\n");
      for (int c : {8, 4, 2, 1}) run<3>("48-byte pieces, 32 B + 16 B", out, nullptr, n, c);
      for (int c : {8, 4, 2, 1}) run<0>("contiguous 512 B per instruction", out, nullptr, n, c);
      for (int delay : {0, 1, 4}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const size_t smem = 220 * 1024 / 2 - 2048;
        cudaFuncSetAttribute(k_store_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_store_split<<<148 * 2, 256, smem>>>(out, n, delay);
        cudaEventRecord(e0);
        for (int it = 0; it < 5; ++it) k_store_split<<<148 * 2, 256, smem>>>(out, n, delay);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("rows split over two warps, second part %d sweeps later   %7.3f ms  %7.1f GB/s (2 CTAs/SM)\n", delay, ms, (double)n * ROW / ms * 1e-6);
      }
    }
    run_tma("TMA bulk store 2304 B per row", out, rows, n, 1);
    run_tma("TMA bulk store 2304 B per row", out, rows, n, 2);
    run_tma("TMA bulk store 2304 B per row", out, rows, n, 4);
    std::mt19937_64 g(1); std::shuffle(h.begin(), h.end(), g);
  }
  cudaMemset(out, 0, n * ROW);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); for (int i = 0; i < 5; ++i) cudaMemsetAsync(out, 1, n * ROW); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemset %7.1f GB/s\n", (double)n * ROW / (ms / 5) * 1e-6);
  return 0;
}
