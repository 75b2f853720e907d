// HBM write-only bandwidth on this GPU: (a) linear fill, (b) the series-major result pattern of the AC kernels
// (each thread = one point writes NS double2 values at stride P), for several grid shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu && ./write_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void fill_linear(double2* out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = make_double2(1.0, (double)i);
}

template <int NS>
__global__ void fill_series(double2* out, long long P) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
#pragma unroll 8
    for (int s = 0; s < NS; ++s) out[(size_t)s * P + p] = make_double2(1.0, (double)s);
  }
}

template <int NS>
__global__ void fill_point_major(double2* out, long long P) {   // x[P][NS]: each thread writes NS contiguous values
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
#pragma unroll 8
    for (int s = 0; s < NS; ++s) out[(size_t)p * NS + s] = make_double2(1.0, (double)s);
  }
}

template <typename F>
double time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) f();
  float best = 1e30f;
  for (int i = 0; i < 5; ++i) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  return best;
}

int main(int argc, char** argv) {
  const long long P = argc > 1 ? atoll(argv[1]) : 1000001;
  constexpr int NS = 192;
  const size_t n = (size_t)P * NS;
  double2* buf;
  if (cudaMalloc(&buf, n * sizeof(double2)) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  setvbuf(stdout, nullptr, _IOLBF, 0);
  const double gb = n * 16.0 / 1e9;
  for (int blocks : {148, 296}) {
    for (int threads : {96, 128, 160, 192, 224, 256, 512}) {
      double t1 = time_ms([&] { fill_linear<<<blocks, threads>>>(buf, n); });
      double t2 = time_ms([&] { fill_series<NS><<<blocks, threads>>>(buf, P); });
      double t3 = time_ms([&] { fill_point_major<NS><<<blocks, threads>>>(buf, P); });
      printf("grid %5d x %4d : linear %.3f ms %.0f GB/s | series-major %.3f ms %.0f GB/s | point-major %.3f ms %.0f GB/s\n", blocks, threads, t1,
             gb / t1 * 1e3, t2, gb / t2 * 1e3, t3, gb / t3 * 1e3);
    }
  }
  // Store rate of ONE SM for the series-major pattern (few CTAs, one per SM): what an SM whose warps are all in
  // their store phase can push, against its 1/148 share of the HBM rate.
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int blocks : {1, 4, 16, 37, 74}) {
    for (int threads : {32, 96, 192, 384}) {
      long long Ps = (long long)blocks * threads * 256;
      if (Ps > P) Ps = P / ((long long)blocks * threads) * blocks * threads;   // stay inside the buffer
      const double gbs = (double)Ps * NS * 16.0 / 1e9;
      double t = time_ms([&] { fill_series<NS><<<blocks, threads>>>(buf, Ps); });
      printf("per-SM: grid %3d x %3d : %.3f ms, %.1f GB/s per SM = %.1f B/clk at %.0f MHz\n", blocks, threads, t, gbs / t * 1e3 / blocks,
             gbs / t * 1e3 / blocks * 1e9 / (clk_khz * 1e3), clk_khz / 1e3);
    }
  }
  double t4 = time_ms([&] { cudaMemsetAsync(buf, 0, n * 16); });
  printf("cudaMemset: %.3f ms %.0f GB/s\n", t4, gb / t4 * 1e3);
  cudaFree(buf);
  return 0;
}
