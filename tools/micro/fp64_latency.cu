// Dependent-chain latencies of the FP64 instructions the LU kernels are made of (one warp, one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain(double* out, long long* cyc, double seed, int iters) {
  double a = seed + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9, d = a;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) a = fma(a, b, c);                         // DFMA
      if (MODE == 1) a = a * b;                                // DMUL
      if (MODE == 2) a = a + c;                                // DADD
      if (MODE == 3) a = 1.0 / a + 0.5;                        // full division (+DADD)
      if (MODE == 4) a = (a < 1e300) ? fma(a, b, c) : d;       // DFMA || DSETP -> FSEL
      if (MODE == 5) { a = fma(a, b, c); d = fma(d, b, c); }   // two independent chains
      if (MODE == 6) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 0.5; }  // MUFU.RCP64H (+DADD)
      if (MODE == 7) a = sqrt(a) + 0.5;
      if (MODE == 8) { a = fma(a, b, c); d = fma(d, b, c); c = fma(c, b, 1e-12); b = fma(b, 1.0, 1e-12);}   // four independent chains
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + b + c;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_unroll, int warps) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&cyc, 8 * 148);
  const int iters = 2000;
  chain<MODE><<<1, 32 * warps>>>(out, cyc, 1.0, iters);
  chain<MODE><<<1, 32 * warps>>>(out, cyc, 1.0, iters);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/SM=%2d  %.2f cycles per iteration step (%d dependent op(s) each)\n", name, warps, (double)h / (iters * 16.0), ops_per_unroll);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16, 32}) {
    run<0>("DFMA dependent chain", 1, w);
    run<1>("DMUL dependent chain", 1, w);
    run<2>("DADD dependent chain", 1, w);
    run<5>("2 independent DFMA chains", 1, w);
    run<8>("4 independent DFMA chains", 1, w);
    run<4>("DFMA + DSETP -> FSEL", 2, w);
    run<6>("MUFU.RCP64H + DADD", 2, w);
    run<3>("1.0/a (full IEEE division) + DADD", 1, w);
    run<7>("sqrt + DADD", 1, w);
  }
  return 0;
}
