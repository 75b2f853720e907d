// Cost of the pieces of a shared->global bulk-copy epilogue on sm_100a:
//  (a) fence.proxy.async.shared::cta per thread, (b) cp.async.bulk.global.shared::cta issue rate for small copies,
//  (c) commit_group / wait_group.read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store bulk_store.cu && ./bulk_store
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fence(long long* cyc, int iters) {
  extern __shared__ double2 sm[];
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * 16;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(sb), "d"(1.0 * i), "d"(2.0) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// every warp (or the whole block when BLOCKWIDE) copies `bytes` from its staging row to global, `per_group` copies per commit
template <bool BLOCKWIDE>
__global__ void k_bulk(char* out, size_t out_stride, long long* cyc, int iters, int per_group, unsigned bytes, int ring) {
  extern __shared__ double2 sm[];
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm) + (BLOCKWIDE ? 0 : warp * 512);
  char* g = out + ((size_t)blockIdx.x * (blockDim.x >> 5) + (BLOCKWIDE ? 0 : warp)) * out_stride;
  const bool issuer = BLOCKWIDE ? threadIdx.x == 0 : lane == 0;
  long long t0 = clock64();
  unsigned slot = 0;
  for (int i = 0; i < iters; ++i) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(sb + slot * blockDim.x * 16 + (BLOCKWIDE ? threadIdx.x : lane) * 16), "d"(1.0 * i), "d"(2.0) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (BLOCKWIDE) __syncthreads(); else __syncwarp();
    if (issuer) {
      for (int q = 0; q < per_group; ++q)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + (size_t)((i * per_group + q) & 1023) * bytes),
                     "r"(sb + slot * blockDim.x * 16), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    slot = (slot + 1) % ring;
  }
  if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 8 * 1024);
  char* out;
  const size_t stride = 1024 * 4096;
  cudaMalloc(&out, stride * 148 * 8);
  const int iters = 2000;
  long long h;
  cudaFuncSetAttribute(k_fence, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_bulk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int threads : {32, 160}) {
    k_fence<<<148, threads, 64 * 1024>>>(cyc, iters);
    k_fence<<<148, threads, 64 * 1024>>>(cyc, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS + fence.proxy.async: %d threads/SM: %.1f cycles per iteration\n", threads, (double)h / iters);
  }
  for (int per_group : {1, 3, 12}) {
    for (int threads : {32, 160}) {
      k_bulk<false><<<148, threads, 160 * 1024>>>(out, stride, cyc, iters, per_group, 512, 8);
      k_bulk<false><<<148, threads, 160 * 1024>>>(out, stride, cyc, iters, per_group, 512, 8);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("per-warp 512 B bulk copies, %2d per group, %3d threads/SM: %.1f cycles per group (%.1f per copy per warp)\n", per_group, threads,
             (double)h / iters, (double)h / iters / per_group);
    }
    k_bulk<true><<<148, 160, 160 * 1024>>>(out, stride, cyc, iters, per_group, 2560, 8);
    k_bulk<true><<<148, 160, 160 * 1024>>>(out, stride, cyc, iters, per_group, 2560, 8);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("per-block 2560 B bulk copies, %2d per group, 160 threads/SM: %.1f cycles per group (%.1f per copy)\n", per_group, (double)h / iters,
           (double)h / iters / per_group);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
