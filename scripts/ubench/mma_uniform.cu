// Micro-benchmark (tuning aid, not product): cost of issuing tcgen05.mma from (a) a divergent single-thread branch
// (`if (threadIdx.x == 32)`: nvcc wraps every MMA in an ELECT / R2UR.BROADCAST uniformisation loop) versus (b) a branch the compiler
// knows to be warp-uniform (warp index taken through __shfl_sync) with the MMA under elect.sync: operands live in uniform registers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I krasis_b200/csrc scripts/ubench/mma_uniform.cu -o krasis_b200/_lib/mma_uniform
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace kb2;

template <int N, bool UNIFORM>
__global__ void bench(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr;
  const int warp = UNIFORM ? __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) : (int)(threadIdx.x >> 5);
  const uint32_t idesc = umma_idesc_bf16_m128(N);
  const uint32_t b0 = smem_u32(smem);
  if (UNIFORM ? (warp == 1) : (threadIdx.x == 32)) {
    long long t0 = clock64();
    for (int ks = 0; ks < 256; ++ks) {
      const uint64_t bd = umma_desc_k_sw128(b0 + (ks % 3) * 24576);
      const uint32_t at = tb + 384 + (ks & 1) * 64;
      if (!UNIFORM || elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (ks > 0 || k > 0) ? 1u : 0u;
          umma_bf16_ts(tb, at + 8 * k, bd + 2 * k, idesc, acc);
          umma_bf16_ts(tb + 192, at + 32 + 8 * k, bd + 2 * k, idesc, acc);
        }
      }
      if (UNIFORM) __syncwarp();
    }
    long long t1 = clock64();
    if (!UNIFORM || elect_one()) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int N, bool UNIFORM>
void run() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<N, UNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[2], best[2] = {1LL << 60, 1LL << 60};
  for (int r = 0; r < 3; ++r) {
    bench<N, UNIFORM><<<1, 64, 200 * 1024>>>(d);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (h[1] < best[1]) { best[0] = h[0]; best[1] = h[1]; }
  }
  printf("N=%3d %-34s issue %.1f cyc/MMA   done %.1f cyc/MMA   %s\n", N, UNIFORM ? "uniform warp branch + elect.sync" : "divergent single-thread branch", best[0] / 2048.0,
         best[1] / 2048.0, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<32, false>(); run<32, true>();
  run<64, false>(); run<64, true>();
  run<128, false>(); run<128, true>();
  run<160, false>(); run<160, true>();
  return 0;
}
