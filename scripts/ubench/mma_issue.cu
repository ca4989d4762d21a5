// Micro-benchmark (tuning aid, not product): cost of issuing small tcgen05.mma instructions from one thread.
//   for N in {32, 64, 128, 256}: 64 back-to-back MMAs (M=128, K=16, BF16) with (a) descriptors recomputed per MMA, (b) one
//   descriptor pair reused; clock64 from first issue to last issue ("issue") and to the commit's mbarrier completion ("done").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I krasis_b200/csrc scripts/ubench/mma_issue.cu -o gpurun_out/mma_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace kb2;

template <int N, bool REUSE, bool DEP>
__global__ void bench(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr;
  if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
    const uint64_t ad0 = umma_desc_k_sw128(a0), bd0 = umma_desc_k_sw128(b0);
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const uint64_t ad = REUSE ? ad0 : umma_desc_k_sw128(a0 + (i & 1) * 16384) + 2 * (i & 3);
      const uint64_t bd = REUSE ? bd0 : umma_desc_k_sw128(b0 + (i & 1) * 16384) + 2 * (i & 3);
      umma_bf16(tb + (DEP ? 0 : ((i & 1) * 256)), ad, bd, idesc, i > 1 ? 1u : 0u);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int N, bool REUSE, bool DEP>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<N, REUSE, DEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[2] = {0, 0}, best[2] = {1LL << 60, 1LL << 60};
  for (int r = 0; r < 5; ++r) {
    bench<N, REUSE, DEP><<<1, 64, 200 * 1024>>>(d);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (h[1] < best[1]) { best[0] = h[0]; best[1] = h[1]; }
  }
  printf("%-28s N=%3d  issue %6lld cyc (%.1f / MMA)   done %6lld cyc (%.1f / MMA)   err=%s\n", name, N, best[0], best[0] / 64.0, best[1],
         best[1] / 64.0, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<32, false, true>("recompute desc, dependent");
  run<32, true, true>("reuse desc, dependent");
  run<32, true, false>("reuse desc, 2 accumulators");
  run<64, true, true>("reuse desc, dependent");
  run<128, true, true>("reuse desc, dependent");
  run<256, true, true>("reuse desc, dependent");
  run<256, false, true>("recompute desc, dependent");
  return 0;
}
