// Micro-benchmark (tuning aid, not product): does the bookkeeping between K steps break the issue/execute overlap of tcgen05.mma?
//   256 K steps x 8 MMAs (M=128, N=160, K=16, TS form).  Between K steps, per `mode` bit:
//     1  two tcgen05.commit to (unwaited) mbarriers, as the grouped GEMM does for its A and B stage "empty" barriers
//     2  tcgen05.fence::after_thread_sync
//     4  an mbarrier try_wait on a barrier whose phase already completed
//     8  one commit only
//    16  commit every second K step only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I krasis_b200/csrc scripts/ubench/mma_commit.cu -o krasis_b200/_lib/mma_commit
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace kb2;
constexpr int N = 160;

__global__ void bench(long long* out, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, dummy[4], done_bar;
  __shared__ uint32_t tptr;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1); mbar_init(&done_bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1);
    fence_mbar_init();
    mbar_arrive(&done_bar);                       // phase 0 of done_bar completes here: waits on parity 0 pass immediately
  }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr;
  if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint32_t b0 = smem_u32(smem);
    long long t0 = clock64();
    for (int ks = 0; ks < 256; ++ks) {
      if (mode & 4) mbar_wait(&done_bar, 0);
      if (mode & 2) tc_fence_after_sync();
      const uint64_t bd = umma_desc_k_sw128(b0 + (ks % 3) * 24576);
      const uint32_t at = tb + 384 + (ks & 1) * 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = (ks > 0 || k > 0) ? 1u : 0u;
        umma_bf16_ts(tb, at + 8 * k, bd + 2 * k, idesc, acc);
        umma_bf16_ts(tb + 192, at + 32 + 8 * k, bd + 2 * k, idesc, acc);
      }
      if (mode & 1) { umma_commit(&dummy[ks & 1]); umma_commit(&dummy[2 + (ks & 1)]); }
      if (mode & 8) umma_commit(&dummy[ks & 1]);
      if ((mode & 16) && (ks & 1)) umma_commit(&dummy[0]);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int modes[] = {0, 1, 2, 4, 8, 16, 3, 7};
  const char* names[] = {"plain", "2 commits / K step", "fence::after_thread_sync", "mbarrier wait (done)", "1 commit / K step", "1 commit / 2 K steps", "2 commits + fence", "2 commits + fence + wait"};
  for (int i = 0; i < 8; ++i) {
    long long h = 0, best = 1LL << 60;
    for (int r = 0; r < 3; ++r) {
      bench<<<1, 64, 200 * 1024>>>(d, modes[i]);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (h < best) best = h;
    }
    printf("%-28s %.1f cyc/MMA  (%.0f cyc per K step of 8 MMAs)   %s\n", names[i], best / 2048.0, best / 256.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
