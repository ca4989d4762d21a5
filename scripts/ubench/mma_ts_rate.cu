// Micro-benchmark (tuning aid, not product): sustained rate of the grouped GEMM's MMA pattern.
//   per "K step": 4 x (acc0 += A0[k] B[k]; acc1 += A1[k] B[k]) with M=128, K=16, N tokens; A from TMEM (TS) or from shared memory (SS)
//   256 K steps back to back from one thread, one commit at the end.  Reports cycles per MMA and MAC/clk.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I krasis_b200/csrc scripts/ubench/mma_ts_rate.cu -o gpurun_out/mma_ts_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace kb2;

template <int N, bool TS, bool TWO>
__global__ void bench(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr;
  if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc_bf16_m128(N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    long long t0 = clock64();
    for (int ks = 0; ks < 256; ++ks) {
      const uint64_t bd = umma_desc_k_sw128(b0 + (ks % 3) * 24576 % 49152);
      const uint64_t ad = umma_desc_k_sw128(a0 + (ks & 1) * 32768);
      const uint32_t at = tb + 384 + (ks & 1) * 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = (ks > 0 || k > 0) ? 1u : 0u;
        if (TS) {
          umma_bf16_ts(tb, at + 8 * k, bd + 2 * k, idesc, acc);
          if (TWO) umma_bf16_ts(tb + 192, at + 32 + 8 * k, bd + 2 * k, idesc, acc);
        } else {
          umma_bf16(tb, ad + 2 * k, bd + 2 * k, idesc, acc);
          if (TWO) umma_bf16(tb + 192, ad + 1024 + 2 * k, bd + 2 * k, idesc, acc);
        }
      }
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

template <int N, bool TS, bool TWO>
void run(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<N, TS, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[2] = {0, 0}, best[2] = {1LL << 60, 1LL << 60};
  for (int r = 0; r < 5; ++r) {
    bench<N, TS, TWO><<<grid, 64, 200 * 1024>>>(d);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (h[1] < best[1]) { best[0] = h[0]; best[1] = h[1]; }
  }
  const int n_mma = 256 * 4 * (TWO ? 2 : 1);
  printf("%-26s N=%3d grid=%3d  issue %.1f cyc/MMA   done %.1f cyc/MMA   %.0f MAC/clk/SM   err=%s\n", name, N, grid, (double)best[0] / n_mma,
         (double)best[1] / n_mma, 128.0 * N * 16 * n_mma / best[1], cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<128, false, false>("SS one acc", grid);
    run<128, true, false>("TS one acc", grid);
    run<160, false, true>("SS two acc, shared B", grid);
    run<160, true, true>("TS two acc, shared B", grid);
    run<192, false, true>("SS two acc, shared B", grid);
    run<192, true, true>("TS two acc, shared B", grid);
    run<160, true, false>("TS one acc", grid);
    run<256, false, false>("SS one acc", grid);
    run<256, true, false>("TS one acc", grid);
    run<64, true, true>("TS two acc, shared B", grid);
    run<32, true, true>("TS two acc, shared B", grid);
  }
  return 0;
}
