// Micro-benchmark (tuning aid, not product): what slows the grouped GEMM's TS-form MMA stream down in situ?
//   MMA thread: 256 K steps x 8 tcgen05.mma (M=128, N=160, K=16, A from TMEM, B from shared memory), as in grouped_gemm.cu.
//   Background load selected by `mode` bits, running until the MMA thread sets a flag:
//     1  eight warps store the A operand with tcgen05.st (32 columns each, back to back)           -> TMEM write port
//     2  eight warps stream 16 B shared-memory loads + stores (each warp its own 4 KB window)          -> shared-memory port
//     4  one thread streams 24 KB bulk global->shared copies (TMA engine writes into shared memory)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I krasis_b200/csrc scripts/ubench/mma_contention.cu -o krasis_b200/_lib/mma_contention
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace kb2;

constexpr int N = 160;

__global__ void __launch_bounds__(384, 1) bench(long long* out, const uint8_t* gsrc, int mode, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, tbar[2];
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  __shared__ unsigned long long work[12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) tmem_alloc(&tptr, 512);
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tptr;
  unsigned long long cnt = 0;
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128(N);
      const uint32_t b0 = smem_u32(smem);
      long long t0 = clock64();
      for (int ks = 0; ks < 256; ++ks) {
        const uint64_t bd = umma_desc_k_sw128(b0 + (ks % 3) * 24576);
        const uint32_t at = tb + 384 + (ks & 1) * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (ks > 0 || k > 0) ? 1u : 0u;
          umma_bf16_ts(tb, at + 8 * k, bd + 2 * k, idesc, acc);
          umma_bf16_ts(tb + 192, at + 32 + 8 * k, bd + 2 * k, idesc, acc);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (warp == 2) {
    if ((mode & 4) && lane == 0) {                       // bulk copies into the region behind the B stages, two in flight
      uint32_t ph[2] = {0, 0};
      auto issue = [&](int s, unsigned long long n) {
        mbar_arrive_expect_tx(&tbar[s], 24576);
        bulk_g2s(smem + 98304 + s * 24576, gsrc + (n % 512) * 24576 + (size_t)blockIdx.x * 512 * 24576, 24576, &tbar[s]);
      };
      issue(0, 0);
      issue(1, 1);
      int s = 0;
      while (!stop) {
        mbar_wait(&tbar[s], ph[s]);
        ph[s] ^= 1;
        ++cnt;
        issue(s, cnt + 1);
        s ^= 1;
      }
      mbar_wait(&tbar[s], ph[s]);
      mbar_wait(&tbar[s ^ 1], ph[s ^ 1]);
      work[2] = cnt;
    }
  } else if (warp >= 4) {
    const int w = warp - 4;                              // 0..7
    const bool both = (mode & 3) == 3;
    if ((mode & 1) && (!both || w < 4)) {
      uint32_t o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0x3c003c00u + i;
      // warps w and w+4 share lane quadrant w&3; they write different unused columns [448, 512) so the MMA's A data is untouched
      const uint32_t addr = tb + ((uint32_t)((w & 3) * 32) << 16) + 448 + (w >> 2) * 32;
      while (!stop) {
        tmem_st32(addr, o);
        tmem_st_wait();
        ++cnt;
      }
    }
    if ((mode & 2) && (!both || w >= 4)) {
      uint4* win = reinterpret_cast<uint4*>(smem + 147456 + w * 4096);
      uint4 v = make_uint4(1, 2, 3, 4);
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 x = win[lane + 32 * (i & 7)];
          v.x ^= x.x; v.y += x.y;
          win[lane + 32 * ((i + 1) & 7)] = v;
        }
        cnt += 16;                                       // 16 warp-wide 512 B accesses
      }
      if (v.x == 0x12345) sink[0] = v.y;
    }
    if (lane == 0) work[4 + w] = cnt;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long tot = 0;
    for (int i = 4; i < 12; ++i) tot += work[i];
    out[200] = (long long)tot;
    const bool both = (mode & 3) == 3;
    unsigned long long st = 0, ls = 0;
    for (int w = 0; w < 8; ++w) {
      if ((mode & 1) && (!both || w < 4)) st += work[4 + w];
      if ((mode & 2) && (!both || w >= 4)) ls += work[4 + w];
    }
    out[202] = (long long)st;
    out[203] = (long long)ls;
    out[201] = (long long)work[2];
  }
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; uint8_t* g; unsigned long long* sink;
  cudaMalloc(&d, 8 * 256); cudaMalloc(&sink, 64);
  cudaMalloc(&g, (size_t)148 * 512 * 24576);
  cudaMemset(g, 0x3c, (size_t)148 * 512 * 24576);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[8] = {"MMA alone", "+ tcgen05.st (8 warps)", "+ LDS/STS stream (8 warps)", "+ st + LDS/STS", "+ bulk g2s stream", "+ st + bulk", "+ LDS/STS + bulk", "+ all"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 8; ++mode) {
      long long h[256];
      cudaMemset(d, 0, 8 * 256);
      bench<<<grid, 384, 200 * 1024>>>(d, g, mode, sink);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 8 * 256, cudaMemcpyDeviceToHost);
      const double cyc = (double)h[0];
      printf("grid=%3d %-28s %.1f cyc/MMA  %.0f MAC/clk/SM | background: %.1f B/clk shared ld+st, %.1f B/clk tmem st, %.1f B/clk bulk   %s\n", grid, names[mode],
             cyc / 2048, 128.0 * N * 16 * 2048 / cyc, (mode & 2) ? h[203] * 512.0 / cyc : 0.0, (mode & 1) ? h[202] * 4096.0 / cyc : 0.0,
             (mode & 4) ? h[201] * 24576.0 / cyc : 0.0, cudaGetErrorString(e));
    }
  return 0;
}
