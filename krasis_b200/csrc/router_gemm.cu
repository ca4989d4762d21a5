// Router logits on tcgen05: logits[M,E] = hidden[M,H] (bf16) x gate[E,H]^T (bf16), fp32 accumulate in TMEM.
//
// Restates python/krasis/layer.py:532-534 (`torch.matmul(hidden.float(), gate.float().t())` + gate_bias).
// hidden and gate are BF16, so every product is exact in fp32; the tensor core's fp32 accumulation differs
// from cuBLAS sgemm only in summation order / rounding of partial sums — the same class of difference as
// between any two fp32 GEMMs (see oracle/router.py).  Id parity is asserted by tests away from near-ties.
//
// One CTA per (128-token tile, 256-expert column block): UMMA M = 128, N <= 256.  With E = 512 that is 2 x 64 = 128 CTAs at 8192
// tokens (one CTA per token tile and all 512 columns left 84 of the 148 SMs idle and ran at 0.27 of the tensor peak).  Both operands
// arrive by 2-D TMA with 128 B swizzle, 64-wide K blocks, 4-stage ring; one warp produces, one elected lane issues tcgen05.mma (SS),
// four warps drain TMEM and write fp32 logits rows (64 B contiguous per thread per tcgen05.ld).
#include <cuda.h>

#include "moe_common.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace kb2 {

constexpr int kRThreads = 256;
constexpr int kRStages = 4;
constexpr int kRATile = 128 * kBlockK * 2;        // 16 KB
constexpr int kRBMax = 256 * kBlockK * 2;         // 32 KB
constexpr int kRStage = kRATile + kRBMax;         // 48 KB
constexpr int kROffBar = kRStages * kRStage;
constexpr int kRSmem = kROffBar + 64;

__global__ void __launch_bounds__(kRThreads, 1)
    router_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
                       const float* __restrict__ gate_bias, float* __restrict__ logits, int M, int E, int H) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kROffBar);
  uint64_t* empty = full + kRStages;
  uint64_t* done = empty + kRStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done + 1);
  // shuffle-broadcast warp index: role branches provably warp-uniform -> TMA / tcgen05 operands stay in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int nkb = H / kBlockK;
  const int col0 = blockIdx.y * 256;         // first expert column of this CTA
  const int n0 = E - col0 > 256 ? 256 : E - col0;

  if (threadIdx.x == 32) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int i = 0; i < kRStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, 256);
  if (threadIdx.x == 64) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_g);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = kRATile + (uint32_t)n0 * kBlockK * 2;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* a = smem + stage * kRStage;
        uint8_t* b = a + kRATile;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], bytes);
          tma_load_2d(a, &tmap_x, kb * kBlockK, m0, &full[stage]);                 // 128 token rows (OOB rows -> 0)
          for (int r = 0; r < n0; r += 64) tma_load_2d(b + r * 128, &tmap_g, kb * kBlockK, col0 + r, &full[stage]);
        }
        __syncwarp();
        if (++stage == kRStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t id0 = umma_idesc_bf16_m128((uint32_t)n0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + stage * kRStage);
        const uint64_t ad = umma_desc_k_sw128(a_addr);
        const uint64_t bd0 = umma_desc_k_sw128(a_addr + kRATile);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            umma_bf16(tmem_base, ad + 2 * k, bd0 + 2 * k, id0, acc);
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == kRStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* out = logits + (long long)m * E + col0;
    const float* bias = gate_bias ? gate_bias + col0 : nullptr;
    for (int c0 = 0; c0 < n0; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + c0, r);
      tmem_ld_wait();
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 v;
          v.x = __uint_as_float(r[j]) + (bias ? bias[c0 + j] : 0.f);
          v.y = __uint_as_float(r[j + 1]) + (bias ? bias[c0 + j + 1] : 0.f);
          v.z = __uint_as_float(r[j + 2]) + (bias ? bias[c0 + j + 2] : 0.f);
          v.w = __uint_as_float(r[j + 3]) + (bias ? bias[c0 + j + 3] : 0.f);
          *reinterpret_cast<float4*>(out + c0 + j) = v;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

cudaError_t make_tmap_bf16_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows);

bool router_gemm_supported(int E, int H) { return E >= 16 && E <= 512 && E % 64 == 0 && H % kBlockK == 0; }

// tmap_g: tensor map over gate [E][H] with box rows 64 (built once per layer); x: [M][H] bf16
cudaError_t launch_router_gemm(const void* x, const void* tmap_g, const float* bias, float* logits, int M, int E,
                               int H, cudaStream_t s) {
  KernelSpan ks(K_ROUTER_GEMM, s);
  alignas(64) CUtensorMap tx;
  cudaError_t e = make_tmap_bf16_rows(&tx, x, M, H, 128);
  if (e != cudaSuccess) return e;
  static PerDeviceOnce once;
  if (const int dev = once.pending(); dev >= 0) {
    e = cudaFuncSetAttribute(router_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRSmem);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  router_gemm_kernel<<<dim3((M + 127) / 128, (E + 255) / 256), kRThreads, kRSmem, s>>>(tx, *reinterpret_cast<const CUtensorMap*>(tmap_g), bias,
                                                               logits, M, E, H);
  return cudaGetLastError();
}

}  // namespace kb2
