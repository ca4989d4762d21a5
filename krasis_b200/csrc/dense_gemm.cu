// Dense BF16 linear layer on tcgen05:  Y[M,N] = X[M,K] (bf16) x W[N,K]^T (bf16), fp32 accumulate in TMEM,
// BF16 (or FP32) output.  This is `torch.nn.functional.linear` for the attention projections the reference
// keeps in BF16 (python/krasis/attention.py:526-529,672; python/krasis/linear_attention.py:709-710,815 —
// "attention weights BF16 only", python/krasis/config.py:209).
//
// Persistent CTAs, 128 x 256 output tiles, 64-wide K blocks, 4-stage TMA ring (2-D tensor maps, 128 B swizzle),
// one MMA thread (SS-MMA, M=128, N=256), two 256-column TMEM accumulators so the epilogue of tile i overlaps the
// main loop of tile i+1, four epilogue warps writing 32 B per thread per tcgen05.ld.
#include <cuda.h>

#include "moe_common.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace kb2 {

constexpr int kDThreads = 256;
constexpr int kDStages = 4;
constexpr int kDTileM = 128, kDTileN = 256;
constexpr int kDABytes = kDTileM * kBlockK * 2;   // 16 KB
constexpr int kDBBytes = kDTileN * kBlockK * 2;   // 32 KB
constexpr int kDStage = kDABytes + kDBBytes;
constexpr int kDOffBar = kDStages * kDStage;
constexpr int kDSmem = kDOffBar + 128;

// MODE 0: bf16 x bf16 -> bf16   1: bf16 x bf16 -> f32   2: int8 x int8 -> int32 -> bf16( float(acc) * (xs[m] * ws[n]) )
// (W8A8 `int8_linear`, python/krasis/weight_loader.py:46-99: exact integer accumulation, fp32 dequant).
template <int MODE>
__global__ void __launch_bounds__(kDThreads, 1)
    dense_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      void* __restrict__ out, const float* __restrict__ bias, int M, int N, int K, long long ldo,
                      const float* __restrict__ row_scale, const __nv_bfloat16* __restrict__ col_scale, CombineScatter sc) {
  constexpr bool kOutF32 = MODE == 1;
  constexpr int kElemsPerBlock = MODE == 2 ? 128 : 64;     // 128 B of K per smem row
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kDOffBar);
  uint64_t* empty = full + kDStages;
  uint64_t* acc_full = empty + kDStages;     // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // warp index through a shuffle so that the role branches are provably warp-uniform: the single-thread instructions (TMA, tcgen05.mma,
  // commit) then take their operands from uniform registers instead of an ELECT + R2UR.BROADCAST loop per instruction
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_tiles = (N + kDTileN - 1) / kDTileN, m_tiles = (M + kDTileM - 1) / kDTileM;
  const int total = n_tiles * m_tiles;
  const int nkb = K / kElemsPerBlock;

  if (threadIdx.x == 32) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int i = 0; i < kDStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, 512);
  if (threadIdx.x == 64) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m0 = (t / n_tiles) * kDTileM, n0 = (t % n_tiles) * kDTileN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a = smem + stage * kDStage;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], kDStage);
            tma_load_2d(a, &tmap_x, kb * kElemsPerBlock, m0, &full[stage]);
            tma_load_2d(a + kDABytes, &tmap_w, kb * kElemsPerBlock, n0, &full[stage]);               // rows n0..n0+127
            tma_load_2d(a + kDABytes + kDABytes, &tmap_w, kb * kElemsPerBlock, n0 + 128, &full[stage]); // rows n0+128..+255
          }
          __syncwarp();
          if (++stage == kDStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t aphase[2] = {0, 0};
      // kind::i8: c_format S32 (2) at [4,6), a/b format INT8 (1) at [7,10)/[10,13)
      const uint32_t idesc = MODE == 2 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((kDTileN >> 3) << 17) | ((128u >> 4) << 24))
                                       : umma_idesc_bf16_m128(kDTileN);
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&acc_empty[b], aphase[b] ^ 1);
        aphase[b] ^= 1;
        tc_fence_after_sync();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * kDStage);
          const uint64_t ad = umma_desc_k_sw128(a_addr);
          const uint64_t bd = umma_desc_k_sw128(a_addr + kDABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // 4 x 32 B of K per 128 B row (K=16 bf16 or K=32 int8 per MMA)
              if constexpr (MODE == 2)
                umma_i8(tmem_base + b * kDTileN, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16(tmem_base + b * kDTileN, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == kDStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&acc_full[b]);
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    uint32_t fphase[2] = {0, 0};
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int b = it & 1;
      const int m = (t / n_tiles) * kDTileM + q * 32 + lane, n0 = (t % n_tiles) * kDTileN;
      mbar_wait(&acc_full[b], fphase[b]);
      fphase[b] ^= 1;
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * kDTileN;
#pragma unroll 2
      for (int c0 = 0; c0 < kDTileN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        const int n = n0 + c0;
        if (m < M && n < N) {          // N % 16 == 0 is required by the launcher
          float v[16];
          if constexpr (MODE == 2) {
            const float xs = row_scale[m];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (float)(int)r[j] * (xs * __bfloat162float(col_scale[n + j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + (bias ? bias[n + j] : 0.f);
          }
          if constexpr (kOutF32) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (long long)m * ldo + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(out) + (long long)m * ldo;
            if (sc.n_ranks > 0) {
              // GEMM -> reduce-scatter fused: this rank's partial output row goes straight into the receive buffer of the rank
              // that owns token m (peer-mapped memory, [rows_per_rank][n_ranks][ldo], slot = source rank); see CombineScatter
              const int owner = m / sc.rows_per_rank, ml = m - owner * sc.rows_per_rank;
              orow = sc.peer_out[owner] + ((long long)ml * sc.n_ranks + sc.src_rank) * ldo;
            }
            uint4* o = reinterpret_cast<uint4*>(orow + n);
            o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[b]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

cudaError_t make_tmap_bf16_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows);
cudaError_t make_tmap_u8_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows);

static cudaError_t dense_configure() {
  static PerDeviceOnce once;
  const int dev = once.pending();
  if (dev < 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(dense_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dense_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dense_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem);
  if (e != cudaSuccess) return e;
  once.mark(dev);
  return cudaSuccess;
}

// X [M][K] bf16 row-major, W [N][K] bf16 row-major, out [M][ldo] (bf16 or f32).  K % 64 == 0, N % 16 == 0.
cudaError_t launch_dense_gemm_scatter(const void* x, const void* w, void* out, const float* bias, int M, int N, int K,
                                      long long ldo, bool out_f32, int num_sms, cudaStream_t s, const CombineScatter& sc);
cudaError_t launch_dense_gemm(const void* x, const void* w, void* out, const float* bias, int M, int N, int K,
                              long long ldo, bool out_f32, int num_sms, cudaStream_t s) {
  return launch_dense_gemm_scatter(x, w, out, bias, M, N, K, ldo, out_f32, num_sms, s, CombineScatter{});
}
// sc.n_ranks > 0 (BF16 output only): rows are stored into the token owners' receive buffers instead of `out`
cudaError_t launch_dense_gemm_scatter(const void* x, const void* w, void* out, const float* bias, int M, int N, int K,
                                      long long ldo, bool out_f32, int num_sms, cudaStream_t s, const CombineScatter& sc) {
  if (sc.n_ranks > 0 && (out_f32 || M % sc.n_ranks || sc.rows_per_rank * sc.n_ranks != M)) return cudaErrorInvalidValue;
  KernelSpan ks(K_DENSE_BF16, s);
  if (K % kBlockK || N % 16 || M <= 0) return cudaErrorInvalidValue;
  alignas(64) CUtensorMap tx, tw;
  cudaError_t e = make_tmap_bf16_rows(&tx, x, M, K, 128);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tw, w, N, K, 128);
  if (e != cudaSuccess) return e;
  e = dense_configure();
  if (e != cudaSuccess) return e;
  const int total = ((M + kDTileM - 1) / kDTileM) * ((N + kDTileN - 1) / kDTileN);
  const int grid = total < num_sms ? total : num_sms;
  if (out_f32)
    dense_gemm_kernel<1><<<grid, kDThreads, kDSmem, s>>>(tx, tw, out, bias, M, N, K, ldo, nullptr, nullptr, CombineScatter{});
  else
    dense_gemm_kernel<0><<<grid, kDThreads, kDSmem, s>>>(tx, tw, out, bias, M, N, K, ldo, nullptr, nullptr, sc);
  return cudaGetLastError();
}

// W8A8: xq [M][K] int8, x_scale [M] f32, wq [N][K] int8, w_scale [N] bf16 -> out [M][ldo] bf16.  K % 128 == 0, N % 16 == 0.
cudaError_t launch_dense_gemm_i8(const void* xq, const float* x_scale, const void* wq, const void* w_scale, void* out,
                                 int M, int N, int K, long long ldo, int num_sms, cudaStream_t s) {
  KernelSpan ks(K_DENSE_I8, s);
  if (K % 128 || N % 16 || M <= 0) return cudaErrorInvalidValue;
  alignas(64) CUtensorMap tx, tw;
  cudaError_t e = make_tmap_u8_rows(&tx, xq, M, K, 128);
  if (e != cudaSuccess) return e;
  e = make_tmap_u8_rows(&tw, wq, N, K, 128);
  if (e != cudaSuccess) return e;
  e = dense_configure();
  if (e != cudaSuccess) return e;
  const int total = ((M + kDTileM - 1) / kDTileM) * ((N + kDTileN - 1) / kDTileN);
  const int grid = total < num_sms ? total : num_sms;
  dense_gemm_kernel<2><<<grid, kDThreads, kDSmem, s>>>(tx, tw, out, nullptr, M, N, K, ldo, x_scale,
                                                      (const __nv_bfloat16*)w_scale, CombineScatter{});
  return cudaGetLastError();
}

}  // namespace kb2
