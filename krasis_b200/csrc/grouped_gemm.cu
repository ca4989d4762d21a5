// Grouped W4A16 / W8A16 expert GEMM for MoE prefill on sm_100a (tcgen05 + TMEM + bulk TMA).
//
// Replaces the two `moe_wna16_marlin_gemm` launches + `silu_and_mul` of the reference's
// `fused_marlin_moe` call sequence (python/krasis/gpu_prefill.py:152-226; survey K3/K4/K5).
//
// Orientation ("swap-AB"): per expert the weight matrix is the UMMA *A* operand (M = 128 weight rows
// per MMA) and the expert's token rows are the *B* operand (N = tokens, any multiple of 16 up to 256),
// so an expert with 160 routed tokens costs exactly N=160, not a 128-row token tile padded to 256.
// The accumulator D^T[weight_row, token] lives in TMEM (lane = weight row, column = token).
//
// One CTA per SM, persistent over work items (chunk of tokens of one expert) x (pair of weight tiles):
//   GEMM1: the pair is (gate tile t, up tile t) -> SiLU(gate)*up fused in the epilogue, BF16 act out
//   GEMM2: the pair is two consecutive down-proj tiles -> x routing weight, BF16 c3 out
// Warp roles (384 threads):
//   warp 0      : bulk-TMA producer of packed weight tiles + scale tiles (one elected lane)
//   warp 1      : tcgen05.mma issuer (one elected lane)
//   warps 2-3   : token-row gather (cp.async 16 B, writes the 128B-swizzled K-major B operand)
//   warps 4-7   : dequantisers: packed INT4/INT8 (smem) -> BF16 x group scale -> 128B-swizzled A operand
//   warps 8-11  : epilogue: tcgen05.ld -> activation / weight -> global
// Numerics (oracle/moe.py, GPU-path): W = bf16((nib-8)*scale) exactly as Marlin's BF16 dequant,
// fp32 accumulation in TMEM, BF16 rounding at C1, A and C3.
#include "moe_common.cuh"
#include "ptx.cuh"

namespace kb2 {

constexpr int kNumThreads = 384;
constexpr int kStagesA = 2;   // dequantised A ring
constexpr int kStagesB = 3;   // token (B operand) ring
constexpr int kATileBytes = kTileRows * kBlockK * 2;            // 16 KB (bf16)
constexpr int kAStageBytes = 2 * kATileBytes;                   // two weight tiles per stage
constexpr int kBStageBytes = kMaxChunkTokens * kBlockK * 2;     // 32 KB
constexpr int kTmemCols = 512;
constexpr int kAcc1Col = 256;
constexpr int kNumLoaderThreads = 64;
constexpr int kNumDequantThreads = 128;
constexpr int kNumEpiThreads = 128;

template <int FMT>
struct Fmt;
template <>
struct Fmt<kFmtInt4G128> {
  static constexpr int kTileBytes = kInt4TileBytes;
  static constexpr int kStagesW = 4;   // packed-weight ring depth
};
template <>
struct Fmt<kFmtInt8G128> {
  static constexpr int kTileBytes = kInt8TileBytes;
  static constexpr int kStagesW = 2;   // 16.5 KB per stage: keep the CTA under 227 KB
};

template <int FMT>
struct SmemLayout {
  static constexpr int kStagesW = Fmt<FMT>::kStagesW;
  static constexpr int kWStageBytes = 2 * Fmt<FMT>::kTileBytes + 2 * kScaleTileBytes;
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kOffA + kStagesA * kAStageBytes;
  static constexpr int kOffW = kOffB + kStagesB * kBStageBytes;
  static constexpr int kOffBar = kOffW + kStagesW * kWStageBytes;
  static constexpr int kNumBars = 2 * kStagesW + 2 * kStagesA + 2 * kStagesB + 2;
  static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr int kTotal = kOffTmemPtr + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024 B alignment
  static_assert(kDynamic <= 227 * 1024, "CTA shared memory exceeds the 227 KB sm_100 limit");
};

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) {
      stage = 0;
      phase ^= 1;
    }
  }
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// (w & 0x000F000F) | 0x43004300 == bf16x2 {128 + nib_lo, 128 + nib_hi}; subtracting 136 gives the
// signed value exactly, the multiply by the BF16 group scale then rounds once (RNE) — identical to
// bf16((nib - 8) * scale) of the oracle.
__device__ __forceinline__ uint32_t deq_pair_int4(uint32_t w_shifted, __nv_bfloat162 scale2) {
  uint32_t t = (w_shifted & 0x000F000Fu) | 0x43004300u;
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&t);
  const uint32_t k136 = 0x43084308u;  // bf16 136.0 x2
  v = __hsub2(v, *reinterpret_cast<const __nv_bfloat162*>(&k136));
  v = __hmul2(v, scale2);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int FMT, bool kGemm1>
__global__ void __launch_bounds__(kNumThreads, 1) grouped_gemm_kernel(const GemmParams p) {
  using L = SmemLayout<FMT>;
  constexpr int kStagesW = L::kStagesW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + kStagesW;
  uint64_t* a_full = w_empty + kStagesW;
  uint64_t* a_empty = a_full + kStagesA;
  uint64_t* b_full = a_empty + kStagesA;
  uint64_t* b_empty = b_full + kStagesB;
  uint64_t* tmem_full = b_empty + kStagesB;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kOffTmemPtr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 32) {
    for (int i = 0; i < kStagesW; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], kNumDequantThreads);
    }
    for (int i = 0; i < kStagesA; ++i) {
      mbar_init(&a_full[i], kNumDequantThreads);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kStagesB; ++i) {
      mbar_init(&b_full[i], kNumLoaderThreads);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kNumEpiThreads);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_chunks = *p.n_chunks;
  const int n_items = n_chunks * p.items_per_chunk;
  const int nkb = p.n_kblocks;

  if (warp == 0) {
    // ------------------------------------------------------------ weight-tile producer (bulk TMA)
    if (lane == 0) {
      Ring rw;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
        const int rt = item % p.items_per_chunk;
        const int t0 = rt * p.tile0_mul, t1 = t0 + p.tile1_offset;
        const uint8_t* wq_e = p.wq + (long long)cd.expert * p.wq_expert_stride;
        const uint8_t* ws_e = p.ws + (long long)cd.expert * p.ws_expert_stride;
        const int ngroups = nkb / 2;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&w_empty[rw.stage], rw.phase ^ 1);
          uint8_t* dst = smem + L::kOffW + rw.stage * L::kWStageBytes;
          mbar_arrive_expect_tx(&w_full[rw.stage], L::kWStageBytes);
          constexpr int TB = Fmt<FMT>::kTileBytes;
          bulk_g2s(dst, wq_e + ((long long)t0 * nkb + kb) * TB, TB, &w_full[rw.stage]);
          bulk_g2s(dst + TB, wq_e + ((long long)t1 * nkb + kb) * TB, TB, &w_full[rw.stage]);
          bulk_g2s(dst + 2 * TB, ws_e + ((long long)t0 * ngroups + (kb >> 1)) * kScaleTileBytes, kScaleTileBytes,
                   &w_full[rw.stage]);
          bulk_g2s(dst + 2 * TB + kScaleTileBytes, ws_e + ((long long)t1 * ngroups + (kb >> 1)) * kScaleTileBytes,
                   kScaleTileBytes, &w_full[rw.stage]);
          rw.advance(kStagesW);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      Ring ra, rb;
      uint32_t tphase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
        const uint32_t n_pad = (uint32_t)((cd.n_tok + 15) & ~15);
        const uint32_t idesc = umma_idesc_bf16_m128(n_pad);
        mbar_wait(tmem_empty, tphase ^ 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&a_full[ra.stage], ra.phase);
          mbar_wait(&b_full[rb.stage], rb.phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + L::kOffA + ra.stage * kAStageBytes);
          const uint32_t b_addr = smem_u32(smem + L::kOffB + rb.stage * kBStageBytes);
          const uint64_t a0 = umma_desc_k_sw128(a_addr);
          const uint64_t a1 = umma_desc_k_sw128(a_addr + kATileBytes);
          const uint64_t b0 = umma_desc_k_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            umma_bf16(tmem_base, a0 + 2 * k, b0 + 2 * k, idesc, acc);
            umma_bf16(tmem_base + kAcc1Col, a1 + 2 * k, b0 + 2 * k, idesc, acc);
          }
          umma_commit(&a_empty[ra.stage]);
          umma_commit(&b_empty[rb.stage]);
          ra.advance(kStagesA);
          rb.advance(kStagesB);
        }
        umma_commit(tmem_full);
        tphase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ------------------------------------------------------------ token-row gather (B operand)
    const int lt = threadIdx.x - 64;   // 0..63
    const int c = lt & 7;              // 16-byte chunk within the 128 B k-block row
    const int r0 = lt >> 3;            // row within the 8-row swizzle group
    Ring rb;
    int pending_stage = -1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
      const int n_pad = (cd.n_tok + 15) & ~15;
      const int n_pass = n_pad >> 3;
      long long roff[kMaxChunkTokens / 8];
#pragma unroll
      for (int i = 0; i < kMaxChunkTokens / 8; ++i) {
        int row = i * 8 + r0;
        int slot = cd.slot_begin + (row < cd.n_tok ? row : 0);
        int src_row = p.b_row_index ? p.b_row_index[slot] : slot;
        roff[i] = (long long)src_row * p.b_ld + c * 8;
        if (i >= n_pass) roff[i] = roff[0];
      }
      const uint32_t dst_off = r0 * 128 + ((c ^ r0) << 4);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&b_empty[rb.stage], rb.phase ^ 1);
        uint8_t* dst = smem + L::kOffB + rb.stage * kBStageBytes + dst_off;
        const __nv_bfloat16* src = p.b_src + kb * kBlockK;
#pragma unroll
        for (int i = 0; i < kMaxChunkTokens / 8; ++i) {
          if (i < n_pass) cp_async16(dst + i * 1024, src + roff[i]);
        }
        cp_async_commit();
        if (pending_stage >= 0) {
          cp_async_wait<1>();
          fence_proxy_async_smem();
          mbar_arrive(&b_full[pending_stage]);
        }
        pending_stage = rb.stage;
        rb.advance(kStagesB);
      }
    }
    if (pending_stage >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(&b_full[pending_stage]);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ dequantisers (A operand)
    const int t = threadIdx.x - 128;   // weight row within the tile, 0..127
    Ring rw, ra;
    const uint32_t row_off = (t >> 3) * 1024 + (t & 7) * 128;
    const uint32_t sw = (t & 7);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&w_full[rw.stage], rw.phase);
        mbar_wait(&a_empty[ra.stage], ra.phase ^ 1);
        const uint8_t* wsrc = smem + L::kOffW + rw.stage * L::kWStageBytes;
        uint8_t* adst = smem + L::kOffA + ra.stage * kAStageBytes + row_off;
        constexpr int TB = Fmt<FMT>::kTileBytes;
        const __nv_bfloat16* sc = reinterpret_cast<const __nv_bfloat16*>(wsrc + 2 * TB);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const __nv_bfloat16 s = sc[tile * kTileRows + t];
          const __nv_bfloat162 s2 = __halves2bfloat162(s, s);
          if constexpr (FMT == kFmtInt4G128) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint4 w = *reinterpret_cast<const uint4*>(wsrc + tile * TB + h * 2048 + t * 16);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = deq_pair_int4(ww[j], s2);
                o.y = deq_pair_int4(ww[j] >> 4, s2);
                o.z = deq_pair_int4(ww[j] >> 8, s2);
                o.w = deq_pair_int4(ww[j] >> 12, s2);
                const uint32_t chunk = h * 4 + j;
                *reinterpret_cast<uint4*>(adst + tile * kATileBytes + ((chunk ^ sw) << 4)) = o;
              }
            }
          } else {
            // INT8: [quarter q in 0..3][row][16 B] = 16 consecutive K columns per 16 B
            const float sf = __bfloat162float(s);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 w = *reinterpret_cast<const uint4*>(wsrc + tile * TB + q * 2048 + t * 16);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t word = ww[hh * 2 + (e >> 1)];
                  const int b0 = (int)(int8_t)((word >> ((e & 1) * 16)) & 0xFF);
                  const int b1 = (int)(int8_t)((word >> ((e & 1) * 16 + 8)) & 0xFF);
                  __nv_bfloat162 v = __floats2bfloat162_rn((float)b0 * sf, (float)b1 * sf);
                  o[e] = *reinterpret_cast<uint32_t*>(&v);
                }
                const uint32_t chunk = q * 2 + hh;
                *reinterpret_cast<uint4*>(adst + tile * kATileBytes + ((chunk ^ sw) << 4)) =
                    make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_full[ra.stage]);
        mbar_arrive(&w_empty[rw.stage]);
        rw.advance(kStagesW);
        ra.advance(kStagesA);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;             // weight row within the 128-row tile
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
      const int rt = item % p.items_per_chunk;
      const int n_pad = (cd.n_tok + 15) & ~15;
      mbar_wait(tmem_full, tphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < n_pad; c0 += 16) {
        uint32_t r0v[16], r1v[16];
        tmem_ld16(taddr + c0, r0v);
        tmem_ld16(taddr + kAcc1Col + c0, r1v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int tok = c0 + j;
          if (tok < cd.n_tok) {
            const long long slot = cd.slot_begin + tok;
            const float v0 = __uint_as_float(r0v[j]);
            const float v1 = __uint_as_float(r1v[j]);
            if constexpr (kGemm1) {
              const float g = __bfloat162float(__float2bfloat16_rn(v0));
              const float u = __bfloat162float(__float2bfloat16_rn(v1));
              p.out[slot * p.out_ld + rt * kTileRows + row] = __float2bfloat16_rn(silu_f(g) * u);
            } else {
              const float wgt = p.slot_weight[slot];
              __nv_bfloat16* o = p.out + slot * p.out_ld + rt * 2 * kTileRows + row;
              o[0] = __float2bfloat16_rn(wgt * v0);
              o[kTileRows] = __float2bfloat16_rn(wgt * v1);
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(tmem_empty);
      tphase ^= 1;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

template <int FMT, bool kGemm1>
static cudaError_t launch_one(const GemmParams& p, int num_sms, cudaStream_t stream) {
  auto kern = grouped_gemm_kernel<FMT, kGemm1>;
  constexpr int smem = SmemLayout<FMT>::kDynamic;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<num_sms, kNumThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_grouped_gemm(int fmt, bool gemm1, const GemmParams& p, int num_sms, cudaStream_t stream) {
  if (fmt == kFmtInt4G128) {
    return gemm1 ? launch_one<kFmtInt4G128, true>(p, num_sms, stream)
                 : launch_one<kFmtInt4G128, false>(p, num_sms, stream);
  }
  if (fmt == kFmtInt8G128) {
    return gemm1 ? launch_one<kFmtInt8G128, true>(p, num_sms, stream)
                 : launch_one<kFmtInt8G128, false>(p, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace kb2
