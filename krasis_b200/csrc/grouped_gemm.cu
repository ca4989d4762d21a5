// Grouped W4A16 / W8A16 expert GEMM for MoE prefill on sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces the two `moe_wna16_marlin_gemm` launches + `silu_and_mul` of the reference's
// `fused_marlin_moe` call sequence (python/krasis/gpu_prefill.py:152-226; survey K3/K4/K5).
//
// Orientation ("swap-AB"): per expert the weight matrix is the UMMA *A* operand (M = 128 weight rows
// per MMA) and the expert's token rows are the *B* operand (N = tokens, any multiple of 16 up to 192),
// so an expert with 160 routed tokens costs exactly N=160, not a 128-row token tile padded to 256.
// The accumulator D^T[weight_row, token] lives in TMEM (lane = weight row, column = token).
//
// Dataflow per CTA (one per SM, persistent over (token chunk of one expert) x (pair of weight tiles)):
//   HBM --bulk TMA--> packed INT4/INT8 tiles in smem --dequant warps--> BF16 x group scale
//       --tcgen05.st--> A operand in TENSOR MEMORY (never touches shared memory as BF16)
//   sorted token rows --2-D TMA, 128B swizzle--> B operand in smem
//   tcgen05.mma (A from TMEM, B from smem) --> two FP32 accumulators in TMEM
//   epilogue warps: tcgen05.ld -> SiLU(gate)*up | x routing weight -> BF16 -> smem staging -> 16 B coalesced stores
// Keeping the dequantised A operand out of shared memory matters: a BF16 SS-MMA at M=128 reads
// (128+N)*32 B of smem per N/2 cycles (~115 B/clk at N=160), which together with the dequant stores
// exceeded the 128 B/clk smem port in the first version of this kernel (profiles/r01a_*).
//   GEMM1: the pair is (gate tile t, up tile t) -> SiLU(gate)*up fused in the epilogue, BF16 act out
//   GEMM2: the pair is two consecutive down-proj tiles -> x routing weight, BF16 c3 out
// Warp roles (512 threads):
//   warp 0      : bulk-TMA producer of packed weight tiles + scale tiles (one elected lane)
//   warp 1      : tcgen05.mma issuer (one elected lane)
//   warp 2      : 2-D TMA producer of the token (B) operand (one elected lane)
//   warps 4-7   : dequantisers of weight tile 0 (thread t owns weight row t = TMEM lane t)
//   warps 8-11  : epilogue
//   warps 12-15 : dequantisers of weight tile 1
// TMEM columns: [0,192) acc0 | [192,384) acc1 | [384,512) two A stages x (tile0 32 cols | tile1 32 cols)
// Numerics (oracle/moe.py, GPU-path): W = bf16((nib-8)*scale) exactly as Marlin's BF16 dequant,
// fp32 accumulation in TMEM, BF16 rounding at C1, A and C3.
#include <cuda.h>

#include <cstdlib>

#include "moe_common.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace kb2 {

constexpr int kNumThreads = 512;
constexpr int kStagesA = 2;                                     // A operand stages in TMEM
constexpr int kBStageBytes = kMaxChunkTokens * kBlockK * 2;     // 24 KB
constexpr int kBBoxRows = 32;                                   // TMA box = 32 token rows x 64 K
constexpr int kBBoxBytes = kBBoxRows * kBlockK * 2;             // 4 KB
constexpr int kTmemCols = 512;
constexpr int kAcc1Col = kMaxChunkTokens;                       // 192
constexpr int kATmemCol = 2 * kMaxChunkTokens;                  // 384
constexpr int kATileCols = kBlockK / 2;                         // 32 columns = 128 rows x 64 bf16
constexpr int kNumDequantThreads = 256;
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
  static constexpr int kStagesW = 2;
};
template <>
struct Fmt<kFmtQ8_0> {
  static constexpr int kTileBytes = kQ8_0TileBytes;
  static constexpr int kStagesW = 2;
};
template <>
struct Fmt<kFmtQ4_K> {
  static constexpr int kTileBytes = kQ4KTileBytes;
  static constexpr int kStagesW = 3;
};
template <>
struct Fmt<kFmtAffine8> {
  static constexpr int kTileBytes = kAffine8TileBytes;
  static constexpr int kStagesW = 2;
};
template <int FMT>
constexpr bool kHasScaleTiles = (FMT == kFmtInt4G128 || FMT == kFmtInt8G128);

template <int FMT, bool kGemm1>
struct SmemLayout {
  static constexpr int kStagesW = Fmt<FMT>::kStagesW;
  // token (B operand) ring: 4 stages wherever they fit the 227 KB (everything but the 12 KB Affine8 tiles).  With tcgen05.mma issued
  // at full rate the K loop waited ~200 cycles per k-block on token rows with 3 stages (profiles/r02i_gemm_item_timeline_*).
  static constexpr int kStagesB = (FMT == kFmtAffine8) ? 3 : 4;
  static constexpr int kWStageBytes = 2 * Fmt<FMT>::kTileBytes + (kHasScaleTiles<FMT> ? 2 * kScaleTileBytes : 0);
  static constexpr int kStageRowBytes = 2 * kTileRows * 2;   // epilogue staging: 512 B per token (gate|up or two down tiles)
  static constexpr int kOffB = 0;                                               // 1024-aligned (swizzle atoms)
  static constexpr int kOffStage = kOffB + kStagesB * kBStageBytes;
  static constexpr int kOffW = kOffStage + kMaxChunkTokens * kStageRowBytes;
  static constexpr int kOffSlotW = kOffW + kStagesW * kWStageBytes;             // GEMM2: routing weight of the chunk's token slots (f32)
  static constexpr int kOffBar = kOffSlotW + kMaxChunkTokens * 4;
  static constexpr int kNumBars = 2 * kStagesW + 2 * kStagesA + 2 * kStagesB + 2;
  static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
  static constexpr int kTotal = kOffTmemPtr + 16;
  static_assert(kTotal <= 227 * 1024, "CTA shared memory exceeds the 227 KB sm_100 limit");
  static_assert(kOffW % 16 == 0 && kOffBar % 8 == 0, "alignment");
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

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

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
__global__ void __launch_bounds__(kNumThreads, 1)
    grouped_gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_b) {
  using L = SmemLayout<FMT, kGemm1>;
  constexpr int kStagesW = L::kStagesW;
  constexpr int TB = Fmt<FMT>::kTileBytes;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + kStagesW;
  uint64_t* a_full = w_empty + kStagesW;
  uint64_t* a_empty = a_full + kStagesA;
  uint64_t* b_full = a_empty + kStagesA;
  constexpr int kStagesB = L::kStagesB;
  uint64_t* b_empty = b_full + kStagesB;
  uint64_t* tmem_full = b_empty + kStagesB;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kOffTmemPtr);

  // warp index through a shuffle: the compiler then KNOWS the role branches below are warp-uniform and keeps the operands of the
  // single-thread instructions (tcgen05.mma / commit, TMA) in uniform registers.  Issued from a divergent `lane == 0` branch every
  // tcgen05.mma was wrapped in an ELECT + 5 x R2UR.BROADCAST loop: ~50 cycles of issue per MMA, which made the K loop issue-bound
  // (135 cycles per N=160 MMA in situ against 80 of execution: profiles/r02h_gemm_item_timeline.txt, r02h_mma_uniform_ubench.txt).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 32) {
    if (smem_u32(smem) & 1023u) __trap();   // the 128B-swizzle atoms need a 1024 B aligned base
    for (int i = 0; i < kStagesW; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], kNumDequantThreads);
    }
    for (int i = 0; i < kStagesA; ++i) {
      mbar_init(&a_full[i], kNumDequantThreads);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kStagesB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kNumEpiThreads);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr_smem, kTmemCols);
  if (threadIdx.x == 64) prefetch_tmap(&tmap_b);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_chunks = *p.n_chunks;
  const int n_items = n_chunks * p.items_per_chunk;
  const int nkb = p.n_kblocks;
  // tuning aid: per work item of CTA 0, slot = clock64 stamp or accumulated wait cycles (see scripts/gemm_trace.py)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int it, int slot, long long v) {
    if (tracing && it < 16 && (lane == 0 || threadIdx.x == 256)) p.trace[it * 16 + slot] = v;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ weight-tile producer (bulk TMA)
    {
      Ring rw;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
        const int rt = item % p.items_per_chunk;
        const int t0 = rt * p.tile0_mul, t1 = t0 + p.tile1_offset;
        const uint8_t* wq_e = p.wq + (long long)cd.expert * p.wq_expert_stride;
        const uint8_t* ws_e = p.ws + (long long)cd.expert * p.ws_expert_stride;
        const int ngroups = nkb / 2;
        long long wacc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const long long c0 = tracing ? clock64() : 0;
          mbar_wait(&w_empty[rw.stage], rw.phase ^ 1);
          if (tracing) wacc += clock64() - c0;
          uint8_t* dst = smem + L::kOffW + rw.stage * L::kWStageBytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(&w_full[rw.stage], L::kWStageBytes);
            bulk_g2s(dst, wq_e + ((long long)t0 * nkb + kb) * TB, TB, &w_full[rw.stage]);
            bulk_g2s(dst + TB, wq_e + ((long long)t1 * nkb + kb) * TB, TB, &w_full[rw.stage]);
            if constexpr (kHasScaleTiles<FMT>) {
              bulk_g2s(dst + 2 * TB, ws_e + ((long long)t0 * ngroups + (kb >> 1)) * kScaleTileBytes, kScaleTileBytes,
                       &w_full[rw.stage]);
              bulk_g2s(dst + 2 * TB + kScaleTileBytes, ws_e + ((long long)t1 * ngroups + (kb >> 1)) * kScaleTileBytes,
                       kScaleTileBytes, &w_full[rw.stage]);
            }
          }
          __syncwarp();
          rw.advance(kStagesW);
        }
        stamp(item / (int)gridDim.x, 12, wacc);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp runs the loop, one elected lane issues)
    {
      Ring ra, rb;
      uint32_t tphase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
        const uint32_t n_pad = (uint32_t)((cd.n_tok + 15) & ~15);
        const uint32_t idesc = umma_idesc_bf16_m128(n_pad);
        const int ti = item / (int)gridDim.x;
        stamp(ti, 0, tracing ? clock64() : 0);
        mbar_wait(tmem_empty, tphase ^ 1);
        tc_fence_after_sync();
        stamp(ti, 1, tracing ? clock64() : 0);
        long long wa = 0, wb = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const long long c0 = tracing ? clock64() : 0;
          mbar_wait(&a_full[ra.stage], ra.phase);
          const long long c1 = tracing ? clock64() : 0;
          mbar_wait(&b_full[rb.stage], rb.phase);
          if (tracing) { wa += c1 - c0; wb += clock64() - c1; }
          tc_fence_after_sync();
          const uint32_t a_t = tmem_base + kATmemCol + ra.stage * 2 * kATileCols;
          const uint64_t b0 = umma_desc_k_sw128(smem_u32(smem + L::kOffB + rb.stage * kBStageBytes));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
              umma_bf16_ts(tmem_base, a_t + 8 * k, b0 + 2 * k, idesc, acc);
              umma_bf16_ts(tmem_base + kAcc1Col, a_t + kATileCols + 8 * k, b0 + 2 * k, idesc, acc);
            }
            umma_commit(&a_empty[ra.stage]);
            umma_commit(&b_empty[rb.stage]);
          }
          __syncwarp();
          ra.advance(kStagesA);
          rb.advance(kStagesB);
        }
        if (elect_one()) umma_commit(tmem_full);
        __syncwarp();
        stamp(ti, 2, tracing ? clock64() : 0);
        stamp(ti, 3, wa);
        stamp(ti, 4, wb);
        stamp(ti, 14, (long long)cd.n_tok);
        tphase ^= 1;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ token (B operand) producer: 2-D TMA
    if (p.gather_rows != nullptr) {
      // gather mode: the whole warp keeps the chunk's row indices in registers (6 per lane cover 192 slots); lane 0 issues
      // one gather4 per 4 token rows and k-block.  Slots past n_tok (padding up to a multiple of 16) read whatever valid
      // row index follows in the table: those accumulator columns are never stored.
      Ring rb;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
        const int n_pad = (cd.n_tok + 15) & ~15;
        int idx[kMaxChunkTokens / 32];
#pragma unroll
        for (int q = 0; q < kMaxChunkTokens / 32; ++q) idx[q] = (q * 32 + lane < n_pad) ? p.gather_rows[cd.slot_begin + q * 32 + lane] : 0;
        for (int kb = 0; kb < nkb; ++kb) {
          if (lane == 0) {
            mbar_wait(&b_empty[rb.stage], rb.phase ^ 1);
            mbar_arrive_expect_tx(&b_full[rb.stage], n_pad * kBlockK * 2);
          }
          __syncwarp();
          uint8_t* dst = smem + L::kOffB + rb.stage * kBStageBytes;
#pragma unroll
          for (int q = 0; q < kMaxChunkTokens / 32; ++q) {
            if (q * 32 < n_pad) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const int r0 = __shfl_sync(0xffffffffu, idx[q], 4 * g), r1 = __shfl_sync(0xffffffffu, idx[q], 4 * g + 1);
                const int r2 = __shfl_sync(0xffffffffu, idx[q], 4 * g + 2), r3 = __shfl_sync(0xffffffffu, idx[q], 4 * g + 3);
                if (lane == 0 && q * 32 + 4 * g < n_pad)
                  tma_gather4_2d(dst + (q * 32 + 4 * g) * (kBlockK * 2), &tmap_b, kb * kBlockK, r0, r1, r2, r3, &b_full[rb.stage]);
              }
            }
          }
          rb.advance(kStagesB);
        }
      }
    } else {
      Ring rb;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
        const int n_pad = (cd.n_tok + 15) & ~15;
        const int n_box = (n_pad + kBBoxRows - 1) / kBBoxRows;
        long long bacc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const long long c0 = tracing ? clock64() : 0;
          mbar_wait(&b_empty[rb.stage], rb.phase ^ 1);
          if (tracing) bacc += clock64() - c0;
          uint8_t* dst = smem + L::kOffB + rb.stage * kBStageBytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[rb.stage], n_box * kBBoxBytes);
            for (int b = 0; b < n_box; ++b)
              tma_load_2d(dst + b * kBBoxBytes, &tmap_b, kb * kBlockK, cd.slot_begin + b * kBBoxRows, &b_full[rb.stage]);
          }
          __syncwarp();
          rb.advance(kStagesB);
        }
        stamp(item / (int)gridDim.x, 13, bacc);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // idle
  } else if (warp < 8 || warp >= 12) {
    // ------------------------------------------------------------ dequantisers -> A operand in TMEM
    const int tile = warp >= 12 ? 1 : 0;                 // warps 4-7: weight tile 0, warps 12-15: tile 1
    const int t = (warp & 3) * 32 + lane;                // weight row within the tile == TMEM lane
    Ring rw, ra;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kATmemCol + tile * kATileCols;
    const bool dtr = tracing && threadIdx.x == 128;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      long long ww_ = 0, wae = 0, wst = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const long long c0 = dtr ? clock64() : 0;
        mbar_wait(&w_full[rw.stage], rw.phase);
        if (dtr) ww_ += clock64() - c0;
        const uint8_t* wsrc = smem + L::kOffW + rw.stage * L::kWStageBytes;
        __nv_bfloat16 s = __float2bfloat16_rn(0.f);
        if constexpr (kHasScaleTiles<FMT>) s = reinterpret_cast<const __nv_bfloat16*>(wsrc + 2 * TB)[tile * kTileRows + t];
        uint32_t o[32];
        if constexpr (FMT == kFmtQ8_0) {
          // w = bf16(f32(d) * q): the f32 product of an fp16 and an int8 is exact, so this is one rounding (src/gguf.rs:574-593)
          const uint32_t dd = *reinterpret_cast<const uint32_t*>(wsrc + tile * TB + kTileRows * kBlockK + t * 4);
          const float d0 = __half2float(__ushort_as_half((unsigned short)(dd & 0xFFFF)));
          const float d1 = __half2float(__ushort_as_half((unsigned short)(dd >> 16)));
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 w = *reinterpret_cast<const uint4*>(wsrc + tile * TB + q * 2048 + t * 16);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            const float sf = q < 2 ? d0 : d1;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t word = ww[e >> 1];
              const int b0 = (int)(int8_t)((word >> ((e & 1) * 16)) & 0xFF);
              const int b1 = (int)(int8_t)((word >> ((e & 1) * 16 + 8)) & 0xFF);
              __nv_bfloat162 v = __floats2bfloat162_rn((float)b0 * sf, (float)b1 * sf);
              o[q * 8 + e] = *reinterpret_cast<uint32_t*>(&v);
            }
          }
        } else if constexpr (FMT == kFmtAffine8) {
          // GGUF Q6_K / Q5_K / Q5_0 / Q4_0 decoded at load time into int8 codes + (a, b) per 16 elements:
          // w = bf16(fma(a, code, -b)) — a, b and a*code are exact in f32, so this is the reference's single f32 rounding
          const float4 p0 = *reinterpret_cast<const float4*>(wsrc + tile * TB + kTileRows * kBlockK + t * 32);
          const float4 p1 = *reinterpret_cast<const float4*>(wsrc + tile * TB + kTileRows * kBlockK + t * 32 + 16);
          const float av[4] = {p0.x, p0.z, p1.x, p1.z}, bv[4] = {p0.y, p0.w, p1.y, p1.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 w = *reinterpret_cast<const uint4*>(wsrc + tile * TB + q * 2048 + t * 16);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t word = ww[e >> 1];
              const int b0 = (int)(int8_t)((word >> ((e & 1) * 16)) & 0xFF);
              const int b1 = (int)(int8_t)((word >> ((e & 1) * 16 + 8)) & 0xFF);
              __nv_bfloat162 v = __floats2bfloat162_rn(fmaf(av[q], (float)b0, -bv[q]), fmaf(av[q], (float)b1, -bv[q]));
              o[q * 8 + e] = *reinterpret_cast<uint32_t*>(&v);
            }
          }
        } else if constexpr (FMT == kFmtQ4_K) {
          // w = bf16((d*sc)*q - (dmin*mn)) evaluated like src/gguf.rs:711-730: d*sc and dmin*mn exact in f32,
          // (d*sc)*q exact, one f32 rounding at the subtraction (fmaf), then BF16.
          const uint2 hd = *reinterpret_cast<const uint2*>(wsrc + tile * TB + kTileRows * kBlockK / 2 + t * 8);
          const float d = __half2float(__ushort_as_half((unsigned short)(hd.x & 0xFFFF)));
          const float dmin = __half2float(__ushort_as_half((unsigned short)(hd.x >> 16)));
          const float dl = d * (float)(hd.y & 0xFF), ml = dmin * (float)((hd.y >> 8) & 0xFF);
          const float dh = d * (float)((hd.y >> 16) & 0xFF), mh = dmin * (float)(hd.y >> 24);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 w = *reinterpret_cast<const uint4*>(wsrc + tile * TB + h * 2048 + t * 16);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int bp = 0; bp < 2; ++bp) {           // byte pair (2bp, 2bp+1) of word j = elements l, l+1 (and 32+l, 33+l)
                const uint32_t b0 = (ww[j] >> (16 * bp)) & 0xFF, b1 = (ww[j] >> (16 * bp + 8)) & 0xFF;
                const float l0 = __uint_as_float(0x4B000000u | (b0 & 0xF)) - 8388608.0f;
                const float l1 = __uint_as_float(0x4B000000u | (b1 & 0xF)) - 8388608.0f;
                const float h0 = __uint_as_float(0x4B000000u | (b0 >> 4)) - 8388608.0f;
                const float h1 = __uint_as_float(0x4B000000u | (b1 >> 4)) - 8388608.0f;
                __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf(dl, l0, -ml), fmaf(dl, l1, -ml));
                __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf(dh, h0, -mh), fmaf(dh, h1, -mh));
                const int l = h * 16 + j * 4 + bp * 2;                       // element index of b0's low nibble
                o[l / 2] = *reinterpret_cast<uint32_t*>(&lo);
                o[16 + l / 2] = *reinterpret_cast<uint32_t*>(&hi);
              }
            }
          }
        } else if constexpr (FMT == kFmtInt4G128) {
          const __nv_bfloat162 s2 = __halves2bfloat162(s, s);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 w = *reinterpret_cast<const uint4*>(wsrc + tile * TB + h * 2048 + t * 16);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o[h * 16 + j * 4 + 0] = deq_pair_int4(ww[j], s2);
              o[h * 16 + j * 4 + 1] = deq_pair_int4(ww[j] >> 4, s2);
              o[h * 16 + j * 4 + 2] = deq_pair_int4(ww[j] >> 8, s2);
              o[h * 16 + j * 4 + 3] = deq_pair_int4(ww[j] >> 12, s2);
            }
          }
        } else {
          // INT8 tile: [quarter q in 0..3][row][16 B], 16 consecutive K columns per 16 B
          const float sf = __bfloat162float(s);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 w = *reinterpret_cast<const uint4*>(wsrc + tile * TB + q * 2048 + t * 16);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t word = ww[e >> 1];
              const int b0 = (int)(int8_t)((word >> ((e & 1) * 16)) & 0xFF);
              const int b1 = (int)(int8_t)((word >> ((e & 1) * 16 + 8)) & 0xFF);
              __nv_bfloat162 v = __floats2bfloat162_rn((float)b0 * sf, (float)b1 * sf);
              o[q * 8 + e] = *reinterpret_cast<uint32_t*>(&v);
            }
          }
        }
        mbar_arrive(&w_empty[rw.stage]);     // packed words are now in registers
        const long long c1 = dtr ? clock64() : 0;
        mbar_wait(&a_empty[ra.stage], ra.phase ^ 1);
        const long long c2 = dtr ? clock64() : 0;
        tc_fence_after_sync();
        tmem_st32(lane_base + ra.stage * 2 * kATileCols, o);
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(&a_full[ra.stage]);
        if (dtr) { wae += c2 - c1; wst += clock64() - c2; }
        rw.advance(kStagesW);
        ra.advance(kStagesA);
      }
      if (dtr) {
        const int ti = item / (int)gridDim.x;
        stamp(ti, 8, ww_); stamp(ti, 9, wae); stamp(ti, 10, wst); stamp(ti, 11, clock64());
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    // Phase 1 (holds TMEM): accumulators -> BF16 -> smem staging [token][tile0 128 | tile1 128]; then TMEM is
    // released so the next tile's MMAs overlap phase 2.
    // Phase 2: staging -> (GEMM1: SiLU(gate)*up) -> 16 B coalesced global stores.
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;             // weight row within the 128-row tile
    const int et = threadIdx.x - 256;          // 0..127
    uint8_t* stage = smem + L::kOffStage;
    constexpr int kRowB = L::kStageRowBytes;   // 512
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const ChunkDesc cd = p.chunks[item / p.items_per_chunk];
      const int rt = item % p.items_per_chunk;
      const int n_pad = (cd.n_tok + 15) & ~15;
      if constexpr (!kGemm1) {
        // routing weights of this chunk -> shared memory while the MMAs run (a global load per accumulator column inside the
        // drain loop made phase 1 of the down projection 3x longer than the gate/up one: profiles/r02h_gemm_item_timeline.txt)
        float* sw = reinterpret_cast<float*>(smem + L::kOffSlotW);
        for (int i = et; i < n_pad; i += kNumEpiThreads) sw[i] = (i < cd.n_tok) ? p.slot_weight[cd.slot_begin + i] : 0.f;
        named_bar_sync(1, kNumEpiThreads);
      }
      mbar_wait(tmem_full, tphase);
      tc_fence_after_sync();
      const bool etr = tracing && et == 0;
      const int ti = item / (int)gridDim.x;
      if (etr) stamp(ti, 5, clock64());
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      // 16 accumulator columns (tokens) per step, software-pipelined over two register sets: the tcgen05.ld of the next step is in
      // flight while this step is rounded to BF16 and staged (the exposed TMEM load latency was ~1/3 of this phase, which holds TMEM)
      auto load_cols = [&](int c0, uint32_t (&a)[16], uint32_t (&b)[16]) {
        tmem_ld16(taddr + c0, a);
        tmem_ld16(taddr + kAcc1Col + c0, b);
      };
      auto load_weights = [&](int c0, float (&wv)[16]) {
        if constexpr (!kGemm1) {
          // routing weights of the column group, loaded BEFORE the staging stores of the step (a shared-memory load cannot be moved
          // across those stores by the compiler: possible alias)
          const float4* sw4 = reinterpret_cast<const float4*>(smem + L::kOffSlotW) + (c0 >> 2);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 w4 = sw4[q4];
            wv[4 * q4] = w4.x; wv[4 * q4 + 1] = w4.y; wv[4 * q4 + 2] = w4.z; wv[4 * q4 + 3] = w4.w;
          }
        }
      };
      auto stage_cols = [&](int c0, const uint32_t (&a)[16], const uint32_t (&b)[16], const float (&wv)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v0 = __uint_as_float(a[j]);
          float v1 = __uint_as_float(b[j]);
          if constexpr (!kGemm1) {
            v0 *= wv[j];
            v1 *= wv[j];
          }
          __nv_bfloat16* srow = reinterpret_cast<__nv_bfloat16*>(stage + (c0 + j) * kRowB);
          srow[row] = __float2bfloat16_rn(v0);
          srow[kTileRows + row] = __float2bfloat16_rn(v1);
        }
      };
      {
        uint32_t a0[16], b0[16], a1[16], b1[16];
        float w0[16], w1[16];
        load_cols(0, a0, b0);
        for (int c0 = 0; c0 < n_pad; c0 += 32) {
          const bool second = c0 + 16 < n_pad;
          load_weights(c0, w0);
          if (second) load_weights(c0 + 16, w1);
          tmem_ld_wait();                                      // set 0 is in registers
          if (second) load_cols(c0 + 16, a1, b1);
          stage_cols(c0, a0, b0, w0);
          if (second) {
            tmem_ld_wait();                                    // set 1 is in registers
            if (c0 + 32 < n_pad) load_cols(c0 + 32, a0, b0);
            stage_cols(c0 + 16, a1, b1, w1);
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(tmem_empty);                 // accumulators drained: the next tile's MMAs may start
      if (etr) stamp(ti, 6, clock64());
      tphase ^= 1;
      named_bar_sync(1, kNumEpiThreads);       // staging tile complete
      if constexpr (kGemm1) {
        // 16 outputs per step: gate chunk v and up chunk v of one token -> 8 activations (16 B store)
        constexpr int kVecPerRow = kTileRows / 8;          // 16
        const int n_vec = cd.n_tok * kVecPerRow;
        for (int idx = et; idx < n_vec; idx += kNumEpiThreads) {
          const int tok = idx / kVecPerRow, v = idx % kVecPerRow;
          const uint4 g4 = *reinterpret_cast<const uint4*>(stage + tok * kRowB + v * 16);
          const uint4 u4 = *reinterpret_cast<const uint4*>(stage + tok * kRowB + kTileRows * 2 + v * 16);
          const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g4);
          const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&u4);
          uint4 o4;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 g = __bfloat1622float2(g2[i]), u = __bfloat1622float2(u2[i]);
            o2[i] = __floats2bfloat162_rn(silu_f(g.x) * u.x, silu_f(g.y) * u.y);
          }
          *reinterpret_cast<uint4*>(p.out + (long long)(cd.slot_begin + tok) * p.out_ld + (long long)rt * kTileRows + v * 8) = o4;
        }
      } else {
        constexpr int kVecPerRow = 2 * kTileRows / 8;      // 32
        const int n_vec = cd.n_tok * kVecPerRow;
        for (int idx = et; idx < n_vec; idx += kNumEpiThreads) {
          const int tok = idx / kVecPerRow, v = idx % kVecPerRow;
          const uint4 val = *reinterpret_cast<const uint4*>(stage + tok * kRowB + v * 16);
          *reinterpret_cast<uint4*>(p.out + (long long)(cd.slot_begin + tok) * p.out_ld + (long long)rt * 2 * kTileRows + v * 8) = val;
        }
      }
      named_bar_sync(1, kNumEpiThreads);       // staging free again
      if (etr) stamp(ti, 7, clock64());
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// rows x cols BF16 row-major matrix, box = box_rows x 64 columns, 128 B swizzle (matches the UMMA K-major SW128 atom)
cudaError_t make_tmap_bf16_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return cudaErrorNotSupported;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                  gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// rows x cols 8-bit row-major matrix, box = box_rows x 128 columns (128 B), 128 B swizzle
cudaError_t make_tmap_u8_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return cudaErrorNotSupported;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols};
  cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_tmap), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base),
                  gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

int gemm_b_box_rows() { return kBBoxRows; }

// rows x cols BF16 row-major matrix for TMA gather4: box = 1 row x 64 columns, 128 B swizzle
cudaError_t make_tmap_bf16_gather(void* out_tmap, const void* base, long long rows, long long cols) {
  return make_tmap_bf16_rows(out_tmap, base, rows, cols, 1);
}

template <int FMT, bool kGemm1>
static cudaError_t launch_one(const GemmParams& p, const CUtensorMap& tmap, int num_sms, cudaStream_t stream) {
  auto kern = grouped_gemm_kernel<FMT, kGemm1>;
  constexpr int smem = SmemLayout<FMT, kGemm1>::kTotal;
  static PerDeviceOnce once;
  if (const int dev = once.pending(); dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  KernelSpan ks(kGemm1 ? K_GEMM1_GATE_UP : K_GEMM2_DOWN, stream);
  GemmParams pp = p;
  const char* tv = getenv(kGemm1 ? "KB2_GEMM1_TRACE" : "KB2_GEMM2_TRACE");      // tuning only: device pointer to [16][16] int64
  pp.trace = tv ? reinterpret_cast<long long*>(strtoull(tv, nullptr, 0)) : nullptr;
  kern<<<num_sms, kNumThreads, smem, stream>>>(pp, tmap);
  return cudaGetLastError();
}

cudaError_t launch_grouped_gemm(int fmt, bool gemm1, const GemmParams& p, const void* tmap_b, int num_sms,
                                cudaStream_t stream) {
  const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(tmap_b);
  if (fmt == kFmtInt4G128) {
    return gemm1 ? launch_one<kFmtInt4G128, true>(p, tm, num_sms, stream)
                 : launch_one<kFmtInt4G128, false>(p, tm, num_sms, stream);
  }
  if (fmt == kFmtInt8G128) {
    return gemm1 ? launch_one<kFmtInt8G128, true>(p, tm, num_sms, stream)
                 : launch_one<kFmtInt8G128, false>(p, tm, num_sms, stream);
  }
  if (fmt == kFmtQ8_0) {
    return gemm1 ? launch_one<kFmtQ8_0, true>(p, tm, num_sms, stream) : launch_one<kFmtQ8_0, false>(p, tm, num_sms, stream);
  }
  if (fmt == kFmtQ4_K) {
    return gemm1 ? launch_one<kFmtQ4_K, true>(p, tm, num_sms, stream) : launch_one<kFmtQ4_K, false>(p, tm, num_sms, stream);
  }
  if (fmt >= kFmtAffine8 && fmt <= kFmtQ4_0) {      // Q6_K / Q5_K / Q5_0 / Q4_0 share the decoded-affine tile kernel
    return gemm1 ? launch_one<kFmtAffine8, true>(p, tm, num_sms, stream) : launch_one<kFmtAffine8, false>(p, tm, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace kb2
