// GQA prefill attention for sm_100a: per-head RMSNorm + partial RoPE + FP8 paged-KV append, then a causal
// flash-attention kernel on tcgen05 (QK^T and PV on tensor cores, S/P/O in tensor memory).
//
// Restates python/krasis/attention.py:496-687 (GQAAttention.forward):
//   gqa_prep_kernel   split [q | gate] per head (:531-535), flashinfer.norm.rmsnorm per head (:555-559),
//                     half-split partial RoPE with BF16 tables and BF16 arithmetic (:443-494), K/V cast to the cache
//                     dtype (FP8 E4M3, unscaled) and append to the paged cache, page = 16 tokens, NHD (:582-590, kv_cache.py)
//   gqa_fmha_kernel   BatchPrefillWithPagedKVCacheWrapper.run: causal softmax(q k^T / sqrt(d)) v over the paged FP8
//                     cache (:596-642), then attn * sigmoid(gate) (:665-666)
// The q/k/v/o projections are dense_gemm_kernel.
//
// FMHA layout: one CTA = 128 query rows of one head; KV tiles of 64 keys.  12 warps:
//   warps 0-3  KV loaders: FP8 pages -> BF16 -> 128B-swizzled smem (K as K-major B operand, V as MN-major B operand)
//   warp  4    tcgen05.mma issuer: S = Q K^T (SS), O += P V (P from TMEM)
//   warp  5    Q loader (2-D TMA)
//   warps 8-11 softmax: S (TMEM) -> exp2 -> P (BF16, TMEM); running max with lazy rescale of O in TMEM; epilogue
// TMEM: S [0,64) | P [64,96) | O [128,128+D).
#include <cuda.h>
#include <cuda_fp8.h>

#include "moe_common.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace kb2 {

struct GqaDims {
  int H, nh, nkv, d, rotary_dim, gated;
  float theta, eps;
};

__device__ __forceinline__ float bf16r_(float x) { return bf16_round_rn(x); }

// ------------------------------------------------------------------------------------------------
// prep: one CTA per token, one warp per head (q heads then k heads then v heads).  The BF16 cos / sin of the token's position are
// computed once per CTA (powf + cosf + sinf per (head, pair) was most of this kernel: 20 heads recomputed the same 32 angles at QCN
// geometry, 72 heads the same 64 at Qwen3-235B); rows move as BF16 pairs / FP8 pairs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gqa_prep_kernel(GqaDims g, const __nv_bfloat16* __restrict__ q_raw,   // [M][nh*d*(1+gated)]
                                                       const __nv_bfloat16* __restrict__ k_raw,             // [M][nkv*d]
                                                       const __nv_bfloat16* __restrict__ v_raw,
                                                       const float* __restrict__ q_norm, const float* __restrict__ k_norm,
                                                       const int* __restrict__ positions, const int* __restrict__ kv_indices,
                                                       __nv_bfloat16* __restrict__ q_out,                    // [M][nh*d]
                                                       uint8_t* __restrict__ k_cache, uint8_t* __restrict__ v_cache,  // [P][16][nkv][d] fp8
                                                       int M) {
  const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int d = g.d, d2 = g.rotary_dim / 2;
  const int pos = positions[t];
  const long long slot = ((long long)kv_indices[pos >> 4] * 16 + (pos & 15)) * g.nkv;
  extern __shared__ float sx[];              // [nw][d] rows | cos[d2] | sin[d2]
  float* x = sx + warp * d;
  float* s_cos = sx + nw * d;
  float* s_sin = s_cos + d2;
  for (int i = threadIdx.x; i < d2; i += blockDim.x) {
    const float freq = 1.0f / powf(g.theta, (float)(2 * i) / (float)g.rotary_dim);
    const float ang = (float)pos * freq;
    s_cos[i] = bf16r_(cosf(ang));
    s_sin[i] = bf16r_(sinf(ang));
  }
  __syncthreads();
  for (int hh = warp; hh < g.nh + 2 * g.nkv; hh += nw) {
    const bool is_q = hh < g.nh, is_k = !is_q && hh < g.nh + g.nkv;
    const int h = is_q ? hh : (is_k ? hh - g.nh : hh - g.nh - g.nkv);
    const __nv_bfloat16* src = is_q ? q_raw + (long long)t * g.nh * d * (1 + g.gated) + (long long)h * d * (1 + g.gated)
                                    : (is_k ? k_raw : v_raw) + (long long)t * g.nkv * d + (long long)h * d;
    float ss = 0.f;
    for (int i = 2 * lane; i < d; i += 64) {                        // d is even, rows are 4-byte aligned (d % 8 == 0)
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + i));
      x[i] = v.x;
      x[i + 1] = v.y;
      ss += v.x * v.x;
      ss += v.y * v.y;
    }
    if (!is_q && !is_k) {                    // V: plain cast to FP8
      __syncwarp();
      for (int i = 2 * lane; i < d; i += 64) {
        const uint8_t b0 = (uint8_t)__nv_cvt_float_to_fp8(x[i], __NV_SATFINITE, __NV_E4M3);
        const uint8_t b1 = (uint8_t)__nv_cvt_float_to_fp8(x[i + 1], __NV_SATFINITE, __NV_E4M3);
        *reinterpret_cast<uchar2*>(v_cache + (slot + h) * d + i) = make_uchar2(b0, b1);
      }
      __syncwarp();
      continue;
    }
    const float* nwt = is_q ? q_norm : k_norm;
    if (nwt) {
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = rsqrtf(ss / d + g.eps);
      __syncwarp();
      for (int i = lane; i < d; i += 32) x[i] = bf16r_(x[i] * inv * nwt[i]);
    }
    __syncwarp();
    // RoPE on the first rotary_dim dims: (x1, x2) = (x[i], x[i + d2]); BF16 tables, every product / sum rounded to BF16
    for (int i = lane; i < d2; i += 32) {
      const float c = s_cos[i], s = s_sin[i];
      const float x1 = x[i], x2 = x[i + d2];
      const float r1 = bf16r_(bf16r_(x1 * c) - bf16r_(x2 * s));
      const float r2 = bf16r_(bf16r_(x2 * c) + bf16r_(x1 * s));
      x[i] = r1;
      x[i + d2] = r2;
    }
    __syncwarp();
    if (is_q) {
      __nv_bfloat16* dst = q_out + (long long)t * g.nh * d + (long long)h * d;
      for (int i = 2 * lane; i < d; i += 64) *reinterpret_cast<__nv_bfloat162*>(dst + i) = __floats2bfloat162_rn(x[i], x[i + 1]);
    } else {
      for (int i = 2 * lane; i < d; i += 64) {
        const uint8_t b0 = (uint8_t)__nv_cvt_float_to_fp8(x[i], __NV_SATFINITE, __NV_E4M3);
        const uint8_t b1 = (uint8_t)__nv_cvt_float_to_fp8(x[i + 1], __NV_SATFINITE, __NV_E4M3);
        *reinterpret_cast<uchar2*>(k_cache + (slot + h) * d + i) = make_uchar2(b0, b1);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// FMHA
// ------------------------------------------------------------------------------------------------
constexpr int kFQ = 128;        // query rows per CTA
constexpr int kFK = 64;         // keys per KV tile
constexpr int kFThreads = 512;   // 16 warps: 0 KV TMA, 4 MMA, 5 Q TMA + TMEM, 8-15 softmax (two threads per query row)
constexpr int kFStages = 2;

template <int DQ, int DV>
struct FmhaSmem {
  static constexpr int kQBytes = kFQ * DQ * 2;
  static constexpr int kKBytes = kFK * DQ * 2;
  static constexpr int kVBytes = kFK * DV * 2;
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + kQBytes;
  static constexpr int kOffV = kOffK + kFStages * kKBytes;
  static constexpr int kOffBar = kOffV + kFStages * kVBytes;
  static constexpr int kOffXchg = kOffBar + 128;          // float [2 tile parities][2 halves][128 rows]
  static constexpr int kTotal = kOffXchg + 2 * 2 * kFQ * 4;
  static_assert(kTotal <= 227 * 1024, "smem");
  static_assert(kQBytes % 1024 == 0 && kKBytes % 1024 == 0 && kVBytes % 1024 == 0, "swizzle atoms");
};

__device__ __forceinline__ uint32_t umma_idesc_bf16_m128_bmn(uint32_t n) {   // B operand MN-major
  return umma_idesc_bf16_m128(n) | (1u << 16);
}
// MN-major, SWIZZLE_128B: 64-element (128 B) rows along MN, 8 K-rows per 1024 B atom; LBO = stride between
// 64-wide MN chunks, SBO = stride between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t dsc = 0;
  dsc |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  dsc |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  dsc |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  dsc |= static_cast<uint64_t>(1) << 46;
  dsc |= static_cast<uint64_t>(2) << 61;
  return dsc;
}

__device__ __forceinline__ float ex2_approx(float x) {       // 2^x, MUFU.EX2 (2 ulp); ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// DQ = q/k head dim, DV = v head dim.  K and V are dense BF16 rows (the FP8 pages of the sequence are upcast once per call
// by *_gather_kernel, exactly like the reference's page upcast attention.py:324-337) and stream in by TMA:
//   tmap_k   [kv_len][*]: this head's keys at column k_col0 + kvh * (PE ? DQ - 64 : DQ)
//   tmap_kpe [kv_len][64] (PE only, MLA): the rope part of the key, shared by all heads, is the last 64-wide chunk
//   tmap_v   [kv_len][*]: this head's values at column v_col0 + kvh * DV
template <int DQ, int DV, bool PE>
__global__ void __launch_bounds__(kFThreads, 1)
    gqa_fmha_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_kpe, const __grid_constant__ CUtensorMap tmap_v, GqaDims g,
                    const __nv_bfloat16* __restrict__ q_raw,     // for the output gate (g.gated)
                    __nv_bfloat16* __restrict__ out,             // [M][nh*DV]
                    int M, int q_start, int kv_len, float sm_scale_log2, int k_col0, int v_col0) {
  using L = FmhaSmem<DQ, DV>;
  constexpr int D = DQ;
  constexpr int NC = DQ / 64;                      // 64-wide d chunks of q/k
  constexpr int NKC = PE ? NC - 1 : NC;            // of which from tmap_k
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;       // [2]  K and V stages are filled and released separately: K(it) is free as soon as
  uint64_t* k_empty = bars + 3;      // [2]  QK^T(it) has run, one PV earlier than V(it), so the next K tile's TMA starts
  uint64_t* v_full = bars + 10;      // [2]  a whole tile period earlier (the 64 KB tile loads were the latency chain)
  uint64_t* v_empty = bars + 12;     // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_cons = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;      // [2]: PV(it) commits to pv_done[it & 1] (P is double-buffered)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);
  // shuffle-broadcast warp index: role branches provably warp-uniform -> TMA / tcgen05 operands stay in uniform registers (issued
  // from a divergent `lane == 0` branch each tcgen05.mma pays an ELECT + R2UR.BROADCAST loop of ~50 cycles, more than an N=64 MMA runs)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;       // heavy (late) query tiles first
  const int head = blockIdx.y, kvh = head / (g.nh / g.nkv);
  const int q0 = qt * kFQ;
  // keys visible to this tile: positions <= q_start + min(q0 + 127, M - 1)
  const int last_q = min(q0 + kFQ - 1, M - 1);
  const int n_keys = min(kv_len, q_start + last_q + 1);
  const int n_tiles = (n_keys + kFK - 1) / kFK;

  if (threadIdx.x == 128) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_cons, 256);
    mbar_init(p_full, 256);
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kColS = 0, kColP = 64, kColO = 128;      // P: two buffers of 32 columns at 64 and 96

  if (warp < 4) {
    // ------------------------------------------------------------ KV producer: one thread issues the TMA boxes of a tile
    if (warp == 0) {
      const int kx = k_col0 + kvh * (PE ? DQ - 64 : DQ), vx = v_col0 + kvh * DV;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_tiles; ++it) {
        uint8_t* ks = smem + L::kOffK + stage * L::kKBytes;
        uint8_t* vs = smem + L::kOffV + stage * L::kVBytes;
        mbar_wait(&k_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[stage], L::kKBytes);
#pragma unroll
          for (int c = 0; c < NKC; ++c) tma_load_2d(ks + c * (kFK * 128), &tmap_k, kx + c * 64, it * kFK, &k_full[stage]);
          if constexpr (PE) tma_load_2d(ks + NKC * (kFK * 128), &tmap_kpe, 0, it * kFK, &k_full[stage]);
        }
        __syncwarp();
        mbar_wait(&v_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[stage], L::kVBytes);
#pragma unroll
          for (int c = 0; c < DV / 64; ++c) tma_load_2d(vs + c * (kFK * 128), &tmap_v, vx + c * 64, it * kFK, &v_full[stage]);
        }
        __syncwarp();
        if (++stage == kFStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ------------------------------------------------------------ MMA issuer
    if (n_tiles > 0) {
      const uint32_t id_s = umma_idesc_bf16_m128(kFK);
      const uint32_t id_o = umma_idesc_bf16_m128_bmn(DV);
      mbar_wait(q_full, 0);
      int stage = 0;
      uint32_t phase = 0, sc_phase = 0, pf_phase = 0;
      const uint32_t q_addr = smem_u32(smem + L::kOffQ);
      auto issue_pv = [&](int st, bool first, int pbuf) {
        const uint32_t v_addr = smem_u32(smem + L::kOffV + st * L::kVBytes);
#pragma unroll
        for (int k = 0; k < kFK / 16; ++k) {
          const uint64_t bd = umma_desc_mn_sw128(v_addr + k * 2048, kFK * 128, 1024);
          umma_bf16_ts(tmem_base + kColO, tmem_base + kColP + pbuf * 32 + 8 * k, bd, id_o, (first && k == 0) ? 0u : 1u);
        }
      };
      int prev_stage = 0;
      uint32_t prev_phase = 0;
      for (int it = 0; it < n_tiles; ++it) {
        mbar_wait(&k_full[stage], phase);
        if (it > 0) {
          mbar_wait(s_cons, sc_phase);
          sc_phase ^= 1;
        }
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(smem + L::kOffK + stage * L::kKBytes);
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const uint64_t ad = umma_desc_k_sw128(q_addr + c * (kFQ * 128));
            const uint64_t bd = umma_desc_k_sw128(k_addr + c * (kFK * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + kColS, ad + 2 * k, bd + 2 * k, id_s, (c > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(s_full);
          umma_commit(&k_empty[stage]);              // K(it) is free once QK^T(it) has run
        }
        __syncwarp();
        if (it > 0) {
          mbar_wait(&v_full[prev_stage], prev_phase);
          mbar_wait(p_full, pf_phase);
          pf_phase ^= 1;
          tc_fence_after_sync();
          if (elect_one()) {
            issue_pv(prev_stage, it == 1, (it - 1) & 1);
            umma_commit(&pv_done[(it - 1) & 1]);
            umma_commit(&v_empty[prev_stage]);
          }
          __syncwarp();
        }
        prev_stage = stage;
        prev_phase = phase;
        if (++stage == kFStages) { stage = 0; phase ^= 1; }
      }
      mbar_wait(&v_full[prev_stage], prev_phase);
      mbar_wait(p_full, pf_phase);
      tc_fence_after_sync();
      if (elect_one()) {
        issue_pv(prev_stage, n_tiles == 1, (n_tiles - 1) & 1);
        umma_commit(&pv_done[(n_tiles - 1) & 1]);
        umma_commit(&v_empty[prev_stage]);
      }
      __syncwarp();
    }
    __syncwarp();
  } else if (warp == 5) {
    if (n_tiles > 0 && elect_one()) {
      mbar_arrive_expect_tx(q_full, L::kQBytes);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        tma_load_2d(smem + L::kOffQ + c * (kFQ * 128), &tmap_q, head * D + c * 64, q0, q_full);
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ------------------------------------------------------------ softmax + epilogue: two threads per query row.
    // Warp w serves TMEM lane quarter q = w & 3; warps 8-11 take key columns [0,32) of each tile and the first half of
    // the O columns, warps 12-15 the second halves.  The two threads of a row exchange their tile maxima through smem
    // (one 64-thread named barrier per tile), make identical running-max decisions, and keep partial row sums that
    // are added once at the end.
    const int q = warp & 3, hf = (warp - 8) >> 2;
    const int row = q * 32 + lane;
    const int qi = q0 + row;                       // query index within this call
    const int q_pos = q_start + qi;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* xchg = reinterpret_cast<float*>(smem + L::kOffXchg);
    constexpr int KH = kFK / 2, OH = DV / 2;
    float m_ref = -INFINITY, l_sum = 0.f;
    uint32_t sf_phase = 0;
    // PV completions observed so far on each of the two pv_done barriers.  P is double-buffered, so a tile only has to
    // wait for PV(it-2) (its P buffer is free again) — and for PV(it-1) in the rare case it rescales O.  A thread never
    // lags two completions behind on one barrier: the next PV on it cannot be issued before this thread's p_full arrive.
    uint32_t pv_seen[2] = {0u, 0u};
    auto pv_wait = [&](int b, uint32_t need) {
      while (pv_seen[b] < need) {
        mbar_wait(&pv_done[b], pv_seen[b] & 1);
        ++pv_seen[b];
      }
      tc_fence_after_sync();
    };
    for (int it = 0; it < n_tiles; ++it) {
      mbar_wait(s_full, sf_phase);
      sf_phase ^= 1;
      tc_fence_after_sync();
      float s[KH];                                   // raw scores (the softmax scale is folded into the exponent FMA)
      {
        uint32_t r0[16], r1[16];
        tmem_ld16(lane_addr + kColS + hf * KH, r0);
        tmem_ld16(lane_addr + kColS + hf * KH + 16, r1);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          s[jj] = __uint_as_float(r0[jj]);
          s[16 + jj] = __uint_as_float(r1[jj]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(s_cons);                         // S may be overwritten by the next QK^T
      const int key0 = it * kFK + hf * KH;
      // causal / length mask only where a key of this half-tile can be invisible to a row of this warp (warp-uniform)
      if (key0 + KH - 1 > q_start + q0 + q * 32 || key0 + KH > kv_len) {
#pragma unroll
        for (int jj = 0; jj < KH; ++jj)
          if (key0 + jj > q_pos || key0 + jj >= kv_len) s[jj] = -INFINITY;
      }
      float m_tile = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < KH; ++jj) m_tile = fmaxf(m_tile, s[jj]);
      float* xb = xchg + (it & 1) * (2 * kFQ);
      xb[hf * kFQ + row] = m_tile;
      named_bar_sync(1 + q, 64);
      m_tile = fmaxf(m_tile, xb[(hf ^ 1) * kFQ + row]) * sm_scale_log2;      // sm_scale > 0: max commutes with the scale
      // lazy rescale: keep the reference max until the true max exceeds it by 8 (p stays <= 2^8)
      float scale_o = 1.f;
      bool rescale = false;
      if (m_tile > m_ref + 8.f || m_ref == -INFINITY) {
        const float m_new = fmaxf(m_tile, m_ref);
        if (m_new != -INFINITY) {
          scale_o = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);
          rescale = (m_ref != -INFINITY);
          m_ref = m_new;
        }
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;   // row fully masked so far: every s is -inf -> p = 0
      float psum = 0.f;
      uint32_t pk[KH / 2];
#pragma unroll
      for (int jj = 0; jj < KH; jj += 2) {
        const float p0 = ex2_approx(fmaf(s[jj], sm_scale_log2, neg_m));
        const float p1 = ex2_approx(fmaf(s[jj + 1], sm_scale_log2, neg_m));
        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        psum += __low2float(pb) + __high2float(pb);     // normalise with the values the MMA actually uses
        pk[jj / 2] = *reinterpret_cast<uint32_t*>(&pb);
      }
      l_sum = l_sum * scale_o + psum;
      // tcgen05.ld/st are warp-collective (.sync.aligned): the decision must be warp-uniform; rows that keep their
      // reference max multiply by 1
      if (__any_sync(0xffffffffu, rescale)) {
        pv_wait((it - 1) & 1, (uint32_t)((it - 1) / 2 + 1));     // O must hold every PV up to tile it-1 (rescale => it >= 1)
        const float sc = rescale ? scale_o : 1.f;
#pragma unroll 2
        for (int c0 = hf * OH; c0 < (hf + 1) * OH; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + kColO + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) r[jj] = __float_as_uint(__uint_as_float(r[jj]) * sc);
          tmem_st16(lane_addr + kColO + c0, r);
        }
      }
      if (it >= 2) pv_wait(it & 1, (uint32_t)(it / 2));          // PV(it-2) has read this P buffer
      tmem_st16(lane_addr + kColP + (it & 1) * 32 + hf * (KH / 2), pk);
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(p_full);
    }
    if (n_tiles > 0) {
      float* xb = xchg + (n_tiles & 1) * (2 * kFQ);   // the buffer the last tile did not use
      xb[hf * kFQ + row] = l_sum;
      named_bar_sync(1 + q, 64);
      l_sum += xb[(hf ^ 1) * kFQ + row];
      if (n_tiles >= 2) pv_wait((n_tiles - 2) & 1, (uint32_t)((n_tiles - 2) / 2 + 1));
      pv_wait((n_tiles - 1) & 1, (uint32_t)((n_tiles - 1) / 2 + 1));
      const float inv_l = l_sum > 0.f ? 1.0f / l_sum : 0.f;
      const bool live = qi < M;
      const long long obase = (long long)qi * g.nh * DV + (long long)head * DV;
      const long long gbase = (long long)qi * g.nh * D * 2 + (long long)head * 2 * D + D;   // gate half of q_raw
      const bool gated = g.gated != 0;
#pragma unroll 2
      for (int c0 = hf * OH; c0 < (hf + 1) * OH; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(lane_addr + kColO + c0, r);
        tmem_ld_wait();
        if (live) {
          uint32_t o[8];
          uint4 gq[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
          if (gated) {
            gq[0] = *reinterpret_cast<const uint4*>(q_raw + gbase + c0);
            gq[1] = *reinterpret_cast<const uint4*>(q_raw + gbase + c0 + 8);
          }
          const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(gq);
#pragma unroll
          for (int jj = 0; jj < 16; jj += 2) {
            float a0 = bf16r_(__uint_as_float(r[jj]) * inv_l), a1 = bf16r_(__uint_as_float(r[jj + 1]) * inv_l);
            if (gated) {
              a0 *= bf16r_(1.0f / (1.0f + expf(-__bfloat162float(gp[jj]))));
              a1 *= bf16r_(1.0f / (1.0f + expf(-__bfloat162float(gp[jj + 1]))));
            }
            __nv_bfloat162 ob = __floats2bfloat162_rn(a0, a1);
            o[jj / 2] = *reinterpret_cast<uint32_t*>(&ob);
          }
          uint4* dst = reinterpret_cast<uint4*>(out + obase + c0);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

__device__ __forceinline__ void fp8x16_to_bf16(const uint4& q, uint4& lo, uint4& hi) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  uint32_t o[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int hp = 0; hp < 2; ++hp) {
      const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((w[i] >> (16 * hp)) & 0xFFFF), __NV_E4M3);
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
      __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);        // every E4M3 value is exact in BF16
      o[i * 2 + hp] = *reinterpret_cast<uint32_t*>(&b);
    }
  }
  lo = make_uint4(o[0], o[1], o[2], o[3]);
  hi = make_uint4(o[4], o[5], o[6], o[7]);
}

// FP8 pages of the sequence -> dense BF16 rows k_out/v_out [T][nkv*d] (every E4M3 value is exact in BF16)
__global__ void __launch_bounds__(256) gqa_gather_kernel(const uint8_t* __restrict__ k_cache, const uint8_t* __restrict__ v_cache,
                                                         const int* __restrict__ kv_indices, int row_bytes, int T,
                                                         __nv_bfloat16* __restrict__ k_out, __nv_bfloat16* __restrict__ v_out) {
  const int pieces = row_bytes / 16;                       // 16 fp8 per piece
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)T * pieces * 2) return;
  const int which = (int)(idx / ((long long)T * pieces));  // 0 = K, 1 = V
  const long long r = idx % ((long long)T * pieces);
  const int key = (int)(r / pieces), p = (int)(r % pieces);
  const long long slot = (long long)kv_indices[key >> 4] * 16 + (key & 15);
  const uint4 q = *reinterpret_cast<const uint4*>((which ? v_cache : k_cache) + slot * row_bytes + p * 16);
  uint4 lo, hi;
  fp8x16_to_bf16(q, lo, hi);
  uint4* dst = reinterpret_cast<uint4*>((which ? v_out : k_out) + (long long)key * row_bytes + p * 16);
  dst[0] = lo;
  dst[1] = hi;
}

cudaError_t make_tmap_bf16_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows);

// k_bf / v_bf: scratch [kv_len][nkv*d] bf16
cudaError_t launch_gqa_core(const GqaDims& g, const void* q_raw, const void* k_raw, const void* v_raw, const float* q_norm,
                            const float* k_norm, const int* positions, const int* kv_indices, void* q_rot, void* k_cache,
                            void* v_cache, void* k_bf, void* v_bf, void* attn_out, int M, int q_start, int kv_len,
                            cudaStream_t s) {
  if (g.d != 128 && g.d != 256) return cudaErrorInvalidValue;
  static PerDeviceOnce once;
  if (const int dev = once.pending(); dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(gqa_fmha_kernel<128, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FmhaSmem<128, 128>::kTotal);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gqa_fmha_kernel<256, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FmhaSmem<256, 256>::kTotal);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  { KernelSpan ks(K_GQA_PREP, s);
  gqa_prep_kernel<<<M, 256, (8 * g.d + g.rotary_dim) * sizeof(float), s>>>(g, (const __nv_bfloat16*)q_raw, (const __nv_bfloat16*)k_raw,
                                                          (const __nv_bfloat16*)v_raw, q_norm, k_norm, positions,
                                                          kv_indices, (__nv_bfloat16*)q_rot, (uint8_t*)k_cache,
                                                          (uint8_t*)v_cache, M); }
  const int row = g.nkv * g.d;
  const long long n_thr = (long long)kv_len * (row / 16) * 2;
  { KernelSpan ks(K_KV_GATHER, s);
  gqa_gather_kernel<<<(unsigned)((n_thr + 255) / 256), 256, 0, s>>>((const uint8_t*)k_cache, (const uint8_t*)v_cache, kv_indices,
                                                                    row, kv_len, (__nv_bfloat16*)k_bf, (__nv_bfloat16*)v_bf); }
  alignas(64) CUtensorMap tq, tk, tv;
  cudaError_t e = make_tmap_bf16_rows(&tq, q_rot, M, (long long)g.nh * g.d, kFQ);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tk, k_bf, kv_len, row, kFK);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tv, v_bf, kv_len, row, kFK);
  if (e != cudaSuccess) return e;
  const float sl2 = (1.0f / sqrtf((float)g.d)) * 1.4426950408889634f;
  dim3 grid((M + kFQ - 1) / kFQ, g.nh);
  KernelSpan ks(K_FMHA, s);
  if (g.d == 256)
    gqa_fmha_kernel<256, 256, false><<<grid, kFThreads, FmhaSmem<256, 256>::kTotal, s>>>(
        tq, tk, tk, tv, g, (const __nv_bfloat16*)q_raw, (__nv_bfloat16*)attn_out, M, q_start, kv_len, sl2, 0, 0);
  else
    gqa_fmha_kernel<128, 128, false><<<grid, kFThreads, FmhaSmem<128, 128>::kTotal, s>>>(
        tq, tk, tk, tv, g, (const __nv_bfloat16*)q_raw, (__nv_bfloat16*)attn_out, M, q_start, kv_len, sl2, 0, 0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// MLA (DeepSeek-V2 / Kimi) prefill — python/krasis/attention.py:213-374 (MLAAttention.forward)
//   mla_prep_kernel    KV LayerNorm on the compressed latent (:243-246), de-interleave + RoPE (YaRN inv_freq table from the
//                      host, :119-163,165-211) of k_pe and of every head's q_pe (in place in q_full), cast to FP8 and append
//                      to the paged latent cache ckv [P][16][lora] / kpe [P][16][rope] (:282-309)
//   mla_gather_kernel  FP8 pages of the whole sequence -> dense BF16 rows (the reference's per-call page upcast, :324-337)
//   dense_gemm         k_nope | v = ckv @ [w_kc ; w_vc]^T  — the non-absorbed form: for prefill it needs (192+128)/(576+512)
//                      of the attention FLOPs of the absorbed form the reference runs (:268-271,364-368); same math, the
//                      BF16 rounding moves from q_nope_absorbed / attn_out to k_nope / v
//   gqa_fmha_kernel<192,128,1>  causal softmax(q k^T * sm_scale) v, 16 heads, shared rope key
// ------------------------------------------------------------------------------------------------
struct MlaDims {
  int H, nh, nope, rope, dv, lora;
  float eps;
};

__global__ void __launch_bounds__(256) mla_prep_kernel(MlaDims m, const __nv_bfloat16* __restrict__ kv_a,   // [M][lora+rope]
                                                       __nv_bfloat16* __restrict__ q_full,                 // [M][nh*(nope+rope)]
                                                       const float* __restrict__ kv_norm_w,
                                                       const float* __restrict__ inv_freq,                 // [rope/2]
                                                       const int* __restrict__ positions,
                                                       const int* __restrict__ kv_indices,
                                                       uint8_t* __restrict__ ckv_cache, uint8_t* __restrict__ kpe_cache,
                                                       int M) {
  __shared__ float red[8];
  const int t = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = positions[t];
  const long long slot = (long long)kv_indices[pos >> 4] * 16 + (pos & 15);
  const __nv_bfloat16* row = kv_a + (long long)t * (m.lora + m.rope);
  // KV LayerNorm (flashinfer.norm.rmsnorm: fp32, output BF16) then the cache cast
  float ss = 0.f;
  for (int i = tid; i < m.lora; i += 256) {
    const float v = __bfloat162float(row[i]);
    ss += v * v;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss += red[i];
  const float inv = rsqrtf(ss / m.lora + m.eps);
  for (int i = tid; i < m.lora; i += 256) {
    const float y = bf16r_(__bfloat162float(row[i]) * inv * kv_norm_w[i]);
    ckv_cache[slot * m.lora + i] = (uint8_t)__nv_cvt_float_to_fp8(y, __NV_SATFINITE, __NV_E4M3);
  }
  // RoPE: interleaved (re, im) pairs -> half-split, BF16 tables, every product / sum rounded to BF16 (:190-211)
  const int half = m.rope / 2, hd = m.nope + m.rope;
  for (int u = warp; u < m.nh + 1; u += 8) {         // unit nh = the shared key
    __nv_bfloat16* src = u < m.nh ? q_full + (long long)t * m.nh * hd + (long long)u * hd + m.nope : nullptr;
    for (int i0 = 0; i0 < half; i0 += 32) {
      const int i = i0 + lane;
      float x1 = 0.f, x2 = 0.f;
      if (i < half) {
        x1 = __bfloat162float(u < m.nh ? src[2 * i] : row[m.lora + 2 * i]);
        x2 = __bfloat162float(u < m.nh ? src[2 * i + 1] : row[m.lora + 2 * i + 1]);
      }
      __syncwarp();                                  // all lanes have read before anyone overwrites (in place for q)
      if (i < half) {
        const float ang = (float)pos * inv_freq[i];
        const float c = bf16r_(cosf(ang)), sn = bf16r_(sinf(ang));
        const float r1 = bf16r_(bf16r_(x1 * c) - bf16r_(x2 * sn));
        const float r2 = bf16r_(bf16r_(x2 * c) + bf16r_(x1 * sn));
        if (u < m.nh) {
          src[i] = __float2bfloat16_rn(r1);
          src[half + i] = __float2bfloat16_rn(r2);
        } else {
          kpe_cache[slot * m.rope + i] = (uint8_t)__nv_cvt_float_to_fp8(r1, __NV_SATFINITE, __NV_E4M3);
          kpe_cache[slot * m.rope + half + i] = (uint8_t)__nv_cvt_float_to_fp8(r2, __NV_SATFINITE, __NV_E4M3);
        }
      }
    }
  }
}

// one warp per key: ckv_bf16 [T][lora], kpe_bf16 [T][rope]   (lora % 16 == 0, rope % 16 == 0)
__global__ void __launch_bounds__(256) mla_gather_kernel(const uint8_t* __restrict__ ckv_cache,
                                                         const uint8_t* __restrict__ kpe_cache,
                                                         const int* __restrict__ kv_indices, int lora, int rope, int T,
                                                         __nv_bfloat16* __restrict__ ckv_out,
                                                         __nv_bfloat16* __restrict__ kpe_out) {
  const int key = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (key >= T) return;
  const long long slot = (long long)kv_indices[key >> 4] * 16 + (key & 15);
  for (int p = lane; p < (lora + rope) / 16; p += 32) {
    const bool is_c = p < lora / 16;
    const int off = is_c ? p * 16 : (p - lora / 16) * 16;
    const uint4 q = *reinterpret_cast<const uint4*>(is_c ? ckv_cache + slot * lora + off : kpe_cache + slot * rope + off);
    uint4 lo, hi;
    fp8x16_to_bf16(q, lo, hi);
    uint4* dst = reinterpret_cast<uint4*>(is_c ? ckv_out + (long long)key * lora + off : kpe_out + (long long)key * rope + off);
    dst[0] = lo;
    dst[1] = hi;
  }
}

cudaError_t launch_dense_gemm(const void* x, const void* w, void* out, const float* bias, int M, int N, int K,
                              long long ldo, bool out_f32, int num_sms, cudaStream_t s);

// q_full [M][nh*192] (q_proj output; rotated in place), kv_a [M][lora+rope]; w_kv [nh*(nope+dv)][lora] = [w_kc ; w_vc];
// scratch: ckv_bf16 [T][lora], kpe_bf16 [T][rope], kv_up [T][nh*(nope+dv)]; attn_out [M][nh*dv]
cudaError_t launch_mla_core(const MlaDims& m, void* q_full, const void* kv_a, const float* kv_norm_w, const float* inv_freq,
                            const void* w_kv, const int* positions, const int* kv_indices, void* ckv_cache, void* kpe_cache,
                            void* ckv_bf16, void* kpe_bf16, void* kv_up, void* attn_out, int M, int q_start, int kv_len,
                            float sm_scale, int num_sms, cudaStream_t s) {
  if (m.nope != 128 || m.rope != 64 || m.dv != 128 || m.lora % 64) return cudaErrorInvalidValue;
  static PerDeviceOnce once;
  if (const int dev = once.pending(); dev >= 0) {
    cudaError_t e0 = cudaFuncSetAttribute(gqa_fmha_kernel<192, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FmhaSmem<192, 128>::kTotal);
    if (e0 != cudaSuccess) return e0;
    once.mark(dev);
  }
  { KernelSpan ks(K_MLA_PREP, s);
  mla_prep_kernel<<<M, 256, 0, s>>>(m, (const __nv_bfloat16*)kv_a, (__nv_bfloat16*)q_full, kv_norm_w, inv_freq, positions,
                                    kv_indices, (uint8_t*)ckv_cache, (uint8_t*)kpe_cache, M); }
  { KernelSpan ks(K_KV_GATHER, s);
  mla_gather_kernel<<<(kv_len + 7) / 8, 256, 0, s>>>((const uint8_t*)ckv_cache, (const uint8_t*)kpe_cache, kv_indices, m.lora,
                                                     m.rope, kv_len, (__nv_bfloat16*)ckv_bf16, (__nv_bfloat16*)kpe_bf16); }
  const int up = m.nh * (m.nope + m.dv);
  cudaError_t e = launch_dense_gemm(ckv_bf16, w_kv, kv_up, nullptr, kv_len, up, m.lora, up, false, num_sms, s);
  if (e != cudaSuccess) return e;
  alignas(64) CUtensorMap tq, tk, tpe;
  e = make_tmap_bf16_rows(&tq, q_full, M, (long long)m.nh * (m.nope + m.rope), kFQ);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tk, kv_up, kv_len, up, kFK);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tpe, kpe_bf16, kv_len, m.rope, kFK);
  if (e != cudaSuccess) return e;
  GqaDims g{m.H, m.nh, m.nh, m.nope + m.rope, 0, 0, 0.f, m.eps};
  const float sl2 = sm_scale * 1.4426950408889634f;
  dim3 grid((M + kFQ - 1) / kFQ, m.nh);
  KernelSpan ks(K_FMHA, s);
  gqa_fmha_kernel<192, 128, true><<<grid, kFThreads, FmhaSmem<192, 128>::kTotal, s>>>(
      tq, tk, tpe, tk, g, nullptr, (__nv_bfloat16*)attn_out, M, q_start, kv_len, sl2, 0, m.nh * m.nope);
  return cudaGetLastError();
}

}  // namespace kb2
