// Gated DeltaNet chunk scan on tcgen05 (dk == dv == 128): the sequential pass of python/krasis/linear_attention.py:43-60
// (_chunk_step) over the chunks of one (value head, 32-wide dv slice) per CTA.
//
//   per chunk c (64 tokens):   VP  = kcd_c  S_c                       (64 x 32)      \  one M=128 tile: rows 0-63 = kcd, rows
//                              IT  = q_c    S_c                       (64 x 32)      /  64-127 = q          [G1]
//                              v   = vcorr_c - VP                                       CUDA cores
//                              dS  = k_c^T (v . e^{g_last - gcum_j})  (128 x 32)      [G2]
//                              O2  = intra_c v                        (64 x 32)       [G3]
//                              S_{c+1} = e^{g_last} S_c + dS                            CUDA cores (S row per thread, fp32 registers)
//                              out = e^{gcum_i} IT + O2  -> BF16
//
// All contractions run as tcgen05.mma kind::f16 (BF16 inputs, fp32 accumulation in TMEM).  fp32 operands (kcd, intra, S, v)
// are carried as a BF16 pair hi + lo (hi = bf16(x), lo = bf16(x - hi), |x - hi - lo| <= 2^-17 |x|) and every product is the
// three-term sum hi*hi + hi*lo + lo*hi, which keeps the recurrence at fp32-grade accuracy like the reference (it runs these
// matmuls in fp32, linear_attention.py:776-779); q and k hold BF16 values and are exact single operands.
//
// Operand staging: q, k tiles by 2-D TMA (128 B swizzle) straight from the prep kernel's outputs (k is consumed MN-major as the
// A operand of G2, so no transposed copy exists); kcd / intra arrive as ready-made 128B-swizzled K-major images written by
// gdn_chunk_prepare_kernel<true> and move with 1-D bulk copies; S and v B-operands are written by the CUDA cores in the same
// swizzled K-major layout.  64-row A operands are issued as M=128 MMAs whose rows 64-127 read whatever follows in shared
// memory: those rows only produce accumulator lanes 64-127 of a column range nobody reads.
//
// 2-stage operand ring; warp roles and the TMEM map are described at the kernel below.
#include <cuda.h>

#include <cstdlib>

#include "mma_sync.cuh"
#include "moe_common.cuh"
#include "ptx.cuh"

namespace kb2 {

cudaError_t make_tmap_bf16_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows);

constexpr int kTC = 64;            // chunk length (linear_attention.py:702)
constexpr int kTSV = 32;           // dv slice per CTA
constexpr int kTD = 128;           // dk == dv
constexpr int kVcLd = 36;          // padded row of the vcorr slice (floats): conflict-free float4 rows

// per-stage byte offsets
constexpr int kOffA1 = 0;                  // [kcd_hi c0 8K][q c0 8K][kcd_hi c1 8K][q c1 8K]
constexpr int kOffA1L = 32768;             // [kcd_lo c0 8K][kcd_lo c1 8K]   (rows 64-127 of the M=128 reads run into A3: ignored lanes)
constexpr int kOffA3 = 49152;              // [intra_hi 8K][intra_lo 8K]     a true 128-row tile: lanes 0-63 = hi part, 64-127 = lo part
constexpr int kOffA2 = 65536;              // k tile, MN-major: [dk 0-63: 64 tokens x 128 B][dk 64-127]
constexpr int kOffVC = 81920;              // vcorr slice fp32 [64][36]
constexpr int kOffG = kOffVC + kTC * kVcLd * 4;   // gcum fp32 [64]
constexpr int kStageBytes = 92160;         // 90 KB (1 KB multiple: every tile base stays 1024-aligned)
static_assert(kOffG + kTC * 4 <= kStageBytes, "stage layout");
constexpr int kTxBytes = 16384 * 5 + kTC * kVcLd * 4 + kTC * 4;
// static region
constexpr int kOffSH = 2 * kStageBytes;    // S hi: [c0: 32 rows x 128 B][c1]
constexpr int kOffSL = kOffSH + 8192;
constexpr int kOffVH = kOffSL + 8192;      // v hi   [32 rows (dv) x 64 tokens]
constexpr int kOffVL = kOffVH + 4096;
constexpr int kOffVDH = kOffVL + 4096;     // decayed v hi
constexpr int kOffVDL = kOffVDH + 4096;
constexpr int kOffX = kOffVDL + 4096;      // epilogue exchange fp32 [64][36]
constexpr int kOffBar = kOffX + kTC * kVcLd * 4;
constexpr int kOffDec = kOffBar + 128;      // e^(g_last - g_t) per token of the running chunk (layout 6)
constexpr int kTcSmem = kOffDec + kTC * 4;
static_assert(kTcSmem <= 227 * 1024, "smem");

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// MN-major operand, SWIZZLE_128B: 64-element (128 B) rows along MN, 8 K-rows per 1024 B atom
__device__ __forceinline__ uint64_t tc_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// byte offset of element (row, k) in a K-major SW128 tile whose 64-element K chunks are `chunk_stride` bytes apart
__device__ __forceinline__ uint32_t sw128_off(int row, int k, int chunk_stride) {
  const int kk = k & 63;
  return (uint32_t)((k >> 6) * chunk_stride + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4) | ((kk & 7) << 1)));
}

// x = hi + lo with hi = bf16_rne(x), lo = bf16_rne(x - hi): |x - hi - lo| <= 2^-17 |x|.  Both conversions go through the packed
// converter (bf16_bits_rn, ptx.cuh): the F2F.BF16.F32 nvcc emitted for the hi half issues at quarter rate and was the bound of the
// tile-writing phases.  (An integer-arithmetic variant with a truncated lo was measured slower, 340 vs 318 us per scan.)
__device__ __forceinline__ void split_bf16(float x, unsigned short& hi, unsigned short& lo) {
  const uint32_t h = bf16_bits_rn(x);
  const uint32_t l = bf16_bits_rn(x - __uint_as_float(h << 16));
  hi = (unsigned short)h;
  lo = (unsigned short)l;
}

// two values that are K-neighbours in an operand image: hi / lo words {low half = element 0, high half = element 1} with one packed
// conversion each (6 instructions per pair instead of 2 x 4 + 2 packs)
__device__ __forceinline__ void split_bf16_pair(float x0, float x1, uint32_t& hi2, uint32_t& lo2) {
  hi2 = bf16x2_bits_rn(x1, x0);
  lo2 = bf16x2_bits_rn(x1 - __uint_as_float(hi2 & 0xFFFF0000u), x0 - __uint_as_float(hi2 << 16));
}

struct GdnTcParams {
  const uint8_t* kcd_img;    // [nv][n_chunks][hi c0 | hi c1 | lo c0 | lo c1] 8 KB each
  const uint8_t* intra_img;  // [nv][n_chunks][hi | lo] 8 KB each
  const float* vcorr;        // [nv][n_chunks][dv/32][64][36]
  const float* gcum;         // [nv][n_chunks][64]
  float* state;              // [nv][dk][dv] in/out
  __nv_bfloat16* core_out;   // [M][nv*dv]
  int M, n_chunks, nv, nk;
  long long* trace;          // optional (tests / tuning): clock64 stamps of CTA (0,0), chunks [8, 16): [chunk][16]
};

// --------------------------------------------------------------------------------------------------------------------
// The scan kernel.  320 threads: warps 0-7 CUDA cores, warp 8 TMA producer, warp 9 MMA issuer (whole warp runs the role loop, one
// elected lane issues: see ptx.cuh / ENGINEERING_NOTES.md on warp-uniform issue).  Core thread = (TMEM lane L, column half): warps w
// and w + 4 share lane quadrant w & 3 and take 16 of the slice's 32 columns each, so every scheduler has two core warps.
//   TMEM columns   D1 [0,64)  = [q ; kcd_hi] [S_hi | S_lo]   (lanes 0-63 = IT parts, 64-127 = VP parts; the two 32-column halves are
//                  D1b [64,96) = [* ; kcd_lo] S_hi             added by the cores: hi/lo halves of a B operand are STACKED ALONG N, one
//                  D2 [96,160) = k^T [vdec_hi | vdec_lo]       N = 64 instruction stream instead of two N = 32 ones: 16 + 4 + 8 MMAs
//                  D3 [160,224) = intra_hi [v_hi | v_lo] (+ intra_lo v_hi on the first half)                      per chunk, not 24 + 8 + 12)
//   per chunk      G1 -> cores: IT rows keep e^g (IT) in registers; VP rows park v = vcorr - VP as fp32 in shared memory
//                  barrier -> every thread (column n = lane, token octet = warp) converts 8 tokens of one column pairwise
//                  (cvt.rn.bf16x2: the packed word is the K-contiguous memory order) and writes v_hi / v_lo / vdec_hi / vdec_lo with
//                  one 16-byte store each -> G2, G3 -> S = e^{g_last} S + dS in registers, new S_hi / S_lo tiles -> epilogue (local).
// History of the layout with the measurements that drove each step: profiles/README.md (r02a ... r02l), ENGINEERING_NOTES.md.
// --------------------------------------------------------------------------------------------------------------------
constexpr int kT3Threads = 320;

__global__ void __launch_bounds__(kT3Threads, 1)
    gdn_scan_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, GdnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* full = bars;           // [2]
  uint64_t* empty = bars + 2;      // [2]
  uint64_t* s_ready = bars + 4;
  uint64_t* g1_done = bars + 5;
  uint64_t* v_ready = bars + 6;
  uint64_t* g2_done = bars + 7;
  uint64_t* g3_done = bars + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);
  // warp index through a shuffle: the role branches are then provably warp-uniform and the operands of tcgen05.mma / commit / TMA
  // stay in uniform registers (from a divergent `tid == 288` branch every MMA paid an ELECT + R2UR.BROADCAST loop, ~50 cycles)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int h = blockIdx.x, sl = blockIdx.y;
  const int kh = h / (p.nv / p.nk);
  const int n_chunks = p.n_chunks;
  const int vd = p.nv * kTD;
  constexpr int kHC = kTSV / 2;      // columns per thread

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
    mbar_init(s_ready, 256);
    mbar_init(g1_done, 1);
    mbar_init(v_ready, 256);
    mbar_init(g2_done, 1);
    mbar_init(g3_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_smem, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kColD1 = 0, kColD1b = 64, kColD2 = 96, kColD3 = 160;     // D1, D2, D3: 64 columns (hi-part | lo-part)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  auto stamp = [&](int c, int slot) {
    if (tracing && c >= 8 && c < 16 && ((tid & 31) == 0 || tid == 64)) p.trace[(c - 8) * 16 + slot] = clock64();
  };

  if (warp == 8) {
    {
      if (elect_one()) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
      }
      const long long hc0 = (long long)h * n_chunks;
      for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        const uint32_t ph = (uint32_t)(c >> 1) & 1u;
        uint8_t* sb = smem + st * kStageBytes;
        mbar_wait(&empty[st], ph ^ 1u);
        const long long hc = hc0 + c;
        const uint8_t* kimg = p.kcd_img + hc * 32768;
        if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], kTxBytes);
        tma_load_2d(sb + kOffA1, &tmap_q, kh * kTD, c * kTC, &full[st]);
        bulk_g2s(sb + kOffA1 + 8192, kimg, 8192, &full[st]);
        tma_load_2d(sb + kOffA1 + 16384, &tmap_q, kh * kTD + 64, c * kTC, &full[st]);
        bulk_g2s(sb + kOffA1 + 24576, kimg + 8192, 8192, &full[st]);
        bulk_g2s(sb + kOffA1L, kimg + 16384, 16384, &full[st]);
        bulk_g2s(sb + kOffA3, p.intra_img + hc * 16384, 16384, &full[st]);
        bulk_g2s(sb + kOffVC, p.vcorr + ((hc * (kTD / kTSV) + sl) * kTC) * kVcLd, kTC * kVcLd * 4, &full[st]);
        bulk_g2s(sb + kOffG, p.gcum + hc * kTC, kTC * 4, &full[st]);
        tma_load_2d(sb + kOffA2, &tmap_k, kh * kTD, c * kTC, &full[st]);
        tma_load_2d(sb + kOffA2 + 8192, &tmap_k, kh * kTD + 64, c * kTC, &full[st]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    {
      const uint32_t id_k = umma_idesc_bf16_m128(kTSV), id_k2 = umma_idesc_bf16_m128(2 * kTSV);
      const uint32_t id_amn2 = umma_idesc_bf16_m128(2 * kTSV) | (1u << 15);
      const uint32_t sS = smem_u32(smem + kOffSH);                 // per K chunk: [S_hi 32 rows | S_lo 32 rows] = one N=64 tile
      const uint32_t vh = smem_u32(smem + kOffVH), vdh = smem_u32(smem + kOffVDH);   // [v_hi | v_lo], [vdec_hi | vdec_lo]
      for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        const uint32_t ph = (uint32_t)(c >> 1) & 1u, cp = (uint32_t)c & 1u;
        const uint32_t sb = smem_u32(smem + st * kStageBytes);
        mbar_wait(&full[st], ph);
        mbar_wait(s_ready, cp);
        tc_fence_after_sync();
        stamp(c, 0);
        // G1: [q ; kcd_hi] [S_hi ; S_lo]^T -> D1 (64 columns);  [* ; kcd_lo] S_hi^T -> D1b (32 columns, lanes 64-127)
        if (elect_one()) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const uint64_t ad = umma_desc_k_sw128(sb + kOffA1 + ch * 16384);
          const uint64_t bd = umma_desc_k_sw128(sS + ch * 8192);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + kColD1, ad + 2 * ks, bd + 2 * ks, id_k2, (ch > 0 || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const uint64_t ad = umma_desc_k_sw128(sb + kOffA1L + ch * 8192 - 8192);
          const uint64_t bd = umma_desc_k_sw128(sS + ch * 8192);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + kColD1b, ad + 2 * ks, bd + 2 * ks, id_k, (ch > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(g1_done);
        }
        __syncwarp();
        stamp(c, 1);
        mbar_wait(v_ready, cp);
        tc_fence_after_sync();
        stamp(c, 2);
        if (elect_one()) {
        {                                                 // G2: dS = k^T [vdec_hi ; vdec_lo]^T  (64 columns)
          const uint64_t bd = umma_desc_k_sw128(vdh);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = tc_desc_mn_sw128(sb + kOffA2 + ks * 2048, 8192, 1024);
            umma_bf16(tmem_base + kColD2, ad, bd + 2 * ks, id_amn2, ks > 0 ? 1u : 0u);
          }
        }
        umma_commit(g2_done);
        {                                                 // G3 (lanes 0-63): intra_hi [v_hi ; v_lo]^T, then intra_lo v_hi^T onto the hi half
          const uint64_t ah = umma_desc_k_sw128(sb + kOffA3), al = umma_desc_k_sw128(sb + kOffA3 + 8192);
          const uint64_t bd = umma_desc_k_sw128(vh);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + kColD3, ah + 2 * ks, bd + 2 * ks, id_k2, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + kColD3, al + 2 * ks, bd + 2 * ks, id_k, 1u);
        }
        umma_commit(g3_done);
        umma_commit(&empty[st]);
        }
        __syncwarp();
        stamp(c, 3);
      }
    }
    __syncwarp();
  } else {
    const int L = (warp & 3) * 32 + (tid & 31);        // TMEM lane of this thread
    const int hcol = warp >> 2;                        // which 16 of the 32 columns
    const int j0 = hcol * kHC;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + j0;
    float s[kHC];                                      // state row L, columns j0 .. j0+15
    {
      const float4* src = reinterpret_cast<const float4*>(p.state + ((long long)h * kTD + L) * kTD + sl * kTSV + j0);
#pragma unroll
      for (int j4 = 0; j4 < kHC / 4; ++j4) {
        const float4 v = src[j4];
        s[4 * j4] = v.x; s[4 * j4 + 1] = v.y; s[4 * j4 + 2] = v.z; s[4 * j4 + 3] = v.w;
      }
    }
    // operand tile addressing: row n = j0 + jj (j0 is a multiple of 8, so (n & 7) == (jj & 7)), k = L or token index
    uint8_t* s_hi = smem + kOffSH + j0 * 128;                       // K chunk (L >> 6): [S_hi rows 0-31 | S_lo rows 0-31]
    uint8_t* s_lo = s_hi + 4096;
    const uint32_t k_off_s = (uint32_t)((L >> 6) * 8192 + ((L & 7) << 1));
    const int kc_s = (L & 63) >> 3;
    auto write_s_tiles = [&]() {
#pragma unroll
      for (int jj = 0; jj < kHC; ++jj) {
        unsigned short hi, lo;
        split_bf16(s[jj], hi, lo);
        const uint32_t off = k_off_s + jj * 128 + ((kc_s ^ (jj & 7)) << 4);
        *reinterpret_cast<unsigned short*>(s_hi + off) = hi;
        *reinterpret_cast<unsigned short*>(s_lo + off) = lo;
      }
    };
    write_s_tiles();
    fence_proxy_async_smem();
    mbar_arrive(s_ready);
    const int i = L & 63;                              // token row (IT for L < 64, VP for L >= 64)
    const int kc_v = i >> 3;
    const uint32_t k_off_v = (uint32_t)(j0 * 128 + ((i & 7) << 1));
    for (int c = 0; c < n_chunks; ++c) {
      const int st = c & 1;
      const uint32_t ph = (uint32_t)(c >> 1) & 1u, cp = (uint32_t)c & 1u;
      const uint8_t* sb = smem + st * kStageBytes;
      const float* sg = reinterpret_cast<const float*>(sb + kOffG);
      const float* svc = reinterpret_cast<const float*>(sb + kOffVC);
      mbar_wait(&full[st], ph);
      const float g_last = sg[kTC - 1], g_i = sg[i];
      const float d_last = expf(g_last);
      float it[kHC];
#pragma unroll
      for (int jj = 0; jj < kHC; ++jj) it[jj] = 0.f;
      mbar_wait(g1_done, cp);
      tc_fence_after_sync();
      if (tid == 0) stamp(c, 4);
      float* xv = reinterpret_cast<float*>(smem + kOffX);            // parked v: fp32 [64 tokens][36]
      if (L >= 64) {                           // VP rows: v = vcorr - VP for this thread's 16 columns, parked as fp32
        uint32_t a[16], a2[16], b[16];
        tmem_ld16(lane_addr + kColD1, a);
        tmem_ld16(lane_addr + kColD1 + kTSV, a2);
        tmem_ld16(lane_addr + kColD1b, b);
        tmem_ld_wait();
        if (tid == 64) stamp(c, 10);
#pragma unroll
        for (int j4 = 0; j4 < kHC / 4; ++j4) {
          const float4 vc4 = *reinterpret_cast<const float4*>(svc + i * kVcLd + j0 + 4 * j4);
          float4 o;
          o.x = vc4.x - ((__uint_as_float(a[4 * j4]) + __uint_as_float(a2[4 * j4])) + __uint_as_float(b[4 * j4]));
          o.y = vc4.y - ((__uint_as_float(a[4 * j4 + 1]) + __uint_as_float(a2[4 * j4 + 1])) + __uint_as_float(b[4 * j4 + 1]));
          o.z = vc4.z - ((__uint_as_float(a[4 * j4 + 2]) + __uint_as_float(a2[4 * j4 + 2])) + __uint_as_float(b[4 * j4 + 2]));
          o.w = vc4.w - ((__uint_as_float(a[4 * j4 + 3]) + __uint_as_float(a2[4 * j4 + 3])) + __uint_as_float(b[4 * j4 + 3]));
          *reinterpret_cast<float4*>(xv + i * kVcLd + j0 + 4 * j4) = o;
        }
        if (hcol == 0) reinterpret_cast<float*>(smem + kOffDec)[i] = expf(g_last - g_i);
      } else {                                 // IT rows
        uint32_t a[16], a2[16];
        tmem_ld16(lane_addr + kColD1, a);
        tmem_ld16(lane_addr + kColD1 + kTSV, a2);
        tmem_ld_wait();
        const float eg = expf(g_i);
#pragma unroll
        for (int jj = 0; jj < kHC; ++jj) it[jj] = eg * (__uint_as_float(a[jj]) + __uint_as_float(a2[jj]));
      }
      if (tid == 64) stamp(c, 11);
      named_bar_sync(1, 256);
      if (tid == 64) stamp(c, 12);
      {                                        // every thread: column n, tokens 8 t8 .. 8 t8 + 7 of v -> one 16-byte store per tile
        const int n = tid & 31, t8 = tid >> 5;
        const float* sdec = reinterpret_cast<const float*>(smem + kOffDec);
        const float4 d0 = *reinterpret_cast<const float4*>(sdec + 8 * t8), d1 = *reinterpret_cast<const float4*>(sdec + 8 * t8 + 4);
        const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = xv[(8 * t8 + e) * kVcLd + n];
        uint32_t vh[4], vl[4], dh[4], dl[4];
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
          const float x0 = x[2 * pr], x1 = x[2 * pr + 1];
          vh[pr] = bf16x2_bits_rn(x1, x0);
          vl[pr] = bf16x2_bits_rn(x1 - __uint_as_float(vh[pr] & 0xFFFF0000u), x0 - __uint_as_float(vh[pr] << 16));
          const float y0 = x0 * dv[2 * pr], y1 = x1 * dv[2 * pr + 1];
          dh[pr] = bf16x2_bits_rn(y1, y0);
          dl[pr] = bf16x2_bits_rn(y1 - __uint_as_float(dh[pr] & 0xFFFF0000u), y0 - __uint_as_float(dh[pr] << 16));
        }
        const uint32_t off = (uint32_t)(n * 128 + ((t8 ^ (n & 7)) << 4));
        *reinterpret_cast<uint4*>(smem + kOffVH + off) = make_uint4(vh[0], vh[1], vh[2], vh[3]);
        *reinterpret_cast<uint4*>(smem + kOffVL + off) = make_uint4(vl[0], vl[1], vl[2], vl[3]);
        *reinterpret_cast<uint4*>(smem + kOffVDH + off) = make_uint4(dh[0], dh[1], dh[2], dh[3]);
        *reinterpret_cast<uint4*>(smem + kOffVDL + off) = make_uint4(dl[0], dl[1], dl[2], dl[3]);
      }
      if (tid == 64) stamp(c, 13);
      tc_fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(v_ready);
      if (tid == 64) stamp(c, 5);
      mbar_wait(g2_done, cp);
      tc_fence_after_sync();
      if (tid == 0) stamp(c, 6);
      {
        uint32_t a[16], a2[16];
        tmem_ld16(lane_addr + kColD2, a);
        tmem_ld16(lane_addr + kColD2 + kTSV, a2);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < kHC; ++jj) s[jj] = fmaf(d_last, s[jj], __uint_as_float(a[jj]) + __uint_as_float(a2[jj]));
      }
      if (c + 1 < n_chunks) {
        write_s_tiles();
        tc_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(s_ready);
      }
      if (tid == 0) stamp(c, 7);
      if (L < 64) {                            // output rows (IT and intra.v share lanes 0-63)
        mbar_wait(g3_done, cp);
        tc_fence_after_sync();
        if (tid == 0) stamp(c, 8);
        uint32_t a[16], a2[16];
        tmem_ld16(lane_addr + kColD3, a);
        tmem_ld16(lane_addr + kColD3 + kTSV, a2);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < kHC; ++jj) a[jj] = __float_as_uint(__uint_as_float(a[jj]) + __uint_as_float(a2[jj]));
        const int t = c * kTC + i;
        if (t < p.M) {
          uint32_t o[kHC / 2];
#pragma unroll
          for (int j2 = 0; j2 < kHC / 2; ++j2) {
            __nv_bfloat162 pr = __floats2bfloat162_rn(it[2 * j2] + __uint_as_float(a[2 * j2]), it[2 * j2 + 1] + __uint_as_float(a[2 * j2 + 1]));
            o[j2] = *reinterpret_cast<uint32_t*>(&pr);
          }
          uint4* dst = reinterpret_cast<uint4*>(p.core_out + (long long)t * vd + h * kTD + sl * kTSV + j0);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        tc_fence_before_sync();
        if (tid == 0) stamp(c, 9);
      }
    }
    {
      float4* dst = reinterpret_cast<float4*>(p.state + ((long long)h * kTD + L) * kTD + sl * kTSV + j0);
#pragma unroll
      for (int j4 = 0; j4 < kHC / 4; ++j4) dst[j4] = make_float4(s[4 * j4], s[4 * j4 + 1], s[4 * j4 + 2], s[4 * j4 + 3]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256);
}




// ====================================================================================================================
// Chunk prepare on tcgen05 (dk == dv == 128): everything of linear_attention.py:593-646 that does not depend on the carried
// state, for one (value head, 64-token chunk) per loop iteration of a persistent CTA:
//   MMA-A   [k ; q] k^T  (M=128 stacked, N=64, K=128; BF16 operands, exact products)  -> k k^T (lanes 0-63), q k^T (lanes 64-127)
//   cores   A = -(beta_i k_i.k_j) e^{g_i-g_j} (j<i);  intra = (q_i.k_j) e^{g_i-g_j} (j<=i) -> hi/lo images to global
//   cores   T = (I - A)^-1 by substitution on T^T (thread c owns ROW c of T, A^T rows are broadcast float4 loads)
//   MMA-B   vcorr = [T'_hi ; T'_lo] v,  T' = T diag(beta)            (v exact BF16, consumed MN-major straight from the prep output)
//   MMA-C   kcd   = [T''_hi ; T''_lo] k, T'' = T diag(beta e^{g})     (k tile reused MN-major)
//   cores   hi-part lanes + lo-part lanes summed through shared memory; vcorr slices / kcd hi-lo images / gcum to global
// The reference solves with the right-hand sides (solve_triangular); forming T explicitly and multiplying is the same linear
// map evaluated in fp32-grade arithmetic (BF16 hi/lo pairs of T, exact BF16 v and k).
// Warps 0-3 CUDA cores (thread = TMEM lane), warp 4 TMA producer, warp 5 MMA issuer; 2-stage input ring.
// ====================================================================================================================
constexpr int kPStage = 49152;                       // [k c0 8K][q c0 8K][k c1 8K][q c1 8K][v dv0-63 8K][v dv64-127 8K]
constexpr int kPOffV = 32768;
constexpr int kPLdAT = 68;                           // row stride (floats) of the fp32 A and T matrices
constexpr int kPOffImg1 = 2 * kPStage;               // [T'_hi 8K][T'_lo 8K]
static_assert(kPOffImg1 % 1024 == 0, "image alignment");
constexpr int kPOffImg2 = kPOffImg1 + 16384;         // [T''_hi][T''_lo]
constexpr int kPLdX = 132;
constexpr int kPOffXB = kPOffImg2 + 16384;           // exchange: lo part of vcorr rows   fp32 [64][132]
constexpr int kPOffXC = kPOffXB + 64 * kPLdX * 4;    // exchange: hi part of kcd rows
// A (fp32 [64][68]) lives in the XB region, T (fp32 [64][68]) and the product scratch P (fp32 [32][36]) in the XC region:
// both matrices are dead once the operand images exist, long before the epilogue uses the exchange buffers.
constexpr int kPOffA = kPOffXB;
constexpr int kPOffT = kPOffXC;
constexpr int kPOffP = kPOffXC + 64 * kPLdAT * 4;
constexpr int kPLdP = 36;
static_assert(64 * kPLdAT * 4 <= 64 * kPLdX * 4 && 64 * kPLdAT * 4 + 32 * kPLdP * 4 <= 64 * kPLdX * 4, "A / T / P fit the exchange buffers");
constexpr int kPOffSc = kPOffXC + 64 * kPLdX * 4;    // gcum[64] | beta[64] | beta*e^gcum[64] | scan scratch[4]
constexpr int kPOffBar = kPOffSc + 1024;
constexpr int kPSmem = kPOffBar + 128;
// version 2: outputs are staged in shared memory and leave by bulk copies (per-lane 16-byte global stores touch 32 different 128 B
// lines per warp instruction: the output phase was 4 K cycles of LSU wavefronts).  vcorr + kcd staging alias A / T / P (dead by then).
constexpr int kP2OffOutV = kPOffXB;                         // fp32 [4 slices][64][36] = 36,864 B, the global layout
constexpr int kP2OffOutK = kP2OffOutV + 4 * kTC * kVcLd * 4;  // [hi c0 | hi c1 | lo c0 | lo c1] 32 KB, the global layout
constexpr int kP2OffSc = kP2OffOutK + 32768;
static_assert(kP2OffSc >= kPOffP + 32 * 36 * 4, "A / T / P stay inside the aliased region");
constexpr int kP2OffBar = kP2OffSc + 1024;
constexpr int kP2OffIntra = kP2OffBar + 128;                // intra [hi 8K | lo 8K]
constexpr int kP2Smem = kP2OffIntra + 16384;
static_assert(kP2Smem <= 227 * 1024 && kP2OffOutK % 16 == 0 && kP2OffIntra % 16 == 0, "smem");
static_assert(kPSmem <= 227 * 1024, "smem");

struct GdnPrepParams {
  long long* trace;          // optional clock64 stamps of CTA 0, loop iterations [2, 10): [iteration][16]
  const float* beta;         // [M][nv]
  const float* g;            // [M][nv]
  uint8_t* kcd_img;
  uint8_t* intra_img;
  float* vcorr;
  float* gcum;
  int M, n_chunks, nv, nk;
};

// --------------------------------------------------------------------------------------------------------------------
// The kernel: EIGHT CUDA-core warps on one unit (thread = TMEM lane x column half, two warps per scheduler: the first version had
// four and every phase ran at IPC ~0.2, profiles/r02f_*), and MMA-B/C stacked so that no partial sums have to be exchanged:  D_B = [T'_hi ; T''_hi] v + [T'_lo ; T''_lo] v  (lanes 0-63 = T' v = vcorr),  D_C = the same A operands against k
// (lanes 64-127 = T'' k = kcd).  Twice the MMAs (16 of N = 128 per unit, still a few hundred cycles), no shared-memory round trip.
// Warp 8 = TMA producer, warp 9 = MMA issuer.
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT3Threads, 1)
    gdn_prepare_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, GdnPrepParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kP2OffBar);
  uint64_t* full = bars;           // [2]
  uint64_t* empty = bars + 2;      // [2]
  uint64_t* a_done = bars + 4;
  uint64_t* img_ready = bars + 5;
  uint64_t* bc_done = bars + 6;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform role branches
  const int n_units = p.nv * p.n_chunks, r = p.nv / p.nk;
  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
    mbar_init(a_done, 1);
    mbar_init(img_ready, 256);
    mbar_init(bc_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kColA = 0, kColB = 128, kColC = 256;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int it, int slot) {
    if (tracing && it >= 2 && it < 10 && (tid & 31) == 0) p.trace[(it - 2) * 16 + slot] = clock64();
  };

  if (warp == 8) {
    {
      if (elect_one()) { prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_k); prefetch_tmap(&tmap_v); }
      int it = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        const int h = u / p.n_chunks, ch = u % p.n_chunks, kh = h / r;
        uint8_t* sb = smem + st * kPStage;
        mbar_wait(&empty[st], ph ^ 1u);
        if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], kPStage);
        tma_load_2d(sb, &tmap_k, kh * kTD, ch * kTC, &full[st]);
        tma_load_2d(sb + 8192, &tmap_q, kh * kTD, ch * kTC, &full[st]);
        tma_load_2d(sb + 16384, &tmap_k, kh * kTD + 64, ch * kTC, &full[st]);
        tma_load_2d(sb + 24576, &tmap_q, kh * kTD + 64, ch * kTC, &full[st]);
        tma_load_2d(sb + kPOffV, &tmap_v, h * kTD, ch * kTC, &full[st]);
        tma_load_2d(sb + kPOffV + 8192, &tmap_v, h * kTD + 64, ch * kTC, &full[st]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    {
      const uint32_t id_a = umma_idesc_bf16_m128(64);                       // A, B K-major, N = 64
      const uint32_t id_bc = umma_idesc_bf16_m128(128) | (1u << 16);        // B MN-major, N = 128
      const uint32_t img1 = smem_u32(smem + kPOffImg1), img2 = smem_u32(smem + kPOffImg2);
      int it = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u, up = (uint32_t)it & 1u;
        const uint32_t sb = smem_u32(smem + st * kPStage);
        mbar_wait(&full[st], ph);
        tc_fence_after_sync();
        stamp(it, 0);
        if (elect_one()) {
#pragma unroll
        for (int chn = 0; chn < 2; ++chn) {
          const uint64_t ad = umma_desc_k_sw128(sb + chn * 16384);          // rows 0-63 k, rows 64-127 q
          const uint64_t bd = umma_desc_k_sw128(sb + chn * 16384);          // N = 64: the k rows only
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem_base + kColA, ad + 2 * ks, bd + 2 * ks, id_a, (chn > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(a_done);
        }
        __syncwarp();
        stamp(it, 1);
        mbar_wait(img_ready, up);
        tc_fence_after_sync();
        stamp(it, 2);
        if (elect_one()) {
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {            // pass 0: [T'_hi ; T''_hi], pass 1: [T'_lo ; T''_lo] onto the same accumulators
          const uint64_t ai = umma_desc_k_sw128(pass ? img2 : img1);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (pass > 0 || ks > 0) ? 1u : 0u;
            umma_bf16(tmem_base + kColB, ai + 2 * ks, tc_desc_mn_sw128(sb + kPOffV + ks * 2048, 8192, 1024), id_bc, acc);   // lanes 0-63: T' v
            umma_bf16(tmem_base + kColC, ai + 2 * ks, tc_desc_mn_sw128(sb + ks * 2048, 16384, 1024), id_bc, acc);            // lanes 64-127: T'' k
          }
        }
        umma_commit(bc_done);
        umma_commit(&empty[st]);
        }
        __syncwarp();
        stamp(it, 3);
      }
    }
    __syncwarp();
  } else {
    const int L = (warp & 3) * 32 + (tid & 31);          // TMEM lane of this thread; warps w and w+4 share lane quadrant w & 3
    const int part = warp >> 2;                          // which half of the columns this thread takes in every phase
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int i = L & 63, lane = tid & 31;
    float* sA = reinterpret_cast<float*>(smem + kPOffA);
    float* sT = reinterpret_cast<float*>(smem + kPOffT);
    float* sP = reinterpret_cast<float*>(smem + kPOffP);
    float* sg = reinterpret_cast<float*>(smem + kP2OffSc);
    float* sbeta = sg + 64;
    float* secol = sg + 128;
    float* sscan = sg + 192;
    auto load_gate = [&](int u, float& b, float& gg) {
      b = 0.f; gg = 0.f;
      if (u < n_units && tid < 64) {
        const int h = u / p.n_chunks, t = (u % p.n_chunks) * kTC + i;
        if (t < p.M) { b = p.beta[(long long)t * p.nv + h]; gg = p.g[(long long)t * p.nv + h]; }
      }
    };
    float b_nxt, g_nxt;
    load_gate(blockIdx.x, b_nxt, g_nxt);
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t up = (uint32_t)it & 1u;
      const int h = u / p.n_chunks, ch = u % p.n_chunks;
      const long long hc = (long long)h * p.n_chunks + ch;
      const float b_i = b_nxt, g_in = g_nxt;
      load_gate(u + gridDim.x, b_nxt, g_nxt);                      // next unit's gates are in flight during this one
      // inclusive scan of g over the 64 tokens (warps 0, 1)
      float gc = g_in;
      if (tid < 64) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float v = __shfl_up_sync(0xffffffffu, gc, o);
          if (lane >= o) gc += v;
        }
        if (tid == 31) sscan[0] = gc;
      }
      named_bar_sync(2, 256);
      // bulk groups of this thread, oldest first: intra(u-1), vcorr(u-1), kcd(u-1).  The A matrix and the intra staging written next
      // alias / reuse only the first two; the kcd copy may still be draining (it is waited for before T is zeroed)
      if (tid == 0) bulk_wait_group_read1();
      if (tid >= 32 && tid < 64) gc += sscan[0];
      if (tid < 64) {
        sg[i] = gc;
        sbeta[i] = b_i;
        secol[i] = b_i * expf(gc);
      }
      named_bar_sync(2, 256);
      if (tid == 0) stamp(it, 4);
      mbar_wait(a_done, up);
      tc_fence_after_sync();
      if (tid == 0) stamp(it, 5);
      const float g_i = sg[i];
      if (L < 64) {                           // k k^T row i, columns [32 part, 32 part + 32)  ->  A
        const float b_row = sbeta[i];
        uint32_t a[32];
        tmem_ld32(lane_addr + kColA + part * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = part * 32 + 4 * j4 + e;
            const float dec = exp_fast_nobranch(g_i - sg[j]);             // unconditional (select below): see ptx.cuh
            v[e] = j < i ? -(__uint_as_float(a[4 * j4 + e]) * b_row) * dec : 0.f;
          }
          *reinterpret_cast<float4*>(sA + i * kPLdAT + part * 32 + 4 * j4) = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {                                // q k^T row i, columns [32 part, 32 part + 32)  ->  intra hi/lo images (global)
        uint8_t* img = smem + kP2OffIntra + i * 128;          // staged: [hi 8K | lo 8K], leaves by one bulk copy
        uint32_t a[32];
        tmem_ld32(lane_addr + kColA + part * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = part * 32 + q8 * 8 + 2 * e;
            const float d0 = exp_fast_nobranch(g_i - sg[j]), d1 = exp_fast_nobranch(g_i - sg[j + 1]);
            const float v0 = j <= i ? __uint_as_float(a[q8 * 8 + 2 * e]) * d0 : 0.f;
            const float v1 = j + 1 <= i ? __uint_as_float(a[q8 * 8 + 2 * e + 1]) * d1 : 0.f;
            split_bf16_pair(v0, v1, hw[e], lw[e]);
          }
          const int chunk = ((part * 4 + q8) ^ (i & 7)) << 4;
          *reinterpret_cast<uint4*>(img + chunk) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(img + 8192 + chunk) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();               // intra image: generic-proxy stores -> visible to the bulk copy
      named_bar_sync(2, 256);                 // A^T complete; D_A fully read
      if (tid == 0) {
        bulk_s2g(p.intra_img + hc * 16384, smem + kP2OffIntra, 16384);
        bulk_commit_group();
        bulk_wait_group_read1();              // everything older than this intra copy (i.e. kcd(u-1)) has left its staging buffer
      }
      named_bar_sync(2, 256);                 // ... which T / P alias
      if (tid == 0) stamp(it, 6);
      // ---- T = (I - A)^-1 for the unit-lower-triangular 64x64 system, blocked 16 -> 32 -> 64, all 128 threads, fp32:
      //   level 0  the four diagonal blocks D_b = (I - A_bb)^-1 by forward substitution (thread = one column of one block)
      //   level 1  T[1][0] = D_1 (A_10 D_0),  T[3][2] = D_3 (A_32 D_2)                      (16x16x16 products)
      //   level 2  T[2:4][0:2] = T_hi (A_lo T_lo)                                           (32x32x32 products)
      for (int idx = tid; idx < 64 * (kPLdAT / 4); idx += 256) reinterpret_cast<float4*>(sT)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
      named_bar_sync(2, 256);
      if (tid == 0) stamp(it, 11);
      if (tid < 64) {
        const int b = tid >> 4, cc = tid & 15;
        const float* Ab = sA + (16 * b) * kPLdAT + 16 * b;
        float x[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          float a0 = (r == cc) ? 1.f : 0.f, a1 = 0.f;
#pragma unroll
          for (int j = 0; j < r; ++j) {
            if (j & 1) a1 = fmaf(Ab[r * kPLdAT + j], x[j], a1);
            else a0 = fmaf(Ab[r * kPLdAT + j], x[j], a0);
          }
          x[r] = a0 + a1;
          sT[(16 * b + r) * kPLdAT + 16 * b + cc] = x[r];
        }
      }
      named_bar_sync(2, 256);
      if (tid == 0) stamp(it, 12);
      // levels 1 and 2 are small dense products: 16x8 output tiles on mma.sync m16n8k8 with 3xTF32 operands (fp32-grade; the CUDA-core
      // dot products they replace took 770 + 2700 of the unit's 10.7 K cycles, profiles/r02t_gdn_prepare_inverse_phases.txt)
      const int fg = lane >> 2, ft = lane & 3;                              // accumulator fragment: rows fg, fg + 8; columns 2 ft, 2 ft + 1
      {                                       // level 1, stage 1: P_p = A[2p+1][2p] D_2p ; stage 2: T[2p+1][2p] = D_2p+1 P_p   (16x16x16)
        const int pp = warp >> 1, ct = warp & 1;                            // warps 0-3: pair pp, 8-column tile ct
        float c1[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        if (warp < 4) {
          warp_mma_tiles<1, true, true>(c1, sA + (32 * pp + 16) * kPLdAT + 32 * pp, kPLdAT, 1, sT + (32 * pp) * kPLdAT + 32 * pp + 8 * ct, kPLdAT, 1,
                                        0, 16);
          float* pd = sP + (16 * pp) * kPLdP + 8 * ct + 2 * ft;
          *reinterpret_cast<float2*>(pd + fg * kPLdP) = make_float2(c1[0][0], c1[0][1]);
          *reinterpret_cast<float2*>(pd + (fg + 8) * kPLdP) = make_float2(c1[0][2], c1[0][3]);
        }
        named_bar_sync(2, 256);
        float c2[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        if (warp < 4)
          warp_mma_tiles<1, true, true>(c2, sT + (32 * pp + 16) * kPLdAT + 32 * pp + 16, kPLdAT, 1, sP + (16 * pp) * kPLdP + 8 * ct, kPLdP, 1, 0, 16);
        named_bar_sync(2, 256);               // every warp has read D / P before T[2p+1][2p] is written next to D
        if (warp < 4) {
          float* td = sT + (32 * pp + 16) * kPLdAT + 32 * pp + 8 * ct + 2 * ft;
          *reinterpret_cast<float2*>(td + fg * kPLdAT) = make_float2(c2[0][0], c2[0][1]);
          *reinterpret_cast<float2*>(td + (fg + 8) * kPLdAT) = make_float2(c2[0][2], c2[0][3]);
        }
      }
      named_bar_sync(2, 256);
      if (tid == 0) stamp(it, 13);
      {                                       // level 2: P = A[32:64][0:32] T[0:32][0:32] ; T[32:64][0:32] = T[32:64][32:64] P   (32x32x32)
        const int rt = warp >> 2, ct = warp & 3;                            // 8 warps = 2 x 4 output tiles of 16 x 8
        float c1[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        warp_mma_tiles<1, true, true>(c1, sA + (32 + 16 * rt) * kPLdAT, kPLdAT, 1, sT + 8 * ct, kPLdAT, 1, 0, 32);
        float* pd = sP + (16 * rt) * kPLdP + 8 * ct + 2 * ft;
        *reinterpret_cast<float2*>(pd + fg * kPLdP) = make_float2(c1[0][0], c1[0][1]);
        *reinterpret_cast<float2*>(pd + (fg + 8) * kPLdP) = make_float2(c1[0][2], c1[0][3]);
        named_bar_sync(2, 256);
        float c2[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        warp_mma_tiles<1, true, true>(c2, sT + (32 + 16 * rt) * kPLdAT + 32, kPLdAT, 1, sP + 8 * ct, kPLdP, 1, 0, 32);
        float* td = sT + (32 + 16 * rt) * kPLdAT + 8 * ct + 2 * ft;
        *reinterpret_cast<float2*>(td + fg * kPLdAT) = make_float2(c2[0][0], c2[0][1]);
        *reinterpret_cast<float2*>(td + (fg + 8) * kPLdAT) = make_float2(c2[0][2], c2[0][3]);
      }
      named_bar_sync(2, 256);
      if (tid == 0) stamp(it, 7);
      {
        // T' = T diag(beta), T'' = T diag(beta e^gcum) as hi/lo K-major SW128 rows: thread (row c = tid & 63, column half tid >> 6)
        const int c = tid & 63, qr = tid >> 6;                   // row c of T, columns [16 qr, 16 qr + 16)
        const float* trow = sT + c * kPLdAT + qr * 16;
        uint8_t* ih = smem + kPOffImg1 + c * 128;                // [T'_hi rows | T''_hi rows]  (A operand of pass 0)
        uint8_t* il = smem + kPOffImg2 + c * 128;                // [T'_lo rows | T''_lo rows]  (A operand of pass 1)
#pragma unroll
        for (int q8 = 0; q8 < 2; ++q8) {
          const float4 ta = *reinterpret_cast<const float4*>(trow + q8 * 8), tb = *reinterpret_cast<const float4*>(trow + q8 * 8 + 4);
          const float tx[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
          uint32_t h1[4], l1[4], h2[4], l2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = qr * 16 + q8 * 8 + 2 * e;
            split_bf16_pair(tx[2 * e] * sbeta[j], tx[2 * e + 1] * sbeta[j + 1], h1[e], l1[e]);
            split_bf16_pair(tx[2 * e] * secol[j], tx[2 * e + 1] * secol[j + 1], h2[e], l2[e]);
          }
          const int chunk = ((qr * 2 + q8) ^ (c & 7)) << 4;
          *reinterpret_cast<uint4*>(ih + chunk) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
          *reinterpret_cast<uint4*>(ih + 8192 + chunk) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
          *reinterpret_cast<uint4*>(il + chunk) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
          *reinterpret_cast<uint4*>(il + 8192 + chunk) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(img_ready);
        if (tid < 64) p.gcum[hc * kTC + i] = g_i;
        if (tid == 0) stamp(it, 8);
      }
      mbar_wait(bc_done, up);
      tc_fence_after_sync();
      if (tid == 0) stamp(it, 9);
      // lanes 0-63 of D_B hold T' v = vcorr, lanes 64-127 of D_C hold T'' k = kcd (the other halves are unused by-products of the
      // stacked A operand): every thread stores 64 columns of its own row, no exchange
      if (L < 64) {
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int q = part * 2 + qq;
          uint32_t a[32];
          tmem_ld32(lane_addr + kColB + q * 32, a);
          tmem_ld_wait();
          float* dst = reinterpret_cast<float*>(smem + kP2OffOutV) + (q * kTC + i) * kVcLd;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(dst + 4 * j4) =
                make_float4(__uint_as_float(a[4 * j4]), __uint_as_float(a[4 * j4 + 1]), __uint_as_float(a[4 * j4 + 2]), __uint_as_float(a[4 * j4 + 3]));
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 128);                 // the four vcorr warps (0, 1, 4, 5): their half leaves without waiting for the kcd splits
        if (tid == 0) {
          bulk_s2g(p.vcorr + hc * (kTD / kTSV) * kTC * kVcLd, smem + kP2OffOutV, 4 * kTC * kVcLd * 4);
          bulk_commit_group();
        }
      } else {
        uint8_t* img = smem + kP2OffOutK + i * 128;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int q = part * 2 + qq;
          uint32_t a[32];
          tmem_ld32(lane_addr + kColC + q * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = q8 * 8 + 2 * e;
              split_bf16_pair(__uint_as_float(a[jj]), __uint_as_float(a[jj + 1]), hw[e], lw[e]);
            }
            const int k8 = q * 4 + q8;          // 8-element group along dk: chunk c = k8 / 8, 16-byte slot k8 % 8
            const int off = (k8 >> 3) * 8192 + (((k8 & 7) ^ (i & 7)) << 4);
            *reinterpret_cast<uint4*>(img + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(img + 16384 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      named_bar_sync(2, 256);                 // staged outputs complete
      if (tid == 0) {
        bulk_s2g(p.kcd_img + hc * 32768, smem + kP2OffOutK, 32768);
        bulk_commit_group();
      }
      if (tid == 0) stamp(it, 10);
    }
  }
  if (tid == 0) bulk_wait_group0();           // every output copy has completed before the CTA (and its shared memory) goes away
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}


cudaError_t launch_gdn_prepare_tc(const void* qn, const void* kn, const void* vc, const float* beta, const float* g,
                                  void* kcd_img, void* intra_img, float* vcorr, float* gcum, int M, int n_chunks, int nk, int nv,
                                  int num_sms, cudaStream_t s) {
  static PerDeviceOnce once;
  if (const int dev = once.pending(); dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(gdn_prepare_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kP2Smem);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  alignas(64) CUtensorMap tq, tk, tv;
  cudaError_t e = make_tmap_bf16_rows(&tq, qn, M, (long long)nk * kTD, kTC);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tk, kn, M, (long long)nk * kTD, kTC);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tv, vc, M, (long long)nv * kTD, kTC);
  if (e != cudaSuccess) return e;
  const char* trv = getenv("KB2_GDN_PREPARE_TRACE");
  long long* trace = trv ? reinterpret_cast<long long*>(strtoull(trv, nullptr, 0)) : nullptr;
  GdnPrepParams p{trace, beta, g, (uint8_t*)kcd_img, (uint8_t*)intra_img, vcorr, gcum, M, n_chunks, nv, nk};
  const int n_units = nv * n_chunks;
  gdn_prepare_tc_kernel<<<n_units < num_sms ? n_units : num_sms, kT3Threads, kP2Smem, s>>>(tq, tk, tv, p);
  return cudaGetLastError();
}

// qn, kn: [M][nk*128] bf16 (prep kernel outputs).  The prepared operands are in the layouts documented in GdnTcParams.
cudaError_t launch_gdn_scan_tc(const void* qn, const void* kn, const void* kcd_img, const void* intra_img, const float* vcorr,
                               const float* gcum, float* state, void* core_out, int M, int n_chunks, int nk, int nv,
                               cudaStream_t s) {
  static PerDeviceOnce once;
  if (const int dev = once.pending(); dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(gdn_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  // tuning only: KB2_GDN_SCAN_TRACE=<device pointer> receives clock64 stamps of CTA (0,0) (scripts/gdn_scan_tune.py)
  const char* tv = getenv("KB2_GDN_SCAN_TRACE");
  long long* trace = tv ? reinterpret_cast<long long*>(strtoull(tv, nullptr, 0)) : nullptr;
  alignas(64) CUtensorMap tq, tk;
  cudaError_t e = make_tmap_bf16_rows(&tq, qn, M, (long long)nk * kTD, kTC);
  if (e != cudaSuccess) return e;
  e = make_tmap_bf16_rows(&tk, kn, M, (long long)nk * kTD, kTC);
  if (e != cudaSuccess) return e;
  GdnTcParams p{(const uint8_t*)kcd_img, (const uint8_t*)intra_img, vcorr, gcum, state, (__nv_bfloat16*)core_out, M, n_chunks, nv, nk, trace};
  gdn_scan_tc_kernel<<<dim3(nv, kTD / kTSV), kT3Threads, kTcSmem, s>>>(tq, tk, p);
  return cudaGetLastError();
}

}  // namespace kb2
