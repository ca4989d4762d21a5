// Router, token binning and combine kernels for MoE prefill (HBM-bound integer / elementwise work),
// plus the load-time re-tiling of the reference's quantised expert weights.
//
//   router_logits_kernel   layer.py:532-534   logits = hidden.float() @ gate.float().T (+bias), fp32 FMA
//   router_topk_kernel     layer.py:536-560   softmax|sigmoid|gpt-oss scoring, top-k (ties -> lower index,
//                                              src/moe.rs:3116-3128), optional renorm
//   count / scan / scatter                     replaces sglang moe_align_block_size (gpu_prefill.py:118):
//                                              bins (token, k) slots into per-expert contiguous runs
//   combine_kernel                             moe_sum_reduce (gpu_prefill.py:238) + rsf*out + shared
//                                              (gpu_prefill.py:4471-4482)
#include "moe_common.cuh"
#include "prof.cuh"
#include "ptx.cuh"

namespace kb2 {

// ------------------------------------------------------------------------------------------------
// Router logits: C[M,E] = A[M,H] (bf16) * B[E,H]^T (bf16), fp32 FMA, k ascending per output.
// 64x64 tile, 256 threads, 4x4 micro-tile.
// ------------------------------------------------------------------------------------------------
constexpr int RT = 64, RK = 32;

__global__ void __launch_bounds__(256) router_logits_kernel(const __nv_bfloat16* __restrict__ h,
                                                            const __nv_bfloat16* __restrict__ gate,
                                                            const float* __restrict__ gate_bias,
                                                            float* __restrict__ logits, int M, int E, int H) {
  __shared__ float hs[RK][RT + 4];
  __shared__ float gs[RK][RT + 4];
  const int m0 = blockIdx.y * RT, e0 = blockIdx.x * RT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < H; k0 += RK) {
    // 64 rows x 32 k of each operand: 2048 elements / 256 threads = 8 each (one 16 B load)
    {
      const int r = threadIdx.x >> 2, kc = (threadIdx.x & 3) * 8;
      uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
      if (m0 + r < M) va = *reinterpret_cast<const uint4*>(h + (long long)(m0 + r) * H + k0 + kc);
      if (e0 + r < E) vb = *reinterpret_cast<const uint4*>(gate + (long long)(e0 + r) * H + k0 + kc);
      const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&va);
      const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&vb);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hs[kc + i][r] = __bfloat162float(pa[i]);
        gs[kc + i][r] = __bfloat162float(pb[i]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&hs[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&gs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = e0 + tx * 4 + j;
      if (e < E) logits[(long long)m * E + e] = acc[i][j] + (gate_bias ? gate_bias[e] : 0.0f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Scoring + top-k: one warp per token; E <= 1024, E % 32 == 0 not required.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPerLane = 32;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int PER_LANE>
__global__ void __launch_bounds__(128) router_topk_kernel(const float* __restrict__ logits,
                                                          const float* __restrict__ corr_bias, int M, int E,
                                                          int top_k, int scoring, int renorm,
                                                          int* __restrict__ ids, float* __restrict__ wts) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const float* lg = logits + (long long)m * E;
  float score[PER_LANE], key[PER_LANE];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int e = i * 32 + lane;
    score[i] = e < E ? lg[e] : -INFINITY;
    mx = fmaxf(mx, score[i]);
  }
  if (scoring == 0) {           // softmax over all experts
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
      const int e = i * 32 + lane;
      score[i] = e < E ? expf(score[i] - mx) : 0.f;
      s += score[i];
    }
    s = warp_sum(s);
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) score[i] = score[i] / s;
  } else if (scoring == 1) {    // sigmoid
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) score[i] = 1.0f / (1.0f + expf(-score[i]));
  }                              // scoring == 2: raw logits (GPT-OSS), softmax over the selected k below
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int e = i * 32 + lane;
    key[i] = e < E ? score[i] + ((corr_bias && scoring != 2) ? corr_bias[e] : 0.f) : -INFINITY;
  }
  float wsum = 0.f, my_w = 0.f;
  int my_id = 0;
  float first_val = 0.f;
  for (int j = 0; j < top_k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    float bs = 0.f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {       // ascending expert index within the lane: strict '>' keeps the lower index
      if (key[i] > bv) {
        bv = key[i];
        bi = i * 32 + lane;
        bs = score[i];
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
        bs = os;
      }
    }
    if ((bi & 31) == lane) {
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i)
        if (i == (bi >> 5)) key[i] = -INFINITY;
    }
    float wj = bs;
    if (scoring == 2) {
      if (j == 0) first_val = bs;
      wj = expf(bs - first_val);
    }
    wsum += wj;
    if (lane == j) {
      my_w = wj;
      my_id = bi;
    }
  }
  if (lane < top_k) {
    if (renorm || scoring == 2) my_w = my_w / wsum;
    ids[(long long)m * top_k + lane] = my_id;
    wts[(long long)m * top_k + lane] = my_w;
  }
}

// ------------------------------------------------------------------------------------------------
// Binning: count -> scan (+ chunk descriptors) -> scatter
// ------------------------------------------------------------------------------------------------
__global__ void count_kernel(const int* __restrict__ ids, int n, int e_start, int e_end, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = ids[i];
  if (e >= e_start && e < e_end) atomicAdd(&counts[e - e_start], 1);
}

// single block of 1024 threads; E_local <= 1024
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ counts, int E, int* __restrict__ offsets,
                                                    int* __restrict__ cursor, ChunkDesc* __restrict__ chunks,
                                                    int* __restrict__ n_chunks) {
  __shared__ int s_cnt[1024], s_chk[1024];
  const int t = threadIdx.x;
  const int c = t < E ? counts[t] : 0;
  const int nch = (c + kMaxChunkTokens - 1) / kMaxChunkTokens;
  s_cnt[t] = c;
  s_chk[t] = nch;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan (tiny)
    int a = 0, b = 0;
    if (t >= o) {
      a = s_cnt[t - o];
      b = s_chk[t - o];
    }
    __syncthreads();
    s_cnt[t] += a;
    s_chk[t] += b;
    __syncthreads();
  }
  const int off = s_cnt[t] - c, choff = s_chk[t] - nch;
  if (t < E) {
    offsets[t] = off;
    cursor[t] = 0;
    if (nch > 0) {
      int per = (c + nch - 1) / nch;
      per = (per + 15) & ~15;
      for (int i = 0; i < nch; ++i) {
        ChunkDesc d;
        d.expert = t;
        d.slot_begin = off + i * per;
        d.n_tok = min(per, c - i * per);
        d.pad_ = 0;
        chunks[choff + i] = d;
      }
    }
  }
  if (t == 1023) {
    offsets[E] = s_cnt[1023];
    *n_chunks = s_chk[1023];
  }
}

// Token scatter: one warp per (token, k) routing entry.  Lane 0 claims a slot in the expert's contiguous run,
// then the warp copies the token's hidden row into x_sorted[slot] (16 B per lane per step, coalesced) so that
// every expert's tokens form one contiguous K-major tile the GEMM can fetch with 2-D TMA.
__global__ void __launch_bounds__(256) scatter_kernel(const int* __restrict__ ids, const float* __restrict__ wts, int n,
                                                      int top_k, int e_start, int e_end,
                                                      const int* __restrict__ offsets, int* __restrict__ cursor,
                                                      float* __restrict__ sorted_w, int* __restrict__ slot_of,
                                                      int* __restrict__ sorted_ids,
                                                      const uint4* __restrict__ x, uint4* __restrict__ x_sorted,
                                                      int vec_per_row) {
  const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int e = ids[i];
  int slot = -1;
  if (e >= e_start && e < e_end) {
    if (lane == 0) {
      const int le = e - e_start;
      slot = offsets[le] + atomicAdd(&cursor[le], 1);
      sorted_w[slot] = wts[i];
      if (sorted_ids) sorted_ids[slot] = e;
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    const uint4* src = x + (long long)(i / top_k) * vec_per_row;
    uint4* dst = x_sorted + (long long)slot * vec_per_row;
    for (int v = lane; v < vec_per_row; v += 32) dst[v] = src[v];
  }
  if (lane == 0) slot_of[i] = slot;
}

// Gather mode: no row copy — one THREAD per routing entry records which token row sits in which sorted slot; the grouped
// GEMM then fetches the rows from the caller's activation matrix with TMA gather4.
__global__ void __launch_bounds__(256) scatter_index_kernel(const int* __restrict__ ids, const float* __restrict__ wts, int n,
                                                            int top_k, int e_start, int e_end, const int* __restrict__ offsets,
                                                            int* __restrict__ cursor, float* __restrict__ sorted_w,
                                                            int* __restrict__ slot_of, int* __restrict__ sorted_tok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = ids[i];
  int slot = -1;
  if (e >= e_start && e < e_end) {
    const int le = e - e_start;
    slot = offsets[le] + atomicAdd(&cursor[le], 1);
    sorted_w[slot] = wts[i];
    sorted_tok[slot] = i / top_k;
  }
  slot_of[i] = slot;
}

// ------------------------------------------------------------------------------------------------
// Combine: out[m] = bf16( sum_j f32(c3[slot_of[m][j]]) ) in j order; then bf16(rsf*out) (+ shared)
// one thread = 8 consecutive h (16 B)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_kernel(const __nv_bfloat16* __restrict__ c3,
                                                      const int* __restrict__ slot_of, int M, int H, int top_k,
                                                      float rsf, int apply_rsf,
                                                      const __nv_bfloat16* __restrict__ shared,
                                                      __nv_bfloat16* __restrict__ out, CombineScatter sc) {
  const int vec_per_row = H >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * vec_per_row) return;
  const int m = (int)(idx / vec_per_row), v = (int)(idx % vec_per_row);
  float acc[8] = {};
  for (int j = 0; j < top_k; ++j) {
    const int slot = slot_of ? slot_of[(long long)m * top_k + j] : m * top_k + j;     // no table: rows already in token order
    if (slot < 0) continue;
    const uint4 x = *reinterpret_cast<const uint4*>(c3 + (long long)slot * H + v * 8);
    const __nv_bfloat16* px = reinterpret_cast<const __nv_bfloat16*>(&x);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += __bfloat162float(px[i]);
  }
  uint4 sh = make_uint4(0, 0, 0, 0);
  if (shared) sh = *reinterpret_cast<const uint4*>(shared + (long long)m * H + v * 8);
  const __nv_bfloat16* ps = reinterpret_cast<const __nv_bfloat16*>(&sh);
  uint4 o;
  __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float r = bf16_round_rn(acc[i]);                       // every stage rounds to BF16 like the reference's tensor ops
    if (apply_rsf) r = bf16_round_rn(rsf * r);
    if (shared) r = bf16_round_rn(r + __bfloat162float(ps[i]));
    po[i] = __ushort_as_bfloat16((unsigned short)(__float_as_uint(r) >> 16));
  }
  if (sc.n_ranks > 0) {
    // fused reduce-scatter, producer side: the row goes straight into its OWNER rank's receive buffer over NVLink (peer-mapped
    // memory), slot `src_rank` of token (m - owner * rows_per_rank); the owner sums its n_ranks slots with this same kernel
    // (slot_of == nullptr, top_k == n_ranks) once every rank has passed the barrier that follows
    const int owner = m / sc.rows_per_rank, ml = m - owner * sc.rows_per_rank;
    __nv_bfloat16* dst = sc.peer_out[owner] + ((long long)ml * sc.n_ranks + sc.src_rank) * H + v * 8;
    *reinterpret_cast<uint4*>(dst) = o;
    return;
  }
  *reinterpret_cast<uint4*>(out + (long long)m * H + v * 8) = o;
}

// ------------------------------------------------------------------------------------------------
// Load-time re-tiling of the reference quantiser's output into KB2 tiles (see moe_common.cuh)
// ------------------------------------------------------------------------------------------------
// INT4: src packed[E][N][K/8] u32 -> dst tiles; one thread per destination word
__global__ void repack_int4_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int E, int N, int K) {
  const long long total = (long long)E * N * (K / 8);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int words_per_expert_row = K / 8;
  const long long per_expert = (long long)N * words_per_expert_row;
  const int e = (int)(i / per_expert);
  long long r = i % per_expert;          // destination word index within the expert
  const int nkb = K / kBlockK;
  const int tile = (int)(r / (nkb * 1024));          // 1024 words per (tile, kb)
  r %= (long long)nkb * 1024;
  const int kb = (int)(r / 1024);
  const int w = (int)(r % 1024);
  const int h = w / 512, row = (w % 512) / 4, j = w % 4;
  const int n = tile * kTileRows + row;
  const int kw = kb * 8 + h * 4 + j;
  const uint32_t s = src[((long long)e * N + n) * words_per_expert_row + kw];
  // destination nibble p holds source nibble {0,2,4,6,1,3,5,7}[p]
  uint32_t d = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int sp = (p < 4) ? 2 * p : 2 * (p - 4) + 1;
    d |= ((s >> (4 * sp)) & 0xFu) << (4 * p);
  }
  dst[i] = d;
}

// INT8: src data[E][N][K] i8 -> dst tiles [tile][kb][quarter][row][16 B]; one thread per 16 B
__global__ void repack_int8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int E, int N, int K) {
  const long long total = (long long)E * N * (K / 16);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long per_expert = (long long)N * (K / 16);
  const int e = (int)(i / per_expert);
  long long r = i % per_expert;
  const int nkb = K / kBlockK;
  const int tile = (int)(r / (nkb * 512));           // 512 x 16 B per (tile, kb)
  r %= (long long)nkb * 512;
  const int kb = (int)(r / 512);
  const int w = (int)(r % 512);
  const int q = w / 128, row = w % 128;
  const int n = tile * kTileRows + row;
  dst[i] = src[((long long)e * N + n) * (K / 16) + kb * 4 + q];
}

// scales: src[E][N][K/128] bf16 -> dst[E][N/128][K/128][128]
__global__ void repack_scales_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int E, int N,
                                     int G) {
  const long long total = (long long)E * N * G;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long per_expert = (long long)N * G;
  const int e = (int)(i / per_expert);
  long long r = i % per_expert;
  const int tile = (int)(r / (G * kTileRows));
  r %= (long long)G * kTileRows;
  const int g = (int)(r / kTileRows), row = (int)(r % kTileRows);
  dst[i] = src[((long long)e * N + tile * kTileRows + row) * G + g];
}

// ------------------------------------------------------------------------------------------------
// Krasis symmetric group quantiser on the device — bit-exact restatement of src/weights/marlin.rs:145-207 (INT4) and
// :65-114 (INT8): per row and 128-column group  scale = bf16_rne(amax / qmax)  (1.0 if amax == 0);
// q = clamp(round_half_away(w * (1 / f32(scale))), qmin, qmax).  All steps are single IEEE f32 operations.
// One warp per (row, group); output in the reference quantiser's layout (packed [N][K/8] u32 | i8 [N][K], scales [N][K/128]).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quantize_group_kernel(const __nv_bfloat16* __restrict__ w, int bits,
                                                             void* __restrict__ q_out, uint16_t* __restrict__ scales,
                                                             long long rows, int K) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, G = K / kGroup;
  if (wid >= rows * G) return;
  const long long row = wid / G;
  const int g = (int)(wid % G);
  const __nv_bfloat16* src = w + row * K + g * kGroup + lane * 4;
  float v[4];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = __bfloat162float(src[i]);
    amax = fmaxf(amax, fabsf(v[i]));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float qmax = bits == 4 ? 7.0f : 127.0f;
  const float sc_f = amax == 0.f ? 1.0f : __fdiv_rn(amax, qmax);
  const __nv_bfloat16 sc_b = __float2bfloat16_rn(sc_f);
  const float sc = __bfloat162float(sc_b);
  const float inv = sc == 0.f ? 0.f : __fdiv_rn(1.0f, sc);
  int q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float r = roundf(__fmul_rn(v[i], inv));                         // f32::round: half away from zero
    q[i] = (int)fminf(fmaxf(r, bits == 4 ? -8.f : -128.f), qmax);
  }
  if (lane == 0) scales[row * G + g] = __bfloat16_as_ushort(sc_b);
  if (bits == 4) {
    uint32_t nib = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) nib |= (uint32_t)((q[i] + 8) & 0xF) << (4 * i);   // 4 nibbles of this lane
    const uint32_t hi = __shfl_down_sync(0xffffffffu, nib, 1);                      // odd lane's nibbles -> upper half
    if ((lane & 1) == 0)
      reinterpret_cast<uint32_t*>(q_out)[row * (K / 8) + g * (kGroup / 8) + (lane >> 1)] = nib | (hi << 16);
  } else {
    char4 c = make_char4((signed char)q[0], (signed char)q[1], (signed char)q[2], (signed char)q[3]);
    reinterpret_cast<char4*>(q_out)[(row * K + g * kGroup) / 4 + lane] = c;
  }
}

cudaError_t launch_quantize_group(const void* w, int bits, void* q_out, void* scales, long long rows, int K, cudaStream_t s) {
  KernelSpan ks(K_QUANTIZE_GROUP, s);
  if ((bits != 4 && bits != 8) || K % kGroup || rows <= 0) return cudaErrorInvalidValue;
  const long long warps = rows * (K / kGroup);
  quantize_group_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)w, bits, q_out,
                                                                            (uint16_t*)scales, rows, K);
  return cudaGetLastError();
}

// GGUF re-tiling (lossless): rows come from `a` (first n_a rows, e.g. gate) then `b` (e.g. up); src rows are native
// GGUF block rows [N][K/bs * bb] (src/gguf_kernels.rs:9).  One thread per destination (expert, tile, k-block, row).
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t* sc, uint8_t& s, uint8_t& m) {   // src/gguf.rs:666-674
  if (j < 4) {
    s = sc[j] & 63;
    m = sc[j + 4] & 63;
  } else {
    s = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4);
    m = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4);
  }
}

__global__ void retile_gguf_kernel(int fmt, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n_a,
                                   uint8_t* __restrict__ dst, int E, int N, int K) {
  const int nkb = K / kBlockK, ntile = N / kTileRows;
  const long long total = (long long)E * ntile * nkb * kTileRows;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int row = (int)(i % kTileRows);
  const int kb = (int)((i / kTileRows) % nkb);
  const int tile = (int)((i / kTileRows / nkb) % ntile);
  const int e = (int)(i / kTileRows / nkb / ntile);
  const int n = tile * kTileRows + row;
  const bool from_a = n < n_a;
  const int n_src = from_a ? n : n - n_a, rows_src = from_a ? n_a : N - n_a;
  if (fmt == kFmtQ8_0) {
    const long long row_bytes = (long long)K / 32 * 34;
    const uint8_t* src = (from_a ? a : b) + ((long long)e * rows_src + n_src) * row_bytes + (long long)kb * 2 * 34;
    uint8_t* blob = dst + (((long long)e * ntile + tile) * nkb + kb) * kQ8_0TileBytes;
    for (int blk = 0; blk < 2; ++blk) {
      const uint8_t* sb = src + blk * 34;
      blob[kTileRows * kBlockK + row * 4 + blk * 2] = sb[0];
      blob[kTileRows * kBlockK + row * 4 + blk * 2 + 1] = sb[1];
      for (int l = 0; l < 32; ++l) {
        const int k = blk * 32 + l;                      // element within the 64-wide k-block
        blob[(k / 16) * 2048 + row * 16 + (k % 16)] = sb[2 + l];
      }
    }
  } else if (fmt >= kFmtAffine8) {
    // Decode one row's 64 elements of this k-block into int8 codes + (a, b) per 16 elements.  d*sc, dmin*mn are exact in
    // f32 (fp16 x small integer), so w = fma(a, code, -b) in the GEMM reproduces the reference's single rounding.
    uint8_t* blob = dst + (((long long)e * ntile + tile) * nkb + kb) * kAffine8TileBytes;
    float* prm = reinterpret_cast<float*>(blob + kTileRows * kBlockK + row * 32);
    const uint8_t* rows = from_a ? a : b;
    auto put = [&](int k, int code) { blob[(k / 16) * 2048 + row * 16 + (k % 16)] = (uint8_t)(int8_t)code; };
    auto f16 = [](const uint8_t* p) { return __half2float(__ushort_as_half((unsigned short)(p[0] | (p[1] << 8)))); };
    if (fmt == kFmtQ6_K) {
      const uint8_t* sb = rows + ((long long)e * rows_src + n_src) * ((long long)K / 256 * 210) + (long long)(kb / 4) * 210;
      const float d = f16(sb + 208);
      const int8_t* sc = reinterpret_cast<const int8_t*>(sb + 192);
      for (int k = 0; k < 64; ++k) {
        const int el = (kb % 4) * 64 + k, half = el / 128, r = el % 128, sub = r / 32, l = r % 32;
        const uint8_t qlb = sb[half * 64 + (sub & 1) * 32 + l];
        const int q4 = sub < 2 ? (qlb & 0xF) : (qlb >> 4);
        const int q = q4 | (((sb[128 + half * 32 + l] >> (2 * sub)) & 3) << 4);
        put(k, q - 32);
        if (k % 16 == 0) { prm[(k / 16) * 2] = d * (float)sc[half * 8 + l / 16 + 2 * sub]; prm[(k / 16) * 2 + 1] = 0.f; }
      }
    } else if (fmt == kFmtQ5_K) {
      const uint8_t* sb = rows + ((long long)e * rows_src + n_src) * ((long long)K / 256 * 176) + (long long)(kb / 4) * 176;
      const float d = f16(sb), dmin = f16(sb + 2);
      const int c = kb % 4;
      for (int h = 0; h < 2; ++h) {
        uint8_t sj, mj;
        k4_scale_min(2 * c + h, sb + 4, sj, mj);
        for (int l = 0; l < 32; ++l) {
          const uint8_t qb = sb[48 + c * 32 + l];
          const int q4 = h ? (qb >> 4) : (qb & 0xF);
          put(h * 32 + l, q4 + 16 * ((sb[16 + l] >> (2 * c + h)) & 1));
        }
        for (int g = 0; g < 2; ++g) { prm[(h * 2 + g) * 2] = d * (float)sj; prm[(h * 2 + g) * 2 + 1] = dmin * (float)mj; }
      }
    } else {   // Q5_0 / Q4_0: two 32-element blocks per k-block
      const int bb = fmt == kFmtQ5_0 ? 22 : 18;
      const uint8_t* src = rows + ((long long)e * rows_src + n_src) * ((long long)K / 32 * bb) + (long long)kb * 2 * bb;
      for (int blk = 0; blk < 2; ++blk) {
        const uint8_t* sb = src + blk * bb;
        const float d = f16(sb);
        const uint8_t* qs = sb + (fmt == kFmtQ5_0 ? 6 : 2);
        const uint32_t qh = fmt == kFmtQ5_0 ? (uint32_t)sb[2] | ((uint32_t)sb[3] << 8) | ((uint32_t)sb[4] << 16) | ((uint32_t)sb[5] << 24) : 0u;
        for (int l = 0; l < 32; ++l) {
          const int q4 = l < 16 ? (qs[l] & 0xF) : (qs[l - 16] >> 4);
          put(blk * 32 + l, fmt == kFmtQ5_0 ? ((q4 | (int)(((qh >> l) & 1u) << 4)) - 16) : (q4 - 8));
        }
        for (int g = 0; g < 2; ++g) { prm[(blk * 2 + g) * 2] = d; prm[(blk * 2 + g) * 2 + 1] = 0.f; }
      }
    }
  } else {   // Q4_K
    const long long row_bytes = (long long)K / 256 * 144;
    const uint8_t* sb = (from_a ? a : b) + ((long long)e * rows_src + n_src) * row_bytes + (long long)(kb / 4) * 144;
    const int j = kb % 4;
    uint8_t* blob = dst + (((long long)e * ntile + tile) * nkb + kb) * kQ4KTileBytes;
    const uint8_t* qs = sb + 16 + j * 32;
    for (int l = 0; l < 32; ++l) blob[(l / 16) * 2048 + row * 16 + (l % 16)] = qs[l];
    uint8_t* hdr = blob + kTileRows * kBlockK / 2 + row * 8;
    hdr[0] = sb[0]; hdr[1] = sb[1]; hdr[2] = sb[2]; hdr[3] = sb[3];
    uint8_t s0, m0, s1, m1;
    k4_scale_min(2 * j, sb + 4, s0, m0);
    k4_scale_min(2 * j + 1, sb + 4, s1, m1);
    hdr[4] = s0; hdr[5] = m0; hdr[6] = s1; hdr[7] = m1;
  }
}

cudaError_t launch_retile_gguf(int fmt, const void* a, const void* b, int n_a, void* dst, int E, int N, int K,
                               cudaStream_t s) {
  KernelSpan ks(K_RETILE, s);
  const bool k256 = fmt == kFmtQ4_K || fmt == kFmtQ6_K || fmt == kFmtQ5_K;
  if (N % kTileRows || K % kBlockK || (k256 && K % 256) || fmt < kFmtQ8_0 || fmt > kFmtQ4_0) return cudaErrorInvalidValue;
  const long long total = (long long)E * (N / kTileRows) * (K / kBlockK) * kTileRows;
  retile_gguf_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(fmt, (const uint8_t*)a, (const uint8_t*)b, n_a,
                                                                    (uint8_t*)dst, E, N, K);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
cudaError_t launch_router_logits(const void* h, const void* gate, const float* bias, float* logits, int M, int E,
                                 int H, cudaStream_t s) {
  KernelSpan ks(K_ROUTER_GEMM, s);
  dim3 grid((E + RT - 1) / RT, (M + RT - 1) / RT);
  router_logits_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)h, (const __nv_bfloat16*)gate, bias, logits, M, E, H);
  return cudaGetLastError();
}

cudaError_t launch_router_topk(const float* logits, const float* corr_bias, int M, int E, int top_k, int scoring,
                               int renorm, int* ids, float* wts, cudaStream_t s) {
  KernelSpan ks(K_ROUTER_TOPK, s);
  const int warps = 4;
  dim3 grid((M + warps - 1) / warps);
  if (top_k > 32 || E > 1024) return cudaErrorInvalidValue;
  if (E <= 64)
    router_topk_kernel<2><<<grid, warps * 32, 0, s>>>(logits, corr_bias, M, E, top_k, scoring, renorm, ids, wts);
  else if (E <= 128)
    router_topk_kernel<4><<<grid, warps * 32, 0, s>>>(logits, corr_bias, M, E, top_k, scoring, renorm, ids, wts);
  else if (E <= 256)
    router_topk_kernel<8><<<grid, warps * 32, 0, s>>>(logits, corr_bias, M, E, top_k, scoring, renorm, ids, wts);
  else if (E <= 512)
    router_topk_kernel<16><<<grid, warps * 32, 0, s>>>(logits, corr_bias, M, E, top_k, scoring, renorm, ids, wts);
  else
    router_topk_kernel<32><<<grid, warps * 32, 0, s>>>(logits, corr_bias, M, E, top_k, scoring, renorm, ids, wts);
  return cudaGetLastError();
}

cudaError_t launch_binning(const int* ids, const float* wts, int M, int top_k, int e_start, int e_end, int* counts,
                           int* offsets, int* cursor, ChunkDesc* chunks, int* n_chunks, float* sorted_w, int* slot_of,
                           int* sorted_ids, const void* x, void* x_sorted, int H, cudaStream_t s) {
  KernelSpan ks(K_BINNING, s, 3);
  const int n = M * top_k, E = e_end - e_start;
  if (E > 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * E, s);
  if (e != cudaSuccess) return e;
  if (n > 0) count_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, n, e_start, e_end, counts);
  scan_kernel<<<1, 1024, 0, s>>>(counts, E, offsets, cursor, chunks, n_chunks);
  if (n > 0)
    scatter_kernel<<<(n + 7) / 8, 256, 0, s>>>(ids, wts, n, top_k, e_start, e_end, offsets, cursor, sorted_w, slot_of,
                                               sorted_ids, (const uint4*)x, (uint4*)x_sorted, H / 8);
  return cudaGetLastError();
}

cudaError_t launch_binning_index(const int* ids, const float* wts, int M, int top_k, int e_start, int e_end, int* counts,
                                 int* offsets, int* cursor, ChunkDesc* chunks, int* n_chunks, float* sorted_w, int* slot_of,
                                 int* sorted_tok, cudaStream_t s) {
  KernelSpan ks(K_BINNING, s, 3);
  const int n = M * top_k, E = e_end - e_start;
  if (E > 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * E, s);
  if (e != cudaSuccess) return e;
  if (n > 0) count_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, n, e_start, e_end, counts);
  scan_kernel<<<1, 1024, 0, s>>>(counts, E, offsets, cursor, chunks, n_chunks);
  if (n > 0)
    scatter_index_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, wts, n, top_k, e_start, e_end, offsets, cursor, sorted_w, slot_of,
                                                         sorted_tok);
  return cudaGetLastError();
}

cudaError_t launch_combine_scatter(const void* c3, const int* slot_of, int M, int H, int top_k, float rsf, int apply_rsf,
                                   const void* shared, void* out, const CombineScatter& sc, cudaStream_t s) {
  KernelSpan ks(K_COMBINE, s);
  const long long total = (long long)M * (H / 8);
  if (total == 0) return cudaSuccess;
  combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)c3, slot_of, M, H, top_k, rsf,
                                                                apply_rsf, (const __nv_bfloat16*)shared,
                                                                (__nv_bfloat16*)out, sc);
  return cudaGetLastError();
}
cudaError_t launch_combine(const void* c3, const int* slot_of, int M, int H, int top_k, float rsf, int apply_rsf,
                           const void* shared, void* out, cudaStream_t s) {
  return launch_combine_scatter(c3, slot_of, M, H, top_k, rsf, apply_rsf, shared, out, CombineScatter{}, s);
}

cudaError_t launch_repack(int fmt, const void* src_q, const void* src_s, void* dst_q, void* dst_s, int E, int N, int K,
                          cudaStream_t s) {
  KernelSpan ks(K_RETILE, s, 2);
  if (N % kTileRows || K % kGroup) return cudaErrorInvalidValue;
  if (fmt == kFmtInt4G128) {
    const long long total = (long long)E * N * (K / 8);
    repack_int4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const uint32_t*)src_q, (uint32_t*)dst_q, E, N, K);
  } else if (fmt == kFmtInt8G128) {
    const long long total = (long long)E * N * (K / 16);
    repack_int8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const uint4*)src_q, (uint4*)dst_q, E, N, K);
  } else {
    return cudaErrorInvalidValue;
  }
  const long long ts = (long long)E * N * (K / kGroup);
  repack_scales_kernel<<<(unsigned)((ts + 255) / 256), 256, 0, s>>>((const uint16_t*)src_s, (uint16_t*)dst_s, E, N,
                                                                   K / kGroup);
  return cudaGetLastError();
}

}  // namespace kb2
