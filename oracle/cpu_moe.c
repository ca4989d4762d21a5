/*
 * C/AVX2 restatement of the reference's CPU expert path — TEST INFRASTRUCTURE + timed CPU baseline only
 * (bench.py `cpu_baseline` / `--impl reference`).  The reference is Rust and cannot be compiled here
 * (no cargo/rustc), so this "port" follows, function by function:
 *
 *   quantize_activation_int16            src/kernel/avx2.rs:234-268   (scale = amax/32767, round half away)
 *   expert_matmul_int4_transposed_integer src/kernel/avx2.rs:1066-1206 (INT16 x INT4 -> INT32 per group via
 *                                         madd_epi16 on nibble pairs; per group fma(float(isum), ws*as, out))
 *   fast_exp / fast_sigmoid              src/kernel/avx2.rs:2229-2291 (2^(x log2 e), degree-5 poly, rcp + 1 NR)
 *   silu_quantize_int16_avx2             src/kernel/avx2.rs:2310-2366 (SiLU(g)*u -> INT16, cvtps (RNE) in pass 2)
 *   expert_forward_unified               src/moe.rs:184-380
 *   moe_forward_unified / _flattened     src/moe.rs:572-715, 727-  (one token at a time, experts x N-chunks
 *                                         spread over the thread pool, out = sum_i w_i * expert_i)
 *   moe_worker batch loop                src/moe.rs:1287-1343 (tokens processed sequentially)
 *
 * Weight layout = the reference's CPU "unified transposed" form (src/weights/mod.rs:287-397):
 *   w13 packed [K/8][2I] u32 (word (kw, n): nibble j = column 8kw+j of output row n, value q+8),
 *   w13 scales [K/gs][2I] bf16, w2 packed [I/8][H], w2 scales [I/gs][H].
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float bf16_to_f32(uint16_t v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* src/kernel/avx2.rs:234-304 */
static void quant_act_i16(const float* x, int k, int gs, int16_t* q, float* scales) {
  for (int g = 0; g < k / gs; ++g) {
    float mx = 0.f;
    for (int i = 0; i < gs; ++i) mx = fmaxf(mx, fabsf(x[g * gs + i]));
    const float scale = mx > 0.f ? mx / 32767.0f : 1.0f;
    const float inv = mx > 0.f ? 32767.0f / mx : 0.0f;
    scales[g] = scale;
    for (int i = 0; i < gs; ++i) {
      int v = (int)roundf(x[g * gs + i] * inv); /* Rust f32::round: half away from zero */
      if (v > 32767) v = 32767;
      if (v < -32768) v = -32768;
      q[g * gs + i] = (int16_t)v;
    }
  }
}

/* src/kernel/avx2.rs:1066-1206 for output columns [n0, n0+n_out), n_out % 8 == 0, n_out <= 256 */
static void matvec_int4_t(const uint32_t* packed, const uint16_t* wscales, const int16_t* a, const float* ascales,
                          float* out, int k, int n_stride, int n0, int n_out, int gs) {
  const int nb = n_out / 8, ppg = gs / 8;
  const __m256i m0f = _mm256_set1_epi32(0xF), off8 = _mm256_set1_epi32(8), mffff = _mm256_set1_epi32(0xFFFF);
  __m256i acc[32];
  for (int b = 0; b < nb; ++b) _mm256_storeu_ps(out + b * 8, _mm256_setzero_ps());
  for (int g = 0; g < k / gs; ++g) {
    for (int b = 0; b < nb; ++b) acc[b] = _mm256_setzero_si256();
    for (int p = 0; p < ppg; ++p) {
      const int kr = g * ppg + p;
      __m256i ap[4];
      for (int j = 0; j < 4; ++j) {
        uint32_t pair = (uint16_t)a[kr * 8 + 2 * j] | ((uint32_t)(uint16_t)a[kr * 8 + 2 * j + 1] << 16);
        ap[j] = _mm256_set1_epi32((int)pair);
      }
      const uint32_t* row = packed + (size_t)kr * n_stride + n0;
      for (int b = 0; b < nb; ++b) {
        const __m256i w = _mm256_loadu_si256((const __m256i*)(row + b * 8));
        __m256i s = acc[b];
        for (int j = 0; j < 4; ++j) {
          const __m256i lo = _mm256_sub_epi32(_mm256_and_si256(_mm256_srli_epi32(w, 8 * j), m0f), off8);
          const __m256i hi = _mm256_sub_epi32(_mm256_and_si256(_mm256_srli_epi32(w, 8 * j + 4), m0f), off8);
          const __m256i wp = _mm256_or_si256(_mm256_and_si256(lo, mffff), _mm256_slli_epi32(hi, 16));
          s = _mm256_add_epi32(s, _mm256_madd_epi16(wp, ap[j]));
        }
        acc[b] = s;
      }
    }
    const __m256 as = _mm256_set1_ps(ascales[g]);
    for (int b = 0; b < nb; ++b) {
      const __m128i sb = _mm_loadu_si128((const __m128i*)(wscales + (size_t)g * n_stride + n0 + b * 8));
      const __m256 ws = _mm256_castsi256_ps(_mm256_slli_epi32(_mm256_cvtepu16_epi32(sb), 16));
      const __m256 comb = _mm256_mul_ps(ws, as);
      _mm256_storeu_ps(out + b * 8, _mm256_fmadd_ps(_mm256_cvtepi32_ps(acc[b]), comb, _mm256_loadu_ps(out + b * 8)));
    }
  }
}

/* src/kernel/avx2.rs:2229-2291 */
static inline __m256 fast_exp(__m256 x) {
  const __m256 t = _mm256_mul_ps(x, _mm256_set1_ps(1.4426950408889634f));
  const __m256 n = _mm256_floor_ps(t);
  const __m256 f = _mm256_sub_ps(t, n);
  __m256 p = _mm256_fmadd_ps(_mm256_set1_ps(0.0013333558f), f, _mm256_set1_ps(0.009618129f));
  p = _mm256_fmadd_ps(p, f, _mm256_set1_ps(0.0555041f));
  p = _mm256_fmadd_ps(p, f, _mm256_set1_ps(0.2402265f));
  p = _mm256_fmadd_ps(p, f, _mm256_set1_ps(0.6931472f));
  p = _mm256_fmadd_ps(p, f, _mm256_set1_ps(1.0f));
  const __m256i e = _mm256_slli_epi32(_mm256_add_epi32(_mm256_cvtps_epi32(n), _mm256_set1_epi32(127)), 23);
  return _mm256_mul_ps(p, _mm256_castsi256_ps(e));
}
static inline __m256 fast_sigmoid(__m256 x) {
  __m256 nx = _mm256_sub_ps(_mm256_setzero_ps(), x);
  nx = _mm256_max_ps(_mm256_min_ps(nx, _mm256_set1_ps(20.f)), _mm256_set1_ps(-20.f));
  const __m256 d = _mm256_add_ps(_mm256_set1_ps(1.f), fast_exp(nx));
  const __m256 r = _mm256_rcp_ps(d);
  return _mm256_mul_ps(r, _mm256_fnmadd_ps(d, r, _mm256_set1_ps(2.f)));
}

/* src/kernel/avx2.rs:2310-2366 */
static void silu_quant_i16(float* gate, const float* up, int16_t* q, float* scales, int n, int gs) {
  const __m256 sign = _mm256_set1_ps(-0.0f);
  for (int g = 0; g < n / gs; ++g) {
    __m256 mx = _mm256_setzero_ps();
    for (int i = 0; i < gs; i += 8) {
      const __m256 gv = _mm256_loadu_ps(gate + g * gs + i);
      const __m256 h = _mm256_mul_ps(_mm256_mul_ps(gv, fast_sigmoid(gv)), _mm256_loadu_ps(up + g * gs + i));
      _mm256_storeu_ps(gate + g * gs + i, h);
      mx = _mm256_max_ps(mx, _mm256_andnot_ps(sign, h));
    }
    float t[8];
    _mm256_storeu_ps(t, mx);
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, t[i]);
    const float scale = m > 0.f ? m / 32767.0f : 1.0f, inv = m > 0.f ? 32767.0f / m : 0.0f;
    scales[g] = scale;
    const __m256 iv = _mm256_set1_ps(inv);
    for (int i = 0; i < gs; i += 8) {
      const __m256i v = _mm256_cvtps_epi32(_mm256_mul_ps(_mm256_loadu_ps(gate + g * gs + i), iv));
      const __m128i p = _mm_packs_epi32(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
      _mm_storeu_si128((__m128i*)(q + g * gs + i), p);
    }
  }
}

#define NCHUNK 256

/*
 * Batch MoE forward, INT4 g128 unified weights.  Tokens sequentially (moe_worker), per token three flat
 * parallel phases over (expert, N-chunk) work items (moe_forward_flattened), then out = sum_i w_i * e_i.
 * ids < 0 are skipped (src/moe.rs:2722).  out: [M][H] f32.  Returns 0.
 */
int kcpu_moe_forward_int4(const uint32_t* w13, const uint16_t* s13, const uint32_t* w2, const uint16_t* s2, int E,
                          int H, int I, int gs, const uint16_t* x_bf16, const int32_t* ids, const float* wts, int M,
                          int topk, float* out, int nthreads) {
  (void)E;
  const size_t w13_e = (size_t)(H / 8) * 2 * I, s13_e = (size_t)(H / gs) * 2 * I;
  const size_t w2_e = (size_t)(I / 8) * H, s2_e = (size_t)(I / gs) * H;
  float* xf = (float*)malloc(sizeof(float) * H);
  int16_t* xa = (int16_t*)malloc(sizeof(int16_t) * H);
  float* xs = (float*)malloc(sizeof(float) * (H / gs));
  float* w13o = (float*)malloc(sizeof(float) * (size_t)topk * 2 * I);
  int16_t* ha = (int16_t*)malloc(sizeof(int16_t) * (size_t)topk * I);
  float* hs = (float*)malloc(sizeof(float) * (size_t)topk * (I / gs));
  float* eo = (float*)malloc(sizeof(float) * (size_t)topk * H);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const int c13 = (2 * I + NCHUNK - 1) / NCHUNK, c2 = (H + NCHUNK - 1) / NCHUNK;
#pragma omp parallel
  for (int m = 0; m < M; ++m) {
#pragma omp single
    {
      for (int i = 0; i < H; ++i) xf[i] = bf16_to_f32(x_bf16[(size_t)m * H + i]);
      quant_act_i16(xf, H, gs, xa, xs);
    }
    const int32_t* id = ids + (size_t)m * topk;
#pragma omp for schedule(static)
    for (int it = 0; it < topk * c13; ++it) { /* phase 1: w13 chunks */
      const int j = it / c13, c = it % c13, e = id[j];
      if (e < 0) continue;
      const int n0 = c * NCHUNK, nn = (2 * I - n0 < NCHUNK) ? 2 * I - n0 : NCHUNK;
      matvec_int4_t(w13 + e * w13_e, s13 + e * s13_e, xa, xs, w13o + (size_t)j * 2 * I + n0, H, 2 * I, n0, nn, gs);
    }
#pragma omp for schedule(static)
    for (int j = 0; j < topk; ++j) { /* phase 2: SiLU*up -> INT16 */
      if (id[j] < 0) continue;
      float* g = w13o + (size_t)j * 2 * I;
      silu_quant_i16(g, g + I, ha + (size_t)j * I, hs + (size_t)j * (I / gs), I, gs);
    }
#pragma omp for schedule(static)
    for (int it = 0; it < topk * c2; ++it) { /* phase 3: w2 chunks */
      const int j = it / c2, c = it % c2, e = id[j];
      if (e < 0) continue;
      const int n0 = c * NCHUNK, nn = (H - n0 < NCHUNK) ? H - n0 : NCHUNK;
      matvec_int4_t(w2 + e * w2_e, s2 + e * s2_e, ha + (size_t)j * I, hs + (size_t)j * (I / gs),
                    eo + (size_t)j * H + n0, I, H, n0, nn, gs);
    }
#pragma omp for schedule(static)
    for (int h = 0; h < H; ++h) { /* weighted sum in expert order (src/moe.rs:663-669) */
      float s = 0.f;
      for (int j = 0; j < topk; ++j)
        if (id[j] >= 0) s += wts[(size_t)m * topk + j] * eo[(size_t)j * H + h];
      out[(size_t)m * H + h] = s;
    }
  }
  free(xf); free(xa); free(xs); free(w13o); free(ha); free(hs); free(eo);
  return 0;
}

/* ====================================================================================================================
 * Native-GGUF CPU expert path: moe_forward_gguf (src/moe.rs:990-1110, non-NUMA parallel branch: one expert per task, then the
 * weighted sum in expert order) over expert_forward_gguf (src/gguf_kernels.rs:690-756) with the INT16 activation path:
 *   quantize_{bf16,f32}_to_int16   src/gguf_kernels.rs:108-172  (groups of 32, scale = amax/32767, round half away, group sums)
 *   matvec_q4_k_avx2               src/gguf_kernels.rs:271-370  (madd_epi16 on zero-extended nibbles; fp32 lane accumulation;
 *                                   dmin*mn*ascale*sum correction subtracted once per row)
 *   matvec_q8_0_avx2               src/gguf_kernels.rs:376-425
 *   SiLU in plain f32 (g / (1 + exp(-g)))  :727-731
 * Types: 8 = Q8_0 (34 B / 32), 12 = Q4_K (144 B / 256)  (src/gguf.rs:15-31,56-85).
 * ==================================================================================================================== */
static inline float f16_to_f32(const uint8_t* p) {
  uint16_t h;
  memcpy(&h, p, 2);
  return _cvtsh_ss(h);
}

static void quant_act_i16_g32(const float* x, int k, int16_t* q, float* scales, int32_t* sums) {
  for (int g = 0; g < k / 32; ++g) {
    float mx = 0.f;
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fabsf(x[g * 32 + i]));
    const float scale = mx > 0.f ? mx / 32767.0f : 1.0f, inv = mx > 0.f ? 32767.0f / mx : 0.0f;
    scales[g] = scale;
    int32_t sum = 0;
    for (int i = 0; i < 32; ++i) {
      int v = (int)roundf(x[g * 32 + i] * inv);
      if (v > 32767) v = 32767;
      if (v < -32768) v = -32768;
      q[g * 32 + i] = (int16_t)v;
      sum += v;
    }
    sums[g] = sum;
  }
}

static inline float hsum_avx(__m256 v) {        /* same reduction order as gguf_kernels.rs:hsum_avx */
  const __m128 s = _mm_add_ps(_mm256_castps256_ps128(v), _mm256_extractf128_ps(v, 1));
  const __m128 s64 = _mm_add_ps(s, _mm_movehdup_ps(s));
  return _mm_cvtss_f32(_mm_add_ss(s64, _mm_movehl_ps(s64, s64)));
}

static inline void scale_min_k4(int j, const uint8_t* sc, uint8_t* s, uint8_t* m) {   /* src/gguf_kernels.rs:640-648 */
  if (j < 4) { *s = sc[j] & 63; *m = sc[j + 4] & 63; }
  else { *s = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); *m = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
}

static void matvec_q4_k(const uint8_t* w, const int16_t* a, const float* as, const int32_t* asum, int n, int k, float* out) {
  const int bpr = k / 256, row_bytes = bpr * 144;
  const __m128i m0f = _mm_set1_epi8(0x0F);
  for (int row = 0; row < n; ++row) {
    const uint8_t* rd = w + (size_t)row * row_bytes;
    __m256 facc = _mm256_setzero_ps();
    float corr = 0.f;
    for (int b = 0; b < bpr; ++b) {
      const uint8_t* blk = rd + b * 144;
      const float d = f16_to_f32(blk), dmin = f16_to_f32(blk + 2);
      for (int j = 0; j < 4; ++j) {
        uint8_t scl, mnl, sch, mnh;
        scale_min_k4(2 * j, blk + 4, &scl, &mnl);
        scale_min_k4(2 * j + 1, blk + 4, &sch, &mnh);
        const uint8_t* qs = blk + 16 + j * 32;
        const int gl = b * 8 + 2 * j, gh = gl + 1, ab = b * 256 + j * 64;
        const __m128i r0 = _mm_loadu_si128((const __m128i*)qs), r1 = _mm_loadu_si128((const __m128i*)(qs + 16));
        const __m128i l0 = _mm_and_si128(r0, m0f), l1 = _mm_and_si128(r1, m0f);
        const __m128i h0 = _mm_and_si128(_mm_srli_epi16(r0, 4), m0f), h1 = _mm_and_si128(_mm_srli_epi16(r1, 4), m0f);
        __m256i il = _mm256_madd_epi16(_mm256_cvtepu8_epi16(l0), _mm256_loadu_si256((const __m256i*)(a + ab)));
        il = _mm256_add_epi32(il, _mm256_madd_epi16(_mm256_cvtepu8_epi16(l1), _mm256_loadu_si256((const __m256i*)(a + ab + 16))));
        facc = _mm256_fmadd_ps(_mm256_cvtepi32_ps(il), _mm256_set1_ps(d * (float)scl * as[gl]), facc);
        corr += dmin * (float)mnl * as[gl] * (float)asum[gl];
        __m256i ih = _mm256_madd_epi16(_mm256_cvtepu8_epi16(h0), _mm256_loadu_si256((const __m256i*)(a + ab + 32)));
        ih = _mm256_add_epi32(ih, _mm256_madd_epi16(_mm256_cvtepu8_epi16(h1), _mm256_loadu_si256((const __m256i*)(a + ab + 48))));
        facc = _mm256_fmadd_ps(_mm256_cvtepi32_ps(ih), _mm256_set1_ps(d * (float)sch * as[gh]), facc);
        corr += dmin * (float)mnh * as[gh] * (float)asum[gh];
      }
    }
    out[row] = hsum_avx(facc) - corr;
  }
}

static void matvec_q8_0(const uint8_t* w, const int16_t* a, const float* as, int n, int k, float* out) {
  const int bpr = k / 32, row_bytes = bpr * 34;
  for (int row = 0; row < n; ++row) {
    const uint8_t* rd = w + (size_t)row * row_bytes;
    __m256 facc = _mm256_setzero_ps();
    for (int b = 0; b < bpr; ++b) {
      const uint8_t* blk = rd + b * 34;
      const float comb = f16_to_f32(blk) * as[b];
      __m256i ia = _mm256_madd_epi16(_mm256_cvtepi8_epi16(_mm_loadu_si128((const __m128i*)(blk + 2))),
                                     _mm256_loadu_si256((const __m256i*)(a + b * 32)));
      ia = _mm256_add_epi32(ia, _mm256_madd_epi16(_mm256_cvtepi8_epi16(_mm_loadu_si128((const __m128i*)(blk + 18))),
                                                  _mm256_loadu_si256((const __m256i*)(a + b * 32 + 16))));
      facc = _mm256_fmadd_ps(_mm256_cvtepi32_ps(ia), _mm256_set1_ps(comb), facc);
    }
    out[row] = hsum_avx(facc);
  }
}

static int gguf_row_bytes(int type, int k) { return type == 8 ? k / 32 * 34 : (type == 12 ? k / 256 * 144 : -1); }

static void gguf_matvec(int type, const uint8_t* w, const int16_t* a, const float* as, const int32_t* asum, int n, int k, float* out) {
  if (type == 8) matvec_q8_0(w, a, as, n, k, out);
  else matvec_q4_k(w, a, as, asum, n, k, out);
}

/*
 * gate/up: [E][I][row_bytes(t13, H)], down: [E][H][row_bytes(t2, I)] raw GGUF blocks.  Tokens sequentially; per token the
 * selected experts run in parallel (one task each, rayon par_iter in the reference), then out = sum_i w_i * expert_i in
 * expert order.  ids < 0 skipped.  Returns 0, or -1 for an unsupported type.
 */
int kcpu_moe_forward_gguf(const uint8_t* gate, const uint8_t* up, const uint8_t* down, int t13, int t2, int E, int H, int I,
                          const uint16_t* x_bf16, const int32_t* ids, const float* wts, int M, int topk, float* out,
                          int nthreads) {
  (void)E;
  const int rb13 = gguf_row_bytes(t13, H), rb2 = gguf_row_bytes(t2, I);
  if (rb13 < 0 || rb2 < 0 || H % 32 || I % 32) return -1;
  const size_t gu_e = (size_t)I * rb13, dn_e = (size_t)H * rb2;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  float* xf = (float*)malloc(sizeof(float) * H);
  int16_t* xa = (int16_t*)malloc(sizeof(int16_t) * H);
  float* xs = (float*)malloc(sizeof(float) * (H / 32));
  int32_t* xsum = (int32_t*)malloc(sizeof(int32_t) * (H / 32));
  float* go = (float*)malloc(sizeof(float) * (size_t)topk * 2 * I);       /* gate_out | up_out -> hidden (in place) */
  int16_t* ha = (int16_t*)malloc(sizeof(int16_t) * (size_t)topk * I);
  float* hs = (float*)malloc(sizeof(float) * (size_t)topk * (I / 32));
  int32_t* hsum = (int32_t*)malloc(sizeof(int32_t) * (size_t)topk * (I / 32));
  float* eo = (float*)malloc(sizeof(float) * (size_t)topk * H);
#pragma omp parallel
  for (int m = 0; m < M; ++m) {
#pragma omp single
    {
      for (int i = 0; i < H; ++i) xf[i] = bf16_to_f32(x_bf16[(size_t)m * H + i]);
      quant_act_i16_g32(xf, H, xa, xs, xsum);
    }
    const int32_t* id = ids + (size_t)m * topk;
#pragma omp for schedule(dynamic, 1)
    for (int j = 0; j < topk; ++j) {
      const int e = id[j];
      if (e < 0) continue;
      float* g = go + (size_t)j * 2 * I;
      float* u = g + I;
      gguf_matvec(t13, gate + e * gu_e, xa, xs, xsum, I, H, g);
      gguf_matvec(t13, up + e * gu_e, xa, xs, xsum, I, H, u);
      for (int i = 0; i < I; ++i) g[i] = g[i] / (1.0f + expf(-g[i])) * u[i];
      quant_act_i16_g32(g, I, ha + (size_t)j * I, hs + (size_t)j * (I / 32), hsum + (size_t)j * (I / 32));
      gguf_matvec(t2, down + e * dn_e, ha + (size_t)j * I, hs + (size_t)j * (I / 32), hsum + (size_t)j * (I / 32), H, I,
                  eo + (size_t)j * H);
    }
#pragma omp for schedule(static)
    for (int h = 0; h < H; ++h) {
      float s = 0.f;
      for (int j = 0; j < topk; ++j)
        if (id[j] >= 0) s += wts[(size_t)m * topk + j] * eo[(size_t)j * H + h];
      out[(size_t)m * H + h] = s;
    }
  }
  free(xf); free(xa); free(xs); free(xsum); free(go); free(ha); free(hs); free(hsum); free(eo);
  return 0;
}

int kcpu_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
