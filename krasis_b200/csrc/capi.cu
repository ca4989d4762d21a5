// C-ABI implementation (include/krasis_b200.h).  Host-side engine state + kernel orchestration.
// The reference keeps this state in Rust (`KrasisEngine`, src/moe.rs:1377-3296) and Python
// (`GpuPrefillManager`, python/krasis/gpu_prefill.py:326-4484); here it is C++ behind a C boundary.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/krasis_b200.h"
#include "moe_common.cuh"
#include "prof.cuh"
#include <atomic>
#include <mutex>

namespace kb2 {
cudaError_t launch_grouped_gemm(int fmt, bool gemm1, const GemmParams& p, const void* tmap_b, int num_sms,
                                cudaStream_t stream);
cudaError_t make_tmap_bf16_rows(void* out_tmap, const void* base, long long rows, long long cols, int box_rows);
int gemm_b_box_rows();
cudaError_t launch_router_logits(const void* h, const void* gate, const float* bias, float* logits, int M, int E,
                                 int H, cudaStream_t s);
bool router_gemm_supported(int E, int H);
cudaError_t launch_router_gemm(const void* x, const void* tmap_g, const float* bias, float* logits, int M, int E,
                               int H, cudaStream_t s);
cudaError_t launch_router_topk(const float* logits, const float* corr_bias, int M, int E, int top_k, int scoring,
                               int renorm, int* ids, float* wts, cudaStream_t s);
cudaError_t launch_binning(const int* ids, const float* wts, int M, int top_k, int e_start, int e_end, int* counts,
                           int* offsets, int* cursor, ChunkDesc* chunks, int* n_chunks, float* sorted_w, int* slot_of,
                           int* sorted_ids, const void* x, void* x_sorted, int H, cudaStream_t s);
cudaError_t launch_binning_index(const int* ids, const float* wts, int M, int top_k, int e_start, int e_end, int* counts,
                                 int* offsets, int* cursor, ChunkDesc* chunks, int* n_chunks, float* sorted_w, int* slot_of,
                                 int* sorted_tok, cudaStream_t s);
cudaError_t make_tmap_bf16_gather(void* out_tmap, const void* base, long long rows, long long cols);
cudaError_t launch_combine_scatter(const void* c3, const int* slot_of, int M, int H, int top_k, float rsf, int apply_rsf,
                                   const void* shared, void* out, const CombineScatter& sc, cudaStream_t s);
cudaError_t launch_combine(const void* c3, const int* slot_of, int M, int H, int top_k, float rsf, int apply_rsf,
                           const void* shared, void* out, cudaStream_t s);
cudaError_t launch_quantize_group(const void* w, int bits, void* q_out, void* scales, long long rows, int K, cudaStream_t s);
cudaError_t launch_retile_gguf(int fmt, const void* a, const void* b, int n_a, void* dst, int E, int N, int K,
                               cudaStream_t s);
cudaError_t launch_repack(int fmt, const void* src_q, const void* src_s, void* dst_q, void* dst_s, int E, int N, int K,
                          cudaStream_t s);
}  // namespace kb2

using namespace kb2;

// ---- process-wide kernel accounting (prof.cuh) --------------------------------------------------------------------------
namespace {
std::atomic<long long> g_launches{0};
std::atomic<bool> g_kprof_on{false};
std::mutex g_kprof_mu;
std::vector<cudaEvent_t> g_kev;                                   // event pool, reused across collects
size_t g_kev_used = 0;
std::vector<std::pair<int, int>> g_kspans[K_NUM];                 // (begin, end) event indices per kernel class
const char* const kKernelNames[K_NUM] = {
    "router_gemm", "router_topk", "binning(count+scan+scatter)", "grouped_gemm<gate_up+silu_mul>", "grouped_gemm<down>", "combine",
    "dense_gemm<bf16>", "dense_gemm<int8>", "gdn_prep(conv+l2norm+gates)", "gdn_conv_state", "gdn_chunk_prepare", "gdn_chunk_scan",
    "gdn_post(gated_rmsnorm)", "gqa_prep(norm+rope+fp8_append)", "kv_gather(fp8->bf16)", "fmha", "mla_prep", "rmsnorm", "quant_rows_int8",
    "silu_and_mul", "sigmoid_gate_mul", "add_bf16", "quantize_group", "retile"};
}  // namespace

namespace kb2 {
void kernel_span_begin(int id, cudaStream_t s, int n_launches, int* slot) {
  g_launches.fetch_add(n_launches, std::memory_order_relaxed);
  *slot = -1;
  if (!g_kprof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_kprof_mu);
  if (g_kev_used + 2 > g_kev.size())
    for (int i = 0; i < 512; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); g_kev.push_back(ev); }
  *slot = (int)g_kev_used;
  g_kev_used += 2;
  cudaEventRecord(g_kev[*slot], s);
}
void kernel_span_end(int id, cudaStream_t s, int slot) {
  std::lock_guard<std::mutex> lk(g_kprof_mu);
  cudaEventRecord(g_kev[slot + 1], s);
  g_kspans[id].push_back({slot, slot + 1});
}
}  // namespace kb2

constexpr bool kDefaultMoeGather = false;   // flip after the gather4 path is measured (KB2_MOE_GATHER overrides per call)

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

extern "C" int kb2_set_error_(int code, const char* msg) {   // shared with capi_attn.cu (not exported)
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return fail(KB2_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

struct LayerWeights {
  const uint8_t* w13_q = nullptr;
  const uint8_t* w13_s = nullptr;
  const uint8_t* w2_q = nullptr;
  const uint8_t* w2_s = nullptr;
  bool owned = false;
  void* gate = nullptr;        // [E][H] bf16
  float* gate_bias = nullptr;  // [E]
  float* corr_bias = nullptr;  // [E]
  alignas(64) unsigned char tmap_gate[128];   // CUtensorMap over gate (router GEMM B operand)
  bool has_tmap_gate = false;
};

struct kb2_engine {
  kb2_config cfg{};
  int e_start = 0, e_end = 0, e_local = 0;
  int num_sms = 0;
  std::vector<LayerWeights> layers;
  // scratch (sized for cfg.max_tokens)
  int *counts = nullptr, *offsets = nullptr, *cursor = nullptr, *n_chunks = nullptr;
  ChunkDesc* chunks = nullptr;
  int* slot_of = nullptr;
  int* sorted_tok = nullptr;                       // gather mode: token row of every sorted slot (+ kMaxChunkTokens of padding)
  void* x_sorted = nullptr;                        // [max_tokens*k][H] bf16: per-expert contiguous token tiles
  alignas(64) unsigned char tmap_x[128];           // CUtensorMap over x_sorted (GEMM1 B operand)
  alignas(64) unsigned char tmap_act[128];         // CUtensorMap over act      (GEMM2 B operand)
  float* sorted_w = nullptr;
  void *act = nullptr, *c3 = nullptr;
  float* logits = nullptr;
  int* ids_tmp = nullptr;
  float* w_tmp = nullptr;
  void *x_tmp = nullptr, *out_tmp = nullptr;
  int64_t launches = 0;
  // optional per-kernel timing (CUDA events on the launching stream)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_spans[KB2_PROF_NUM];   // (begin, end) indices into ev_pool
  size_t ev_used = 0;
};

struct ProfSpan {
  kb2_engine* e; int which; cudaStream_t s; int b = -1;
  ProfSpan(kb2_engine* e_, int w, cudaStream_t s_) : e(e_), which(w), s(s_) {
    if (!e->profiling) return;
    if (e->ev_used + 2 > e->ev_pool.size()) {
      for (int i = 0; i < 256; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); e->ev_pool.push_back(ev); }
    }
    b = (int)e->ev_used; e->ev_used += 2;
    cudaEventRecord(e->ev_pool[b], s);
  }
  ~ProfSpan() {
    if (b < 0) return;
    cudaEventRecord(e->ev_pool[b + 1], s);
    e->ev_spans[which].push_back({b, b + 1});
  }
};

// device buffers that are freed on scope exit unless released (error paths of the loaders must not leak)
struct DevBufs {
  std::vector<void*> p;
  explicit DevBufs(size_t n) : p(n, nullptr) {}
  ~DevBufs() { for (void* q : p) if (q) cudaFree(q); }
  void* release(size_t i) { void* q = p[i]; p[i] = nullptr; return q; }
  DevBufs(const DevBufs&) = delete;
  DevBufs& operator=(const DevBufs&) = delete;
};

static int fmt13(const kb2_engine* e) { return e->cfg.weight_format; }
static int fmt2(const kb2_engine* e) { return e->cfg.w2_weight_format >= 0 ? e->cfg.w2_weight_format : e->cfg.weight_format; }

// bytes of one (128-row x 64-K) tile blob, and of the separate scale tile (0 for self-contained GGUF blobs)
static size_t blob_bytes(int fmt) {
  switch (fmt) {
    case KB2_FMT_INT4_G128: return kInt4TileBytes;
    case KB2_FMT_INT8_G128: return kInt8TileBytes;
    case KB2_FMT_GGUF_Q8_0: return kQ8_0TileBytes;
    case KB2_FMT_GGUF_Q4_K: return kQ4KTileBytes;
    case KB2_FMT_GGUF_Q6_K: case KB2_FMT_GGUF_Q5_K: case KB2_FMT_GGUF_Q5_0: case KB2_FMT_GGUF_Q4_0: return kAffine8TileBytes;
  }
  return 0;
}
static size_t gguf_row_bytes(int fmt, size_t k) {      // src/gguf.rs:56-85
  switch (fmt) {
    case KB2_FMT_GGUF_Q8_0: return k / 32 * 34;
    case KB2_FMT_GGUF_Q4_K: return k / 256 * 144;
    case KB2_FMT_GGUF_Q6_K: return k / 256 * 210;
    case KB2_FMT_GGUF_Q5_K: return k / 256 * 176;
    case KB2_FMT_GGUF_Q5_0: return k / 32 * 22;
    case KB2_FMT_GGUF_Q4_0: return k / 32 * 18;
  }
  return 0;
}
static bool gguf_needs_k256(int fmt) { return fmt == KB2_FMT_GGUF_Q4_K || fmt == KB2_FMT_GGUF_Q6_K || fmt == KB2_FMT_GGUF_Q5_K; }
static bool has_scale_tiles(int fmt) { return fmt == KB2_FMT_INT4_G128 || fmt == KB2_FMT_INT8_G128; }

static size_t tiled_bytes(const kb2_engine* e, int which) {
  const size_t H = e->cfg.hidden_size, I = e->cfg.moe_intermediate_size, E = e->e_local;
  switch (which) {
    case 0: return E * (2 * I / kTileRows) * (H / kBlockK) * blob_bytes(fmt13(e));
    case 1: return has_scale_tiles(fmt13(e)) ? E * 2 * I * (H / kGroup) * 2 : 0;
    case 2: return E * (H / kTileRows) * (I / kBlockK) * blob_bytes(fmt2(e));
    case 3: return has_scale_tiles(fmt2(e)) ? E * H * (I / kGroup) * 2 : 0;
  }
  return 0;
}

extern "C" {

KB2_API const char* kb2_last_error(void) { return g_err.c_str(); }
KB2_API const char* kb2_version(void) { return "krasis_b200 0.1 (sm_100a)"; }

KB2_API int kb2_create(const kb2_config* c, kb2_engine** out) {
  if (!c || !out) return fail(KB2_ERR_VALUE, "null argument");
  if (c->hidden_size <= 0 || c->hidden_size % 256)
    return fail(KB2_ERR_VALUE, "hidden_size (%d) must be a positive multiple of 256", c->hidden_size);
  if (c->moe_intermediate_size <= 0 || c->moe_intermediate_size % 128)
    return fail(KB2_ERR_VALUE, "moe_intermediate_size (%d) must be a positive multiple of 128", c->moe_intermediate_size);
  if (c->num_ranks < 1 || c->rank < 0 || c->rank >= c->num_ranks) return fail(KB2_ERR_VALUE, "bad rank %d/%d", c->rank, c->num_ranks);
  if (c->n_routed_experts < c->num_ranks) return fail(KB2_ERR_VALUE, "fewer experts than ranks");
  if (c->num_experts_per_tok < 1 || c->num_experts_per_tok > 32) return fail(KB2_ERR_VALUE, "top-k must be in [1,32]");
  if (c->weight_format < KB2_FMT_INT4_G128 || c->weight_format > KB2_FMT_GGUF_Q4_0)
    return fail(KB2_ERR_VALUE, "unknown weight_format %d", c->weight_format);
  if (c->w2_weight_format > KB2_FMT_GGUF_Q4_0) return fail(KB2_ERR_VALUE, "unknown w2_weight_format %d", c->w2_weight_format);
  {
    const int f2 = c->w2_weight_format >= 0 ? c->w2_weight_format : c->weight_format;
    if (gguf_needs_k256(c->weight_format) && c->hidden_size % 256) return fail(KB2_ERR_VALUE, "K-quant gate/up needs hidden_size %% 256 == 0");
    if (gguf_needs_k256(f2) && c->moe_intermediate_size % 256) return fail(KB2_ERR_VALUE, "K-quant down needs moe_intermediate_size %% 256 == 0 (use a 32-element type: Q8_0 / Q5_0 / Q4_0)");
  }
  if (c->max_tokens < 1 || c->num_moe_layers < 1) return fail(KB2_ERR_VALUE, "max_tokens and num_moe_layers must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(KB2_ERR_CUDA, "no CUDA device: krasis_b200 has no CPU fallback");
  }
  CUDA_TRY(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major != 10) return fail(KB2_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", c->device, prop.major, prop.minor);

  kb2_engine* e = new kb2_engine();
  e->cfg = *c;
  const int per = c->n_routed_experts / c->num_ranks;            // gpu_prefill.py:353-359
  e->e_start = c->rank * per;
  e->e_end = (c->rank == c->num_ranks - 1) ? c->n_routed_experts : (c->rank + 1) * per;
  e->e_local = e->e_end - e->e_start;
  if (e->e_local > 1024) { delete e; return fail(KB2_ERR_VALUE, "more than 1024 local experts"); }
  if (c->n_routed_experts > 1024) { delete e; return fail(KB2_ERR_VALUE, "more than 1024 experts"); }
  e->num_sms = prop.multiProcessorCount;
  e->layers.resize(c->num_moe_layers);
  const size_t MK = (size_t)c->max_tokens * c->num_experts_per_tok;
  const size_t max_chunks = c->n_routed_experts + MK / kMaxChunkTokens + 1;
  // a failed allocation must not leak the engine or the buffers made so far: kb2_destroy frees whatever exists
#define ALLOC(ptr, bytes)                                                                               \
  do {                                                                                                  \
    cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));                                               \
    if (_e != cudaSuccess) {                                                                            \
      kb2_destroy(e);                                                                                   \
      return fail(KB2_ERR_CUDA, "cudaMalloc(%s, %zu bytes): %s", #ptr, (size_t)(bytes), cudaGetErrorString(_e)); \
    }                                                                                                   \
  } while (0)
  ALLOC(e->counts, sizeof(int) * c->n_routed_experts);
  ALLOC(e->offsets, sizeof(int) * (c->n_routed_experts + 1));
  ALLOC(e->cursor, sizeof(int) * c->n_routed_experts);
  ALLOC(e->n_chunks, sizeof(int));
  ALLOC(e->chunks, sizeof(ChunkDesc) * max_chunks);
  ALLOC(e->x_sorted, MK * c->hidden_size * 2);
  ALLOC(e->slot_of, sizeof(int) * MK);
  ALLOC(e->sorted_tok, sizeof(int) * (MK + kMaxChunkTokens));
  if (cudaMemset(e->sorted_tok, 0, sizeof(int) * (MK + kMaxChunkTokens)) != cudaSuccess) { kb2_destroy(e); return fail(KB2_ERR_CUDA, "cudaMemset failed"); }
  ALLOC(e->sorted_w, sizeof(float) * MK);
  ALLOC(e->act, MK * c->moe_intermediate_size * 2);
  ALLOC(e->c3, MK * c->hidden_size * 2);
  ALLOC(e->logits, sizeof(float) * (size_t)c->max_tokens * c->n_routed_experts);
  ALLOC(e->ids_tmp, sizeof(int) * MK);
  ALLOC(e->w_tmp, sizeof(float) * MK);
  ALLOC(e->x_tmp, (size_t)c->max_tokens * c->hidden_size * 2);
  ALLOC(e->out_tmp, (size_t)c->max_tokens * c->hidden_size * 2);
#undef ALLOC
  cudaError_t te = make_tmap_bf16_rows(e->tmap_x, e->x_sorted, (long long)MK, c->hidden_size, gemm_b_box_rows());
  if (te == cudaSuccess) te = make_tmap_bf16_rows(e->tmap_act, e->act, (long long)MK, c->moe_intermediate_size, gemm_b_box_rows());
  if (te != cudaSuccess) {
    kb2_destroy(e);
    return fail(KB2_ERR_CUDA, "tensor map creation failed: %s", cudaGetErrorString(te));
  }
  *out = e;
  return KB2_OK;
}

static void free_layer(LayerWeights& L) {
  if (L.owned) {
    cudaFree((void*)L.w13_q); cudaFree((void*)L.w13_s); cudaFree((void*)L.w2_q); cudaFree((void*)L.w2_s);
  }
  L.w13_q = L.w13_s = L.w2_q = L.w2_s = nullptr;
  L.owned = false;
}

KB2_API void kb2_destroy(kb2_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  for (auto& L : e->layers) {
    free_layer(L);
    cudaFree(L.gate); cudaFree(L.gate_bias); cudaFree(L.corr_bias);
  }
  cudaFree(e->counts); cudaFree(e->offsets); cudaFree(e->cursor); cudaFree(e->n_chunks); cudaFree(e->chunks);
  cudaFree(e->x_sorted); cudaFree(e->slot_of); cudaFree(e->sorted_tok); cudaFree(e->sorted_w); cudaFree(e->act); cudaFree(e->c3);
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  cudaFree(e->logits); cudaFree(e->ids_tmp); cudaFree(e->w_tmp); cudaFree(e->x_tmp); cudaFree(e->out_tmp);
  delete e;
}

KB2_API int kb2_get_config(const kb2_engine* e, kb2_config* out) {
  if (!e || !out) return fail(KB2_ERR_VALUE, "null argument");
  *out = e->cfg;
  return KB2_OK;
}

KB2_API int kb2_expert_range(const kb2_engine* e, int32_t* s, int32_t* t) {
  if (!e || !s || !t) return fail(KB2_ERR_VALUE, "null argument");
  *s = e->e_start; *t = e->e_end;
  return KB2_OK;
}

KB2_API size_t kb2_tiled_bytes(const kb2_engine* e, int which) { return e ? tiled_bytes(e, which) : 0; }
KB2_API int64_t kb2_launch_count(const kb2_engine* e) { return e ? e->launches : 0; }

static int check_layer(kb2_engine* e, int layer) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  if (layer < 0 || layer >= (int)e->layers.size())
    return fail(KB2_ERR_VALUE, "moe_layer_idx %d out of range [0, %d)", layer, (int)e->layers.size());
  return KB2_OK;
}

KB2_API int kb2_retile_dev(kb2_engine* e, int fmt, const void* sq, const void* ss, void* dq, void* ds, int n_experts,
                   int n_rows, int k_cols, void* stream) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaError_t r = launch_repack(fmt, sq, ss, dq, ds, n_experts, n_rows, k_cols, (cudaStream_t)stream);
  if (r != cudaSuccess) return fail(r == cudaErrorInvalidValue ? KB2_ERR_VALUE : KB2_ERR_CUDA, "retile: %s", cudaGetErrorString(r));
  e->launches += 2;
  return KB2_OK;
}

KB2_API int kb2_load_experts_host(kb2_engine* e, int layer, const void* w13_q, const void* w13_s, const void* w2_q,
                          const void* w2_s) {
  if (int r = check_layer(e, layer)) return r;
  if (!w13_q || !w13_s || !w2_q || !w2_s) return fail(KB2_ERR_VALUE, "null weight pointer");
  if (!has_scale_tiles(fmt13(e)) || !has_scale_tiles(fmt2(e)) || fmt13(e) != fmt2(e))   // validate BEFORE touching the layer
    return fail(KB2_ERR_STATE, "kb2_load_experts_host takes the INT4/INT8 group-quantised layout; this engine was created for GGUF blocks");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const int H = e->cfg.hidden_size, I = e->cfg.moe_intermediate_size, fmt = e->cfg.weight_format;
  DevBufs dst(4), tmp(4);
  const void* src[4] = {w13_q, w13_s, w2_q, w2_s};
  for (int i = 0; i < 4; ++i) {
    const size_t b = tiled_bytes(e, i);   // the re-tiling is a permutation: same byte counts
    CUDA_TRY(cudaMalloc(&dst.p[i], b));
    CUDA_TRY(cudaMalloc(&tmp.p[i], b));
    CUDA_TRY(cudaMemcpy(tmp.p[i], src[i], b, cudaMemcpyHostToDevice));
  }
  cudaError_t r1 = launch_repack(fmt, tmp.p[0], tmp.p[1], dst.p[0], dst.p[1], e->e_local, 2 * I, H, 0);
  cudaError_t r2 = launch_repack(fmt, tmp.p[2], tmp.p[3], dst.p[2], dst.p[3], e->e_local, H, I, 0);
  e->launches += 4;
  CUDA_TRY(cudaDeviceSynchronize());
  if (r1 != cudaSuccess || r2 != cudaSuccess) return fail(KB2_ERR_CUDA, "retile failed");
  LayerWeights& L = e->layers[layer];
  free_layer(L);                          // the old weights go only after the new ones exist
  L.w13_q = (const uint8_t*)dst.release(0); L.w13_s = (const uint8_t*)dst.release(1);
  L.w2_q = (const uint8_t*)dst.release(2); L.w2_s = (const uint8_t*)dst.release(3);
  L.owned = true;
  return KB2_OK;
}

KB2_API int kb2_quantize_group_dev(const void* w_bf16_dev, int32_t num_bits, void* q_dev, void* scales_dev, int64_t rows,
                                   int32_t k_cols, int32_t device, void* stream) {
  if (!w_bf16_dev || !q_dev || !scales_dev) return fail(KB2_ERR_VALUE, "null argument");
  if ((num_bits != 4 && num_bits != 8) || rows <= 0 || k_cols <= 0 || k_cols % kGroup)
    return fail(KB2_ERR_VALUE, "quantize: num_bits must be 4 or 8 and cols (%d) divisible by group_size (128)", k_cols);   // marlin.rs:152
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_quantize_group(w_bf16_dev, num_bits, q_dev, scales_dev, rows, k_cols, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_load_experts_dev(kb2_engine* e, int layer, const void* w13_q_dev, const void* w13_s_dev, const void* w2_q_dev,
                         const void* w2_s_dev, void* stream) {
  if (int r = check_layer(e, layer)) return r;
  if (!w13_q_dev || !w13_s_dev || !w2_q_dev || !w2_s_dev) return fail(KB2_ERR_VALUE, "null weight pointer");
  if (!has_scale_tiles(fmt13(e)) || fmt13(e) != fmt2(e)) return fail(KB2_ERR_STATE, "engine was created for GGUF blocks");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const int H = e->cfg.hidden_size, I = e->cfg.moe_intermediate_size, fmt = e->cfg.weight_format;
  DevBufs dst(4);
  for (int i = 0; i < 4; ++i) CUDA_TRY(cudaMalloc(&dst.p[i], tiled_bytes(e, i)));
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(launch_repack(fmt, w13_q_dev, w13_s_dev, dst.p[0], dst.p[1], e->e_local, 2 * I, H, s));
  CUDA_TRY(launch_repack(fmt, w2_q_dev, w2_s_dev, dst.p[2], dst.p[3], e->e_local, H, I, s));
  e->launches += 4;
  LayerWeights& L = e->layers[layer];
  free_layer(L);
  L.w13_q = (const uint8_t*)dst.release(0); L.w13_s = (const uint8_t*)dst.release(1);
  L.w2_q = (const uint8_t*)dst.release(2); L.w2_s = (const uint8_t*)dst.release(3);
  L.owned = true;
  return KB2_OK;
}

KB2_API int kb2_load_experts_gguf_host(kb2_engine* e, int layer, const void* gate_host, const void* up_host, const void* down_host) {
  if (int r = check_layer(e, layer)) return r;
  if (!gate_host || !up_host || !down_host) return fail(KB2_ERR_VALUE, "null weight pointer");
  const int f13 = fmt13(e), f2 = fmt2(e);
  if (has_scale_tiles(f13) || has_scale_tiles(f2)) return fail(KB2_ERR_STATE, "this engine was not created for GGUF block formats");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const size_t H = e->cfg.hidden_size, I = e->cfg.moe_intermediate_size, E = e->e_local;
  const size_t gu_bytes = E * I * gguf_row_bytes(f13, H), dn_bytes = E * H * gguf_row_bytes(f2, I);
  DevBufs src(3), til(2);      // raw blocks (freed on return), tiles (kept on success)
  CUDA_TRY(cudaMalloc(&src.p[0], gu_bytes)); CUDA_TRY(cudaMalloc(&src.p[1], gu_bytes)); CUDA_TRY(cudaMalloc(&src.p[2], dn_bytes));
  CUDA_TRY(cudaMemcpy(src.p[0], gate_host, gu_bytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(src.p[1], up_host, gu_bytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(src.p[2], down_host, dn_bytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&til.p[0], tiled_bytes(e, 0))); CUDA_TRY(cudaMalloc(&til.p[1], tiled_bytes(e, 2)));
  cudaError_t r1 = launch_retile_gguf(f13, src.p[0], src.p[1], (int)I, til.p[0], (int)E, (int)(2 * I), (int)H, 0);
  cudaError_t r2 = launch_retile_gguf(f2, src.p[2], src.p[2], (int)H, til.p[1], (int)E, (int)H, (int)I, 0);
  e->launches += 2;
  CUDA_TRY(cudaDeviceSynchronize());
  if (r1 != cudaSuccess || r2 != cudaSuccess) return fail(KB2_ERR_CUDA, "GGUF re-tiling failed");
  LayerWeights& L = e->layers[layer];
  free_layer(L);
  L.w13_q = (const uint8_t*)til.release(0); L.w2_q = (const uint8_t*)til.release(1); L.w13_s = nullptr; L.w2_s = nullptr;
  L.owned = true;
  return KB2_OK;
}

KB2_API int kb2_load_experts_tiled_host(kb2_engine* e, int layer, const void* w13_q, const void* w13_s, const void* w2_q,
                                        const void* w2_s) {
  if (int r = check_layer(e, layer)) return r;
  if (!w13_q || !w2_q || (has_scale_tiles(fmt13(e)) && !w13_s) || (has_scale_tiles(fmt2(e)) && !w2_s)) return fail(KB2_ERR_VALUE, "null weight pointer");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  DevBufs dst(4);
  const void* src[4] = {w13_q, w13_s, w2_q, w2_s};
  for (int i = 0; i < 4; ++i) {
    const size_t b = tiled_bytes(e, i);
    if (!b) continue;
    CUDA_TRY(cudaMalloc(&dst.p[i], b));
    CUDA_TRY(cudaMemcpy(dst.p[i], src[i], b, cudaMemcpyHostToDevice));
  }
  LayerWeights& L = e->layers[layer];
  free_layer(L);
  L.w13_q = (const uint8_t*)dst.release(0); L.w13_s = (const uint8_t*)dst.release(1);
  L.w2_q = (const uint8_t*)dst.release(2); L.w2_s = (const uint8_t*)dst.release(3);
  L.owned = true;
  return KB2_OK;
}

KB2_API int kb2_export_experts_tiled_host(kb2_engine* e, int layer, int which, void* dst_host, size_t dst_bytes) {
  if (int r = check_layer(e, layer)) return r;
  if (which < 0 || which > 3 || !dst_host) return fail(KB2_ERR_VALUE, "bad argument");
  LayerWeights& L = e->layers[layer];
  if (!L.w13_q) return fail(KB2_ERR_STATE, "GPU weights not available for layer %d", layer);
  const size_t b = tiled_bytes(e, which);
  if (dst_bytes != b) return fail(KB2_ERR_VALUE, "buffer %d: expected %zu bytes, got %zu", which, b, dst_bytes);   // moe.rs:2285-2300
  const void* src[4] = {L.w13_q, L.w13_s, L.w2_q, L.w2_s};
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  if (b) CUDA_TRY(cudaMemcpy(dst_host, src[which], b, cudaMemcpyDeviceToHost));
  return KB2_OK;
}

KB2_API int kb2_attach_experts_tiled_dev(kb2_engine* e, int layer, const void* w13_q, const void* w13_s, const void* w2_q,
                                 const void* w2_s) {
  if (int r = check_layer(e, layer)) return r;
  if (!w13_q || !w2_q || (has_scale_tiles(fmt13(e)) && !w13_s) || (has_scale_tiles(fmt2(e)) && !w2_s)) return fail(KB2_ERR_VALUE, "null weight pointer");
  LayerWeights& L = e->layers[layer];
  free_layer(L);
  L.w13_q = (const uint8_t*)w13_q; L.w13_s = (const uint8_t*)w13_s;
  L.w2_q = (const uint8_t*)w2_q; L.w2_s = (const uint8_t*)w2_s;
  L.owned = false;
  return KB2_OK;
}

KB2_API int kb2_set_router_host(kb2_engine* e, int layer, const void* gate, const float* bias, const float* corr) {
  if (int r = check_layer(e, layer)) return r;
  if (!gate) return fail(KB2_ERR_VALUE, "null gate weight");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  LayerWeights& L = e->layers[layer];
  const size_t E = e->cfg.n_routed_experts, H = e->cfg.hidden_size;
  cudaFree(L.gate); cudaFree(L.gate_bias); cudaFree(L.corr_bias);
  L.gate = nullptr; L.gate_bias = nullptr; L.corr_bias = nullptr;
  CUDA_TRY(cudaMalloc(&L.gate, E * H * 2));
  CUDA_TRY(cudaMemcpy(L.gate, gate, E * H * 2, cudaMemcpyHostToDevice));
  L.has_tmap_gate = false;
  if (router_gemm_supported((int)E, (int)H)) {
    CUDA_TRY(make_tmap_bf16_rows(L.tmap_gate, L.gate, (long long)E, (long long)H, 64));
    L.has_tmap_gate = true;
  }
  if (bias) {
    CUDA_TRY(cudaMalloc((void**)&L.gate_bias, E * 4));
    CUDA_TRY(cudaMemcpy(L.gate_bias, bias, E * 4, cudaMemcpyHostToDevice));
  }
  if (corr) {
    CUDA_TRY(cudaMalloc((void**)&L.corr_bias, E * 4));
    CUDA_TRY(cudaMemcpy(L.corr_bias, corr, E * 4, cudaMemcpyHostToDevice));
  }
  return KB2_OK;
}

KB2_API int kb2_route(kb2_engine* e, int layer, const void* hidden, int32_t M, int32_t* ids, float* wts, void* stream) {
  if (int r = check_layer(e, layer)) return r;
  LayerWeights& L = e->layers[layer];
  if (!L.gate) return fail(KB2_ERR_STATE, "router weights not set for layer %d", layer);
  if (M < 0 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, e->cfg.max_tokens);
  if (M == 0) return KB2_OK;
  if (!hidden || !ids || !wts) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  { ProfSpan ps(e, KB2_PROF_ROUTER_LOGITS, s);
    if (L.has_tmap_gate)   // tcgen05 path (E <= 512); the fp32-FMA kernel remains for larger / odd expert counts
      CUDA_TRY(launch_router_gemm(hidden, L.tmap_gate, L.gate_bias, e->logits, M, e->cfg.n_routed_experts, e->cfg.hidden_size, s));
    else
      CUDA_TRY(launch_router_logits(hidden, L.gate, L.gate_bias, e->logits, M, e->cfg.n_routed_experts, e->cfg.hidden_size, s)); }
  ProfSpan ps2(e, KB2_PROF_ROUTER_TOPK, s);
  CUDA_TRY(launch_router_topk(e->logits, L.corr_bias, M, e->cfg.n_routed_experts, e->cfg.num_experts_per_tok,
                              e->cfg.scoring_func, e->cfg.norm_topk_prob, ids, wts, s));
  e->launches += 2;
  return KB2_OK;
}

// Shared implementation: `n_rows` activation rows, `K` routing entries per row.
static int moe_forward_impl(kb2_engine* e, int layer, const void* x, const int32_t* ids, const float* wts, void* out,
                            int M, int K, int apply_rsf, const void* shared, cudaStream_t s, const CombineScatter* scatter = nullptr) {
  LayerWeights& L = e->layers[layer];
  const int H = e->cfg.hidden_size, I = e->cfg.moe_intermediate_size;
  const int f13 = fmt13(e), f2 = fmt2(e);

  // KB2_MOE_GATHER=1: the gate/up GEMM fetches token rows from x itself with TMA gather4; no x_sorted copy is made
  const char* gev = getenv("KB2_MOE_GATHER");
  const bool gather = gev ? gev[0] == '1' : kDefaultMoeGather;
  alignas(64) unsigned char tmap_gather[128];
  { ProfSpan ps(e, KB2_PROF_BINNING, s);
    if (gather) {
      CUDA_TRY(launch_binning_index(ids, wts, M, K, e->e_start, e->e_end, e->counts, e->offsets, e->cursor, e->chunks,
                                    e->n_chunks, e->sorted_w, e->slot_of, e->sorted_tok, s));
      CUDA_TRY(make_tmap_bf16_gather(tmap_gather, x, M, H));
    } else {
      CUDA_TRY(launch_binning(ids, wts, M, K, e->e_start, e->e_end, e->counts, e->offsets, e->cursor, e->chunks,
                              e->n_chunks, e->sorted_w, e->slot_of, nullptr, x, e->x_sorted, H, s));
    } }
  e->launches += 3;

  GemmParams g1{};
  g1.wq = L.w13_q; g1.ws = L.w13_s;
  g1.wq_expert_stride = (long long)(tiled_bytes(e, 0) / e->e_local);
  g1.ws_expert_stride = (long long)(tiled_bytes(e, 1) / e->e_local);   // 0 for GGUF blobs
  g1.n_kblocks = H / kBlockK;
  g1.items_per_chunk = I / kTileRows; g1.tile1_offset = I / kTileRows; g1.tile0_mul = 1;
  g1.chunks = e->chunks; g1.n_chunks = e->n_chunks;
  g1.out = (__nv_bfloat16*)e->act; g1.out_ld = I; g1.slot_weight = nullptr;
  g1.gather_rows = gather ? e->sorted_tok : nullptr;
  { ProfSpan ps(e, KB2_PROF_GEMM1, s);
    CUDA_TRY(launch_grouped_gemm(f13, true, g1, gather ? tmap_gather : e->tmap_x, e->num_sms, s)); }

  GemmParams g2{};
  g2.wq = L.w2_q; g2.ws = L.w2_s;
  g2.wq_expert_stride = (long long)(tiled_bytes(e, 2) / e->e_local);
  g2.ws_expert_stride = (long long)(tiled_bytes(e, 3) / e->e_local);
  g2.n_kblocks = I / kBlockK;
  g2.items_per_chunk = H / (2 * kTileRows); g2.tile1_offset = 1; g2.tile0_mul = 2;
  g2.chunks = e->chunks; g2.n_chunks = e->n_chunks;
  g2.out = (__nv_bfloat16*)e->c3; g2.out_ld = H; g2.slot_weight = e->sorted_w;
  { ProfSpan ps(e, KB2_PROF_GEMM2, s);
    CUDA_TRY(launch_grouped_gemm(f2, false, g2, e->tmap_act, e->num_sms, s)); }

  { ProfSpan ps(e, KB2_PROF_COMBINE, s);
    if (scatter) CUDA_TRY(launch_combine_scatter(e->c3, e->slot_of, M, H, K, e->cfg.routed_scaling_factor, 0, nullptr, nullptr, *scatter, s));
    else CUDA_TRY(launch_combine(e->c3, e->slot_of, M, H, K, e->cfg.routed_scaling_factor, apply_rsf, shared, out, s)); }
  e->launches += 3;
  return KB2_OK;
}

KB2_API int kb2_moe_forward(kb2_engine* e, int layer, const void* x, const int32_t* ids, const float* wts, void* out,
                    int32_t M, int32_t routed_only, const void* shared, void* stream) {
  if (int r = check_layer(e, layer)) return r;
  LayerWeights& L = e->layers[layer];
  if (M < 0 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, e->cfg.max_tokens);
  if (!L.w13_q) return fail(KB2_ERR_STATE, "GPU weights not available for layer %d", layer);
  if (M == 0) return KB2_OK;
  if (!x || !ids || !wts || !out) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  return moe_forward_impl(e, layer, x, ids, wts, out, M, e->cfg.num_experts_per_tok, routed_only ? 0 : 1,
                          routed_only ? nullptr : shared, (cudaStream_t)stream);
}

// Expert-parallel forward with the reduce-scatter fused into the combine kernel: this rank's partial sums (local expert slice, ALL
// M = num_ranks * rows_per_rank tokens) are written straight into the receive buffers of the token owners (peer-mapped memory from
// kb2_comm_peer_alloc, layout [rows_per_rank][num_ranks][H] bf16, slot = source rank).  After kb2_comm_barrier the owner calls
// kb2_finish_routed_slots.  Replaces kb2_moe_forward(routed_only) + kb2_comm_reduce_scatter_bf16 + kb2_finish_routed.
KB2_API int kb2_moe_forward_scatter(kb2_engine* e, int layer, const void* x, const int32_t* ids, const float* wts,
                                    void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank, int32_t M, void* stream) {
  if (int r = check_layer(e, layer)) return r;
  LayerWeights& L = e->layers[layer];
  if (M < 0 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, e->cfg.max_tokens);
  if (num_ranks < 1 || num_ranks > kMaxPeers || src_rank < 0 || src_rank >= num_ranks || M % num_ranks)
    return fail(KB2_ERR_VALUE, "scatter: need 1 <= num_ranks <= %d, 0 <= src_rank < num_ranks, num_tokens %% num_ranks == 0", kMaxPeers);
  if (!L.w13_q) return fail(KB2_ERR_STATE, "GPU weights not available for layer %d", layer);
  if (M == 0) return KB2_OK;
  if (!x || !ids || !wts || !peer_recv_host) return fail(KB2_ERR_VALUE, "null argument");
  CombineScatter sc;
  for (int r = 0; r < num_ranks; ++r) {
    if (!peer_recv_host[r]) return fail(KB2_ERR_VALUE, "scatter: null receive buffer for rank %d", r);
    sc.peer_out[r] = (__nv_bfloat16*)peer_recv_host[r];
  }
  sc.rows_per_rank = M / num_ranks; sc.src_rank = src_rank; sc.n_ranks = num_ranks;
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  return moe_forward_impl(e, layer, x, ids, wts, nullptr, M, e->cfg.num_experts_per_tok, 0, nullptr, (cudaStream_t)stream, &sc);
}

// out[m] = bf16(rsf * bf16(sum_r f32(slots[m][r]))) + shared[m]: the consumer side of the fused reduce-scatter (fp32 sum of the
// num_ranks BF16 partials, one rounding — NCCL's ring rounds to BF16 at every hop).
KB2_API int kb2_finish_routed_slots(kb2_engine* e, const void* slots, int32_t num_ranks, const void* shared, void* out, int32_t rows,
                                    void* stream) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  if (rows < 0 || rows > e->cfg.max_tokens || num_ranks < 1 || num_ranks > kMaxPeers) return fail(KB2_ERR_VALUE, "bad rows / num_ranks");
  if (rows == 0) return KB2_OK;
  if (!slots || !out) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  ProfSpan ps(e, KB2_PROF_COMBINE, s);
  CUDA_TRY(launch_combine(slots, nullptr, rows, e->cfg.hidden_size, num_ranks, e->cfg.routed_scaling_factor, 1, shared, out, s));
  e->launches += 1;
  return KB2_OK;
}

// ---- expert-parallel building blocks (all-to-all dispatch / combine) -------------------------
KB2_API int kb2_ep_bin_rows(kb2_engine* e, const void* x, const int32_t* ids, const float* wts, int32_t M, void* x_sorted,
                    float* w_sorted, int32_t* ids_sorted, int32_t* slot_of, int32_t* counts, void* stream) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  if (M < 0 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, e->cfg.max_tokens);
  if (M == 0) return KB2_OK;
  if (!x || !ids || !wts || !x_sorted || !w_sorted || !ids_sorted || !slot_of || !counts) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  ProfSpan ps(e, KB2_PROF_BINNING, s);
  CUDA_TRY(launch_binning(ids, wts, M, e->cfg.num_experts_per_tok, 0, e->cfg.n_routed_experts, counts, e->offsets,
                          e->cursor, e->chunks, e->n_chunks, w_sorted, slot_of, ids_sorted, x, x_sorted,
                          e->cfg.hidden_size, s));
  e->launches += 3;
  return KB2_OK;
}

KB2_API int kb2_moe_forward_rows(kb2_engine* e, int layer, const void* rows, const int32_t* expert_ids, const float* wts,
                         void* out_rows, int32_t n_rows, void* stream) {
  if (int r = check_layer(e, layer)) return r;
  LayerWeights& L = e->layers[layer];
  const long long cap = (long long)e->cfg.max_tokens * e->cfg.num_experts_per_tok;
  if (n_rows < 0 || n_rows > cap) return fail(KB2_ERR_VALUE, "n_rows %d outside [0, max_tokens*top_k=%lld]", n_rows, cap);
  if (!L.w13_q) return fail(KB2_ERR_STATE, "GPU weights not available for layer %d", layer);
  if (n_rows == 0) return KB2_OK;
  if (!rows || !expert_ids || !wts || !out_rows) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  return moe_forward_impl(e, layer, rows, expert_ids, wts, out_rows, n_rows, 1, 0, nullptr, (cudaStream_t)stream);
}

KB2_API int kb2_ep_combine_rows(kb2_engine* e, const void* rows_sorted, const int32_t* slot_of, int32_t M, int32_t routed_only,
                        const void* shared, void* out, void* stream) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  if (M < 0 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, e->cfg.max_tokens);
  if (M == 0) return KB2_OK;
  if (!rows_sorted || !slot_of || !out) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  ProfSpan ps(e, KB2_PROF_COMBINE, s);
  CUDA_TRY(launch_combine(rows_sorted, slot_of, M, e->cfg.hidden_size, e->cfg.num_experts_per_tok,
                          e->cfg.routed_scaling_factor, routed_only ? 0 : 1, routed_only ? nullptr : shared, out, s));
  e->launches += 1;
  return KB2_OK;
}

KB2_API int kb2_finish_routed(kb2_engine* e, const void* routed, const void* shared, void* out, int32_t M, void* stream) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  if (M < 0 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, e->cfg.max_tokens);
  if (M == 0) return KB2_OK;
  if (!routed || !out) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  ProfSpan ps(e, KB2_PROF_COMBINE, s);
  CUDA_TRY(launch_combine(routed, nullptr, M, e->cfg.hidden_size, 1, e->cfg.routed_scaling_factor, 1, shared, out, s));
  e->launches += 1;
  return KB2_OK;
}

KB2_API int kb2_moe_forward_host(kb2_engine* e, int layer, const void* x_host, const int32_t* ids_host, const float* w_host,
                         void* out_host, int32_t M, int32_t routed_only, void* stream) {
  if (int r = check_layer(e, layer)) return r;
  if (M < 0 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, e->cfg.max_tokens);
  if (M == 0) return KB2_OK;
  if (!x_host || !out_host) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t H = e->cfg.hidden_size, K = e->cfg.num_experts_per_tok;
  CUDA_TRY(cudaMemcpyAsync(e->x_tmp, x_host, (size_t)M * H * 2, cudaMemcpyHostToDevice, s));
  if (ids_host) {
    if (!w_host) return fail(KB2_ERR_VALUE, "topk_weights_host is NULL but topk_ids_host is not");
    CUDA_TRY(cudaMemcpyAsync(e->ids_tmp, ids_host, (size_t)M * K * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->w_tmp, w_host, (size_t)M * K * 4, cudaMemcpyHostToDevice, s));
  } else {
    if (int r = kb2_route(e, layer, e->x_tmp, M, e->ids_tmp, e->w_tmp, stream)) return r;
  }
  if (int r = kb2_moe_forward(e, layer, e->x_tmp, e->ids_tmp, e->w_tmp, e->out_tmp, M, routed_only, nullptr, stream)) return r;
  CUDA_TRY(cudaMemcpyAsync(out_host, e->out_tmp, (size_t)M * H * 2, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return KB2_OK;
}

KB2_API int kb2_prefill_moe_stack_host(kb2_engine* e, const void* x_host, void* out_host, int32_t M, int32_t first_layer,
                                       int32_t n_layers, void* stream) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  if (first_layer < 0 || n_layers < 1 || first_layer + n_layers > (int)e->layers.size())
    return fail(KB2_ERR_VALUE, "layer range [%d, %d) outside [0, %d)", first_layer, first_layer + n_layers, (int)e->layers.size());
  if (M < 1 || M > e->cfg.max_tokens) return fail(KB2_ERR_VALUE, "num_tokens %d outside [1, max_tokens=%d]", M, e->cfg.max_tokens);
  if (!x_host || !out_host) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t H = e->cfg.hidden_size;
  CUDA_TRY(cudaMemcpyAsync(e->x_tmp, x_host, (size_t)M * H * 2, cudaMemcpyHostToDevice, s));
  for (int l = first_layer; l < first_layer + n_layers; ++l) {
    if (int r = kb2_route(e, l, e->x_tmp, M, e->ids_tmp, e->w_tmp, stream)) return r;
    if (int r = kb2_moe_forward(e, l, e->x_tmp, e->ids_tmp, e->w_tmp, e->out_tmp, M, 0, nullptr, stream)) return r;
  }
  CUDA_TRY(cudaMemcpyAsync(out_host, e->out_tmp, (size_t)M * H * 2, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return KB2_OK;
}

KB2_API int kb2_last_expert_counts(kb2_engine* e, int32_t* counts_host, void* stream) {
  if (!e || !counts_host) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  CUDA_TRY(cudaMemcpyAsync(counts_host, e->counts, sizeof(int) * e->e_local, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int64_t kb2_total_launches(void) { return (int64_t)g_launches.load(); }
KB2_API int kb2_kernel_profile_num(void) { return K_NUM; }
KB2_API const char* kb2_kernel_profile_name(int id) { return id >= 0 && id < K_NUM ? kKernelNames[id] : ""; }
KB2_API int kb2_kernel_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_kprof_mu);
  g_kprof_on.store(on != 0);
  g_kev_used = 0;
  for (auto& v : g_kspans) v.clear();
  return KB2_OK;
}
KB2_API int kb2_kernel_profile_collect(double* total_ms, int64_t* n_spans) {
  if (!total_ms || !n_spans) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_kprof_mu);
  for (int k = 0; k < K_NUM; ++k) {
    double t = 0;
    for (auto& sp : g_kspans[k]) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, g_kev[sp.first], g_kev[sp.second]) == cudaSuccess) t += ms;
    }
    total_ms[k] = t;
    n_spans[k] = (int64_t)g_kspans[k].size();
    g_kspans[k].clear();
  }
  g_kev_used = 0;
  return KB2_OK;
}

KB2_API int kb2_profile_enable(kb2_engine* e, int on) {
  if (!e) return fail(KB2_ERR_VALUE, "null engine");
  e->profiling = on != 0;
  e->ev_used = 0;
  for (auto& v : e->ev_spans) v.clear();
  return KB2_OK;
}

KB2_API int kb2_profile_collect(kb2_engine* e, double* total_ms, int64_t* n_spans) {
  if (!e || !total_ms || !n_spans) return fail(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  CUDA_TRY(cudaDeviceSynchronize());
  for (int k = 0; k < KB2_PROF_NUM; ++k) {
    double t = 0;
    for (auto& sp : e->ev_spans[k]) {
      float ms = 0;
      CUDA_TRY(cudaEventElapsedTime(&ms, e->ev_pool[sp.first], e->ev_pool[sp.second]));
      t += ms;
    }
    total_ms[k] = t;
    n_spans[k] = (int64_t)e->ev_spans[k].size();
    e->ev_spans[k].clear();
  }
  e->ev_used = 0;
  return KB2_OK;
}

}  // extern "C"
