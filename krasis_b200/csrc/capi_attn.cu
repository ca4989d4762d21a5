// C-ABI for the attention blocks that feed the MoE path (include/krasis_b200.h, "attention" section).
// Host state mirrors python/krasis/linear_attention.py:GatedDeltaNetAttention (weights + conv/recurrent state).
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/krasis_b200.h"
#include "moe_common.cuh"

namespace kb2 {
struct GdnDims {
  int H, nk, nv, dk, dv, K;
  float eps, scale;
  int ba_ld;
};
cudaError_t launch_dense_gemm(const void* x, const void* w, void* out, const float* bias, int M, int N, int K,
                              long long ldo, bool out_f32, int num_sms, cudaStream_t s);
cudaError_t launch_gdn_core(const GdnDims& d, const void* qkvz, const void* ba, const void* conv_w, void* conv_state,
                            const float* A_log, const float* dt_bias, const float* norm_w, float* rec_state,
                            void* qn, void* kn, void* vc, float* beta, float* g, float* vcorr, float* kcd, float* intra,
                            float* gcum, float* core, void* normed_out, int M, cudaStream_t s);
cudaError_t launch_dense_gemm_scatter(const void* x, const void* w, void* out, const float* bias, int M, int N, int K, long long ldo,
                                      bool out_f32, int num_sms, cudaStream_t s, const CombineScatter& sc);
cudaError_t launch_combine(const void* c3, const int* slot_of, int M, int H, int top_k, float rsf, int apply_rsf, const void* shared,
                           void* out, cudaStream_t s);
cudaError_t launch_dense_gemm_i8(const void* xq, const float* x_scale, const void* wq, const void* w_scale, void* out,
                                 int M, int N, int K, long long ldo, int num_sms, cudaStream_t s);
cudaError_t launch_rmsnorm(void* x, void* residual, const float* w, void* out, int M, int H, float eps, cudaStream_t s);
cudaError_t launch_quant_rows_int8(const void* x, void* q, float* scale_f32, void* scale_bf16, int rows, int K, cudaStream_t s);
bool rmsnorm_q8_supported(int H);
cudaError_t launch_rmsnorm_q8(void* x, void* residual, const float* w, void* out, void* q, float* q_scale, int M, int H, float eps, cudaStream_t s);
bool silu_mul_quant_supported(int N);
cudaError_t launch_silu_mul_quant(const void* x, void* q, float* scale_f32, int rows, int N, cudaStream_t s);
cudaError_t launch_silu_and_mul(const void* x, void* out, int rows, int N, cudaStream_t s);
cudaError_t launch_sigmoid_gate_mul(const void* h, const void* w, void* y, int M, int H, int N, cudaStream_t s);
cudaError_t launch_add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t s);
struct GqaDims {
  int H, nh, nkv, d, rotary_dim, gated;
  float theta, eps;
};
cudaError_t launch_gqa_core(const GqaDims& g, const void* q_raw, const void* k_raw, const void* v_raw, const float* q_norm,
                            const float* k_norm, const int* positions, const int* kv_indices, void* q_rot, void* k_cache,
                            void* v_cache, void* k_bf, void* v_bf, void* attn_out, int M, int q_start, int kv_len,
                            cudaStream_t s);
struct MlaDims {
  int H, nh, nope, rope, dv, lora;
  float eps;
};
cudaError_t launch_mla_core(const MlaDims& m, void* q_full, const void* kv_a, const float* kv_norm_w, const float* inv_freq,
                            const void* w_kv, const int* positions, const int* kv_indices, void* ckv_cache, void* kpe_cache,
                            void* ckv_bf16, void* kpe_bf16, void* kv_up, void* attn_out, int M, int q_start, int kv_len,
                            float sm_scale, int num_sms, cudaStream_t s);
}  // namespace kb2
using namespace kb2;

extern "C" int kb2_set_error_(int code, const char* msg);   // defined in capi.cu

static int failf(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return kb2_set_error_(code, buf);
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return failf(KB2_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

static int device_sms(int device) {
  static int cached[64] = {0};
  if (device < 0 || device >= 64) return 148;
  if (!cached[device]) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) == cudaSuccess) cached[device] = p.multiProcessorCount;
    else cached[device] = 148;
  }
  return cached[device];
}

struct GdnLayer {
  void *w_qkvz = nullptr, *w_ba = nullptr, *w_out = nullptr, *conv_w = nullptr, *conv_state = nullptr;
  float *A_log = nullptr, *dt_bias = nullptr, *norm_w = nullptr, *rec_state = nullptr;
  bool loaded = false;
};

// the handle's output scatter resolved for a call over M tokens (rows_per_rank = M / n_ranks); disabled: all zero
static CombineScatter scatter_for(const CombineScatter& sc, int M) {
  CombineScatter r = sc;
  if (r.n_ranks > 0) r.rows_per_rank = M / r.n_ranks;
  return r;
}
static int set_scatter(CombineScatter& sc, void* const* peer_recv_host, int num_ranks, int src_rank) {
  sc = CombineScatter{};
  if (num_ranks == 0) return KB2_OK;                                      // back to the plain local output
  if (!peer_recv_host || num_ranks < 1 || num_ranks > kMaxPeers || src_rank < 0 || src_rank >= num_ranks)
    return failf(KB2_ERR_VALUE, "output scatter: need 1 <= num_ranks <= %d and 0 <= src_rank < num_ranks", kMaxPeers);
  for (int r = 0; r < num_ranks; ++r) {
    if (!peer_recv_host[r]) return failf(KB2_ERR_VALUE, "output scatter: null receive buffer for rank %d", r);
    sc.peer_out[r] = (__nv_bfloat16*)peer_recv_host[r];
  }
  sc.n_ranks = num_ranks; sc.src_rank = src_rank;
  return KB2_OK;
}

struct kb2_gdn {
  kb2_gdn_config cfg{};
  GdnDims d{};
  std::vector<GdnLayer> layers;
  void *qkvz = nullptr, *ba = nullptr, *qn = nullptr, *kn = nullptr, *vc = nullptr, *normed = nullptr;
  float *beta = nullptr, *g = nullptr, *vcorr = nullptr, *kcd = nullptr, *intra = nullptr, *gcum = nullptr, *core = nullptr;
  CombineScatter out_scatter{};      // kb2_attn_set_output_scatter: out_proj rows go to the token owners' receive buffers
};

static std::vector<float> bf16_to_f32_host(const void* p, size_t n) {
  std::vector<float> out(n);
  const uint16_t* u = (const uint16_t*)p;
  for (size_t i = 0; i < n; ++i) {
    uint32_t b = (uint32_t)u[i] << 16;
    memcpy(&out[i], &b, 4);
  }
  return out;
}

extern "C" {

KB2_API int kb2_linear_bf16(const void* x_dev, const void* w_dev, const float* bias_dev, void* out_dev, int32_t M,
                            int32_t N, int32_t K, int32_t out_f32, int32_t device, void* stream) {
  if (!x_dev || !w_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (M <= 0 || N <= 0 || K <= 0 || K % 64 || N % 16) return failf(KB2_ERR_VALUE, "linear: need M>0, K %% 64 == 0, N %% 16 == 0 (M=%d N=%d K=%d)", M, N, K);
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_dense_gemm(x_dev, w_dev, out_dev, bias_dev, M, N, K, N, out_f32 != 0, device_sms(device), (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_rmsnorm(void* x_dev, void* residual_dev, const float* weight_dev, void* out_dev, int32_t M, int32_t H,
                        float eps, int32_t device, void* stream) {
  if (!x_dev || !weight_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (M <= 0 || H <= 0 || H % 8) return failf(KB2_ERR_VALUE, "rmsnorm: need M > 0 and H %% 8 == 0");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_rmsnorm(x_dev, residual_dev, weight_dev, out_dev, M, H, eps, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_quantize_rows_int8(const void* x_dev, void* q_dev, float* scale_f32_dev, void* scale_bf16_dev, int32_t rows,
                                   int32_t K, int32_t device, void* stream) {
  if (!x_dev || !q_dev || (!scale_f32_dev && !scale_bf16_dev)) return failf(KB2_ERR_VALUE, "null argument");
  if (rows <= 0 || K <= 0) return failf(KB2_ERR_VALUE, "quantize: rows and K must be positive");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_quant_rows_int8(x_dev, q_dev, scale_f32_dev, scale_bf16_dev, rows, K, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_int8_linear(const void* x_dev, const void* wq_dev, const void* w_scale_bf16_dev, void* out_dev, void* xq_scratch_dev,
                            float* xs_scratch_dev, int32_t M, int32_t N, int32_t K, int32_t device, void* stream) {
  if (!x_dev || !wq_dev || !w_scale_bf16_dev || !out_dev || !xq_scratch_dev || !xs_scratch_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (M <= 0 || K % 128 || N % 16) return failf(KB2_ERR_VALUE, "int8_linear: need M > 0, K %% 128 == 0, N %% 16 == 0 (M=%d N=%d K=%d)", M, N, K);
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(launch_quant_rows_int8(x_dev, xq_scratch_dev, xs_scratch_dev, nullptr, M, K, s));
  CUDA_TRY(launch_dense_gemm_i8(xq_scratch_dev, xs_scratch_dev, wq_dev, w_scale_bf16_dev, out_dev, M, N, K, N, device_sms(device), s));
  return KB2_OK;
}

KB2_API int kb2_rmsnorm_q8(void* x_dev, void* residual_dev, const float* weight_dev, void* out_dev, void* q_dev, float* q_scale_dev,
                           int32_t M, int32_t H, float eps, int32_t device, void* stream) {
  if (!x_dev || !weight_dev || !out_dev || !q_dev || !q_scale_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (M <= 0 || H <= 0 || H % 8) return failf(KB2_ERR_VALUE, "rmsnorm: need M > 0 and H %% 8 == 0");
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (rmsnorm_q8_supported(H)) {
    CUDA_TRY(launch_rmsnorm_q8(x_dev, residual_dev, weight_dev, out_dev, q_dev, q_scale_dev, M, H, eps, s));
  } else {                                   // same results, two passes
    CUDA_TRY(launch_rmsnorm(x_dev, residual_dev, weight_dev, out_dev, M, H, eps, s));
    CUDA_TRY(launch_quant_rows_int8(out_dev, q_dev, q_scale_dev, nullptr, M, H, s));
  }
  return KB2_OK;
}

KB2_API int kb2_int8_linear_q8(const void* xq_dev, const float* xs_dev, const void* wq_dev, const void* w_scale_bf16_dev, void* out_dev,
                               int32_t M, int32_t N, int32_t K, int32_t device, void* stream) {
  if (!xq_dev || !xs_dev || !wq_dev || !w_scale_bf16_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (M <= 0 || K % 128 || N % 16) return failf(KB2_ERR_VALUE, "int8_linear: need M > 0, K %% 128 == 0, N %% 16 == 0 (M=%d N=%d K=%d)", M, N, K);
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_dense_gemm_i8(xq_dev, xs_dev, wq_dev, w_scale_bf16_dev, out_dev, M, N, K, N, device_sms(device), (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_silu_mul_int8_linear(const void* gate_up_dev, const void* wq_dev, const void* w_scale_bf16_dev, void* out_dev,
                                     void* act_scratch_dev, void* xq_scratch_dev, float* xs_scratch_dev, int32_t M, int32_t N, int32_t K,
                                     int32_t device, void* stream) {
  if (!gate_up_dev || !wq_dev || !w_scale_bf16_dev || !out_dev || !xq_scratch_dev || !xs_scratch_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (M <= 0 || K % 128 || N % 16) return failf(KB2_ERR_VALUE, "int8_linear: need M > 0, K %% 128 == 0, N %% 16 == 0 (M=%d N=%d K=%d)", M, N, K);
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  if (silu_mul_quant_supported(K)) {
    CUDA_TRY(launch_silu_mul_quant(gate_up_dev, xq_scratch_dev, xs_scratch_dev, M, K, s));
  } else {                                   // same results through the BF16 activation
    if (!act_scratch_dev) return failf(KB2_ERR_VALUE, "silu_mul_int8_linear: K=%d needs the BF16 activation scratch", K);
    CUDA_TRY(launch_silu_and_mul(gate_up_dev, act_scratch_dev, M, K, s));
    CUDA_TRY(launch_quant_rows_int8(act_scratch_dev, xq_scratch_dev, xs_scratch_dev, nullptr, M, K, s));
  }
  CUDA_TRY(launch_dense_gemm_i8(xq_scratch_dev, xs_scratch_dev, wq_dev, w_scale_bf16_dev, out_dev, M, N, K, N, device_sms(device), s));
  return KB2_OK;
}

KB2_API int kb2_silu_and_mul(const void* x_dev, void* out_dev, int32_t rows, int32_t N, int32_t device, void* stream) {
  if (!x_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (rows <= 0 || N % 8) return failf(KB2_ERR_VALUE, "silu_and_mul: need rows > 0 and N %% 8 == 0");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_silu_and_mul(x_dev, out_dev, rows, N, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_sigmoid_gate_mul(const void* hidden_dev, const void* gate_w_dev, void* y_dev, int32_t M, int32_t H, int32_t N,
                                 int32_t device, void* stream) {
  if (!hidden_dev || !gate_w_dev || !y_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (M <= 0) return failf(KB2_ERR_VALUE, "sigmoid_gate_mul: M must be positive");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_sigmoid_gate_mul(hidden_dev, gate_w_dev, y_dev, M, H, N, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_add_bf16(const void* a_dev, const void* b_dev, void* out_dev, int64_t n, int32_t device, void* stream) {
  if (!a_dev || !b_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (n <= 0 || n % 8) return failf(KB2_ERR_VALUE, "add_bf16: n must be a positive multiple of 8");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_add_bf16(a_dev, b_dev, out_dev, n, (cudaStream_t)stream));
  return KB2_OK;
}

// out[m] = bf16(sum_r f32(slots[m][r])): the consumer side of a reduce-scatter fused into a producer kernel (slots = this rank's
// receive buffer [rows][num_ranks][H] bf16, filled by every rank's out_proj GEMM / combine kernel; call after kb2_comm_barrier)
KB2_API int kb2_sum_slots_bf16(const void* slots_dev, int32_t num_ranks, void* out_dev, int32_t rows, int32_t H, int32_t device, void* stream) {
  if (!slots_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (rows <= 0 || H <= 0 || H % 8 || num_ranks < 1 || num_ranks > kMaxPeers) return failf(KB2_ERR_VALUE, "sum_slots: bad rows / H / num_ranks");
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(launch_combine(slots_dev, nullptr, rows, H, num_ranks, 1.0f, 0, nullptr, out_dev, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_gdn_create(const kb2_gdn_config* c, kb2_gdn** out) {
  if (!c || !out) return failf(KB2_ERR_VALUE, "null argument");
  if (c->num_k_heads < 1 || c->num_v_heads % c->num_k_heads) return failf(KB2_ERR_VALUE, "num_v_heads must be a multiple of num_k_heads");
  if (c->k_head_dim > 128 || c->v_head_dim > 128 || c->v_head_dim % 32 || c->k_head_dim % 8) return failf(KB2_ERR_VALUE, "head dims must be <= 128 (v_head_dim %% 32 == 0)");
  if (c->conv_kernel < 1 || c->conv_kernel > 8) return failf(KB2_ERR_VALUE, "conv_kernel must be in [1,8]");
  if (c->hidden_size % 64) return failf(KB2_ERR_VALUE, "hidden_size must be a multiple of 64");
  const int kd = c->num_k_heads * c->k_head_dim, vd = c->num_v_heads * c->v_head_dim;
  if ((2 * kd + 2 * vd) % 16 || vd % 64) return failf(KB2_ERR_VALUE, "qkvz width (%d) must be a multiple of 16 and value_dim (%d) of 64", 2 * kd + 2 * vd, vd);
  if (c->max_tokens < 1 || c->num_layers < 1) return failf(KB2_ERR_VALUE, "max_tokens and num_layers must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return failf(KB2_ERR_CUDA, "no CUDA device: krasis_b200 has no CPU fallback");
  }
  CUDA_TRY(cudaSetDevice(c->device));
  kb2_gdn* h = new kb2_gdn();
  h->cfg = *c;
  h->d = GdnDims{c->hidden_size, c->num_k_heads, c->num_v_heads, c->k_head_dim, c->v_head_dim, c->conv_kernel,
                 c->rms_norm_eps, 1.0f / sqrtf((float)c->k_head_dim), (2 * c->num_v_heads + 15) & ~15};
  h->layers.resize(c->num_layers);
  const size_t M = c->max_tokens, nch = (M + 63) / 64, nv = c->num_v_heads;
#define ALLOC(ptr, bytes) CUDA_TRY(cudaMalloc((void**)&(ptr), (bytes)))
  ALLOC(h->qkvz, M * (2 * kd + 2 * vd) * 2);
  ALLOC(h->ba, M * (size_t)h->d.ba_ld * 2);
  ALLOC(h->qn, M * kd * 2);
  ALLOC(h->kn, M * kd * 2);
  ALLOC(h->vc, M * vd * 2);
  ALLOC(h->normed, M * vd * 2);
  ALLOC(h->beta, M * nv * 4);
  ALLOC(h->g, M * nv * 4);
  ALLOC(h->vcorr, nv * nch * 64 * c->v_head_dim * 4 * 9 / 8);   // tcgen05 scan: rows padded 32 -> 36 floats per dv slice
  ALLOC(h->kcd, nv * nch * 64 * c->k_head_dim * 4);
  ALLOC(h->intra, nv * nch * 64 * 64 * 4);
  ALLOC(h->gcum, nv * nch * 64 * 4);
  ALLOC(h->core, M * vd * 4);
#undef ALLOC
  *out = h;
  return KB2_OK;
}

KB2_API int kb2_gdn_set_output_scatter(kb2_gdn* h, void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  return set_scatter(h->out_scatter, peer_recv_host, num_ranks, src_rank);
}

KB2_API void kb2_gdn_destroy(kb2_gdn* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (auto& L : h->layers) {
    cudaFree(L.w_qkvz); cudaFree(L.w_ba); cudaFree(L.w_out); cudaFree(L.conv_w); cudaFree(L.conv_state);
    cudaFree(L.A_log); cudaFree(L.dt_bias); cudaFree(L.norm_w); cudaFree(L.rec_state);
  }
  cudaFree(h->qkvz); cudaFree(h->ba); cudaFree(h->qn); cudaFree(h->kn); cudaFree(h->vc); cudaFree(h->normed);
  cudaFree(h->beta); cudaFree(h->g); cudaFree(h->vcorr); cudaFree(h->kcd); cudaFree(h->intra); cudaFree(h->gcum);
  cudaFree(h->core);
  delete h;
}

static int gdn_check(kb2_gdn* h, int layer) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  if (layer < 0 || layer >= (int)h->layers.size()) return failf(KB2_ERR_VALUE, "layer %d out of range [0, %d)", layer, (int)h->layers.size());
  return KB2_OK;
}

KB2_API int kb2_gdn_set_weights_host(kb2_gdn* h, int layer, const void* in_proj_qkvz, const void* in_proj_ba,
                                     const void* conv1d_weight, const void* A_log, const void* dt_bias,
                                     const void* norm_weight, const void* out_proj) {
  if (int r = gdn_check(h, layer)) return r;
  if (!in_proj_qkvz || !in_proj_ba || !conv1d_weight || !A_log || !dt_bias || !norm_weight || !out_proj) return failf(KB2_ERR_VALUE, "null weight pointer");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const auto& c = h->cfg;
  const size_t kd = c.num_k_heads * c.k_head_dim, vd = c.num_v_heads * c.v_head_dim, C = 2 * kd + vd, H = c.hidden_size;
  GdnLayer& L = h->layers[layer];
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    if (!*dst) { cudaError_t e = cudaMalloc(dst, bytes); if (e != cudaSuccess) return e; }
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  CUDA_TRY(up(&L.w_qkvz, in_proj_qkvz, (2 * kd + 2 * vd) * H * 2));
  if (!L.w_ba) CUDA_TRY(cudaMalloc(&L.w_ba, (size_t)h->d.ba_ld * H * 2));
  CUDA_TRY(cudaMemset(L.w_ba, 0, (size_t)h->d.ba_ld * H * 2));            // rows padded to a multiple of 16 with zeros
  CUDA_TRY(cudaMemcpy(L.w_ba, in_proj_ba, 2 * c.num_v_heads * H * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(up(&L.w_out, out_proj, H * vd * 2));
  CUDA_TRY(up(&L.conv_w, conv1d_weight, C * c.conv_kernel * 2));
  auto a = bf16_to_f32_host(A_log, c.num_v_heads), b = bf16_to_f32_host(dt_bias, c.num_v_heads), n = bf16_to_f32_host(norm_weight, c.v_head_dim);
  CUDA_TRY(up((void**)&L.A_log, a.data(), a.size() * 4));
  CUDA_TRY(up((void**)&L.dt_bias, b.data(), b.size() * 4));
  CUDA_TRY(up((void**)&L.norm_w, n.data(), n.size() * 4));
  if (!L.conv_state) CUDA_TRY(cudaMalloc(&L.conv_state, C * c.conv_kernel * 2));
  if (!L.rec_state) CUDA_TRY(cudaMalloc((void**)&L.rec_state, (size_t)c.num_v_heads * c.k_head_dim * c.v_head_dim * 4));
  CUDA_TRY(cudaMemset(L.conv_state, 0, C * c.conv_kernel * 2));
  CUDA_TRY(cudaMemset(L.rec_state, 0, (size_t)c.num_v_heads * c.k_head_dim * c.v_head_dim * 4));
  L.loaded = true;
  return KB2_OK;
}

KB2_API int kb2_gdn_reset_state(kb2_gdn* h, int layer, void* stream) {
  if (int r = gdn_check(h, layer)) return r;
  GdnLayer& L = h->layers[layer];
  if (!L.loaded) return failf(KB2_ERR_STATE, "GDN weights not set for layer %d", layer);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const auto& c = h->cfg;
  const size_t C = 2 * c.num_k_heads * c.k_head_dim + c.num_v_heads * c.v_head_dim;
  CUDA_TRY(cudaMemsetAsync(L.conv_state, 0, C * c.conv_kernel * 2, (cudaStream_t)stream));
  CUDA_TRY(cudaMemsetAsync(L.rec_state, 0, (size_t)c.num_v_heads * c.k_head_dim * c.v_head_dim * 4, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_gdn_get_state_host(kb2_gdn* h, int layer, void* conv_state_bf16_host, float* recurrent_state_host) {
  if (int r = gdn_check(h, layer)) return r;
  GdnLayer& L = h->layers[layer];
  if (!L.loaded) return failf(KB2_ERR_STATE, "GDN weights not set for layer %d", layer);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const auto& c = h->cfg;
  const size_t C = 2 * c.num_k_heads * c.k_head_dim + c.num_v_heads * c.v_head_dim;
  CUDA_TRY(cudaDeviceSynchronize());
  if (conv_state_bf16_host) CUDA_TRY(cudaMemcpy(conv_state_bf16_host, L.conv_state, C * c.conv_kernel * 2, cudaMemcpyDeviceToHost));
  if (recurrent_state_host) CUDA_TRY(cudaMemcpy(recurrent_state_host, L.rec_state, (size_t)c.num_v_heads * c.k_head_dim * c.v_head_dim * 4, cudaMemcpyDeviceToHost));
  return KB2_OK;
}

KB2_API int kb2_gdn_forward(kb2_gdn* h, int layer, const void* hidden_dev, void* out_dev, int32_t M, void* stream) {
  if (int r = gdn_check(h, layer)) return r;
  GdnLayer& L = h->layers[layer];
  if (M < 0 || M > h->cfg.max_tokens) return failf(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, h->cfg.max_tokens);
  if (!L.loaded) return failf(KB2_ERR_STATE, "GDN weights not set for layer %d", layer);
  if (M == 0) return KB2_OK;
  if (!hidden_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const auto& c = h->cfg;
  const int kd = c.num_k_heads * c.k_head_dim, vd = c.num_v_heads * c.v_head_dim, H = c.hidden_size;
  const int sms = device_sms(c.device);
  CUDA_TRY(launch_dense_gemm(hidden_dev, L.w_qkvz, h->qkvz, nullptr, M, 2 * kd + 2 * vd, H, 2 * kd + 2 * vd, false, sms, s));
  CUDA_TRY(launch_dense_gemm(hidden_dev, L.w_ba, h->ba, nullptr, M, h->d.ba_ld, H, h->d.ba_ld, false, sms, s));
  CUDA_TRY(launch_gdn_core(h->d, h->qkvz, h->ba, L.conv_w, L.conv_state, L.A_log, L.dt_bias, L.norm_w, L.rec_state,
                           h->qn, h->kn, h->vc, h->beta, h->g, h->vcorr, h->kcd, h->intra, h->gcum, h->core, h->normed, M, s));
  CUDA_TRY(launch_dense_gemm_scatter(h->normed, L.w_out, out_dev, nullptr, M, H, vd, H, false, sms, s, scatter_for(h->out_scatter, M)));
  return KB2_OK;
}

// ------------------------------------------------------------------------------------------------ GQA
struct GqaLayer {
  void *wq = nullptr, *wk = nullptr, *wv = nullptr, *wo = nullptr;
  float *q_norm = nullptr, *k_norm = nullptr;
  bool loaded = false;
};
struct kb2_gqa {
  kb2_gqa_config cfg{};
  GqaDims g{};
  std::vector<GqaLayer> layers;
  void *q_raw = nullptr, *k_raw = nullptr, *v_raw = nullptr, *q_rot = nullptr, *attn = nullptr;
  void *k_bf = nullptr, *v_bf = nullptr;     // dense BF16 upcast of the sequence's FP8 pages, [kv_cap][nkv*d] each
  long long kv_cap = 0;
  CombineScatter out_scatter{};
};

KB2_API int kb2_gqa_create(const kb2_gqa_config* c, kb2_gqa** out) {
  if (!c || !out) return failf(KB2_ERR_VALUE, "null argument");
  if (c->head_dim != 128 && c->head_dim != 256) return failf(KB2_ERR_VALUE, "head_dim must be 128 or 256 (got %d)", c->head_dim);
  if (c->num_kv_heads < 1 || c->num_heads % c->num_kv_heads) return failf(KB2_ERR_VALUE, "num_heads must be a multiple of num_kv_heads");
  if (c->rotary_dim < 0 || c->rotary_dim > c->head_dim || c->rotary_dim % 2) return failf(KB2_ERR_VALUE, "bad rotary_dim %d", c->rotary_dim);
  if (c->hidden_size % 64 || c->page_size != 16) return failf(KB2_ERR_VALUE, "hidden_size %% 64 == 0 and page_size == 16 required");
  if (c->max_tokens < 1 || c->num_layers < 1) return failf(KB2_ERR_VALUE, "max_tokens and num_layers must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return failf(KB2_ERR_CUDA, "no CUDA device: krasis_b200 has no CPU fallback");
  }
  CUDA_TRY(cudaSetDevice(c->device));
  kb2_gqa* h = new kb2_gqa();
  h->cfg = *c;
  h->g = GqaDims{c->hidden_size, c->num_heads, c->num_kv_heads, c->head_dim, c->rotary_dim, c->gated ? 1 : 0, c->rope_theta, c->rms_norm_eps};
  h->layers.resize(c->num_layers);
  const size_t M = c->max_tokens, qd = (size_t)c->num_heads * c->head_dim, kvd = (size_t)c->num_kv_heads * c->head_dim;
  CUDA_TRY(cudaMalloc(&h->q_raw, M * qd * (c->gated ? 2 : 1) * 2));
  CUDA_TRY(cudaMalloc(&h->k_raw, M * kvd * 2));
  CUDA_TRY(cudaMalloc(&h->v_raw, M * kvd * 2));
  CUDA_TRY(cudaMalloc(&h->q_rot, M * qd * 2));
  CUDA_TRY(cudaMalloc(&h->attn, M * qd * 2));
  *out = h;
  return KB2_OK;
}

KB2_API int kb2_gqa_set_output_scatter(kb2_gqa* h, void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  return set_scatter(h->out_scatter, peer_recv_host, num_ranks, src_rank);
}

KB2_API void kb2_gqa_destroy(kb2_gqa* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (auto& L : h->layers) { cudaFree(L.wq); cudaFree(L.wk); cudaFree(L.wv); cudaFree(L.wo); cudaFree(L.q_norm); cudaFree(L.k_norm); }
  cudaFree(h->q_raw); cudaFree(h->k_raw); cudaFree(h->v_raw); cudaFree(h->q_rot); cudaFree(h->attn);
  cudaFree(h->k_bf); cudaFree(h->v_bf);
  delete h;
}

KB2_API int kb2_gqa_set_weights_host(kb2_gqa* h, int layer, const void* q_proj, const void* k_proj, const void* v_proj,
                                     const void* o_proj, const void* q_norm, const void* k_norm) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  if (layer < 0 || layer >= (int)h->layers.size()) return failf(KB2_ERR_VALUE, "layer %d out of range", layer);
  if (!q_proj || !k_proj || !v_proj || !o_proj) return failf(KB2_ERR_VALUE, "null weight pointer");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const auto& c = h->cfg;
  const size_t qd = (size_t)c.num_heads * c.head_dim, kvd = (size_t)c.num_kv_heads * c.head_dim, H = c.hidden_size;
  GqaLayer& L = h->layers[layer];
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    if (!*dst) { cudaError_t e = cudaMalloc(dst, bytes); if (e != cudaSuccess) return e; }
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  CUDA_TRY(up(&L.wq, q_proj, qd * (c.gated ? 2 : 1) * H * 2));
  CUDA_TRY(up(&L.wk, k_proj, kvd * H * 2));
  CUDA_TRY(up(&L.wv, v_proj, kvd * H * 2));
  CUDA_TRY(up(&L.wo, o_proj, H * qd * 2));
  if (q_norm) { auto v = bf16_to_f32_host(q_norm, c.head_dim); CUDA_TRY(up((void**)&L.q_norm, v.data(), v.size() * 4)); }
  if (k_norm) { auto v = bf16_to_f32_host(k_norm, c.head_dim); CUDA_TRY(up((void**)&L.k_norm, v.data(), v.size() * 4)); }
  L.loaded = true;
  return KB2_OK;
}

KB2_API int kb2_gqa_forward(kb2_gqa* h, int layer, const void* hidden_dev, const int32_t* positions_dev, int32_t first_position,
                            void* k_cache_layer_dev, void* v_cache_layer_dev, const int32_t* kv_indices_dev,
                            int32_t kv_len_after, void* out_dev, int32_t M, void* stream) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  if (layer < 0 || layer >= (int)h->layers.size()) return failf(KB2_ERR_VALUE, "layer %d out of range", layer);
  GqaLayer& L = h->layers[layer];
  if (M < 0 || M > h->cfg.max_tokens) return failf(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, h->cfg.max_tokens);
  if (!L.loaded) return failf(KB2_ERR_STATE, "GQA weights not set for layer %d", layer);
  if (M == 0) return KB2_OK;
  if (!hidden_dev || !positions_dev || !k_cache_layer_dev || !v_cache_layer_dev || !kv_indices_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (first_position < 0 || kv_len_after != first_position + M) return failf(KB2_ERR_VALUE, "positions must be contiguous: kv_len_after (%d) != first_position (%d) + M (%d)", kv_len_after, first_position, M);
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const auto& c = h->cfg;
  const int qd = c.num_heads * c.head_dim, kvd = c.num_kv_heads * c.head_dim, H = c.hidden_size, sms = device_sms(c.device);
  const int qw = qd * (c.gated ? 2 : 1);
  CUDA_TRY(launch_dense_gemm(hidden_dev, L.wq, h->q_raw, nullptr, M, qw, H, qw, false, sms, s));
  CUDA_TRY(launch_dense_gemm(hidden_dev, L.wk, h->k_raw, nullptr, M, kvd, H, kvd, false, sms, s));
  CUDA_TRY(launch_dense_gemm(hidden_dev, L.wv, h->v_raw, nullptr, M, kvd, H, kvd, false, sms, s));
  if (kv_len_after > h->kv_cap) {              // grow the upcast scratch (cudaFree synchronises: no kernel still reads it)
    const long long cap = kv_len_after > 2 * h->kv_cap ? kv_len_after : 2 * h->kv_cap;
    cudaFree(h->k_bf); cudaFree(h->v_bf);
    h->k_bf = h->v_bf = nullptr;
    h->kv_cap = 0;
    CUDA_TRY(cudaMalloc(&h->k_bf, (size_t)cap * kvd * 2));
    CUDA_TRY(cudaMalloc(&h->v_bf, (size_t)cap * kvd * 2));
    h->kv_cap = cap;
  }
  CUDA_TRY(launch_gqa_core(h->g, h->q_raw, h->k_raw, h->v_raw, L.q_norm, L.k_norm, positions_dev, kv_indices_dev, h->q_rot,
                           k_cache_layer_dev, v_cache_layer_dev, h->k_bf, h->v_bf, h->attn, M, first_position, kv_len_after, s));
  CUDA_TRY(launch_dense_gemm_scatter(h->attn, L.wo, out_dev, nullptr, M, H, qd, H, false, sms, s, scatter_for(h->out_scatter, M)));
  return KB2_OK;
}


// ---------------------------------------------------------------------------------------------- MLA
struct MlaLayer {
  void *wq = nullptr, *wqa = nullptr, *wkva = nullptr, *wkv = nullptr, *wo = nullptr;
  float *qa_norm = nullptr, *kv_norm = nullptr, *inv_freq = nullptr;
  bool loaded = false;
};
struct kb2_mla {
  kb2_mla_config cfg{};
  MlaDims m{};
  std::vector<MlaLayer> layers;
  void *q_full = nullptr, *q_a = nullptr, *kv_a = nullptr, *ckv_bf16 = nullptr, *kpe_bf16 = nullptr, *kv_up = nullptr, *attn = nullptr;
  CombineScatter out_scatter{};
};

KB2_API int kb2_mla_create(const kb2_mla_config* c, kb2_mla** out) {
  if (!c || !out) return failf(KB2_ERR_VALUE, "null argument");
  if (c->qk_nope_head_dim != 128 || c->qk_rope_head_dim != 64 || c->v_head_dim != 128)
    return failf(KB2_ERR_VALUE, "MLA head geometry must be nope 128 / rope 64 / v 128 (got %d/%d/%d)", c->qk_nope_head_dim,
                 c->qk_rope_head_dim, c->v_head_dim);
  if (c->kv_lora_rank < 64 || c->kv_lora_rank % 64) return failf(KB2_ERR_VALUE, "kv_lora_rank %% 64 == 0 required");
  if (c->q_lora_rank < 0 || c->q_lora_rank % 64) return failf(KB2_ERR_VALUE, "q_lora_rank must be 0 or a multiple of 64");
  if (c->num_heads < 1 || c->hidden_size % 64 || c->page_size != 16) return failf(KB2_ERR_VALUE, "hidden_size %% 64 == 0 and page_size == 16 required");
  if (c->max_tokens < 1 || c->num_layers < 1 || c->max_kv_len < c->max_tokens) return failf(KB2_ERR_VALUE, "max_tokens, num_layers >= 1 and max_kv_len >= max_tokens required");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return failf(KB2_ERR_CUDA, "no CUDA device: krasis_b200 has no CPU fallback");
  }
  CUDA_TRY(cudaSetDevice(c->device));
  kb2_mla* h = new kb2_mla();
  h->cfg = *c;
  h->m = MlaDims{c->hidden_size, c->num_heads, c->qk_nope_head_dim, c->qk_rope_head_dim, c->v_head_dim, c->kv_lora_rank, c->rms_norm_eps};
  h->layers.resize(c->num_layers);
  const size_t M = c->max_tokens, T = c->max_kv_len, nh = c->num_heads;
  CUDA_TRY(cudaMalloc(&h->q_full, M * nh * 192 * 2));
  if (c->q_lora_rank) CUDA_TRY(cudaMalloc(&h->q_a, M * c->q_lora_rank * 2 * 2));     // raw + normed
  CUDA_TRY(cudaMalloc(&h->kv_a, M * (c->kv_lora_rank + 64) * 2));
  CUDA_TRY(cudaMalloc(&h->ckv_bf16, T * c->kv_lora_rank * 2));
  CUDA_TRY(cudaMalloc(&h->kpe_bf16, T * 64 * 2));
  CUDA_TRY(cudaMalloc(&h->kv_up, T * nh * 256 * 2));
  CUDA_TRY(cudaMalloc(&h->attn, M * nh * 128 * 2));
  *out = h;
  return KB2_OK;
}

KB2_API int kb2_mla_set_output_scatter(kb2_mla* h, void* const* peer_recv_host, int32_t num_ranks, int32_t src_rank) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  return set_scatter(h->out_scatter, peer_recv_host, num_ranks, src_rank);
}

KB2_API void kb2_mla_destroy(kb2_mla* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  for (auto& L : h->layers) {
    cudaFree(L.wq); cudaFree(L.wqa); cudaFree(L.wkva); cudaFree(L.wkv); cudaFree(L.wo);
    cudaFree(L.qa_norm); cudaFree(L.kv_norm); cudaFree(L.inv_freq);
  }
  cudaFree(h->q_full); cudaFree(h->q_a); cudaFree(h->kv_a); cudaFree(h->ckv_bf16); cudaFree(h->kpe_bf16); cudaFree(h->kv_up); cudaFree(h->attn);
  delete h;
}

KB2_API int kb2_mla_set_weights_host(kb2_mla* h, int layer, const void* q_proj_or_q_b_proj, const void* q_a_proj,
                                     const void* q_a_layernorm, const void* kv_a_proj_with_mqa, const void* kv_a_layernorm,
                                     const void* w_kc, const void* w_vc, const void* o_proj, const float* rope_inv_freq) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  if (layer < 0 || layer >= (int)h->layers.size()) return failf(KB2_ERR_VALUE, "layer %d out of range", layer);
  const auto& c = h->cfg;
  if (!q_proj_or_q_b_proj || !kv_a_proj_with_mqa || !kv_a_layernorm || !w_kc || !w_vc || !o_proj || !rope_inv_freq)
    return failf(KB2_ERR_VALUE, "null weight pointer");
  if ((c.q_lora_rank > 0) != (q_a_proj != nullptr && q_a_layernorm != nullptr))
    return failf(KB2_ERR_VALUE, "q_a_proj / q_a_layernorm must be given exactly when q_lora_rank > 0");
  CUDA_TRY(cudaSetDevice(c.device));
  MlaLayer& L = h->layers[layer];
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    if (!*dst) { cudaError_t e = cudaMalloc(dst, bytes); if (e != cudaSuccess) return e; }
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  const size_t H = c.hidden_size, nh = c.num_heads, lora = c.kv_lora_rank, qin = c.q_lora_rank ? c.q_lora_rank : H;
  CUDA_TRY(up(&L.wq, q_proj_or_q_b_proj, nh * 192 * qin * 2));
  if (c.q_lora_rank) {
    CUDA_TRY(up(&L.wqa, q_a_proj, (size_t)c.q_lora_rank * H * 2));
    auto v = bf16_to_f32_host(q_a_layernorm, c.q_lora_rank);
    CUDA_TRY(up((void**)&L.qa_norm, v.data(), v.size() * 4));
  }
  CUDA_TRY(up(&L.wkva, kv_a_proj_with_mqa, (lora + 64) * H * 2));
  {
    auto v = bf16_to_f32_host(kv_a_layernorm, (int)lora);
    CUDA_TRY(up((void**)&L.kv_norm, v.data(), v.size() * 4));
  }
  // [w_kc ; w_vc] as one [nh*256][lora] projection: rows [0, nh*128) = k_nope heads, rows [nh*128, nh*256) = v heads
  if (!L.wkv) CUDA_TRY(cudaMalloc(&L.wkv, nh * 256 * lora * 2));
  CUDA_TRY(cudaMemcpy(L.wkv, w_kc, nh * 128 * lora * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy((char*)L.wkv + nh * 128 * lora * 2, w_vc, nh * 128 * lora * 2, cudaMemcpyHostToDevice));
  CUDA_TRY(up(&L.wo, o_proj, H * nh * 128 * 2));
  CUDA_TRY(up((void**)&L.inv_freq, rope_inv_freq, 32 * 4));
  L.loaded = true;
  return KB2_OK;
}

KB2_API int kb2_mla_forward(kb2_mla* h, int layer, const void* hidden_dev, const int32_t* positions_dev, int32_t first_position,
                            void* ckv_cache_layer_dev, void* kpe_cache_layer_dev, const int32_t* kv_indices_dev,
                            int32_t kv_len_after, void* out_dev, int32_t M, void* stream) {
  if (!h) return failf(KB2_ERR_VALUE, "null handle");
  if (layer < 0 || layer >= (int)h->layers.size()) return failf(KB2_ERR_VALUE, "layer %d out of range", layer);
  MlaLayer& L = h->layers[layer];
  const auto& c = h->cfg;
  if (M < 0 || M > c.max_tokens) return failf(KB2_ERR_VALUE, "num_tokens %d outside [0, max_tokens=%d]", M, c.max_tokens);
  if (!L.loaded) return failf(KB2_ERR_STATE, "MLA weights not set for layer %d", layer);
  if (M == 0) return KB2_OK;
  if (!hidden_dev || !positions_dev || !ckv_cache_layer_dev || !kpe_cache_layer_dev || !kv_indices_dev || !out_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (first_position < 0 || kv_len_after != first_position + M) return failf(KB2_ERR_VALUE, "positions must be contiguous: kv_len_after (%d) != first_position (%d) + M (%d)", kv_len_after, first_position, M);
  if (kv_len_after > c.max_kv_len) return failf(KB2_ERR_VALUE, "sequence length %d exceeds max_kv_len %d", kv_len_after, c.max_kv_len);
  CUDA_TRY(cudaSetDevice(c.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int H = c.hidden_size, nh = c.num_heads, lora = c.kv_lora_rank, sms = device_sms(c.device);
  CUDA_TRY(launch_dense_gemm(hidden_dev, L.wkva, h->kv_a, nullptr, M, lora + 64, H, lora + 64, false, sms, s));
  if (c.q_lora_rank) {
    void* q_a_normed = (char*)h->q_a + (size_t)c.max_tokens * c.q_lora_rank * 2;
    CUDA_TRY(launch_dense_gemm(hidden_dev, L.wqa, h->q_a, nullptr, M, c.q_lora_rank, H, c.q_lora_rank, false, sms, s));
    CUDA_TRY(launch_rmsnorm(h->q_a, nullptr, L.qa_norm, q_a_normed, M, c.q_lora_rank, c.rms_norm_eps, s));
    CUDA_TRY(launch_dense_gemm(q_a_normed, L.wq, h->q_full, nullptr, M, nh * 192, c.q_lora_rank, nh * 192, false, sms, s));
  } else {
    CUDA_TRY(launch_dense_gemm(hidden_dev, L.wq, h->q_full, nullptr, M, nh * 192, H, nh * 192, false, sms, s));
  }
  CUDA_TRY(launch_mla_core(h->m, h->q_full, h->kv_a, L.kv_norm, L.inv_freq, L.wkv, positions_dev, kv_indices_dev,
                           ckv_cache_layer_dev, kpe_cache_layer_dev, h->ckv_bf16, h->kpe_bf16, h->kv_up, h->attn, M,
                           first_position, kv_len_after, c.sm_scale, sms, s));
  CUDA_TRY(launch_dense_gemm_scatter(h->attn, L.wo, out_dev, nullptr, M, H, nh * 128, H, false, sms, s, scatter_for(h->out_scatter, M)));
  return KB2_OK;
}

}  // extern "C"
