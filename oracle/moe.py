"""MoE expert-MLP oracle in the two numerics the reference has (test infrastructure only).

GPU-path numerics  (what GpuPrefillManager.forward produces; python/krasis/gpu_prefill.py)
------------------------------------------------------------------------------------------
The reference calls third-party sglang `fused_marlin_moe` (gpu_prefill.py:566-656); its
buffers and call sequence are spelled out in the in-tree GPT-OSS twin
gpu_prefill.py:64-239, which fixes the rounding points this oracle reproduces:
  W_deq   = bf16((nib-8) * scale)            Marlin dequantises to BF16 and multiplies by the
                                             BF16 group scale in BF16 before the BF16 MMA
  C1      = bf16( X[sorted] @ W13_e^T )      intermediate_cache1 is BF16 (gpu_prefill.py:137-144), fp32 accumulate
  A       = bf16( silu(C1[:, :I]) * C1[:, I:] )    silu_and_mul, fp32 inside (gpu_prefill.py:46-61 is the GPT-OSS variant)
  C3      = bf16( w_topk * (A @ W2_e^T) )    mul_topk_weights=True only on GEMM #2 (gpu_prefill.py:199-226)
  routed  = bf16( sum_k C3[m, k, :] )        moe_sum_reduce with factor 1.0 (gpu_prefill.py:238, rsf not forwarded at :4158-4175)
  out     = bf16(rsf * routed) + shared      gpu_prefill.py:4471-4482 (BF16 tensor ops)
EP slicing (gpu_prefill.py:353-359, 4140-4149): non-local ids contribute exactly 0.
Parity with the Marlin kernel itself is UNPINNED (no golden vectors in the reference
tree) — this is the documented math, accumulated in float64.

CPU-path numerics  (what the timed reference CPU baseline computes)
-------------------------------------------------------------------
  moe_forward_unified      src/moe.rs:572-715     INT4/INT8 g128, INT16 activations
  expert_forward_unified   src/moe.rs:184-380
  quantize_activation_int16 src/kernel/avx2.rs:234-268
  transposed INT4 kernel    src/kernel/avx2.rs:1066-1206 (per group: fma(float(isum), ws*as, out))
  moe_forward_gguf / expert_forward_gguf  src/moe.rs:990-1110, src/gguf_kernels.rs:690-756
  Q4_K / Q8_0 INT16 kernels src/gguf_kernels.rs:271-426
"""
import numpy as np

from .bf16 import bf16_bits_to_f32, round_bf16
from . import quant, gguf_blocks as G


# ------------------------------------------------------------------ weights containers

class Int4Layer:
    """One MoE layer of Krasis INT4/INT8 g128 experts in the row-major [N, K] form the
    quantiser emits: w13 = [gate rows ; up rows] (src/weights/mod.rs:346-349)."""

    def __init__(self, w13_q, w13_s, w2_q, w2_s, bits=4, group_size=128):
        self.w13_q, self.w13_s, self.w2_q, self.w2_s = w13_q, w13_s, w2_q, w2_s
        self.bits, self.group_size = bits, group_size
        self.E = w13_q.shape[0]
        self.I = w13_q.shape[1] // 2
        self.H = w2_q.shape[1]

    def _deq(self, q, s):
        if self.bits == 4:
            return quant.dequantize_int4(q, s, self.group_size)
        return quant.dequantize_int8(q, s, self.group_size)

    def w13_f32(self, e):
        return self._deq(self.w13_q[e], self.w13_s[e])

    def w2_f32(self, e):
        return self._deq(self.w2_q[e], self.w2_s[e])

    def w13_ints(self, e):
        return quant.unpack_int4(self.w13_q[e]) if self.bits == 4 else self.w13_q[e]

    def w2_ints(self, e):
        return quant.unpack_int4(self.w2_q[e]) if self.bits == 4 else self.w2_q[e]


class GgufLayer:
    """One MoE layer of native GGUF blocks: gate/up [E, I, row_bytes(H)], down [E, H, row_bytes(I)]."""

    def __init__(self, gate, up, down, gate_up_type, down_type, H, I):
        self.gate, self.up, self.down = gate, up, down
        self.gate_up_type, self.down_type = gate_up_type, down_type
        self.E, self.H, self.I = gate.shape[0], H, I

    def w13_f32(self, e):
        g = G.dequantize(self.gate_up_type, self.gate[e].reshape(-1), self.I * self.H).reshape(self.I, self.H)
        u = G.dequantize(self.gate_up_type, self.up[e].reshape(-1), self.I * self.H).reshape(self.I, self.H)
        return np.concatenate([g, u], axis=0)

    def w2_f32(self, e):
        return G.dequantize(self.down_type, self.down[e].reshape(-1), self.H * self.I).reshape(self.H, self.I)


def make_int_layer(rng, E, H, I, bits=4, std=0.02, group_size=128):
    """BF16 N(0, std^2) weights pushed through the real quantiser."""
    from .bf16 import f32_to_bf16_bits
    qf = quant.quantize_int4 if bits == 4 else quant.quantize_int8
    w13q, w13s, w2q, w2s = [], [], [], []
    for _ in range(E):
        a = f32_to_bf16_bits(rng.normal(0, std, (2 * I, H)).astype(np.float32))
        b = f32_to_bf16_bits(rng.normal(0, std, (H, I)).astype(np.float32))
        q, s = qf(a, group_size); w13q.append(q); w13s.append(s)
        q, s = qf(b, group_size); w2q.append(q); w2s.append(s)
    return Int4Layer(np.stack(w13q), np.stack(w13s), np.stack(w2q), np.stack(w2s), bits, group_size)


def make_gguf_layer(rng, E, H, I, gate_up_type=G.GGML_Q4_K, down_type=G.GGML_Q8_0):
    def mk(t, rows, k):
        if t == G.GGML_Q8_0:   # realistic: quantise N(0, 0.02) floats
            return np.stack([G.quantize_q8_0(rng.normal(0, 0.02, (rows, k)).astype(np.float32)) for _ in range(E)])
        if t == G.GGML_Q4_K:   # d, dmin sized so |w| ~ 0.1 and gate/up pre-activations are O(1), like a trained model
            return np.stack([G.random_q4_k(rng, rows, k, d_range=(1e-4, 6e-4)) for _ in range(E)])
        return np.stack([G.random_blocks(rng, t, rows, k) for _ in range(E)])
    return GgufLayer(mk(gate_up_type, I, H), mk(gate_up_type, I, H), mk(down_type, H, I),
                     gate_up_type, down_type, H, I)


# ------------------------------------------------------------------ GPU-path numerics

def silu_f32(x):
    x = np.asarray(x, np.float32)
    return (x / (np.float32(1.0) + np.exp(-x).astype(np.float32))).astype(np.float32)


def moe_forward_gpu_path(layer, x_bf16_f32, topk_ids, topk_w, expert_start=0, expert_end=None,
                         return_intermediates=False):
    """Routed-experts output in GPU-path numerics: returns float32 holding BF16 values, [M, H].

    x_bf16_f32: [M, H] float32 with BF16-representable values.  ids < 0 or outside
    [expert_start, expert_end) contribute zero (EP masking; also submit_forward's id -1,
    src/moe.rs:2722)."""
    M, H = x_bf16_f32.shape
    k = topk_ids.shape[1]
    E = layer.E
    expert_end = E if expert_end is None else expert_end
    c3 = np.zeros((M, k, H), np.float32)
    inter = {}
    x64 = x_bf16_f32.astype(np.float64)
    for e in range(expert_start, expert_end):
        m_idx, k_idx = np.nonzero(topk_ids == e)
        if m_idx.size == 0:
            continue
        w13 = round_bf16(layer.w13_f32(e)).astype(np.float64)     # W_deq in BF16
        w2 = round_bf16(layer.w2_f32(e)).astype(np.float64)
        c1 = round_bf16((x64[m_idx] @ w13.T).astype(np.float32))
        I = w13.shape[0] // 2
        a = round_bf16(silu_f32(c1[:, :I]) * c1[:, I:])
        y = (a.astype(np.float64) @ w2.T)
        w = topk_w[m_idx, k_idx].astype(np.float32)
        c3[m_idx, k_idx] = round_bf16((w[:, None].astype(np.float64) * y).astype(np.float32))
        if return_intermediates:
            inter[e] = dict(rows=m_idx, kpos=k_idx, c1=c1, a=a)
    out = np.zeros((M, H), np.float32)
    for j in range(k):                       # moe_sum_reduce: fp32 accumulate in k order
        out = (out + c3[:, j]).astype(np.float32)
    out = round_bf16(out)
    return (out, inter) if return_intermediates else out


def finish_gpu_path(routed_bf16, rsf=1.0, shared_bf16=None):
    """gpu_prefill.py:4471-4482: BF16 tensor ops `rsf * output + shared` / `output *= rsf`."""
    out = routed_bf16
    if shared_bf16 is not None:
        out = round_bf16(round_bf16(np.float32(rsf) * out) + shared_bf16)
    elif rsf != 1.0:
        out = round_bf16(np.float32(rsf) * out)
    return out


# ------------------------------------------------------------------ CPU-path numerics (numpy, small sizes)

def _round_half_away(x):
    x = np.asarray(x, np.float64)
    return np.copysign(np.floor(np.abs(x) + 0.5), x)


def quantize_activation_int16(act_f32, group_size):
    """src/kernel/avx2.rs:234-268 (BF16 input) and :274-304 (f32 input): identical math."""
    a = np.asarray(act_f32, np.float32).reshape(-1, group_size)
    mx = np.abs(a).max(axis=1).astype(np.float32)
    scale = np.where(mx > 0, mx / np.float32(32767.0), np.float32(1.0)).astype(np.float32)
    inv = np.where(mx > 0, np.float32(32767.0) / np.where(mx > 0, mx, 1).astype(np.float32),
                   np.float32(0.0)).astype(np.float32)
    q = _round_half_away((a * inv[:, None]).astype(np.float32)).clip(-32768, 32767).astype(np.int16)
    return q.reshape(-1), scale


def _int_matvec(w_int8_rows, w_scales_f32, a_i16, a_scales, gs):
    """avx2.rs:1066-1206: out[n] = sum_g fma(float(isum[n,g]), ws[n,g]*as[g], out) in group order."""
    n, k = w_int8_rows.shape
    isum = (w_int8_rows.astype(np.int64).reshape(n, k // gs, gs)
            * a_i16.astype(np.int64).reshape(1, k // gs, gs)).sum(axis=2)
    out = np.zeros(n, np.float32)
    for g in range(k // gs):
        comb = (w_scales_f32[:, g] * a_scales[g]).astype(np.float32)
        # fused multiply-add: one rounding
        out = (isum[:, g].astype(np.float32).astype(np.float64) * comb.astype(np.float64)
               + out.astype(np.float64)).astype(np.float32)
    return out


def expert_forward_cpu_int(layer: Int4Layer, e, a_i16, a_scales):
    """src/moe.rs:184-380 (standard SiLU branch; scalar sigmoid = exact exp here — the AVX2
    build uses a ~20-bit fast exp, avx2.rs:2229-2291, so compare with tolerance)."""
    gs = layer.group_size
    w13 = _int_matvec(layer.w13_ints(e), bf16_bits_to_f32(layer.w13_s[e]), a_i16, a_scales, gs)
    I = layer.I
    h = (silu_f32(w13[:I]) * w13[I:]).astype(np.float32)
    h_i16, h_s = quantize_activation_int16(h, gs)
    return _int_matvec(layer.w2_ints(e), bf16_bits_to_f32(layer.w2_s[e]), h_i16, h_s, gs)


def moe_forward_cpu_int(layer: Int4Layer, x_bf16_f32, topk_ids, topk_w, rsf=1.0, shared_out=None):
    """src/moe.rs:572-715: one token at a time; out = sum_i w_i * expert_i (f32, in order)."""
    M, H = x_bf16_f32.shape
    out = np.zeros((M, H), np.float32)
    for m in range(M):
        a_i16, a_s = quantize_activation_int16(x_bf16_f32[m], layer.group_size)
        for j in range(topk_ids.shape[1]):
            e = int(topk_ids[m, j])
            if e < 0:
                continue
            out[m] = (out[m] + np.float32(topk_w[m, j]) * expert_forward_cpu_int(layer, e, a_i16, a_s)).astype(np.float32)
    if shared_out is not None:
        out = (np.float32(rsf) * out + shared_out).astype(np.float32)
    return out


def quantize_int16_g32_with_sums(act_f32):
    """src/gguf_kernels.rs:110-172."""
    q, s = quantize_activation_int16(act_f32, 32)
    return q, s, q.reshape(-1, 32).astype(np.int32).sum(axis=1)


def gguf_matvec_int(ggml_type, rows_bytes, a_i16, a_s, a_sum, n, k):
    """src/gguf_kernels.rs:271-426 in exact real arithmetic per term (float64 accumulate);
    the AVX2 kernel's fp32 lane order is not reproduced — compare with tolerance 1e-5 rel."""
    a = a_i16.astype(np.float64).reshape(-1, 32)
    b = np.asarray(rows_bytes, np.uint8).reshape(n, -1)
    if ggml_type == G.GGML_Q8_0:
        blk = b.reshape(n, k // 32, 34)
        d = G._f16(blk[:, :, 0:2]).astype(np.float64)
        q = blk[:, :, 2:].view(np.int8).astype(np.float64)
        return ((q * a[None]).sum(axis=2) * d * a_s[None].astype(np.float64)).sum(axis=1).astype(np.float32)
    if ggml_type == G.GGML_Q4_K:
        blk = b.reshape(n, k // 256, 144)
        d = G._f16(blk[:, :, 0:2]).astype(np.float64)[:, :, None]
        dmin = G._f16(blk[:, :, 2:4]).astype(np.float64)[:, :, None]
        sc, mn = G.q4k_scales_mins(blk[:, :, 4:16])
        qs = blk[:, :, 16:144].reshape(n, k // 256, 4, 32)
        q = np.stack([qs & 0xF, qs >> 4], axis=3).reshape(n, k // 256, 8, 32).astype(np.float64)
        ag = a.reshape(k // 256, 8, 32)
        dot = (q * ag[None]).sum(axis=3)
        asg = a_s.astype(np.float64).reshape(k // 256, 8)[None]
        sm = a_sum.astype(np.float64).reshape(k // 256, 8)[None]
        val = d * sc * asg * dot - dmin * mn * asg * sm
        return val.sum(axis=(1, 2)).astype(np.float32)
    raise NotImplementedError(G.NAMES.get(ggml_type, ggml_type))


def expert_forward_cpu_gguf(layer: GgufLayer, e, x_f32):
    """src/gguf_kernels.rs:690-756 for the INT16 formats (Q4_K, Q8_0)."""
    H, I = layer.H, layer.I
    q, s, sm = quantize_int16_g32_with_sums(x_f32)
    g = gguf_matvec_int(layer.gate_up_type, layer.gate[e], q, s, sm, I, H)
    u = gguf_matvec_int(layer.gate_up_type, layer.up[e], q, s, sm, I, H)
    h = (silu_f32(g) * u).astype(np.float32)
    q2, s2, sm2 = quantize_int16_g32_with_sums(h)
    return gguf_matvec_int(layer.down_type, layer.down[e], q2, s2, sm2, H, I)
