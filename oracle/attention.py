"""Attention-block oracles (torch on CPU, float32/float64; test infrastructure only).

Gated DeltaNet (Qwen3-Next linear attention)
  definitional recurrence        python/krasis/linear_attention.py:542-561 (== SURVEY.md A.4)
  chunk-64 prefill               python/krasis/linear_attention.py:593-693 (_chunked_inner), :695-844 (_forward_chunked)
  un-interleave qkvz / ba        python/krasis/linear_attention.py:337-391
  causal depthwise conv(4)+SiLU  python/krasis/linear_attention.py:722-741
  l2norm                         python/krasis/linear_attention.py:114-117
  gated RMSNorm                  python/krasis/linear_attention.py:987-1004

GQA (Qwen3 / Qwen3-Next gated)   python/krasis/attention.py:496-687
  partial half-split RoPE, bf16 tables   :443-494   (pinned: reference execution, tests/golden/mla_rope_reference.npz)
  per-head RMSNorm (flashinfer.norm.rmsnorm)  :555-559
  KV cast to cache dtype (FP8 E4M3, unscaled) :582-583
  causal softmax(QK^T * d^-1/2) V             :596-642
  sigmoid output gate                          :665-666

MLA (DeepSeek-V2 / Kimi)          python/krasis/attention.py:213-374 (absorbed form, as the reference runs it)
  YaRN inverse frequencies + BF16 cos/sin tables   :119-163   (pinned: tests/golden/mla_rope_reference.npz holds
  de-interleave + half-split rotation in BF16      :165-211    outputs of the reference's own _get_rope_cos_sin/_apply_rope)
  w_kc absorb / w_vc un-absorb einsums in fp32     :268-271,364-368
  attention core = FlashInfer MLA over the BF16 upcast of the FP8 latent pages (:324-361): third-party, absent from
  the tree -> restated as fp32 softmax attention; parity of that core is unpinned beyond the reference's own cos>0.999 check

tests/golden/make_attention_golden.py imports the REFERENCE's own linear_attention module (CPU, eager) to
generate fixtures these functions are pinned against.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------- Gated DeltaNet

def l2norm(x, eps=1e-6):
    return x * torch.rsqrt((x * x).sum(dim=-1, keepdim=True) + eps)


def gdn_unsplit(mixed_qkvz, mixed_ba, nk, nv, dk, dv):
    """linear_attention.py:337-391: per key-head group [q(dk) k(dk) v(r*dv) z(r*dv)], [b(r) a(r)]."""
    M = mixed_qkvz.shape[0]
    r = nv // nk
    g = mixed_qkvz.view(M, nk, 2 * dk + 2 * dv * r)
    q, k, v, z = torch.split(g, [dk, dk, r * dv, r * dv], dim=2)
    v, z = v.reshape(M, nv, dv), z.reshape(M, nv, dv)
    b, a = torch.split(mixed_ba.view(M, nk, 2 * r), [r, r], dim=2)
    return q, k, v, z, b.reshape(M, nv), a.reshape(M, nv)


def gdn_conv_silu(q, k, v, conv_weight, conv_state=None):
    """Depthwise causal conv (kernel K=4) over [q|k|v] channels + SiLU; returns (conv_out [M,C] in the input
    dtype, new_state [C,K]).  conv_weight [C, K]; state = last K inputs (linear_attention.py:725-728)."""
    M = q.shape[0]
    mixed = torch.cat([q.reshape(M, -1), k.reshape(M, -1), v.reshape(M, -1)], dim=-1)     # [M, C]
    C, K = conv_weight.shape
    if conv_state is None:
        conv_state = torch.zeros(C, K, dtype=mixed.dtype)
    inp = torch.cat([conv_state, mixed.t()], dim=-1)                                      # [C, K+M]
    new_state = inp[:, -K:].clone()
    out = F.conv1d(inp.unsqueeze(0).to(conv_weight.dtype), conv_weight.unsqueeze(1), groups=C)[0]   # [C, M+1]
    out = F.silu(out[:, -M:]).to(mixed.dtype)
    return out.t().contiguous(), new_state


def gdn_gates(b, a, A_log, dt_bias):
    """beta = sigmoid(b);  g = -exp(A_log) * softplus(a + dt_bias)   (linear_attention.py:523-524)."""
    return torch.sigmoid(b), -A_log.float().exp() * F.softplus(a.float() + dt_bias)


def gdn_recurrent(q, k, v, beta, g, state=None):
    """Definitional oracle.  q,k [M,nv,dk] (already l2-normed, q scaled), v [M,nv,dv], beta,g [M,nv].
    Returns (out [M,nv,dv], state [nv,dk,dv]); float64 internally."""
    M, nv, dk = q.shape
    dv = v.shape[-1]
    S = torch.zeros(nv, dk, dv, dtype=torch.float64) if state is None else state.double().clone()
    q, k, v, beta, g = q.double(), k.double(), v.double(), beta.double(), g.double()
    out = torch.empty(M, nv, dv, dtype=torch.float64)
    for t in range(M):
        S = S * g[t].exp()[:, None, None]
        mem = (S * k[t][:, :, None]).sum(dim=1)
        delta = (v[t] - mem) * beta[t][:, None]
        S = S + k[t][:, :, None] * delta[:, None, :]
        out[t] = (S * q[t][:, :, None]).sum(dim=1)
    return out, S


def gdn_chunked(q, k, v, beta, g, state=None, chunk=64, dtype=torch.float32):
    """The reference's chunked prefill math (linear_attention.py:593-693,776-811) in `dtype`."""
    M, nv, dk = q.shape
    dv = v.shape[-1]
    pad = (chunk - M % chunk) % chunk
    def padt(x):
        return F.pad(x, (0, 0) * (x.dim() - 1) + (0, pad)) if pad else x
    q_, k_, v_ = [padt(x.to(dtype)).transpose(0, 1) for x in (q, k, v)]                  # [nv, T, d]
    beta_, g_ = padt(beta.to(dtype)).t(), padt(g.to(dtype)).t()                          # [nv, T]
    T = M + pad
    n = T // chunk
    qc, kc, vc = [x.reshape(nv, n, chunk, -1) for x in (q_, k_, v_)]
    bc, gc = beta_.reshape(nv, n, chunk), g_.reshape(nv, n, chunk)
    vb, kb = vc * bc[..., None], kc * bc[..., None]
    gcum = gc.cumsum(dim=-1)
    decay = (gcum[..., :, None] - gcum[..., None, :]).tril().exp().tril()
    A = -(kb @ kc.transpose(-1, -2)) * decay
    A = A.masked_fill(torch.triu(torch.ones(chunk, chunk, dtype=torch.bool), 0), 0)
    value_corr = torch.linalg.solve_triangular(-A, vb, upper=False, unitriangular=True)
    k_cumdecay = torch.linalg.solve_triangular(-A, kb * gcum.exp()[..., None], upper=False, unitriangular=True)
    S = torch.zeros(nv, dk, dv, dtype=dtype) if state is None else state.to(dtype).clone()
    out = torch.zeros(nv, n, chunk, dv, dtype=dtype)
    strict = torch.triu(torch.ones(chunk, chunk, dtype=torch.bool), 1)
    for i in range(n):
        qi, ki, vi, gi, kcd, dm = qc[:, i], kc[:, i], value_corr[:, i], gcum[:, i], k_cumdecay[:, i], decay[:, i]
        intra = ((qi @ ki.transpose(-1, -2)) * dm).masked_fill(strict, 0)
        v_new = vi - kcd @ S
        out[:, i] = (qi * gi[..., None].exp()) @ S + intra @ v_new
        gl = gi[:, -1]
        S = S * gl[:, None, None].exp() + (ki * (gl[:, None] - gi).exp()[..., None]).transpose(-1, -2) @ v_new
    out = out.reshape(nv, T, dv)[:, :M].transpose(0, 1)
    return out, S


def gdn_segment_affine(k, v, beta, g, chunk=64, dtype=torch.float64):
    """The state update of gdn_chunked over a whole token segment as ONE affine map  S_end = P S_start + Q  per head
    (P [nv, dk, dk], Q [nv, dk, dv]).  Per chunk  S' = e^{g_last} S + Kd^T (vcorr - kcd S)  with Kd = k e^{g_last - gcum}, so
    P_c = e^{g_last} I - Kd^T kcd and Q_c = Kd^T vcorr; chunks compose as (P_b, Q_b) o (P_a, Q_a) = (P_b P_a, P_b Q_a + Q_b).
    Groundwork for a sequence-parallel scan across ranks: each rank reduces its own token segment to (P, Q)
    without knowing the incoming state.  The segment length must be a multiple of `chunk` (segments are cut at chunk boundaries)."""
    M, nv, dk = k.shape
    dv = v.shape[-1]
    assert M % chunk == 0
    n = M // chunk
    kc = k.to(dtype).transpose(0, 1).reshape(nv, n, chunk, dk)
    vc = v.to(dtype).transpose(0, 1).reshape(nv, n, chunk, dv)
    bc, gc = beta.to(dtype).t().reshape(nv, n, chunk), g.to(dtype).t().reshape(nv, n, chunk)
    vb, kb = vc * bc[..., None], kc * bc[..., None]
    gcum = gc.cumsum(dim=-1)
    decay = (gcum[..., :, None] - gcum[..., None, :]).tril().exp().tril()
    A = -(kb @ kc.transpose(-1, -2)) * decay
    A = A.masked_fill(torch.triu(torch.ones(chunk, chunk, dtype=torch.bool), 0), 0)
    vcorr = torch.linalg.solve_triangular(-A, vb, upper=False, unitriangular=True)
    kcd = torch.linalg.solve_triangular(-A, kb * gcum.exp()[..., None], upper=False, unitriangular=True)
    P = torch.eye(dk, dtype=dtype).expand(nv, dk, dk).clone()
    Q = torch.zeros(nv, dk, dv, dtype=dtype)
    for i in range(n):
        gl = gcum[:, i, -1]
        kd_t = (kc[:, i] * (gl[:, None] - gcum[:, i]).exp()[..., None]).transpose(-1, -2)          # [nv, dk, chunk]
        Pc = torch.eye(dk, dtype=dtype)[None] * gl.exp()[:, None, None] - kd_t @ kcd[:, i]
        Qc = kd_t @ vcorr[:, i]
        P, Q = Pc @ P, Pc @ Q + Qc
    return P, Q


def gated_rmsnorm(x, gate, weight, eps):
    """linear_attention.py:987-1004 (norm in fp32 -> input dtype; gate silu in fp32 -> input dtype; product)."""
    dt = x.dtype
    xf = x.float()
    xn = (weight.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(dt)
    return xn * F.silu(gate.float()).to(dt)


def gdn_layer_prefill(hidden, w, cfg, conv_state=None, state=None, recurrent=False):
    """Whole GatedDeltaNetAttention._forward_chunked (BF16 weights path): hidden [M,H] bf16 -> [M,H] bf16.
    w: dict(in_proj_qkvz, in_proj_ba, conv1d_weight [C,1,K], A_log, dt_bias, norm_weight, out_proj)."""
    nk, nv, dk, dv = cfg["nk"], cfg["nv"], cfg["dk"], cfg["dv"]
    M = hidden.shape[0]
    qkvz = F.linear(hidden, w["in_proj_qkvz"])
    ba = F.linear(hidden, w["in_proj_ba"])
    q, k, v, z, b, a = gdn_unsplit(qkvz, ba, nk, nv, dk, dv)
    conv_out, new_conv = gdn_conv_silu(q, k, v, w["conv1d_weight"].squeeze(1), conv_state)
    kd = nk * dk
    qa = conv_out[:, :kd].reshape(M, nk, dk)
    ka = conv_out[:, kd:2 * kd].reshape(M, nk, dk)
    va = conv_out[:, 2 * kd:].reshape(M, nv, dv)
    beta, g = gdn_gates(b, a, w["A_log"], w["dt_bias"])
    r = nv // nk
    if r > 1:
        qa, ka = qa.repeat_interleave(r, dim=1), ka.repeat_interleave(r, dim=1)
    qa, ka = l2norm(qa), l2norm(ka)
    qa = qa * (1.0 / dk ** 0.5)
    if recurrent:
        core, S = gdn_recurrent(qa.float(), ka.float(), va.float(), beta.float(), g, state)
        core = core.float()
    else:
        core, S = gdn_chunked(qa.float(), ka.float(), va.float(), beta.float(), g, state)
    core = core.to(hidden.dtype)
    out = gated_rmsnorm(core, z, w["norm_weight"], cfg["eps"])
    flat = out.reshape(M, nv * dv).to(torch.bfloat16)
    return F.linear(flat, w["out_proj"]), new_conv, S


# --------------------------------------------------------------------------------------------- GQA

def rmsnorm(x, weight, eps):
    """flashinfer.norm.rmsnorm semantics: fp32 accumulate, output in x.dtype."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * weight.float()).to(x.dtype)


def rope_tables(max_len, rotary_dim, theta):
    """attention.py:443-459: fp32 angles, tables stored in BF16."""
    freqs = 1.0 / (theta ** (torch.arange(0, rotary_dim, 2).float() / rotary_dim))
    ang = torch.outer(torch.arange(max_len, dtype=torch.float32), freqs)
    return ang.cos().to(torch.bfloat16), ang.sin().to(torch.bfloat16)


def apply_rope(x, cos, sin):
    """attention.py:480-490: half-split rotation of the first 2*d2 dims, rest pass through (bf16 arithmetic)."""
    d2 = cos.shape[-1]
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    x1, x2 = x[..., :d2], x[..., d2:2 * d2]
    rot = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)
    return torch.cat([rot, x[..., 2 * d2:]], dim=-1) if 2 * d2 < x.shape[-1] else rot


def gqa_core(q, k_cache, v_cache, q_pos, sm_scale):
    """Causal attention of M new queries at positions q_pos over a cache of L keys (positions 0..L-1).
    q [M,nh,d] bf16; k_cache/v_cache [L,nkv,d] (already cast through the KV dtype, e.g. fp8 -> float).
    fp32 softmax, like FlashInfer; returns [M,nh,d] bf16."""
    M, nh, d = q.shape
    nkv = k_cache.shape[1]
    grp = nh // nkv
    kf = k_cache.float().repeat_interleave(grp, dim=1)          # [L, nh, d]
    vf = v_cache.float().repeat_interleave(grp, dim=1)
    s = torch.einsum("mhd,lhd->hml", q.float(), kf) * sm_scale
    L = k_cache.shape[0]
    mask = torch.arange(L)[None, :] > q_pos[:, None]
    s = s.masked_fill(mask[None], float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.einsum("hml,lhd->mhd", p, vf).to(torch.bfloat16)


def gqa_layer_prefill(hidden, w, cfg, positions, k_cache=None, v_cache=None, kv_dtype=torch.float8_e4m3fn, rows=None):
    """GQAAttention.forward (attention.py:496-687) for BF16 weights: returns (out [M,H] bf16, k_cache, v_cache)
    with the caches holding values already rounded through kv_dtype (stored as that dtype).
    rows (optional LongTensor): evaluate the attention core and the output projection only for these query rows of
    `hidden` (the projections, norms, RoPE and the cache append still cover every token) -> out [len(rows), H]; this is
    how the 8K-token parity tests stay within seconds of CPU time."""
    nh, nkv, d = cfg["nh"], cfg["nkv"], cfg["d"]
    M = hidden.shape[0]
    q_raw = F.linear(hidden, w["q_proj"])
    k = F.linear(hidden, w["k_proj"]).reshape(M, nkv, d)
    v = F.linear(hidden, w["v_proj"]).reshape(M, nkv, d)
    gated = q_raw.shape[1] == 2 * nh * d
    if gated:
        q, gate = q_raw.view(M, nh, 2 * d).chunk(2, dim=-1)
        gate = gate.reshape(M, nh * d)
    else:
        q = q_raw.reshape(M, nh, d)
    if w.get("q_norm") is not None:
        q = rmsnorm(q, w["q_norm"], cfg["eps"])
    if w.get("k_norm") is not None:
        k = rmsnorm(k, w["k_norm"], cfg["eps"])
    cos, sin = rope_tables(int(positions.max()) + 1, cfg["rotary_dim"], cfg["theta"])
    q = apply_rope(q, cos[positions], sin[positions])
    k = apply_rope(k, cos[positions], sin[positions])
    k_new, v_new = k.to(kv_dtype), v.to(kv_dtype)
    k_cache = k_new if k_cache is None else torch.cat([k_cache, k_new], dim=0)
    v_cache = v_new if v_cache is None else torch.cat([v_cache, v_new], dim=0)
    if rows is not None:
        q, positions, M = q[rows], positions[rows], len(rows)
        if gated:
            gate = gate[rows]
    attn = gqa_core(q.to(torch.bfloat16), k_cache, v_cache, positions, 1.0 / math.sqrt(d))
    flat = attn.reshape(M, nh * d)
    if gated:
        flat = flat * torch.sigmoid(gate)
    return F.linear(flat.to(torch.bfloat16), w["o_proj"]), k_cache, v_cache


# --------------------------------------------------------------------------------------------- MLA

def mla_inv_freq(rope_dim, theta, rope_scaling=None):
    """attention.py:129-157 (YaRN blend as HF DeepseekV2YarnRotaryEmbedding)."""
    dim = rope_dim
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim))
    cfg = rope_scaling or {}
    factor = cfg.get("factor", 1.0)
    if factor > 1.0:
        original_max = cfg.get("original_max_position_embeddings", 4096)
        beta_fast, beta_slow = cfg.get("beta_fast", 32.0), cfg.get("beta_slow", 1.0)
        low = math.floor(dim * math.log(original_max / (beta_fast * 2 * math.pi)) / (2 * math.log(theta)))
        high = math.ceil(dim * math.log(original_max / (beta_slow * 2 * math.pi)) / (2 * math.log(theta)))
        low, high = max(low, 0), min(high, dim // 2 - 1)
        ramp = torch.clamp((torch.arange(dim // 2).float() - low) / max(high - low, 0.001), 0, 1)
        mask = 1.0 - ramp
        freqs = (freqs / factor) * (1 - mask) + freqs * mask
    return freqs


def mla_rope_tables(max_len, rope_dim, theta, rope_scaling=None):
    f = torch.outer(torch.arange(max_len, dtype=torch.float32), mla_inv_freq(rope_dim, theta, rope_scaling))
    return f.cos().to(torch.bfloat16), f.sin().to(torch.bfloat16)


def mla_deinterleave(x):
    d = x.shape[-1]
    return x.view(*x.shape[:-1], d // 2, 2).transpose(-1, -2).reshape(x.shape)


def mla_apply_rope(x, cos, sin):
    """x [M, heads, rope] bf16 interleaved -> de-interleaved and rotated, BF16 arithmetic (attention.py:165-211)."""
    x = mla_deinterleave(x)
    d2 = x.shape[-1] // 2
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    x1, x2 = x[..., :d2], x[..., d2:]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)


def mla_sm_scale(nope, rope, rope_scaling=None):
    s = 1.0 / math.sqrt(nope + rope)
    cfg = rope_scaling or {}
    if cfg.get("factor", 1.0) > 1.0:
        m = 0.1 * cfg.get("mscale_all_dim", 0) * math.log(cfg["factor"]) + 1.0
        s *= m * m
    return s


def mla_layer_prefill(hidden, w, cfg, positions, ckv_cache=None, kpe_cache=None, kv_dtype=torch.float8_e4m3fn):
    """MLAAttention.forward (attention.py:213-374) for BF16 weights, absorbed form.  Returns (out [M,H] bf16, ckv_cache
    [L,lora], kpe_cache [L,rope]) with the caches stored in kv_dtype."""
    nh, nope, rope, dv, lora = cfg["nh"], cfg["nope"], cfg["rope"], cfg["dv"], cfg["lora"]
    M = hidden.shape[0]
    kv_out = F.linear(hidden, w["kv_a_proj_with_mqa"])
    ckv = rmsnorm(kv_out[:, :lora], w["kv_a_layernorm"], cfg["eps"])
    k_pe = kv_out[:, lora:]
    if cfg.get("q_lora"):
        q_c = rmsnorm(F.linear(hidden, w["q_a_proj"]), w["q_a_layernorm"], cfg["eps"])
        q_full = F.linear(q_c, w["q_b_proj"])
    else:
        q_full = F.linear(hidden, w["q_proj"])
    q_full = q_full.reshape(M, nh, nope + rope)
    q_nope, q_pe = q_full[:, :, :nope], q_full[:, :, nope:]
    cos, sin = mla_rope_tables(int(positions.max()) + 1, rope, cfg["theta"], cfg.get("rope_scaling"))
    q_pe = mla_apply_rope(q_pe, cos[positions], sin[positions])
    k_pe = mla_apply_rope(k_pe.unsqueeze(1), cos[positions], sin[positions]).squeeze(1)
    q_abs = torch.einsum("mhi,hid->mhd", q_nope.float(), w["w_kc"].float()).to(torch.bfloat16)
    ckv_new, kpe_new = ckv.to(kv_dtype), k_pe.to(kv_dtype)
    ckv_cache = ckv_new if ckv_cache is None else torch.cat([ckv_cache, ckv_new], dim=0)
    kpe_cache = kpe_new if kpe_cache is None else torch.cat([kpe_cache, kpe_new], dim=0)
    cf, pf = ckv_cache.to(torch.bfloat16).float(), kpe_cache.to(torch.bfloat16).float()
    s = (torch.einsum("mhd,ld->hml", q_abs.float(), cf) + torch.einsum("mhr,lr->hml", q_pe.float(), pf)) \
        * mla_sm_scale(nope, rope, cfg.get("rope_scaling"))
    L = cf.shape[0]
    s = s.masked_fill((torch.arange(L)[None, :] > positions[:, None])[None], float("-inf"))
    attn_out = torch.einsum("hml,ld->mhd", torch.softmax(s, dim=-1), cf).to(torch.bfloat16)
    proj = torch.einsum("mhd,hod->mho", attn_out.float(), w["w_vc"].float()).to(torch.bfloat16)
    return F.linear(proj.reshape(M, nh * dv), w["o_proj"]), ckv_cache, kpe_cache
