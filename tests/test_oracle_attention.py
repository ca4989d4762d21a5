"""Pins oracle/attention.py against fixtures produced by EXECUTING the reference's own
python/krasis/linear_attention.py on CPU (tests/golden/make_attention_golden.py), and checks the
chunked form against the definitional recurrence."""
import os

import numpy as np
import pytest
import torch

from oracle import attention as A

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "gdn_reference.npz"))


def _w():
    bf = torch.bfloat16
    return {k[2:]: torch.from_numpy(G[k]).to(bf) for k in G.files if k.startswith("w_")}


def _cfg():
    nk, nv, dk, dv, H, K = [int(v) for v in G["cfg"]]
    return dict(nk=nk, nv=nv, dk=dk, dv=dv, H=H, K=K, eps=1e-6)


def test_gdn_layer_matches_reference_execution_two_calls_with_state_carry():
    w, cfg = _w(), _cfg()
    x1 = torch.from_numpy(G["x1"]).to(torch.bfloat16)
    x2 = torch.from_numpy(G["x2"]).to(torch.bfloat16)
    y1, conv1, st1 = A.gdn_layer_prefill(x1, w, cfg)
    assert torch.equal(conv1.float(), torch.from_numpy(G["conv1"])[0])           # conv state: exact
    assert np.abs(st1.numpy() - G["st1"][0]).max() < 1e-5 * max(1.0, np.abs(G["st1"]).max())
    d1 = np.abs(y1.float().numpy() - G["y1"])
    assert d1.max() <= 2 ** -7 * np.abs(G["y1"]).max(), d1.max()                # <= 1 bf16 ulp of the max
    y2, conv2, st2 = A.gdn_layer_prefill(x2, w, cfg, conv_state=conv1, state=st1)
    assert torch.equal(conv2.float(), torch.from_numpy(G["conv2"])[0])
    assert np.abs(st2.numpy() - G["st2"][0]).max() < 1e-5 * max(1.0, np.abs(G["st2"]).max())
    assert np.abs(y2.float().numpy() - G["y2"]).max() <= 2 ** -7 * np.abs(G["y2"]).max()


def test_gdn_chunked_equals_definitional_recurrence():
    torch.manual_seed(0)
    M, nv, dk, dv = 200, 3, 16, 24
    q = A.l2norm(torch.randn(M, nv, dk)) / dk ** 0.5
    k = A.l2norm(torch.randn(M, nv, dk))
    v = torch.randn(M, nv, dv)
    beta = torch.sigmoid(torch.randn(M, nv))
    g = -torch.rand(M, nv) * 0.5
    s0 = torch.randn(nv, dk, dv) * 0.1
    o_r, s_r = A.gdn_recurrent(q, k, v, beta, g, s0)
    o_c, s_c = A.gdn_chunked(q, k, v, beta, g, s0, dtype=torch.float64)
    assert (o_r - o_c).abs().max() < 1e-10 and (s_r - s_c).abs().max() < 1e-10
    o_f, s_f = A.gdn_chunked(q, k, v, beta, g, s0, dtype=torch.float32)
    assert (o_r - o_f.double()).abs().max() < 2e-5


def test_gqa_rope_and_cache_semantics():
    torch.manual_seed(1)
    cfg = dict(nh=4, nkv=2, d=32, rotary_dim=8, theta=10000.0, eps=1e-6)
    H = 64
    bf = torch.bfloat16
    w = dict(q_proj=(torch.randn(2 * cfg["nh"] * cfg["d"], H) * 0.2).to(bf), k_proj=(torch.randn(cfg["nkv"] * cfg["d"], H) * 0.2).to(bf),
             v_proj=(torch.randn(cfg["nkv"] * cfg["d"], H) * 0.2).to(bf), o_proj=(torch.randn(H, cfg["nh"] * cfg["d"]) * 0.2).to(bf),
             q_norm=(1 + 0.1 * torch.randn(cfg["d"])).to(bf), k_norm=(1 + 0.1 * torch.randn(cfg["d"])).to(bf))
    x = torch.randn(40, H).to(bf)
    pos = torch.arange(40)
    y, kc, vc = A.gqa_layer_prefill(x, w, cfg, pos)
    assert kc.dtype == torch.float8_e4m3fn and kc.shape == (40, 2, 32)
    # chunked prefill (two calls appending to the cache) equals one call: causal attention over the paged cache
    y_a, kc_a, vc_a = A.gqa_layer_prefill(x[:25], w, cfg, pos[:25])
    y_b, _, _ = A.gqa_layer_prefill(x[25:], w, cfg, pos[25:], kc_a, vc_a)
    assert torch.equal(torch.cat([y_a, y_b]), y)
    # token 0 attends only to itself: output = o_proj(sigmoid(gate) * v0 expanded over the query groups)
    # partial RoPE leaves dims >= rotary_dim untouched
    q = torch.randn(3, 4, 32).to(bf)
    cos, sin = A.rope_tables(8, 8, 10000.0)
    r = A.apply_rope(q, cos[torch.tensor([0, 3, 7])], sin[torch.tensor([0, 3, 7])])
    assert torch.equal(r[..., 8:], q[..., 8:]) and torch.equal(r[0], q[0])


# ------------------------------------------------------------------------------------------------ MLA

MLA_G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mla_rope_reference.npz"))
V2_LITE_ROPE = {"beta_fast": 32, "beta_slow": 1, "factor": 40, "mscale": 0.707, "mscale_all_dim": 0.707,
                "original_max_position_embeddings": 4096, "type": "yarn"}


@pytest.mark.parametrize("tag,scaling", [("yarn", V2_LITE_ROPE), ("plain", None)])
def test_mla_rope_matches_reference_execution(tag, scaling):
    """YaRN tables, de-interleave + BF16 rotation and sm_scale: bit-exact against outputs of the reference's own code."""
    rows = torch.from_numpy(MLA_G[f"{tag}_rows"])
    cos, sin = A.mla_rope_tables(5000, 64, 10000.0, scaling)
    assert np.array_equal(cos[rows].float().numpy(), MLA_G[f"{tag}_cos"])
    assert np.array_equal(sin[rows].float().numpy(), MLA_G[f"{tag}_sin"])
    pos = torch.from_numpy(MLA_G[f"{tag}_pos"])
    q = torch.from_numpy(MLA_G[f"{tag}_q_in"]).to(torch.bfloat16)
    k = torch.from_numpy(MLA_G[f"{tag}_k_in"]).to(torch.bfloat16)
    assert np.array_equal(A.mla_apply_rope(q, cos[pos], sin[pos]).float().numpy(), MLA_G[f"{tag}_q_out"])
    assert np.array_equal(A.mla_apply_rope(k, cos[pos], sin[pos]).float().numpy(), MLA_G[f"{tag}_k_out"])
    assert A.mla_sm_scale(128, 64, scaling) == float(MLA_G[f"{tag}_sm_scale"])


def test_mla_absorbed_equals_explicit_heads():
    """The absorbed form the reference runs == ordinary multi-head attention with k_nope = ckv W_kc^T, v = ckv W_vc^T
    (what the B200 path computes), to BF16-rounding tolerance."""
    torch.manual_seed(3)
    nh, nope, rope, dv, lora, H, M = 4, 128, 64, 128, 512, 256, 70
    bf = torch.bfloat16
    w = dict(q_proj=(torch.randn(nh * (nope + rope), H) * 0.06).to(bf), kv_a_proj_with_mqa=(torch.randn(lora + rope, H) * 0.06).to(bf),
             kv_a_layernorm=(1 + 0.1 * torch.randn(lora)).to(bf), w_kc=(torch.randn(nh, nope, lora) * 0.04).to(bf),
             w_vc=(torch.randn(nh, dv, lora) * 0.04).to(bf), o_proj=(torch.randn(H, nh * dv) * 0.05).to(bf))
    cfg = dict(nh=nh, nope=nope, rope=rope, dv=dv, lora=lora, theta=10000.0, eps=1e-6, rope_scaling=V2_LITE_ROPE)
    x = torch.randn(M, H).to(bf)
    pos = torch.arange(M)
    out, ckv, kpe = A.mla_layer_prefill(x, w, cfg, pos)
    # explicit heads from the same caches
    cf, pf = ckv.to(bf).float(), kpe.to(bf).float()
    k_nope = torch.einsum("ld,hid->lhi", cf, w["w_kc"].float())
    v = torch.einsum("ld,hod->lho", cf, w["w_vc"].float())
    q_full = torch.nn.functional.linear(x, w["q_proj"]).reshape(M, nh, nope + rope)
    cos, sin = A.mla_rope_tables(M, rope, 10000.0, V2_LITE_ROPE)
    q_pe = A.mla_apply_rope(q_full[:, :, nope:], cos[pos], sin[pos]).float()
    s = (torch.einsum("mhi,lhi->hml", q_full[:, :, :nope].float(), k_nope) + torch.einsum("mhr,lr->hml", q_pe, pf)) \
        * A.mla_sm_scale(nope, rope, V2_LITE_ROPE)
    s = s.masked_fill((torch.arange(M)[None, :] > pos[:, None])[None], float("-inf"))
    o = torch.einsum("hml,lho->mho", torch.softmax(s, -1), v).to(bf)
    want = torch.nn.functional.linear(o.reshape(M, nh * dv), w["o_proj"]).float()
    assert (out.float() - want).abs().max() <= 2 * 2 ** -7 * want.abs().max()


def test_mla_product_inv_freq_is_the_oracles():
    from krasis_b200.attention import mla_rope_inv_freq, mla_sm_scale
    for sc in (V2_LITE_ROPE, None):
        assert np.array_equal(mla_rope_inv_freq(64, 10000.0, sc), A.mla_inv_freq(64, 10000.0, sc).numpy())
        assert mla_sm_scale(128, 64, sc) == A.mla_sm_scale(128, 64, sc)


@pytest.mark.parametrize("tag", ["gqa_partial", "gqa_full"])
def test_gqa_rope_matches_reference_execution(tag):
    """GQA RoPE tables and (partial) half-split rotation in BF16: bit-exact against the reference's own
    GQAAttention._get_rope_cos_sin/_apply_rope (tests/golden/make_mla_golden.py)."""
    d, rot, theta = MLA_G[f"{tag}_cfg"]
    pos = torch.from_numpy(MLA_G[f"{tag}_pos"])
    cos, sin = A.rope_tables(8192, int(rot), float(theta))
    assert np.array_equal(cos[pos].float().numpy(), MLA_G[f"{tag}_cos"])
    assert np.array_equal(sin[pos].float().numpy(), MLA_G[f"{tag}_sin"])
    for name in ("q", "k"):
        x = torch.from_numpy(MLA_G[f"{tag}_{name}_in"]).to(torch.bfloat16)
        assert np.array_equal(A.apply_rope(x, cos[pos], sin[pos]).float().numpy(), MLA_G[f"{tag}_{name}_out"])


def test_gdn_segment_state_update_is_affine_and_composes():
    """oracle.attention.gdn_segment_affine: a token segment's effect on the recurrent state is S_end = P S_start + Q, and two
    segments compose as (P2 P1, P2 Q1 + Q2) — checked against the sequential chunked pass with a non-zero incoming state.  This is
    the identity a sequence-parallel scan across ranks rests on."""
    torch.manual_seed(13)
    M, nv, dk, dv = 256, 3, 16, 8
    q = A.l2norm(torch.randn(M, nv, dk))
    k = A.l2norm(torch.randn(M, nv, dk))
    v = torch.randn(M, nv, dv)
    beta = torch.sigmoid(torch.randn(M, nv))
    g = -torch.rand(M, nv) * 0.3
    S0 = torch.randn(nv, dk, dv, dtype=torch.float64) * 0.5
    _, S_seq = A.gdn_chunked(q, k, v, beta, g, state=S0, chunk=64, dtype=torch.float64)
    P, Q = A.gdn_segment_affine(k, v, beta, g, chunk=64)
    assert torch.allclose(P @ S0 + Q, S_seq, rtol=1e-9, atol=1e-11)
    h = M // 2                                                       # two "ranks", each reducing its own half without the incoming state
    P1, Q1 = A.gdn_segment_affine(k[:h], v[:h], beta[:h], g[:h], chunk=64)
    P2, Q2 = A.gdn_segment_affine(k[h:], v[h:], beta[h:], g[h:], chunk=64)
    assert torch.allclose(P2 @ P1, P, rtol=1e-9, atol=1e-11) and torch.allclose(P2 @ Q1 + Q2, Q, rtol=1e-9, atol=1e-11)
    _, S_def = A.gdn_recurrent(q, k, v, beta, g, state=S0)           # and the definitional recurrence agrees
    assert torch.allclose(P @ S0 + Q, S_def, rtol=1e-8, atol=1e-10)
