"""Pins the oracle against every known-answer vector the reference's own unit tests hold
for this path (tests/golden/reference_kats.json; each entry cites the Rust test), and against
the independent gguf-py dequantiser.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import quant, router, moe as omoe, gguf_blocks as G
from oracle.bf16 import f32_to_bf16_bits, bf16_bits_to_f32, round_bf16

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
VEC = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))


def _synth(n, off=0.5, amp=0.2):
    i = np.arange(n, dtype=np.float32)
    return f32_to_bf16_bits(((i / np.float32(n)) - np.float32(off)) * np.float32(amp))


def test_bf16_rne_matches_reference_formula():
    # marlin.rs:25-30 — compare with torch's RNE on random + tie values
    import torch
    x = np.random.default_rng(0).normal(0, 1, 4096).astype(np.float32)
    ties = (np.arange(0x3F80, 0x3F90, dtype=np.uint32) << 16 | 0x8000).view(np.float32)
    x = np.concatenate([x, ties])
    ours = f32_to_bf16_bits(x)
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(ours, ref)


def test_gguf_block_table():
    names = {v: k for k, v in G.NAMES.items()}
    names["F32"] = G.GGML_F32
    for name, (be, bb) in KATS["gguf_block_table"]["entries"].items():
        assert G.BLOCK[names[name]] == (be, bb), name


def test_get_scale_min_k4_kat():
    k = KATS["get_scale_min_k4"]
    sc, mn = G.get_scale_min_k4(k["j"], np.array(k["scales"], np.uint8))
    assert int(sc) == k["expect_sc"] and int(mn) == k["expect_mn"]


def test_q4_0_all_0x88_kat():
    k = KATS["q4_0_all_0x88"]
    blk = np.zeros(18, np.uint8)
    blk[0:2] = np.array([k["d"]], np.float16).view(np.uint8)
    blk[2:] = k["byte"]
    w = G.dequant_q4_0(blk, 32)
    assert abs(float((w * np.float32(k["x"])).sum()) - k["expect"]) < k["tol"]


def test_q8_0_simple_kat():
    k = KATS["q8_0_simple"]
    blk = np.zeros(34, np.uint8)
    blk[0:2] = np.array([k["d"]], np.float16).view(np.uint8)
    blk[2:] = k["q"]
    w = G.dequant_q8_0(blk, 32)
    assert abs(float((w * np.float32(k["x"])).sum()) - k["expect"]) < k["tol"]
    # and through the INT16 activation path of the CPU kernel (gguf_kernels.rs:390-426)
    q, s, sm = omoe.quantize_int16_g32_with_sums(np.ones(32, np.float32))
    y = omoe.gguf_matvec_int(G.GGML_Q8_0, blk[None], q, s, sm, 1, 32)
    assert abs(float(y[0]) - k["expect"]) < k["tol"]


def test_int16_quant_roundtrip_and_sums_kat():
    k = KATS["int16_quant_roundtrip_g32"]
    i = np.arange(k["k"], dtype=np.float32)
    val = (i - 32.0) * np.float32(0.1)
    bits = (val.view(np.uint32) >> 16).astype(np.uint16)          # the test truncates, not RNE
    x = bf16_bits_to_f32(bits)
    q, s, sm = omoe.quantize_int16_g32_with_sums(x)
    rec = q.reshape(-1, 32).astype(np.float32) * s[:, None]
    assert np.abs(rec.reshape(-1) - x).max() < k["max_abs_err"]
    assert np.array_equal(sm, q.reshape(-1, 32).astype(np.int32).sum(axis=1))


def test_int4_roundtrip_synthetic_kat():
    k = KATS["int4_roundtrip_synthetic"]
    w = _synth(k["rows"] * k["cols"]).reshape(k["rows"], k["cols"])
    p, s = quant.quantize_int4(w, k["group_size"])
    assert p.shape == (k["rows"], k["cols"] // 8) and s.shape == (k["rows"], 1)
    err = np.abs(quant.dequantize_int4(p, s, k["group_size"]) - bf16_bits_to_f32(w)).max()
    assert err < k["max_err_lt"]


def test_marlin_tables_and_roundtrip_kat():
    k = KATS["perm_tables"]
    assert sorted(quant.marlin_weight_perm_int4().tolist()) == list(range(k["weight_perm_len"]))
    sp, sp1 = quant.marlin_scale_perms()
    assert sorted(sp.tolist()) == list(range(64)) and sorted(sp1.tolist()) == list(range(32))
    r = KATS["marlin_repack_roundtrip"]
    w = _synth(r["n"] * r["k"]).reshape(r["n"], r["k"])
    p, s = quant.quantize_int4(w, r["group_size"])
    mp, ms = quant.marlin_repack_int4(p, s, r["group_size"])
    assert mp.shape == (r["k"] // 16, 2 * r["n"]) and ms.shape == (r["k"] // r["group_size"], r["n"])
    p2, s2 = quant.marlin_unpack_int4(mp, ms, r["group_size"])
    assert np.array_equal(p, p2) and np.array_equal(s, s2)
    # multi-group / larger shape too (scale_perm path, group_size < K)
    rng = np.random.default_rng(1)
    w = f32_to_bf16_bits(rng.normal(0, 0.02, (128, 512)).astype(np.float32))
    p, s = quant.quantize_int4(w)
    p2, s2 = quant.marlin_unpack_int4(*quant.marlin_repack_int4(p, s))
    assert np.array_equal(p, p2) and np.array_equal(s, s2)


def test_quantiser_scale_rule_and_range():
    # scale = bf16(amax/7) (1.0 if amax == 0); q in [-8, 7]; nibble = q + 8
    w = np.zeros((2, 128), np.float32)
    w[1, 5] = 0.7
    w[1, 9] = -0.7
    p, s = quant.quantize_int4(f32_to_bf16_bits(w))
    assert bf16_bits_to_f32(s)[0, 0] == 1.0
    q = quant.unpack_int4(p)
    assert q[0].max() == 0 and q[0].min() == 0
    assert q[1, 5] == 7 and q[1, 9] == -7 and q.min() >= -8 and q.max() <= 7
    q8, s8 = quant.quantize_int8(f32_to_bf16_bits(w))
    assert q8[1, 5] == 127 and q8[1, 9] == -127


@pytest.mark.parametrize("t", [G.GGML_Q8_0, G.GGML_Q4_0, G.GGML_Q5_0, G.GGML_Q4_K, G.GGML_Q5_K, G.GGML_Q6_K])
def test_gguf_dequant_bit_exact_vs_gguf_py(t):
    gguf = pytest.importorskip("gguf")
    from gguf import quants
    rng = np.random.default_rng(t)
    b = G.random_blocks(rng, t, 8, 1024)
    ours = G.dequantize(t, b.reshape(-1), 8 * 1024)
    ref = quants.dequantize(b, gguf.GGMLQuantizationType(t)).reshape(-1)
    assert np.array_equal(ours, ref)


def test_q6k_reference_dequant_bug_is_visible():
    b = G.random_blocks(np.random.default_rng(3), G.GGML_Q6_K, 2, 256)
    assert not np.array_equal(G.dequant_q6_k(b.reshape(-1), 512), G.dequant_q6_k_gguf_rs(b.reshape(-1), 512))


def test_router_variants_and_tiebreak():
    lg = np.array([[1.0, 3.0, 3.0, 2.0, -1.0]], np.float32)
    ids, w = router.route_from_logits(lg, 3)
    assert ids.tolist() == [[1, 2, 3]]                      # tie -> lower index first (moe.rs:3120)
    sm = np.exp(lg - 3) / np.exp(lg - 3).sum()
    assert np.allclose(w, sm[0, [1, 2, 3]], rtol=1e-6)
    _, wn = router.route_from_logits(lg, 3, norm_topk_prob=True)
    assert abs(wn.sum() - 1) < 1e-6
    ids_b, w_b = router.route_from_logits(lg, 2, scoring_func="sigmoid",
                                          e_score_correction_bias=np.array([10, 0, 0, 0, 0], np.float32))
    assert ids_b.tolist() == [[0, 1]]
    assert np.allclose(w_b[0], 1 / (1 + np.exp(-lg[0, [0, 1]])), rtol=1e-6)   # weights from UNbiased scores
    ids_g, w_g = router.route_from_logits(lg, 2, gpt_oss=True)
    assert ids_g.tolist() == [[1, 2]] and np.allclose(w_g, 0.5)


def test_router_matches_torch_reference_lines():
    # layer.py:532-558 executed literally with torch on the golden inputs
    import torch
    h = torch.from_numpy(VEC["r_hidden"]).to(torch.bfloat16)
    g = torch.from_numpy(VEC["r_gate"]).to(torch.bfloat16)
    logits = torch.matmul(h.float(), g.float().t())
    scores = torch.softmax(logits, dim=-1)
    tw, ti = torch.topk(scores, 6, dim=-1)
    gaps = router.min_topk_gap(logits.numpy(), 6)
    ok = gaps > 1e-5
    assert ok.mean() > 0.9
    assert np.array_equal(ti.numpy()[ok].astype(np.int32), VEC["r_ids"][ok])
    assert np.allclose(tw.numpy()[ok], VEC["r_w"][ok], rtol=2e-6, atol=1e-9)
    twn = tw / tw.sum(dim=-1, keepdim=True)
    assert np.allclose(twn.numpy()[ok], VEC["r_w_norm"][ok], rtol=3e-6)


def test_frozen_oracle_vectors_regression():
    p4, s4 = quant.quantize_int4(VEC["qw"])
    assert np.array_equal(p4, VEC["q4_packed"]) and np.array_equal(s4, VEC["q4_scales"])
    q8, s8 = quant.quantize_int8(VEC["qw"])
    assert np.array_equal(q8, VEC["q8_data"]) and np.array_equal(s8, VEC["q8_scales"])
    layer = omoe.Int4Layer(VEC["m_w13_q"], VEC["m_w13_s"], VEC["m_w2_q"], VEC["m_w2_s"])
    y = omoe.moe_forward_gpu_path(layer, VEC["m_x"], VEC["m_ids"], VEC["m_w"])
    assert np.array_equal(y, VEC["m_y_gpu"])


def test_moe_behavioural_kats():
    # moe.rs:3892-3916: all ids -1 => zeros; identical tokens => identical rows
    layer = omoe.Int4Layer(VEC["m_w13_q"], VEC["m_w13_s"], VEC["m_w2_q"], VEC["m_w2_s"])
    x = np.repeat(VEC["m_x"][:1], 4, axis=0)
    ids = np.repeat(VEC["m_ids"][:1], 4, axis=0)
    w = np.repeat(VEC["m_w"][:1], 4, axis=0)
    for fwd in (omoe.moe_forward_gpu_path, omoe.moe_forward_cpu_int):
        y = fwd(layer, x, ids, w)
        assert np.abs(y - y[0:1]).max() < KATS["moe_behaviour"]["identical_tokens_identical_rows_tol"]
        z = fwd(layer, x, np.full_like(ids, -1), w)
        assert not z.any()


def test_cpu_and_gpu_numerics_agree_loosely():
    # tests/test_gpu_prefill.py:215 only demands cosine > 0.5 GPU-vs-CPU; ours are far closer
    a, b = VEC["m_y_gpu"].reshape(-1).astype(np.float64), VEC["m_y_cpu"].reshape(-1).astype(np.float64)
    cos = a @ b / np.linalg.norm(a) / np.linalg.norm(b)
    assert cos > 0.999


def test_ep_slices_sum_to_full():
    # gpu_prefill.py:353-359,4140-4149 — contiguous expert ranges, non-local => 0
    layer = omoe.Int4Layer(VEC["m_w13_q"], VEC["m_w13_s"], VEC["m_w2_q"], VEC["m_w2_s"])
    x, ids, w = VEC["m_x"], VEC["m_ids"], VEC["m_w"]
    full = omoe.moe_forward_gpu_path(layer, x, ids, w)
    parts = [omoe.moe_forward_gpu_path(layer, x, ids, w, s, s + 2) for s in range(0, 8, 2)]
    # partial sums are each rounded to BF16 before the add (model.py:3195-3211) => 1 ulp-level slack
    tot = np.zeros_like(full)
    for p in parts:
        tot = round_bf16(tot + p)
    assert np.abs(tot - full).max() <= 2 ** -7 * np.abs(full).max()


def test_gguf_cpu_path_close_to_dequant_path():
    rng = np.random.default_rng(5)
    lay = omoe.make_gguf_layer(rng, 2, 256, 256)
    x = round_bf16(rng.normal(0, 1, 256).astype(np.float32))
    y = omoe.expert_forward_cpu_gguf(lay, 0, x)
    w13, w2 = lay.w13_f32(0).astype(np.float64), lay.w2_f32(0).astype(np.float64)
    c1 = w13 @ x.astype(np.float64)
    h = c1[:256] / (1 + np.exp(-c1[:256])) * c1[256:]
    ref = w2 @ h
    assert np.abs(y - ref).max() < 2e-3 * np.abs(ref).max() + 1e-6


# ------------------------------------------------------------------ router pinned on the reference's own compute_routing

ROUTER_G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "router_reference.npz"))
ROUTER_CASES = {   # name: (k, scoring, norm_topk_prob, gpt_oss)   — tests/golden/make_router_golden.py
    "qwen_softmax_norm": (8, "softmax", True, False),
    "deepseek_softmax": (6, "softmax", False, False),
    "kimi_sigmoid_bias_norm": (8, "sigmoid", True, False),
    "gptoss_topk_softmax": (4, "softmax", False, True),
}


@pytest.mark.parametrize("name", sorted(ROUTER_CASES))
def test_router_oracle_matches_reference_execution(name):
    """oracle.router vs outputs of python/krasis/layer.py:compute_routing executed on CPU: ids identical (order included)
    wherever the selection is not a near-tie in fp32, weights to 3e-5 relative (fp32 matmul summation order)."""
    from oracle import router as R
    from oracle.bf16 import bf16_bits_to_f32
    k, scoring, norm, gpt_oss = ROUTER_CASES[name]
    hidden, gate = bf16_bits_to_f32(ROUTER_G[f"{name}_hidden"]), bf16_bits_to_f32(ROUTER_G[f"{name}_gate"])
    gb = ROUTER_G[f"{name}_gate_bias"] if f"{name}_gate_bias" in ROUTER_G.files else None
    cb = bf16_bits_to_f32(ROUTER_G[f"{name}_corr_bias"]) if f"{name}_corr_bias" in ROUTER_G.files else None
    ids, w = R.compute_routing(hidden, gate, k, scoring_func=scoring, norm_topk_prob=norm, e_score_correction_bias=cb,
                               gpt_oss=gpt_oss, gate_bias=gb)
    want_ids, want_w = ROUTER_G[f"{name}_ids"], ROUTER_G[f"{name}_w"]
    same = (ids == want_ids).all(axis=1)
    assert same.mean() >= 0.97, same.mean()                   # near-ties may flip an order in a handful of rows
    assert np.allclose(w[same], want_w[same], rtol=3e-5, atol=1e-7)
    # rows that differ must differ only by a swap of (nearly) equal weights
    for r in np.nonzero(~same)[0]:
        assert sorted(ids[r]) == sorted(want_ids[r]) or np.abs(np.sort(w[r]) - np.sort(want_w[r])).max() < 1e-4


def test_int8_linear_oracle_matches_reference_execution():
    """oracle.dense.quantize_to_int8 / int8_linear vs outputs of the reference's own weight_loader.py functions run on CPU
    (tests/golden/make_int8_golden.py): integer work, bit-exact."""
    import torch
    from oracle import dense as D
    Gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "int8_reference.npz"))
    bf = lambda a: torch.from_numpy(a.view(np.int16)).view(torch.bfloat16)
    x, w = bf(Gd["x"]), bf(Gd["w"])
    wq, ws = D.quantize_to_int8(w)
    assert np.array_equal(wq.numpy(), Gd["wq"]) and np.array_equal(ws.view(torch.int16).numpy().view(np.uint16), Gd["ws"])
    y = D.int8_linear(x, wq, ws)
    assert np.array_equal(y.view(torch.int16).numpy().view(np.uint16), Gd["y"])


def test_division_free_row_quantisation_rule_is_exact():
    """elementwise.cu:quant_code_rhe replaces `round_half_even(x / scale)` (weight_loader.py:25-43, 46-99) by `rint(x * (1 / scale))` and
    takes the true division only when the product lands within 1e-3 of a half-integer.  Every step is a correctly rounded fp32
    operation, so the rule can be checked on the CPU: it must agree with the division for every input, including values constructed to sit
    on rounding boundaries."""
    rng = np.random.default_rng(5)
    f32 = np.float32
    scales = np.concatenate([(rng.random(4000).astype(f32) * f32(3.0) + f32(1e-3)) / f32(127.0),
                             (np.arange(1, 2049, dtype=np.float32) / f32(64.0)) / f32(127.0)]).astype(f32)   # row maxima on the bf16 grid too
    worst = 0.0
    for s in scales:
        mx = s * f32(127.0)
        x = (rng.standard_normal(512).astype(f32) * mx / f32(2.5)).astype(f32)
        x = np.clip(x, -mx, mx)
        bits = x.view(np.uint32)
        x = ((bits + np.uint32(0x7FFF) + ((bits >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).view(f32)   # bf16-valued inputs
        k = rng.integers(-127, 127, 64).astype(f32) + f32(0.5)                                                    # half-integer targets
        x = np.concatenate([x, (k * s).astype(f32), np.nextafter((k * s).astype(f32), f32(np.inf)), np.nextafter((k * s).astype(f32), f32(-np.inf))])
        want = np.rint((x / s).astype(f32))
        y = (x * (f32(1.0) / s)).astype(f32)
        r = np.rint(y)
        near = np.abs(np.abs(y - r) - f32(0.5)) < f32(1e-3)
        got = np.where(near, want, r)
        assert np.array_equal(got, want), (s, x[got != want][:4])
        worst = max(worst, float(np.abs(y - (x / s).astype(f32)).max()))
    assert worst < 1e-4          # the margin the 1e-3 window relies on
