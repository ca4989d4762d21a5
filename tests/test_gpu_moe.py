"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Integer / index work (re-tiling, binning, top-k ids) must be bit-exact; the BF16 MoE output must be
within 2 BF16 ulp of the row maximum and cosine >= 0.9999 of the oracle.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import quant, router, moe as omoe  # noqa: E402
from oracle.bf16 import f32_to_bf16_bits, round_bf16  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))


def bf16_t(x_f32: np.ndarray, device="cuda"):
    return torch.from_numpy(np.ascontiguousarray(x_f32, dtype=np.float32)).to(torch.bfloat16).to(device)


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.float().cpu().numpy()


def make_engine(layer, top_k, M, num_bits=4, rank=0, num_ranks=1, E_global=None, **kw):
    from krasis_b200 import KrasisEngine, QuantizedExperts
    E = E_global or layer.E
    eng = KrasisEngine(hidden_size=layer.H, moe_intermediate_size=layer.I, n_routed_experts=E,
                       num_experts_per_tok=top_k, num_moe_layers=1, num_bits=num_bits, rank=rank,
                       num_ranks=num_ranks, max_tokens=max(M, 1), **kw)
    s, t = eng.expert_start, eng.expert_end
    eng.load_quantized_layer(0, QuantizedExperts(layer.w13_q[s:t], layer.w13_s[s:t], layer.w2_q[s:t], layer.w2_s[s:t]))
    return eng


def assert_close_bf16(got, want, ulps=2, cos_min=0.9999):
    got, want = got.astype(np.float64), want.astype(np.float64)
    rowmax = np.abs(want).max(axis=1, keepdims=True)
    tol = ulps * rowmax * 2.0 ** -8 + 1e-30
    bad = np.abs(got - want) > tol
    assert not bad.any(), f"{bad.sum()} elements beyond {ulps} bf16 ulp of row max; worst {np.abs(got - want).max()}"
    if np.linalg.norm(want) > 0:
        cos = (got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want))
        assert cos >= cos_min, cos


# ------------------------------------------------------------------ layout transform (bit-exact)

def tile_int4_numpy(packed, scales):
    """numpy restatement of the KB2 tile layout (krasis_b200/csrc/moe_common.cuh)."""
    E, N, KW = packed.shape
    K = KW * 8
    nib = (packed[..., None] >> (np.arange(8, dtype=np.uint32) * 4)) & 0xF           # [E,N,KW,8]
    order = np.array([0, 2, 4, 6, 1, 3, 5, 7])
    w = np.bitwise_or.reduce(nib[..., order] << (np.arange(8, dtype=np.uint32) * 4), axis=-1).astype(np.uint32)
    # [E, tile, row, kb, h, j] -> [E, tile, kb, h, row, j]
    w = w.reshape(E, N // 128, 128, K // 64, 2, 4).transpose(0, 1, 3, 4, 2, 5)
    s = scales.reshape(E, N // 128, 128, K // 128).transpose(0, 1, 3, 2)
    return np.ascontiguousarray(w).reshape(-1), np.ascontiguousarray(s).reshape(-1)


def test_retile_int4_bit_exact():
    from krasis_b200 import KrasisEngine
    rng = np.random.default_rng(0)
    E, N, K = 3, 256, 384
    packed = rng.integers(0, 2 ** 32, (E, N, K // 8), dtype=np.uint64).astype(np.uint32)
    scales = rng.integers(0, 2 ** 16, (E, N, K // 128), dtype=np.uint32).astype(np.uint16)
    eng = KrasisEngine(hidden_size=256, moe_intermediate_size=128, n_routed_experts=4, num_experts_per_tok=2,
                       num_moe_layers=1, max_tokens=8)
    sq = torch.from_numpy(packed.view(np.int32)).cuda()
    ss = torch.from_numpy(scales.view(np.int16)).cuda()
    dq, ds = torch.empty_like(sq), torch.empty_like(ss)
    from krasis_b200 import capi
    capi.check(eng._lib.kb2_retile_dev(eng._h, capi.FMT_INT4_G128, sq.data_ptr(), ss.data_ptr(), dq.data_ptr(),
                                       ds.data_ptr(), E, N, K, None))
    torch.cuda.synchronize()
    wq, ws = tile_int4_numpy(packed, scales)
    assert np.array_equal(dq.cpu().numpy().view(np.uint32).reshape(-1), wq)
    assert np.array_equal(ds.cpu().numpy().view(np.uint16).reshape(-1), ws)


# ------------------------------------------------------------------ router

@pytest.mark.parametrize("E,k,H,M,norm", [(64, 6, 256, 200, False), (512, 10, 512, 333, True), (256, 8, 256, 64, True),
                                          (96, 4, 256, 130, True),      # E % 64 != 0 -> fp32-FMA fallback kernel
                                          (384, 8, 2048, 1000, True)])  # two accumulators (256 + 128), M % 128 != 0
def test_router_ids_bit_exact_and_weights_close(E, k, H, M, norm):
    from krasis_b200 import KrasisEngine
    rng = np.random.default_rng(E + k)
    hid = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    gate = round_bf16(rng.normal(0, 0.02, (E, H)).astype(np.float32))
    eng = KrasisEngine(hidden_size=H, moe_intermediate_size=128, n_routed_experts=E, num_experts_per_tok=k,
                       num_moe_layers=1, max_tokens=M, norm_topk_prob=norm)
    eng.set_routing_weights(0, f32_to_bf16_bits(gate))
    ids, w = eng.compute_routing(0, bf16_t(hid))
    lg = router.router_logits(hid, gate)
    ids_o, w_o = router.route_from_logits(lg, k, norm_topk_prob=norm)
    ok = router.min_topk_gap(lg, k) > 1e-5          # documented near-tie policy
    assert ok.mean() > 0.95
    assert np.array_equal(ids.cpu().numpy()[ok], ids_o[ok]), "top-k ids must be bit-exact away from near-ties"
    # weights: fp32 softmax of logits whose fp32 accumulation order differs (tensor-core partial sums vs the
    # oracle's single rounding of an exact sum): |d logit| <~ 2e-6 at H=2048 => relative weight error <~ 1e-5
    rel = np.abs(w.cpu().numpy()[ok] - w_o[ok]) / w_o[ok]
    assert rel.max() < 3e-5, f"max relative routing-weight error {rel.max():.3e}"


def test_router_exact_ties_go_to_lower_index():
    from krasis_b200 import KrasisEngine
    E, H, k = 64, 256, 4
    gate = np.zeros((E, H), np.float32)
    gate[[5, 9, 40], 0] = 1.0                        # three experts tie exactly; the rest tie at 0
    hid = np.zeros((3, H), np.float32)
    hid[:, 0] = [1.0, 2.0, 0.5]
    eng = KrasisEngine(hidden_size=H, moe_intermediate_size=128, n_routed_experts=E, num_experts_per_tok=k,
                       num_moe_layers=1, max_tokens=8)
    eng.set_routing_weights(0, f32_to_bf16_bits(gate))
    ids, _ = eng.compute_routing(0, bf16_t(hid))
    assert ids.cpu().numpy().tolist() == [[5, 9, 40, 0]] * 3


def test_router_sigmoid_with_selection_bias():
    from krasis_b200 import KrasisEngine
    rng = np.random.default_rng(7)
    E, H, k, M = 128, 256, 8, 50
    hid = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    gate = round_bf16(rng.normal(0, 0.05, (E, H)).astype(np.float32))
    cb = rng.normal(0, 0.1, E).astype(np.float32)
    eng = KrasisEngine(hidden_size=H, moe_intermediate_size=128, n_routed_experts=E, num_experts_per_tok=k,
                       num_moe_layers=1, max_tokens=M, scoring_func="sigmoid", norm_topk_prob=True)
    eng.set_routing_weights(0, f32_to_bf16_bits(gate), e_score_correction_bias=cb)
    ids, w = eng.compute_routing(0, bf16_t(hid))
    ids_o, w_o = router.compute_routing(hid, gate, k, scoring_func="sigmoid", norm_topk_prob=True,
                                        e_score_correction_bias=cb)
    same = (ids.cpu().numpy() == ids_o).all(axis=1)
    assert same.mean() > 0.95                       # near-ties on the biased sigmoid are allowed to differ
    assert np.allclose(w.cpu().numpy()[same], w_o[same], rtol=1e-5)


# ------------------------------------------------------------------ MoE forward vs oracle

def test_moe_golden_fixture():
    layer = omoe.Int4Layer(VEC["m_w13_q"], VEC["m_w13_s"], VEC["m_w2_q"], VEC["m_w2_s"])
    eng = make_engine(layer, 2, 24)
    out = eng.moe_forward(0, bf16_t(VEC["m_x"]), torch.from_numpy(VEC["m_ids"]).cuda(),
                          torch.from_numpy(VEC["m_w"]).cuda(), routed_only=True)
    assert_close_bf16(to_np(out), VEC["m_y_gpu"])


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("E,H,I,k,M,skew", [(8, 256, 128, 2, 1, 0), (8, 256, 128, 2, 37, 0),
                                            (16, 512, 256, 4, 300, 3.0), (4, 256, 256, 2, 700, 0)])
def test_moe_forward_matches_oracle(bits, E, H, I, k, M, skew):
    rng = np.random.default_rng(E * 1000 + M + bits)
    layer = omoe.make_int_layer(rng, E, H, I, bits=bits)
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    logits = rng.normal(0, 1, (M, E)).astype(np.float32)
    logits[:, 0] += skew                             # skew => expert 0 gets > 256 tokens (multi-chunk)
    ids, w = router.route_from_logits(logits, k, norm_topk_prob=True)
    eng = make_engine(layer, k, M, num_bits=bits)
    out = eng.moe_forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda(), routed_only=True)
    want = omoe.moe_forward_gpu_path(layer, x, ids, w)
    counts = eng.last_expert_counts()
    assert np.array_equal(counts, np.bincount(ids.reshape(-1), minlength=E))      # binning is integer-exact
    assert_close_bf16(to_np(out), want)


@pytest.mark.parametrize("gu,dn", [("Q4_K", "Q8_0"), ("Q8_0", "Q8_0"), ("Q4_K", "Q4_K"),
                                   ("Q4_K", "Q6_K"),      # the Q4_K_M pairing (src/weights/mod.rs:646-647)
                                   ("Q5_K", "Q5_0"), ("Q4_0", "Q4_0"), ("Q6_K", "Q5_K")])
def test_moe_forward_native_gguf_blocks(gu, dn):
    """GGUF expert tensors consumed as native blocks on the GPU (north_star: Q4_K / Q8_0; Q6_K / Q5_K / Q5_0 / Q4_0 are decoded
    losslessly into int8 codes + per-16 affine pairs at load time).  The in-kernel dequant reproduces src/gguf.rs's single f32
    rounding (W = bf16(dequant)), so the usual MoE tolerance applies."""
    from krasis_b200 import KrasisEngine
    from oracle import gguf_blocks as G
    T = {"Q4_K": G.GGML_Q4_K, "Q8_0": G.GGML_Q8_0, "Q6_K": G.GGML_Q6_K, "Q5_K": G.GGML_Q5_K, "Q5_0": G.GGML_Q5_0, "Q4_0": G.GGML_Q4_0}
    rng = np.random.default_rng(31)
    E, H, I, k, M = 6, 512, 256, 2, 230
    lay = omoe.make_gguf_layer(rng, E, H, I, gate_up_type=T[gu], down_type=T[dn])
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids, w = router.route_from_logits(rng.normal(0, 1, (M, E)).astype(np.float32), k, norm_topk_prob=True)
    eng = KrasisEngine(hidden_size=H, moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k, num_moe_layers=1,
                       max_tokens=M, gguf_gate_up_type=gu, gguf_down_type=dn)
    eng.load_gguf_layer(0, lay.gate, lay.up, lay.down)
    out = eng.moe_forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda(), routed_only=True)
    # random (not quantised-from-real-weights) Q4_K blocks give heavy cancellation in the dot products, so allow 3 ulp
    assert_close_bf16(to_np(out), omoe.moe_forward_gpu_path(lay, x, ids, w), ulps=3)


def test_moe_scaling_and_shared_add():
    rng = np.random.default_rng(11)
    layer = omoe.make_int_layer(rng, 8, 256, 128)
    M, k = 40, 2
    x = round_bf16(rng.normal(0, 1, (M, 256)).astype(np.float32))
    ids, w = router.route_from_logits(rng.normal(0, 1, (M, 8)).astype(np.float32), k)
    shared = round_bf16(rng.normal(0, 0.1, (M, 256)).astype(np.float32))
    eng = make_engine(layer, k, M, routed_scaling_factor=2.5)
    routed = omoe.moe_forward_gpu_path(layer, x, ids, w)
    tx, ti, tw = bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda()
    out = eng.moe_forward(0, tx, ti, tw, shared=bf16_t(shared))
    assert_close_bf16(to_np(out), omoe.finish_gpu_path(routed, 2.5, shared))
    out2 = eng.moe_forward(0, tx, ti, tw)
    assert_close_bf16(to_np(out2), omoe.finish_gpu_path(routed, 2.5))


def test_moe_behavioural_kats():
    """src/moe.rs:3892-3916: all ids -1 => zeros; identical tokens => identical rows (here bit-identical)."""
    rng = np.random.default_rng(12)
    layer = omoe.make_int_layer(rng, 8, 256, 128)
    eng = make_engine(layer, 2, 64)
    x1 = round_bf16(rng.normal(0, 1, (1, 256)).astype(np.float32))
    x = np.repeat(x1, 64, axis=0)
    ids = np.tile(np.array([[3, 6]], np.int32), (64, 1))
    w = np.tile(np.array([[0.7, 0.3]], np.float32), (64, 1))
    out = to_np(eng.moe_forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda(), routed_only=True))
    assert np.abs(out).max() > 1e-4                                      # RMS sanity (tests/test_gpu_prefill.py)
    assert np.array_equal(out, np.repeat(out[:1], 64, axis=0))
    z = eng.moe_forward(0, bf16_t(x), torch.full((64, 2), -1, dtype=torch.int32).cuda(),
                        torch.from_numpy(w).cuda(), routed_only=True)
    assert not to_np(z).any()


def test_ep_slices_sum_to_single_engine():
    """gpu_prefill.py:353-359,4140-4149: each rank computes its expert slice, partial sums add up."""
    rng = np.random.default_rng(13)
    E, H, I, k, M = 12, 256, 128, 3, 90
    layer = omoe.make_int_layer(rng, E, H, I)
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids, w = router.route_from_logits(rng.normal(0, 1, (M, E)).astype(np.float32), k, norm_topk_prob=True)
    tx, ti, tw = bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda()
    full = to_np(make_engine(layer, k, M).moe_forward(0, tx, ti, tw, routed_only=True))
    R = 5                                                # 12 experts / 5 ranks: last rank takes the remainder
    ranges, parts = [], []
    for r in range(R):
        eng = make_engine(layer, k, M, rank=r, num_ranks=R)
        ranges.append((eng.expert_start, eng.expert_end))
        part = to_np(eng.moe_forward(0, tx, ti, tw, routed_only=True))
        assert_close_bf16(part, omoe.moe_forward_gpu_path(layer, x, ids, w, *ranges[-1]))
        parts.append(part)
    assert ranges == [(0, 2), (2, 4), (4, 6), (6, 8), (8, 12)]
    assert np.abs(np.sum(parts, axis=0) - full).max() <= 4 * 2.0 ** -8 * np.abs(full).max()


def test_token_permutation_equivariance_bit_exact():
    rng = np.random.default_rng(14)
    E, H, I, k, M = 16, 512, 256, 4, 257
    layer = omoe.make_int_layer(rng, E, H, I)
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids, w = router.route_from_logits(rng.normal(0, 1, (M, E)).astype(np.float32), k)
    eng = make_engine(layer, k, M)
    perm = rng.permutation(M)
    a = to_np(eng.moe_forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda(), routed_only=True))
    b = to_np(eng.moe_forward(0, bf16_t(x[perm]), torch.from_numpy(ids[perm]).cuda(),
                              torch.from_numpy(w[perm]).cuda(), routed_only=True))
    assert np.array_equal(a[perm], b)


def test_host_buffer_entry_point_routes_and_matches():
    rng = np.random.default_rng(15)
    E, H, I, k, M = 8, 256, 128, 2, 50
    layer = omoe.make_int_layer(rng, E, H, I)
    gate = round_bf16(rng.normal(0, 0.05, (E, H)).astype(np.float32))
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    eng = make_engine(layer, k, M, norm_topk_prob=True)
    eng.set_routing_weights(0, f32_to_bf16_bits(gate))
    xh = torch.from_numpy(x).to(torch.bfloat16).pin_memory()
    out = eng.moe_forward_host(0, xh, routed_only=True)
    ids, w = router.compute_routing(x, gate, k, norm_topk_prob=True)
    assert_close_bf16(out.float().numpy(), omoe.moe_forward_gpu_path(layer, x, ids, w))


def test_error_behaviour_mirrors_reference():
    from krasis_b200 import KrasisEngine
    from krasis_b200.capi import Kb2Error
    eng = KrasisEngine(hidden_size=256, moe_intermediate_size=128, n_routed_experts=8, num_experts_per_tok=2,
                       num_moe_layers=2, max_tokens=16)
    x = torch.zeros(4, 256, dtype=torch.bfloat16, device="cuda")
    ids = torch.zeros(4, 2, dtype=torch.int32, device="cuda")
    w = torch.zeros(4, 2, dtype=torch.float32, device="cuda")
    with pytest.raises(Kb2Error, match="GPU weights not available"):       # PyRuntimeError in the reference
        eng.moe_forward(1, x, ids, w)
    with pytest.raises(ValueError):                                         # PyValueError: bad layer index
        eng.moe_forward(5, x, ids, w)
    with pytest.raises(ValueError):
        eng.moe_forward(0, x[:, :128].contiguous(), ids, w)
    with pytest.raises(ValueError):                                         # more tokens than max_tokens
        eng.moe_forward(0, torch.zeros(17, 256, dtype=torch.bfloat16, device="cuda"),
                        torch.zeros(17, 2, dtype=torch.int32, device="cuda"),
                        torch.zeros(17, 2, dtype=torch.float32, device="cuda"))
    with pytest.raises(Kb2Error, match="router weights not set"):
        eng.compute_routing(0, x)


# ------------------------------------------------------------------ full-size (BASELINE config) properties

def _random_tiled_layer(eng, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ts = []
    for which in range(4):
        n = eng.tiled_bytes(which)
        if which in (0, 2):
            ts.append(torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g))
        else:   # bf16 scales in [0.005, 0.05] like src/decode.rs:4379-4392
            s = torch.rand(n // 2, device="cuda", generator=g) * 0.045 + 0.005
            ts.append(s.to(torch.bfloat16))
    return ts


def untile_expert_int4(wq, ws, e, N, K):
    """Inverse of the KB2 layout for ONE expert (numpy, test-only) -> reference-format (packed, scales)."""
    per_q, per_s = N * K // 2, N * (K // 128) * 2
    q = wq[e * per_q:(e + 1) * per_q].view(np.uint32).reshape(N // 128, K // 64, 2, 128, 4)
    q = q.transpose(0, 3, 1, 2, 4).reshape(N, K // 8)                       # [row][kb,h,j]
    nib = (q[..., None] >> (np.arange(8, dtype=np.uint32) * 4)) & 0xF
    inv = np.argsort(np.array([0, 2, 4, 6, 1, 3, 5, 7]))
    packed = np.bitwise_or.reduce(nib[..., inv] << (np.arange(8, dtype=np.uint32) * 4), axis=-1).astype(np.uint32)
    s = ws[e * per_s:(e + 1) * per_s].view(np.uint16).reshape(N // 128, K // 128, 128).transpose(0, 2, 1).reshape(N, K // 128)
    return packed, np.ascontiguousarray(s)


def test_qcn_full_size_sampled_tokens_match_oracle():
    """BASELINE config C4 (Qwen3-Coder-Next: H2048 I512 E512 k10, M=8192), random packed weights:
    check a sample of tokens against the oracle + size-independent properties."""
    from krasis_b200 import KrasisEngine
    H, I, E, k, M = 2048, 512, 512, 10, 8192
    eng = KrasisEngine(hidden_size=H, moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k,
                       num_moe_layers=1, max_tokens=M, norm_topk_prob=True)
    ts = _random_tiled_layer(eng, 42)
    eng.attach_tiled_layer(0, *ts)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(M, H, device="cuda", generator=g)
    x = (x / x.pow(2).mean(-1, keepdim=True).sqrt()).to(torch.bfloat16)
    gate = (torch.randn(E, H, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    eng.set_routing_weights(0, gate)
    ids, w = eng.compute_routing(0, x)
    out = eng.moe_forward(0, x, ids, w, routed_only=True)
    torch.cuda.synchronize()
    ids_n, w_n = ids.cpu().numpy(), w.cpu().numpy()
    counts = eng.last_expert_counts()
    assert counts.sum() == M * k and np.array_equal(counts, np.bincount(ids_n.reshape(-1), minlength=E))
    assert (np.sort(ids_n, axis=1)[:, 1:] != np.sort(ids_n, axis=1)[:, :-1]).all()       # k distinct experts per token
    assert np.allclose(w_n.sum(axis=1), 1.0, atol=1e-5)
    # sampled tokens vs oracle
    wq13, ws13 = ts[0].cpu().numpy(), ts[1].view(torch.int16).cpu().numpy().view(np.uint8)
    wq2, ws2 = ts[2].cpu().numpy(), ts[3].view(torch.int16).cpu().numpy().view(np.uint8)
    xs = x.float().cpu().numpy()
    sample = [0, 1, 4095, 8191, 1234]
    used = sorted(set(ids_n[sample].reshape(-1).tolist()))
    w13q = np.zeros((E, 2 * I, H // 8), np.uint32); w13s = np.zeros((E, 2 * I, H // 128), np.uint16)
    w2q = np.zeros((E, H, I // 8), np.uint32); w2s = np.zeros((E, H, I // 128), np.uint16)
    for e in used:
        w13q[e], w13s[e] = untile_expert_int4(wq13, ws13, e, 2 * I, H)
        w2q[e], w2s[e] = untile_expert_int4(wq2, ws2, e, H, I)
    layer = omoe.Int4Layer(w13q, w13s, w2q, w2s)
    want = omoe.moe_forward_gpu_path(layer, xs[sample], ids_n[sample], w_n[sample])
    assert_close_bf16(to_np(out)[sample], want)
    # determinism: a second run is bit-identical (slot order inside an expert may differ, results may not)
    out2 = eng.moe_forward(0, x, ids, w, routed_only=True)
    assert torch.equal(out, out2)


def test_manager_owned_int4_shared_expert():
    """Models without a shared-expert gate on one GPU (DeepSeek-V2-Lite): the manager runs the fused shared experts as a
    one-expert INT4 MoE with weight 1 and returns bf16(rsf * routed) + shared (gpu_prefill.py:4738-4801,4471-4480)."""
    from krasis_b200 import GpuPrefillManager, QuantizedExperts
    rng = np.random.default_rng(31)
    E, H, I, k, M, n_sh, rsf = 8, 256, 128, 2, 100, 2, 2.5
    lay = omoe.make_int_layer(rng, E, H, I, 4)
    sh = omoe.make_int_layer(rng, 1, H, n_sh * I, 4)
    mgr = GpuPrefillManager(device="cuda:0", num_experts=E, hidden_size=H, intermediate_size=I, n_shared_experts=n_sh,
                            routed_scaling_factor=rsf, num_bits=4, num_moe_layers=1, top_k=k, max_tokens=M)
    mgr._engine.load_quantized_layer(0, QuantizedExperts(lay.w13_q, lay.w13_s, lay.w2_q, lay.w2_s))
    mgr.load_shared_expert(0, QuantizedExperts(sh.w13_q, sh.w13_s, sh.w2_q, sh.w2_s))
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids = np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int32)
    w = rng.uniform(0.1, 1, (M, k)).astype(np.float32)
    out = mgr.forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda())
    routed = omoe.moe_forward_gpu_path(lay, x, ids, w)
    shared = omoe.moe_forward_gpu_path(sh, x, np.zeros((M, 1), np.int32), np.ones((M, 1), np.float32))
    assert_close_bf16(to_np(out), omoe.finish_gpu_path(routed, rsf, shared), ulps=3)
    ro = mgr.forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda(), routed_only=True)
    assert_close_bf16(to_np(ro), routed)


@pytest.mark.parametrize("bits", [4, 8])
def test_gather4_token_fetch_is_bit_identical_to_the_sorted_copy(bits, monkeypatch):
    """KB2_MOE_GATHER=1: the gate/up GEMM reads token rows straight from the activation matrix with TMA gather4 (row indices
    from the binning pass) instead of from the x_sorted copy.  Same tiles, same arithmetic: outputs must be bit-identical,
    including ragged experts (1 token, 17 tokens, > 192 tokens -> several chunks) and skipped ids."""
    rng = np.random.default_rng(77)
    E, H, I, k, M = 8, 512, 256, 2, 700
    layer = omoe.make_int_layer(rng, E, H, I, bits)
    x = round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids = np.stack([rng.choice(6, k, replace=False) for _ in range(M)]).astype(np.int32)       # experts 0-5 busy (> 192 tokens each)
    ids[:17, 0] = 6                                                                              # 17 tokens on expert 6
    ids[:17, 1] = 0
    ids[40, 1] = 7                                                                               # a single token on expert 7
    ids[41, 1] = -1
    w = rng.dirichlet(np.ones(k), M).astype(np.float32)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("KB2_MOE_GATHER", mode)
        eng = make_engine(layer, k, M, num_bits=bits)
        outs[mode] = eng.moe_forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(w).cuda(), routed_only=True)
        torch.cuda.synchronize()
    assert torch.equal(outs["0"], outs["1"])
    assert_close_bf16(to_np(outs["1"]), omoe.moe_forward_gpu_path(layer, x, ids, w))
