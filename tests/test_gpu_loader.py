"""Device quantiser (bit-exact vs the oracle's restatement of marlin.rs) and the safetensors / GGUF load paths."""
import json
import os
import struct

import numpy as np
import pytest
import torch

from oracle import bf16 as B
from oracle import quant as Q
from oracle import moe as OM
from oracle import gguf_blocks as G

pytestmark = pytest.mark.gpu


def _bf16_tensor(u16: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(u16.view(np.int16)).view(torch.bfloat16)


@pytest.mark.parametrize("bits", [4, 8])
def test_device_quantiser_bit_exact(bits):
    from krasis_b200 import KrasisEngine
    rng = np.random.default_rng(bits)
    rows, K = 257, 1024
    w = rng.normal(0, 0.05, (rows, K)).astype(np.float32)
    w[3, :128] = 0.0                                   # all-zero group -> scale 1.0 path (marlin.rs:176)
    w[5, 128:256] = 1e-30                              # tiny amax
    w[7, 0] = 300.0                                    # large outlier
    # exact half-way cases: w = (n + 0.5) * scale
    w[9, :128] = (np.arange(128) % 15 - 7 + 0.5) * 0.125
    wb = B.f32_to_bf16_bits(w)
    eng = KrasisEngine(hidden_size=256, moe_intermediate_size=128, n_routed_experts=8, num_experts_per_tok=2,
                       num_moe_layers=1, num_bits=bits, max_tokens=64)
    q, s = eng.quantize_group(_bf16_tensor(wb).cuda(), bits)
    torch.cuda.synchronize()
    if bits == 4:
        qr, sr = Q.quantize_int4(wb)
    else:
        qr, sr = Q.quantize_int8(wb)
    assert np.array_equal(s.cpu().numpy().view(np.uint16), sr.view(np.uint16))
    assert np.array_equal(q.cpu().numpy().view(qr.dtype), qr)


def _write_safetensors(path, tensors):
    hdr, blob, off = {}, [], 0
    for name, arr in tensors.items():
        raw = np.ascontiguousarray(arr).tobytes()
        hdr[name] = {"dtype": "BF16", "shape": list(arr.shape), "data_offsets": [off, off + len(raw)]}
        blob.append(raw)
        off += len(raw)
    h = json.dumps(hdr).encode()
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)) + h + b"".join(blob))


@pytest.mark.parametrize("bits", [4, 8])
def test_safetensors_to_moe_matches_oracle(tmp_path, bits):
    """HF BF16 checkpoint -> loader -> device quantiser -> tiles -> MoE forward == oracle on oracle-quantised weights."""
    from krasis_b200 import KrasisEngine
    from krasis_b200.loader import load_experts_from_safetensors
    from tests.test_gpu_moe import assert_close_bf16, bf16_t, to_np
    rng = np.random.default_rng(10 + bits)
    E, H, I, k, M = 8, 256, 128, 2, 96
    qf = Q.quantize_int4 if bits == 4 else Q.quantize_int8
    t, qs = {}, [[], [], [], []]
    for e in range(E):
        w13 = B.f32_to_bf16_bits(rng.normal(0, 0.02, (2 * I, H)).astype(np.float32))
        w2 = B.f32_to_bf16_bits(rng.normal(0, 0.02, (H, I)).astype(np.float32))
        base = f"model.layers.1.mlp.experts.{e}."
        t[base + "gate_proj.weight"], t[base + "up_proj.weight"], t[base + "down_proj.weight"] = w13[:I], w13[I:], w2
        q, sc = qf(w13); qs[0].append(q); qs[1].append(sc)
        q, sc = qf(w2); qs[2].append(q); qs[3].append(sc)
    lay = OM.Int4Layer(*[np.stack(a) for a in qs], bits, 128)
    _write_safetensors(tmp_path / "model.safetensors", t)
    eng = KrasisEngine(hidden_size=H, moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k,
                       num_moe_layers=1, num_bits=bits, max_tokens=M)
    load_experts_from_safetensors(eng, str(tmp_path), first_k_dense=1)
    x = B.round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids = np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int32)
    wts = rng.uniform(0.1, 1, (M, k)).astype(np.float32)
    out = eng.moe_forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(wts).cuda(), routed_only=True)
    assert_close_bf16(to_np(out), OM.moe_forward_gpu_path(lay, x, ids, wts))


def test_gguf_file_to_moe_matches_oracle(tmp_path):
    gguf = pytest.importorskip("gguf")
    from krasis_b200 import KrasisEngine
    from krasis_b200.loader import load_experts_from_gguf
    from tests.test_gpu_moe import assert_close_bf16, bf16_t, to_np
    rng = np.random.default_rng(5)
    E, H, I, k, M = 4, 256, 256, 2, 64
    lay = OM.make_gguf_layer(rng, E, H, I, G.GGML_Q4_K, G.GGML_Q8_0)
    path = str(tmp_path / "m.gguf")
    w = gguf.GGUFWriter(path, "llama")
    w.add_uint32("llama.block_count", 1)
    Qt = gguf.GGMLQuantizationType
    w.add_tensor("blk.0.ffn_gate_exps.weight", lay.gate, raw_shape=lay.gate.shape, raw_dtype=Qt.Q4_K)
    w.add_tensor("blk.0.ffn_up_exps.weight", lay.up, raw_shape=lay.up.shape, raw_dtype=Qt.Q4_K)
    w.add_tensor("blk.0.ffn_down_exps.weight", lay.down, raw_shape=lay.down.shape, raw_dtype=Qt.Q8_0)
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    eng = KrasisEngine(hidden_size=H, moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k,
                       num_moe_layers=1, gguf_gate_up_type="Q4_K", gguf_down_type="Q8_0", max_tokens=M)
    load_experts_from_gguf(eng, path)
    x = B.round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32))
    ids = np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int32)
    wts = rng.uniform(0.1, 1, (M, k)).astype(np.float32)
    out = eng.moe_forward(0, bf16_t(x), torch.from_numpy(ids).cuda(), torch.from_numpy(wts).cuda(), routed_only=True)
    assert_close_bf16(to_np(out), OM.moe_forward_gpu_path(lay, x, ids, wts), ulps=3)


def test_marlin_cache_to_moe_is_bit_identical_to_direct_load(tmp_path):
    """Reference GPU cache file (Marlin order) -> reader -> tiles gives the same MoE output, bit for bit, as loading the
    quantiser's arrays directly."""
    import struct
    from krasis_b200 import KrasisEngine, QuantizedExperts
    from krasis_b200 import marlin_cache as MC
    from tests.test_gpu_moe import bf16_t
    rng = np.random.default_rng(21)
    E, H, I, k, M, gs = 4, 256, 128, 2, 80, 128
    lay = OM.make_int_layer(rng, E, H, I, 4)
    body = []
    for e in range(E):
        for q, s in ((lay.w13_q[e], lay.w13_s[e]), (lay.w2_q[e], lay.w2_s[e])):
            mp, ms = Q.marlin_repack_int4(q, s, gs)
            body += [np.ascontiguousarray(mp).tobytes(), np.ascontiguousarray(ms).tobytes()]
    path = tmp_path / "experts_marlin_int4_g128.bin"
    path.write_bytes(b"KRAS" + struct.pack("<I", 3) + struct.pack("<7Q", H, I, E, 1, gs, MC.fnv1a(b"{}"), 0) + b"".join(body))
    mk = lambda: KrasisEngine(hidden_size=H, moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k,
                              num_moe_layers=1, num_bits=4, max_tokens=M)
    a, b = mk(), mk()
    MC.load_experts_from_marlin_cache(a, str(path), b"{}")
    b.load_quantized_layer(0, QuantizedExperts(lay.w13_q, lay.w13_s, lay.w2_q, lay.w2_s))
    x = bf16_t(B.round_bf16(rng.normal(0, 1, (M, H)).astype(np.float32)))
    ids = torch.from_numpy(np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int32)).cuda()
    wts = torch.from_numpy(rng.uniform(0.1, 1, (M, k)).astype(np.float32)).cuda()
    assert torch.equal(a.moe_forward(0, x, ids, wts, routed_only=True), b.moe_forward(0, x, ids, wts, routed_only=True))


def test_tile_cache_round_trip_is_bit_identical(tmp_path):
    """Quantise + re-tile once, write the KB2 tile cache, load it into a fresh engine (and into the two ranks of an EP pair):
    the tiles and the MoE output are bit-identical and nothing is re-quantised; a wrong config hash / geometry / size is refused."""
    from krasis_b200 import KrasisEngine, tile_cache as T
    rng = np.random.default_rng(5)
    E, H, I, k, M, L = 8, 256, 128, 2, 64, 2
    kw = dict(hidden_size=H, moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k, num_moe_layers=L, max_tokens=M)
    a = KrasisEngine(**kw)
    g = torch.Generator(device="cuda").manual_seed(3)
    for l in range(L):
        a.load_bf16_layer(l, (torch.randn(E, 2 * I, H, device="cuda", generator=g) * 0.05).to(torch.bfloat16),
                          (torch.randn(E, H, I, device="cuda", generator=g) * 0.05).to(torch.bfloat16))
    cfg_json = b'{"hidden_size": 256}'
    path = str(tmp_path / T.cache_file_name(4))
    size = T.write_tile_cache(a, path, cfg_json, n_shared_experts=0)
    assert size == T.expected_size(a) == os.path.getsize(path)
    x = torch.randn(M, H, device="cuda", generator=g).to(torch.bfloat16)
    ids = torch.from_numpy(np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int32)).cuda()
    w = torch.from_numpy(rng.dirichlet(np.ones(k), M).astype(np.float32)).cuda()
    b = KrasisEngine(**kw)
    T.load_tile_cache(b, path, cfg_json)
    for l in range(L):
        assert torch.equal(a.moe_forward(l, x, ids, w), b.moe_forward(l, x, ids, w))
    parts = []
    for r in range(2):                                           # an EP rank reads only its expert slice
        e = KrasisEngine(**kw, rank=r, num_ranks=2)
        T.load_tile_cache(e, path, cfg_json)
        parts.append(e.moe_forward(1, x, ids, w, routed_only=True).float())
    full = a.moe_forward(1, x, ids, w, routed_only=True).float()
    assert (parts[0] + parts[1] - full).abs().max().item() <= 2 ** -7 * full.abs().max().item()
    with pytest.raises(ValueError):
        T.load_tile_cache(b, path, b'{"hidden_size": 257}')      # config.json changed
    c = KrasisEngine(**{**kw, "num_moe_layers": 3})
    with pytest.raises(ValueError):
        T.load_tile_cache(c, path, cfg_json)                      # geometry mismatch
    with open(path, "ab") as f:
        f.write(b"\0")
    with pytest.raises(ValueError):
        T.load_tile_cache(b, path, cfg_json)                      # size mismatch


def test_reference_engine_load_and_marlin_getters(tmp_path):
    """The reference's two-step start-up — KrasisEngine() then load(model_dir, ...) (src/moe.rs:1482,1538) — on a real-format
    checkpoint directory, the KB2 tile cache it leaves behind, and the Marlin-order hand-off (get_expert_*, src/moe.rs:1972-2097):
    the returned bytes are the oracle's quantiser output pushed through the oracle's marlin_repack, bit for bit."""
    from krasis_b200 import KrasisEngine, tile_cache as T
    from tests.test_gpu_pretrained import build_v2lite_checkpoint
    from tests.test_loader_cpu import _write_safetensors
    hf, t, W = build_v2lite_checkpoint()
    json.dump(hf, open(tmp_path / "config.json", "w"))
    _write_safetensors(tmp_path / "model.safetensors", t)
    cdir = str(tmp_path / "cache")
    eng = KrasisEngine(parallel=True, num_threads=None, skip_shared_experts=False)
    eng.load(str(tmp_path), max_tokens=64, cache_dir=cdir)
    E, I, H = hf["n_routed_experts"], hf["moe_intermediate_size"], hf["hidden_size"]
    assert (eng.hidden_size(), eng.intermediate_size(), eng.num_experts(), eng.top_k(), eng.num_moe_layers()) == (H, I, E, 2, 2)
    assert os.path.exists(os.path.join(cdir, T.cache_file_name(4)))
    w13, w2 = W["layers"][2]["experts"]                          # absolute layer 2 = moe layer 1 (first_k_dense_replace = 1)
    for e in (0, E - 1):
        q, s = Q.quantize_int4(w13[e].view(torch.int16).numpy().view(np.uint16))
        mp, ms = Q.marlin_repack_int4(q, s)
        assert eng.get_expert_w13_packed(1, e, e + 1) == mp.tobytes()
        assert eng.get_expert_w13_scales(1, e, e + 1) == np.ascontiguousarray(ms).view(np.uint16).tobytes()
        q, s = Q.quantize_int4(w2[e].view(torch.int16).numpy().view(np.uint16))
        mp, ms = Q.marlin_repack_int4(q, s)
        assert eng.get_expert_w2_packed(1, e, e + 1) == mp.tobytes()
        assert eng.get_expert_w2_scales(1, e, e + 1) == np.ascontiguousarray(ms).view(np.uint16).tobytes()
    assert len(eng.get_expert_w13_packed(0)) == E * (H // 16) * (2 * 2 * I) * 4          # [K/16, 2N] u32 per expert (mod.rs:955-970)
    buf = [np.zeros(len(eng.get_expert_w13_packed(1, 2, 5)), np.uint8), np.zeros(len(eng.get_expert_w13_scales(1, 2, 5)), np.uint8),
           np.zeros(len(eng.get_expert_w2_packed(1, 2, 5)), np.uint8), np.zeros(len(eng.get_expert_w2_scales(1, 2, 5)), np.uint8)]
    args = [v for b in buf for v in (b.ctypes.data, b.size)]
    eng.write_experts_range_into_pinned(1, 2, 5, *args)
    assert buf[0].tobytes() == eng.get_expert_w13_packed(1, 2, 5) and buf[3].tobytes() == eng.get_expert_w2_scales(1, 2, 5)
    with pytest.raises(ValueError):
        eng.write_experts_range_into_pinned(1, 2, 5, buf[0].ctypes.data, buf[0].size - 4, *args[2:])
    # second start: the cache is used (same tiles, bit-identical MoE output) and a stale cache is rebuilt, not trusted
    eng2 = KrasisEngine()
    eng2.load(str(tmp_path), max_tokens=64, cache_dir=cdir)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(64, H, device="cuda", generator=g).to(torch.bfloat16)
    ids = torch.randint(0, E, (64, 2), device="cuda", generator=g, dtype=torch.int32)
    ids[:, 1] = (ids[:, 0] + 1) % E
    w = torch.rand(64, 2, device="cuda", generator=g)
    assert torch.equal(eng.moe_forward(1, x, ids, w), eng2.moe_forward(1, x, ids, w))
