"""Host-side loaders: safetensors reader + expert naming, GGUF v3 parser + expert slicing (CPU only).
The GGUF file is written with the independent `gguf` python package and read back with our parser."""
import json
import os
import struct

import numpy as np
import pytest

from krasis_b200 import loader as Ld
from oracle import gguf_blocks as G


def _write_safetensors(path, tensors):
    hdr, blob, off = {}, b"", 0
    for name, (dtype, arr) in tensors.items():
        raw = np.ascontiguousarray(arr).tobytes()
        hdr[name] = {"dtype": dtype, "shape": list(arr.shape), "data_offsets": [off, off + len(raw)]}
        blob += raw
        off += len(raw)
    h = json.dumps(hdr).encode()
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)) + h + blob)


def test_safetensors_reader_and_expert_naming(tmp_path):
    rng = np.random.default_rng(0)
    E, H, I, L = 3, 128, 128, 2
    t = {}
    ref = {}
    for l in range(L):
        for e in range(E):
            for proj, shp in (("gate_proj", (I, H)), ("up_proj", (I, H)), ("down_proj", (H, I))):
                a = rng.integers(0, 2 ** 16, shp, dtype=np.uint32).astype(np.uint16)
                t[f"model.layers.{l + 1}.mlp.experts.{e}.{proj}.weight"] = ("BF16", a)
                ref[(l, e, proj)] = a
    t["model.embed_tokens.weight"] = ("BF16", np.zeros((4, H), np.uint16))
    _write_safetensors(tmp_path / "model.safetensors", t)
    ts = Ld.open_model_safetensors(str(tmp_path))
    assert Ld.expert_prefix(ts.keys()) == "model"
    w13, w2 = Ld.read_layer_experts_bf16(ts, "model", 2, 1, 3)       # absolute layer 2 = moe layer 1 with first_k_dense=1
    assert w13.shape == (2, 2 * I, H) and w2.shape == (2, H, I)
    assert np.array_equal(w13[0, :I], ref[(1, 1, "gate_proj")]) and np.array_equal(w13[0, I:], ref[(1, 1, "up_proj")])
    assert np.array_equal(w2[1], ref[(1, 2, "down_proj")])


def test_gguf_parser_and_merged_expert_slicing(tmp_path):
    gguf = pytest.importorskip("gguf")
    rng = np.random.default_rng(1)
    E, H, I = 4, 256, 256
    gate = G.random_blocks(rng, G.GGML_Q4_K, E * I, H).reshape(E, I, -1)
    up = G.random_blocks(rng, G.GGML_Q4_K, E * I, H).reshape(E, I, -1)
    down = np.stack([G.quantize_q8_0(rng.normal(0, 0.02, (H, I)).astype(np.float32)) for _ in range(E)])
    path = str(tmp_path / "m.gguf")
    w = gguf.GGUFWriter(path, "llama")
    w.add_uint32("llama.block_count", 2)
    w.add_string("general.name", "synthetic")
    Q = gguf.GGMLQuantizationType
    w.add_tensor("blk.1.ffn_gate_exps.weight", gate, raw_shape=(E, I, gate.shape[2]), raw_dtype=Q.Q4_K)
    w.add_tensor("blk.1.ffn_up_exps.weight", up, raw_shape=(E, I, up.shape[2]), raw_dtype=Q.Q4_K)
    w.add_tensor("blk.1.ffn_down_exps.weight", down, raw_shape=(E, H, down.shape[2]), raw_dtype=Q.Q8_0)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    g = Ld.GgufFile(path)
    assert g.metadata["general.name"] == "synthetic" and g.metadata["llama.block_count"] == 2
    assert g.tensors["blk.1.ffn_gate_exps.weight"]["type"] == 12 and g.tensors["blk.1.ffn_down_exps.weight"]["type"] == 8
    a, b, c, tgu, tdn = Ld.gguf_expert_blocks(g, 1, 1, 3, H, I)
    assert (tgu, tdn) == ("Q4_K", "Q8_0")
    assert np.array_equal(a, gate[1:3]) and np.array_equal(b, up[1:3]) and np.array_equal(c, down[1:3])
    # the bytes we hand to the GPU dequantise (oracle) to what gguf-py says
    ref = gguf.quants.dequantize(gate[2], Q.Q4_K)
    assert np.array_equal(G.dequantize(G.GGML_Q4_K, a[1].reshape(-1), I * H).reshape(I, H), ref)


def test_attention_weight_conventions(tmp_path):
    """Qwen3.5 separate in_proj_* -> fused per-group layout (loop restatement of weight_loader.py:384-414), kv_b_proj split,
    norm +1 shift in BF16, router / norm naming."""
    import torch
    torch.manual_seed(0)
    nk, dk, nv, dv, H = 4, 8, 8, 16, 32
    r, kd, vd = nv // nk, nk * dk, nv * dv
    bf = torch.bfloat16
    qkv, z = torch.randn(2 * kd + vd, H).to(bf), torch.randn(vd, H).to(bf)
    b, a = torch.randn(nv, H).to(bf), torch.randn(nv, H).to(bf)
    parts, ba = [], []
    for i in range(nk):                                       # the reference's loop, restated
        parts += [qkv[i * dk:(i + 1) * dk], qkv[kd + i * dk:kd + (i + 1) * dk],
                  qkv[2 * kd + i * r * dv:2 * kd + (i + 1) * r * dv], z[i * r * dv:(i + 1) * r * dv]]
        ba += [b[i * r:(i + 1) * r], a[i * r:(i + 1) * r]]
    got_qkvz, got_ba = Ld.fuse_qwen35_linear_attn(qkv, z, b, a, nk, dk, nv, dv)
    assert torch.equal(got_qkvz, torch.cat(parts, 0)) and torch.equal(got_ba, torch.cat(ba, 0))
    # and the oracle's un-interleave (linear_attention.py:337-391) undoes it
    from oracle import attention as A
    x = torch.randn(3, H).to(bf)
    q_, k_, v_, z_, b_, a_ = A.gdn_unsplit(torch.nn.functional.linear(x, got_qkvz), torch.nn.functional.linear(x, got_ba), nk, nv, dk, dv)
    assert torch.equal(q_.reshape(3, -1), torch.nn.functional.linear(x, qkv[:kd]))
    assert torch.equal(z_.reshape(3, -1), torch.nn.functional.linear(x, z))
    assert torch.equal(a_.reshape(3, -1), torch.nn.functional.linear(x, a))
    kv_b = torch.randn(4 * (128 + 64), 16).to(bf)
    w_kc, w_vc = Ld.split_kv_b_proj(kv_b, 4, 128, 64)
    assert w_kc.shape == (4, 128, 16) and w_vc.shape == (4, 64, 16)
    assert torch.equal(w_kc[1], kv_b[192:192 + 128]) and torch.equal(w_vc[3], kv_b[3 * 192 + 128:4 * 192])
    w = torch.tensor([0.00390625, -0.5, 0.1]).to(bf)
    assert torch.equal(Ld.norm_plus_one(w), (w.float() + 1.0).to(bf)) and Ld.norm_plus_one(w).dtype == bf
    # file-level naming
    u16 = lambda t: t.view(torch.int16).numpy().view(np.uint16)
    t = {"model.layers.0.linear_attn.in_proj_qkv.weight": ("BF16", u16(qkv)), "model.layers.0.linear_attn.in_proj_z.weight": ("BF16", u16(z)),
         "model.layers.0.linear_attn.in_proj_b.weight": ("BF16", u16(b)), "model.layers.0.linear_attn.in_proj_a.weight": ("BF16", u16(a)),
         "model.layers.0.linear_attn.out_proj.weight": ("BF16", u16(torch.randn(H, vd).to(bf))),
         "model.layers.0.linear_attn.conv1d.weight": ("BF16", u16(torch.randn(2 * kd + vd, 1, 4).to(bf))),
         "model.layers.0.linear_attn.A_log": ("BF16", u16(torch.randn(nv).to(bf))), "model.layers.0.linear_attn.dt_bias": ("BF16", u16(torch.randn(nv).to(bf))),
         "model.layers.0.linear_attn.norm.weight": ("BF16", u16(torch.randn(dv).to(bf))),
         "model.layers.0.input_layernorm.weight": ("BF16", u16(w)), "model.layers.0.post_attention_layernorm.weight": ("BF16", u16(w)),
         "model.layers.0.mlp.gate.weight": ("BF16", u16(torch.randn(6, H).to(bf)))}
    _write_safetensors(tmp_path / "model.safetensors", t)
    ts = Ld.open_model_safetensors(str(tmp_path))
    la = Ld.load_linear_attention_weights(ts, "model", 0, nk, dk, nv, dv)
    assert torch.equal(la["in_proj_qkvz"], got_qkvz) and la["conv1d_weight"].shape == (2 * kd + vd, 1, 4)
    norms = Ld.load_layer_norms(ts, "model", 0, norm_bias_one=True)
    assert torch.equal(norms["input_layernorm"], w + 1.0)
    gate, bias = Ld.load_router(ts, "model", 0)
    assert gate.shape == (6, H) and bias is None


def test_marlin_cache_file_round_trip(tmp_path):
    """A cache file written the way the reference writes it (header :4117-4144, body :2462-2476, Marlin order produced by the
    oracle's restatement of marlin_repack) reads back to exactly the quantiser's arrays; header validation errors."""
    import struct
    from krasis_b200 import marlin_cache as MC
    from oracle import quant as Q, bf16 as B
    assert MC.fnv1a(b"") == 0xCBF29CE484222325 and MC.fnv1a(b"a") == 0xAF63DC4C8601EC8C       # FNV-1a 64 test vectors
    assert np.array_equal(MC._weight_perm_int4(), Q.marlin_weight_perm_int4())
    rng = np.random.default_rng(3)
    H, I, E, L, gs = 256, 128, 4, 2, 128
    cfg_json = b'{"hidden_size": 256}'
    body, ref = [], {}
    for l in range(L):
        for e in range(E):
            w13 = B.f32_to_bf16_bits(rng.normal(0, 0.02, (2 * I, H)).astype(np.float32))
            w2 = B.f32_to_bf16_bits(rng.normal(0, 0.02, (H, I)).astype(np.float32))
            q13, s13 = Q.quantize_int4(w13)
            q2, s2 = Q.quantize_int4(w2)
            ref[(l, e)] = (q13, s13, q2, s2)
            for q, s in ((q13, s13), (q2, s2)):
                mp, ms = Q.marlin_repack_int4(q, s, gs)
                body += [np.ascontiguousarray(mp).tobytes(), np.ascontiguousarray(ms).tobytes()]
    hdr = b"KRAS" + struct.pack("<I", 3) + struct.pack("<7Q", H, I, E, L, gs, MC.fnv1a(cfg_json), 0)
    assert len(hdr) == 64
    path = tmp_path / "experts_marlin_int4_g128.bin"
    path.write_bytes(hdr + b"".join(body))
    c = MC.MarlinCacheFile(str(path), cfg_json)
    assert (c.hidden_size, c.moe_intermediate_size, c.n_routed_experts, c.num_moe_layers, c.bits) == (H, I, E, L, 4)
    q13, s13, q2, s2 = c.layer_quantiser_arrays(1, 1, 3)
    for j, e in enumerate((1, 2)):
        r = ref[(1, e)]
        assert np.array_equal(q13[j], r[0]) and np.array_equal(s13[j].view(np.uint16), r[1].view(np.uint16))
        assert np.array_equal(q2[j], r[2]) and np.array_equal(s2[j].view(np.uint16), r[3].view(np.uint16))
    with pytest.raises(ValueError):
        MC.MarlinCacheFile(str(path), b"other config")
    (tmp_path / "experts_marlin_int4_g128.bin").write_bytes(hdr + b"".join(body)[:-4])
    with pytest.raises(ValueError):
        MC.MarlinCacheFile(str(path))
    assert MC.marlin_expert_byte_sizes(2048, 512, 128, 4) == (1048576, 32768, 524288, 16384)      # 1 MiB + 32 KiB + 0.5 MiB + 16 KiB


def test_marlin_cache_int8_round_trip(tmp_path):
    import struct
    from krasis_b200 import marlin_cache as MC
    from oracle import quant as Q, bf16 as B
    assert np.array_equal(MC._weight_perm_int8(), Q.marlin_weight_perm_int8())
    rng = np.random.default_rng(4)
    H, I, E, gs = 256, 128, 3, 128
    body, ref = [], []
    for e in range(E):
        qs = []
        for shp in ((2 * I, H), (H, I)):
            q, s = Q.quantize_int8(B.f32_to_bf16_bits(rng.normal(0, 0.02, shp).astype(np.float32)))
            mp, ms = Q.marlin_repack_int8(q, s, gs)
            body += [np.ascontiguousarray(mp).tobytes(), np.ascontiguousarray(ms).tobytes()]
            qs += [q, s]
        ref.append(qs)
    path = tmp_path / "experts_marlin_int8_g128.bin"
    path.write_bytes(b"KRAS" + struct.pack("<I", 3) + struct.pack("<7Q", H, I, E, 1, gs, 0, 0) + b"".join(body))
    c = MC.MarlinCacheFile(str(path))
    assert c.bits == 8
    q13, s13, q2, s2 = c.layer_quantiser_arrays(0, 0, E)
    for e in range(E):
        assert np.array_equal(q13[e], ref[e][0]) and np.array_equal(s13[e].view(np.uint16), ref[e][1].view(np.uint16))
        assert np.array_equal(q2[e], ref[e][2]) and np.array_equal(s2[e].view(np.uint16), ref[e][3].view(np.uint16))
    assert MC.marlin_expert_byte_sizes(2048, 512, 128, 8) == (2097152, 32768, 1048576, 16384)


@pytest.mark.parametrize("tag", ["grouped", "single_group"])
def test_marlin_inverse_matches_reference_execution(tag):
    """krasis_b200.marlin_cache.marlin_to_rowmajor_int4 and the oracle's marlin_unpack_int4 / marlin_repack_int4 against
    outputs of the reference's own inverse_marlin_repack / inverse_scale_permute (python/krasis/triton_moe.py:73-180)
    executed on CPU (tests/golden/make_marlin_golden.py): bit-exact, both directions."""
    from krasis_b200 import marlin_cache as MC
    from oracle import quant as Q
    Gm = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "marlin_inverse_reference.npz"))
    K, N, gs = [int(v) for v in Gm[f"{tag}_cfg"]]
    wm, sm, packed, scales = (Gm[f"{tag}_{k}"] for k in ("wm", "sm", "packed", "scales"))
    p, s = MC.marlin_to_rowmajor_int4(wm, sm, gs)                     # batched (leading dim 2)
    assert np.array_equal(p, packed) and np.array_equal(s, scales)
    for b in range(wm.shape[0]):
        po, so = Q.marlin_unpack_int4(wm[b], sm[b], gs)
        assert np.array_equal(po, packed[b]) and np.array_equal(so.view(np.uint16), scales[b])
        mp, ms = Q.marlin_repack_int4(packed[b], scales[b], gs)       # forward direction of the oracle
        assert np.array_equal(mp, wm[b]) and np.array_equal(ms.view(np.uint16), sm[b])
    fp, fs = MC.rowmajor_to_marlin_int4(packed, scales, gs)           # forward direction of the PRODUCT (get_expert_* hand-off)
    assert np.array_equal(fp, wm) and np.array_equal(fs, sm)


def test_marlin_forward_int8_and_host_untile_round_trips():
    """rowmajor_to_marlin_int8 is the exact inverse of the INT8 cache reader; the host untile of the KB2 tile layout is the
    exact inverse of the numpy statement of the device re-tiling (tests/test_gpu_moe.py checks that against the kernel)."""
    from krasis_b200 import marlin_cache as MC, tiles
    rng = np.random.default_rng(3)
    for N, K in ((128, 256), (256, 128)):
        q = rng.integers(-128, 128, (2, N, K)).astype(np.int8)
        s = rng.integers(0, 2 ** 16, (2, N, K // 128), dtype=np.uint32).astype(np.uint16)
        mq, ms = MC.rowmajor_to_marlin_int8(q, s, 128)
        q2, s2 = MC.marlin_to_rowmajor_int8(mq, ms, 128)
        assert np.array_equal(q, q2) and np.array_equal(s, s2)
        # INT4 tile layout: tile (numpy statement used by the GPU tests) then untile (product)
        packed = rng.integers(0, 2 ** 32, (2, N, K // 8), dtype=np.uint64).astype(np.uint32)
        nib = (packed[..., None] >> (np.arange(8, dtype=np.uint32) * 4)) & 0xF
        w = np.bitwise_or.reduce(nib[..., [0, 2, 4, 6, 1, 3, 5, 7]] << (np.arange(8, dtype=np.uint32) * 4), axis=-1).astype(np.uint32)
        wq = np.ascontiguousarray(w.reshape(2, N // 128, 128, K // 64, 2, 4).transpose(0, 1, 3, 4, 2, 5)).reshape(-1)
        ws = np.ascontiguousarray(s.reshape(2, N // 128, 128, K // 128).transpose(0, 1, 3, 2)).reshape(-1)
        p2, s3 = tiles.untile_int4(wq.view(np.uint8), ws.view(np.uint8), 2, N, K)
        assert np.array_equal(p2, packed) and np.array_equal(s3, s)
        t8 = np.ascontiguousarray(q.reshape(2, N // 128, 128, K // 64, 4, 16).transpose(0, 1, 3, 4, 2, 5)).reshape(-1)
        q3, s4 = tiles.untile_int8(t8.view(np.uint8), ws.view(np.uint8), 2, N, K)
        assert np.array_equal(q3, q) and np.array_equal(s4, s)


def test_reference_style_engine_surface_without_a_model():
    """KrasisEngine(parallel, num_threads, skip_shared_experts) constructs unloaded; every accessor then raises the
    reference's PyRuntimeError text (src/moe.rs:1790-1803); is_marlin_format stays truthy unbound (gpu_prefill.py:851)."""
    from krasis_b200 import KrasisEngine
    e = KrasisEngine(parallel=True, num_threads=8, skip_shared_experts=False)
    assert e.is_parallel() and e.is_marlin_format and not e.has_unified()
    for fn in (e.hidden_size, e.num_experts, e.top_k, e.num_moe_layers, e.gpu_num_bits, e.group_size):
        with pytest.raises(RuntimeError, match="Model not loaded"):
            fn()
    with pytest.raises(RuntimeError, match="Model not loaded"):
        e.get_expert_w13_packed(0)


def test_tile_cache_header_discipline():
    """KB2 tile cache header: reference field order / magic, own version, num_bits, and rejection of foreign files."""
    from krasis_b200 import tile_cache as T
    from krasis_b200.marlin_cache import CACHE_MAGIC, fnv1a
    h = T.pack_header(2048, 512, 512, 48, 128, fnv1a(b'{"a": 1}'), 1, 4)
    assert len(h) == 64 and h[:4] == CACHE_MAGIC
    d = T.unpack_header(h)
    assert d == dict(hidden=2048, inter=512, n_experts=512, n_layers=48, group_size=128, config_hash=fnv1a(b'{"a": 1}'),
                     n_shared=1, num_bits=4)
    with pytest.raises(ValueError):
        T.unpack_header(b"XXXX" + h[4:])                       # bad magic
    marlin = struct.pack("<4sI7Q", CACHE_MAGIC, 3, 2048, 512, 512, 48, 128, 0, 1)     # the reference's Marlin cache (version 3)
    with pytest.raises(ValueError):
        T.unpack_header(marlin)
    with pytest.raises(ValueError):
        T.unpack_header(h[:40])
    assert T.cache_file_name(4) == "experts_kb2_int4_g128.bin"
