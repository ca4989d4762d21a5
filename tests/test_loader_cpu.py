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
