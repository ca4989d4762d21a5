"""Weight loading for the prefill path: safetensors (HF checkpoints) and GGUF (native expert blocks).

Mirrors the reference's loader contracts:
  safetensors mmap reader          src/weights/safetensors_io.rs          -> SafetensorsFile (header JSON + zero-copy slices)
  expert tensor naming             src/weights/mod.rs:4648-4684,5023-5168 -> expert_prefix / per-expert or stacked tensors
  quantise + hand to the GPU path  src/weights/mod.rs:5049-5091 + marlin.rs:145-207 -> load_experts_from_safetensors
                                   (quantiser runs on the device: kb2_quantize_group_dev, bit-exact)
  GGUF v3 parser                   src/gguf.rs:315-450 (magic 0x46554747, alignment 32), tensor table, block sizes :56-85
  GGUF expert tensor naming        src/gguf.rs:488-526, src/weights/mod.rs:3435-3511 (absolute layer index, merged `_exps`)
Parsing is host-side byte work (numpy memmap); every arithmetic step happens in libkrasis_b200 kernels.
"""
import json
import mmap
import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------------------------- safetensors

_ST_DTYPES = {"BF16": (np.uint16, 2), "F16": (np.float16, 2), "F32": (np.float32, 4), "I8": (np.int8, 1),
              "U8": (np.uint8, 1), "I32": (np.int32, 4), "I64": (np.int64, 8)}


class SafetensorsFile:
    """8-byte little-endian header length, JSON header {name: {dtype, shape, data_offsets}}, then raw data."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        n = struct.unpack("<Q", self._f.read(8))[0]
        self.header = json.loads(self._f.read(n))
        self.header.pop("__metadata__", None)
        self._base = 8 + n
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)

    def keys(self):
        return self.header.keys()

    def tensor(self, name: str) -> np.ndarray:
        """Zero-copy view (BF16 comes back as raw uint16, like the Rust reader's &[u16])."""
        info = self.header[name]
        dt, sz = _ST_DTYPES[info["dtype"]]
        a, b = info["data_offsets"]
        cnt = int(np.prod(info["shape"])) if info["shape"] else 1
        if b - a != cnt * sz:
            raise ValueError(f"{name}: size mismatch")
        return np.frombuffer(self._mm, dtype=dt, count=cnt, offset=self._base + a).reshape(info["shape"])


def open_model_safetensors(model_dir: str) -> Dict[str, SafetensorsFile]:
    """name -> file, via model.safetensors.index.json when present."""
    idx = os.path.join(model_dir, "model.safetensors.index.json")
    files: Dict[str, SafetensorsFile] = {}
    out = {}
    if os.path.exists(idx):
        wm = json.load(open(idx))["weight_map"]
        for name, fn in wm.items():
            if fn not in files:
                files[fn] = SafetensorsFile(os.path.join(model_dir, fn))
            out[name] = files[fn]
    else:
        for fn in sorted(os.listdir(model_dir)):
            if fn.endswith(".safetensors"):
                f = SafetensorsFile(os.path.join(model_dir, fn))
                for name in f.keys():
                    out[name] = f
    return out


def expert_prefix(names) -> str:
    """src/weights/mod.rs:4648-4662: part before `.layers.` of any key containing `.mlp.experts.` (skipping mtp)."""
    for n in names:
        if ".mlp.experts." in n and "mtp" not in n:
            return n.split(".layers.")[0]
    raise ValueError("no expert tensors found")


def _expert_tensor_bf16(tensors: Dict[str, SafetensorsFile], name: str) -> np.ndarray:
    """Raw BF16 bits of an expert tensor.  The quantiser's input contract is BF16 (src/weights/mod.rs:5023-5046 reads
    `&[u16]`); any other storage dtype is refused instead of being reinterpreted."""
    dt = tensors[name].header[name]["dtype"]
    if dt != "BF16":
        raise ValueError(f"{name}: expert tensors must be BF16 (found {dt}); pre-quantised / FP16 / FP32 checkpoints are out of scope")
    return tensors[name].tensor(name)


def read_layer_experts_bf16(tensors: Dict[str, SafetensorsFile], prefix: str, layer: int, e0: int, e1: int):
    """Returns (w13 [E, 2I, H] uint16 bf16 with gate rows first, w2 [E, H, I]) for experts [e0, e1)."""
    stacked = f"{prefix}.layers.{layer}.mlp.experts.gate_up_proj"
    if stacked in tensors:                                                  # Qwen3.5: mod.rs:4670-4677,5095-5168
        w13 = _expert_tensor_bf16(tensors, stacked)[e0:e1]
        dn = f"{prefix}.layers.{layer}.mlp.experts.down_proj"
        return np.ascontiguousarray(w13), np.ascontiguousarray(_expert_tensor_bf16(tensors, dn)[e0:e1])
    w13, w2 = [], []
    for e in range(e0, e1):                                                 # mod.rs:5023-5046
        base = f"{prefix}.layers.{layer}.mlp.experts.{e}."
        g = _expert_tensor_bf16(tensors, base + "gate_proj.weight")
        u = _expert_tensor_bf16(tensors, base + "up_proj.weight")
        d = _expert_tensor_bf16(tensors, base + "down_proj.weight")
        w13.append(np.concatenate([g, u], axis=0))                         # w13 = [gate ; up] (mod.rs:346-349)
        w2.append(d)
    return np.stack(w13), np.stack(w2)


def load_experts_from_safetensors(engine, model_dir: str, first_k_dense: int = 0, num_bits: Optional[int] = None):
    """HF BF16 experts -> device -> Krasis group quantiser (on the device) -> B200 tiles, layer by layer."""
    import torch
    tensors = open_model_safetensors(model_dir)
    prefix = expert_prefix(tensors.keys())
    if num_bits is not None and num_bits != engine.gpu_num_bits():
        raise ValueError("num_bits does not match the engine's weight format")
    for moe_idx in range(engine.num_moe_layers()):
        w13, w2 = read_layer_experts_bf16(tensors, prefix, moe_idx + first_k_dense, engine.expert_start, engine.expert_end)
        engine.load_bf16_layer(moe_idx, torch.from_numpy(w13.view(np.int16)).view(torch.bfloat16),
                               torch.from_numpy(w2.view(np.int16)).view(torch.bfloat16))


# --------------------------------------------------------------------------------------------- GGUF

GGUF_MAGIC = 0x46554747       # src/gguf.rs:315
GGUF_ALIGN = 32               # src/gguf.rs:316 (general.alignment overrides)
# ggml type -> (block elements, block bytes)   src/gguf.rs:56-85
GGML_BLOCK = {0: (1, 4), 1: (1, 2), 2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34), 9: (32, 40),
              10: (256, 84), 11: (256, 110), 12: (256, 144), 13: (256, 176), 14: (256, 210), 15: (256, 276), 30: (1, 2)}
GGML_NAME = {0: "F32", 1: "F16", 2: "Q4_0", 6: "Q5_0", 8: "Q8_0", 12: "Q4_K", 13: "Q5_K", 14: "Q6_K", 30: "BF16"}
_KV_FIXED = {0: 1, 1: 1, 2: 2, 3: 2, 4: 4, 5: 4, 6: 4, 7: 1, 10: 8, 11: 8, 12: 8}   # gguf metadata value types


class GgufFile:
    """GGUF v2/v3 reader: header, metadata (parsed just enough to skip / expose scalars), tensor table, data slices."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        mm, self._p = self._mm, 0
        magic, version = struct.unpack_from("<II", mm, 0)
        if magic != GGUF_MAGIC:
            raise ValueError(f"{path}: not a GGUF file (magic {magic:#x})")
        if version < 2:
            raise ValueError(f"GGUF version {version} not supported")
        n_tensors, n_kv = struct.unpack_from("<QQ", mm, 8)
        self._p = 24
        self.metadata = {}
        for _ in range(n_kv):
            key = self._str()
            self.metadata[key] = self._value(self._u32())
        self.tensors = {}
        for _ in range(n_tensors):
            name = self._str()
            nd = self._u32()
            dims = [self._u64() for _ in range(nd)]          # GGUF order: fastest dimension first
            typ = self._u32()
            off = self._u64()
            self.tensors[name] = dict(dims=dims, type=typ, offset=off)
        align = int(self.metadata.get("general.alignment", GGUF_ALIGN))
        self.data_start = (self._p + align - 1) // align * align

    def _u32(self):
        v = struct.unpack_from("<I", self._mm, self._p)[0]
        self._p += 4
        return v

    def _u64(self):
        v = struct.unpack_from("<Q", self._mm, self._p)[0]
        self._p += 8
        return v

    def _str(self):
        n = self._u64()
        s = self._mm[self._p:self._p + n].decode("utf-8", "replace")
        self._p += n
        return s

    def _value(self, t):
        if t == 8:
            return self._str()
        if t == 9:                                           # array: elem type, count, elems
            et, n = self._u32(), self._u64()
            return [self._value(et) for _ in range(n)]
        fmt = {0: "B", 1: "b", 2: "H", 3: "h", 4: "I", 5: "i", 6: "f", 7: "?", 10: "Q", 11: "q", 12: "d"}[t]
        v = struct.unpack_from("<" + fmt, self._mm, self._p)[0]
        self._p += _KV_FIXED[t]
        return v

    def n_elements(self, name) -> int:
        return int(np.prod(self.tensors[name]["dims"]))

    def tensor_bytes(self, name: str) -> np.ndarray:
        """Raw block bytes of a tensor (src/gguf.rs:452-464)."""
        t = self.tensors[name]
        be, bb = GGML_BLOCK[t["type"]]
        nbytes = self.n_elements(name) // be * bb
        return np.frombuffer(self._mm, dtype=np.uint8, count=nbytes, offset=self.data_start + t["offset"])


def gguf_expert_blocks(g: GgufFile, abs_layer: int, e0: int, e1: int, H: int, I: int):
    """Returns (gate [E,I,rb], up [E,I,rb], down [E,H,rb'], gate_up_type, down_type) for experts [e0,e1) of layer
    `abs_layer` (= moe_idx + first_k_dense_replace, src/weights/mod.rs:3435), from merged `_exps` tensors
    (expert e = byte slice, mod.rs:3455-3487) or per-expert tensors (src/gguf.rs:488-526)."""
    out, types = [], []
    for kind, rows, k in (("gate", I, H), ("up", I, H), ("down", H, I)):
        merged = f"blk.{abs_layer}.ffn_{kind}_exps.weight"
        if merged in g.tensors:
            typ = g.tensors[merged]["type"]
            be, bb = GGML_BLOCK[typ]
            per = rows * (k // be) * bb
            raw = g.tensor_bytes(merged)
            out.append(raw[e0 * per:e1 * per].reshape(e1 - e0, rows, k // be * bb))
        else:
            parts = []
            for e in range(e0, e1):
                name = f"blk.{abs_layer}.ffn_{kind}.{e}.weight"
                typ = g.tensors[name]["type"]
                be, bb = GGML_BLOCK[typ]
                parts.append(g.tensor_bytes(name).reshape(rows, k // be * bb))
            out.append(np.stack(parts))
        types.append(typ)
    if types[0] != types[1]:
        raise ValueError("gate and up experts use different GGUF types")
    return out[0], out[1], out[2], GGML_NAME.get(types[0], str(types[0])), GGML_NAME.get(types[2], str(types[2]))


def load_experts_from_gguf(engine, gguf_path: str, first_k_dense: int = 0):
    """KrasisEngine.load(..., gguf_path=..., gguf_native=True) for the GPU path: native blocks, no re-quantisation."""
    g = GgufFile(gguf_path)
    H, I = engine.hidden_size(), engine.intermediate_size()
    for moe_idx in range(engine.num_moe_layers()):
        gate, up, down, _, _ = gguf_expert_blocks(g, moe_idx + first_k_dense, engine.expert_start, engine.expert_end, H, I)
        engine.load_gguf_layer(moe_idx, gate, up, down)


# --------------------------------------------------------------------------------------------- attention / norm weights
# Load-time conventions that change numerics (SURVEY.md §8a last row; python/krasis/weight_loader.py).  All host-side
# re-arrangement of BF16 tensors (torch on CPU): no arithmetic except the documented `+ 1.0` in BF16.

def _st_bf16(tensors: Dict[str, SafetensorsFile], name: str):
    """The reference's `_load_bf16` = `_read_tensor(name).to(torch.bfloat16)` (weight_loader.py:147-166): BF16 storage is
    taken as is; F32 / F16 storage (A_log, dt_bias, e_score_correction_bias, some norm weights in real checkpoints) is
    CONVERTED (round to nearest even, like torch); any other dtype is an error rather than a reinterpretation."""
    import torch
    dt = tensors[name].header[name]["dtype"]
    raw = np.array(tensors[name].tensor(name))
    if dt == "BF16":
        return torch.from_numpy(raw.view(np.int16)).view(torch.bfloat16)
    if dt in ("F32", "F16"):
        return torch.from_numpy(raw).to(torch.bfloat16)
    raise ValueError(f"{name}: dtype {dt} cannot be loaded as a BF16 weight")


def norm_plus_one(w):
    """Qwen3-Next / Qwen3.5 RMSNorm weights are stored as `w - 1`: input / post-attention / final / q_norm / k_norm get
    `w + 1.0` at load, computed in BF16 (weight_loader.py:168-169,259-265,286-289); the Gated-DeltaNet gated norm does NOT
    (:421)."""
    return w + 1.0


def fuse_qwen35_linear_attn(qkv, z, b, a, nk: int, dk: int, nv: int, dv: int):
    """Qwen3.5 stores in_proj_{qkv,z,b,a} separately (flat [Q_all, K_all, V_all]); the kernels consume Qwen3-Coder-Next's
    fused per-key-head-group layout [q_i | k_i | v_i (r heads) | z_i (r heads)] and [b_i | a_i] (weight_loader.py:384-414)."""
    import torch
    r, kd = nv // nk, nk * dk
    H = qkv.shape[1]
    q = qkv[:kd].reshape(nk, dk, H)
    k = qkv[kd:2 * kd].reshape(nk, dk, H)
    v = qkv[2 * kd:].reshape(nk, r * dv, H)
    zz = z.reshape(nk, r * dv, H)
    qkvz = torch.cat([q, k, v, zz], dim=1).reshape(nk * (2 * dk + 2 * r * dv), H)
    ba = torch.cat([b.reshape(nk, r, H), a.reshape(nk, r, H)], dim=1).reshape(2 * nv, H)
    return qkvz.contiguous(), ba.contiguous()


def split_kv_b_proj(kv_b, num_heads: int, qk_nope: int, v_head: int):
    """kv_b_proj [heads*(nope+v), lora] -> w_kc [heads, nope, lora], w_vc [heads, v, lora] (weight_loader.py:223-230)."""
    kv = kv_b.reshape(num_heads, qk_nope + v_head, kv_b.shape[1])
    return kv[:, :qk_nope, :].contiguous(), kv[:, qk_nope:, :].contiguous()


def load_linear_attention_weights(tensors, layers_prefix: str, layer: int, nk: int, dk: int, nv: int, dv: int) -> dict:
    """weights dict for krasis_b200.attention.GatedDeltaNetAttention (weight_loader.py:376-425)."""
    p = f"{layers_prefix}.layers.{layer}.linear_attn"
    w = {}
    if f"{p}.in_proj_qkvz.weight" in tensors:
        w["in_proj_qkvz"] = _st_bf16(tensors, f"{p}.in_proj_qkvz.weight")
        w["in_proj_ba"] = _st_bf16(tensors, f"{p}.in_proj_ba.weight")
    else:
        w["in_proj_qkvz"], w["in_proj_ba"] = fuse_qwen35_linear_attn(
            *[_st_bf16(tensors, f"{p}.in_proj_{n}.weight") for n in ("qkv", "z", "b", "a")], nk, dk, nv, dv)
    w["out_proj"] = _st_bf16(tensors, f"{p}.out_proj.weight")
    w["conv1d_weight"] = _st_bf16(tensors, f"{p}.conv1d.weight")
    w["A_log"], w["dt_bias"] = _st_bf16(tensors, f"{p}.A_log"), _st_bf16(tensors, f"{p}.dt_bias")
    w["norm_weight"] = _st_bf16(tensors, f"{p}.norm.weight")            # not shifted
    return w


def load_gqa_weights(tensors, layers_prefix: str, layer: int, norm_bias_one: bool) -> dict:
    """weights dict for krasis_b200.attention.GQAAttention (weight_loader.py:233-272)."""
    p = f"{layers_prefix}.layers.{layer}.self_attn"
    w = {k: _st_bf16(tensors, f"{p}.{k}.weight") for k in ("q_proj", "k_proj", "v_proj", "o_proj")}
    for k in ("q_norm", "k_norm"):
        if f"{p}.{k}.weight" in tensors:
            t = _st_bf16(tensors, f"{p}.{k}.weight")
            w[k] = norm_plus_one(t) if norm_bias_one else t
    return w


def load_mla_weights(tensors, layers_prefix: str, layer: int, num_heads: int, qk_nope: int, v_head: int, has_q_lora: bool) -> dict:
    """weights dict for krasis_b200.attention.MLAAttention (weight_loader.py:199-231)."""
    p = f"{layers_prefix}.layers.{layer}.self_attn"
    names = ["q_a_proj", "q_b_proj", "q_a_layernorm"] if has_q_lora else ["q_proj"]
    w = {k: _st_bf16(tensors, f"{p}.{k}.weight") for k in names + ["kv_a_proj_with_mqa", "o_proj", "kv_a_layernorm"]}
    w["w_kc"], w["w_vc"] = split_kv_b_proj(_st_bf16(tensors, f"{p}.kv_b_proj.weight"), num_heads, qk_nope, v_head)
    return w


def load_layer_norms(tensors, layers_prefix: str, layer: int, norm_bias_one: bool) -> dict:
    """input_layernorm / post_attention_layernorm (weight_loader.py:274-291)."""
    p = f"{layers_prefix}.layers.{layer}"
    out = {k: _st_bf16(tensors, f"{p}.{k}.weight") for k in ("input_layernorm", "post_attention_layernorm")}
    return {k: norm_plus_one(v) for k, v in out.items()} if norm_bias_one else out


def load_router(tensors, layers_prefix: str, layer: int):
    """(gate [E,H] bf16, e_score_correction_bias or None) — `mlp.gate.*` or `mlp.router.*` (weight_loader.py:307-334)."""
    for stem in ("mlp.gate", "mlp.router"):
        name = f"{layers_prefix}.layers.{layer}.{stem}.weight"
        if name in tensors:
            bias = f"{layers_prefix}.layers.{layer}.{stem}.e_score_correction_bias"
            return _st_bf16(tensors, name), (_st_bf16(tensors, bias) if bias in tensors else None)
    raise KeyError(f"no router weight for layer {layer}")
