"""Reader for the reference's on-disk GPU expert cache `~/.krasis/cache/<model>/experts_marlin_int{4,8}_g{gs}.bin`.

Format (src/weights/mod.rs): 64-byte header :857-866,4117-4144 — magic "KRAS", version u32 (3 = Marlin), then u64 LE
hidden_size, moe_intermediate_size, n_routed_experts, num_moe_layers, group_size, config_hash (FNV-1a of config.json,
:884-893), n_shared_experts; body :2462-2476 — for each (layer, expert): w13_packed [K/16, 2·2I] u32, w13_scales
[K/gs, 2I] bf16, w2_packed [I/16, 2·H] u32, w2_scales [I/gs, H] bf16 in the Marlin GPU layout (byte sizes :955-970),
then the shared experts (:2486-2509).

The B200 kernels do not use the Marlin tile order, so the bytes are taken back to the quantiser's row-major form
([N, K/8] u32 nibbles along K, scales [N, K/gs]; src/weights/marlin.rs:145-207) by inverting the permutation of
marlin.rs:256-327,330-491 — a pure byte permutation done here on the host at load time (like the reference's own
marlin_repack runs on its host) — and handed to kb2_load_experts_host, which re-tiles them on the device.
"""
import mmap
import struct
from typing import Optional, Tuple

import numpy as np

CACHE_MAGIC = b"KRAS"
CACHE_VERSION_MARLIN = 3
CACHE_HEADER_SIZE = 64


def fnv1a(data: bytes) -> int:
    """src/weights/mod.rs:884-893."""
    h = 0xCBF29CE484222325
    for b in data:
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def marlin_w2_padded_n(hidden: int, intermediate: int) -> int:
    """src/weights/mod.rs:942-949."""
    return hidden + 64 if hidden == intermediate and hidden % 256 != 0 else hidden


def marlin_expert_byte_sizes(h: int, m: int, group_size: int, bits: int) -> Tuple[int, int, int, int]:
    """src/weights/mod.rs:955-970: (w13_packed, w13_scales, w2_packed, w2_scales) bytes of one expert."""
    h_w2 = marlin_w2_padded_n(h, m)
    div = 8 if bits == 4 else 4
    return (h // div) * (2 * m) * 4, (h // group_size) * (2 * m) * 2, (m // div) * h_w2 * 4, (m // group_size) * h_w2 * 2


def _weight_perm_int4() -> np.ndarray:
    """dest -> src index inside one 1024-element block (marlin.rs:256-295, with the [0,2,4,6,1,3,5,7] interleave)."""
    perm = []
    for i in range(32):
        col, r = i // 4, i % 4
        base = [16 * row + col + 8 * blk for blk in (0, 1) for row in (2 * r, 2 * r + 1, 2 * r + 8, 2 * r + 9)]
        for j in range(4):
            perm.extend(p + 256 * j for p in base)
    return np.asarray(perm, np.int64).reshape(-1, 8)[:, [0, 2, 4, 6, 1, 3, 5, 7]].reshape(-1)


def _weight_perm_int8() -> np.ndarray:
    """Same base permutation, [0,2,1,3] interleave in groups of four (marlin.rs:586-625)."""
    base = _weight_perm_int4().reshape(-1, 8)[:, [0, 4, 1, 5, 2, 6, 3, 7]].reshape(-1)     # undo the INT4 interleave
    return base.reshape(-1, 4)[:, [0, 2, 1, 3]].reshape(-1)


def _scale_perm(grouped: bool) -> np.ndarray:
    """marlin.rs:302-321: 64-entry permutation for group-wise scales, 32-entry for a single group."""
    if grouped:
        return np.asarray([i + 8 * j for i in range(8) for j in range(8)], np.int64)
    return np.asarray([2 * i + o for i in range(4) for o in (0, 1, 8, 9, 16, 17, 24, 25)], np.int64)


def marlin_to_rowmajor_int4(mpacked: np.ndarray, mscales: np.ndarray, group_size: int):
    """([..., K/16, 2N] u32, [..., K/gs, N] u16) in Marlin order -> ([..., N, K/8] u32, [..., N, K/gs] u16)."""
    lead = mpacked.shape[:-2]
    k16, n2 = mpacked.shape[-2:]
    k, n = k16 * 16, n2 // 2
    shifts = np.arange(8, dtype=np.uint32) * np.uint32(4)
    nib = ((mpacked.astype(np.uint32)[..., None] >> shifts) & np.uint32(0xF)).astype(np.uint8)
    nib = nib.reshape(lead + (k16, n * 16 // 1024, 1024))
    src = np.empty_like(nib)
    src[..., _weight_perm_int4()] = nib                                   # undo the in-block permutation
    t = src.reshape(lead + (k16, n // 16, 16, 16))                        # [K/16][N/16][k 16][n 16] tiles
    kn = np.moveaxis(t, -2, -3).reshape(lead + (k, n))                    # [K][N]
    u = np.swapaxes(kn, -1, -2).astype(np.uint32).reshape(lead + (n, k // 8, 8))
    packed = np.bitwise_or.reduce(u << shifts, axis=-1).astype(np.uint32)
    p = _scale_perm(group_size < k)
    s = mscales.astype(np.uint16).reshape(lead + (-1, len(p)))
    sinv = np.empty_like(s)
    sinv[..., p] = s
    scales = np.ascontiguousarray(np.swapaxes(sinv.reshape(lead + (k // group_size, n)), -1, -2))
    return packed, scales


def _scales_to_rowmajor(mscales: np.ndarray, lead, k: int, n: int, group_size: int):
    p = _scale_perm(group_size < k)
    s = mscales.astype(np.uint16).reshape(lead + (-1, len(p)))
    sinv = np.empty_like(s)
    sinv[..., p] = s
    return np.ascontiguousarray(np.swapaxes(sinv.reshape(lead + (k // group_size, n)), -1, -2))


def marlin_to_rowmajor_int8(mpacked: np.ndarray, mscales: np.ndarray, group_size: int):
    """([..., K/16, 4N] u32, [..., K/gs, N] u16) in Marlin INT8 order (q + 128, marlin.rs:639-760) -> ([..., N, K] i8, scales)."""
    lead = mpacked.shape[:-2]
    k16, n4 = mpacked.shape[-2:]
    k, n = k16 * 16, n4 // 4
    shifts = np.arange(4, dtype=np.uint32) * np.uint32(8)
    b = ((mpacked.astype(np.uint32)[..., None] >> shifts) & np.uint32(0xFF)).astype(np.uint8)
    b = b.reshape(lead + (k16, n * 16 // 1024, 1024))
    src = np.empty_like(b)
    src[..., _weight_perm_int8()] = b
    t = src.reshape(lead + (k16, n // 16, 16, 16))
    kn = np.moveaxis(t, -2, -3).reshape(lead + (k, n))
    q = np.ascontiguousarray((np.swapaxes(kn, -1, -2).astype(np.int16) - 128).astype(np.int8))
    return q, _scales_to_rowmajor(mscales, lead, k, n, group_size)


def rowmajor_to_marlin_int4(packed: np.ndarray, scales: np.ndarray, group_size: int):
    """The forward direction (marlin.rs:330-491 `marlin_repack`): ([..., N, K/8] u32, [..., N, K/gs] u16) as the quantiser
    emits them -> ([..., K/16, 2N] u32, [..., K/gs, N] u16) in the Marlin GPU order.  Exact inverse of marlin_to_rowmajor_int4;
    used to serve the reference's get_expert_* hand-off (src/moe.rs:1972-2097) from this engine."""
    lead = packed.shape[:-2]
    n, k8 = packed.shape[-2:]
    k = k8 * 8
    shifts = np.arange(8, dtype=np.uint32) * np.uint32(4)
    nib = ((packed.astype(np.uint32)[..., None] >> shifts) & np.uint32(0xF)).astype(np.uint8).reshape(lead + (n, k))
    kn = np.swapaxes(nib, -1, -2)                                              # [K][N]
    t = np.moveaxis(kn.reshape(lead + (k // 16, 16, n // 16, 16)), -3, -2)     # [K/16][N/16][k 16][n 16]
    blk = t.reshape(lead + (k // 16, n * 16 // 1024, 1024))[..., _weight_perm_int4()]
    u = blk.reshape(lead + (k // 16, 2 * n, 8)).astype(np.uint32)
    mpacked = np.bitwise_or.reduce(u << shifts, axis=-1).astype(np.uint32)
    return mpacked, _scales_to_marlin(scales, lead, k, n, group_size)


def _scales_to_marlin(scales: np.ndarray, lead, k: int, n: int, group_size: int):
    p = _scale_perm(group_size < k)
    s = np.swapaxes(scales.astype(np.uint16), -1, -2).reshape(lead + (-1, len(p)))
    return np.ascontiguousarray(s[..., p].reshape(lead + (k // group_size, n)))


def rowmajor_to_marlin_int8(q: np.ndarray, scales: np.ndarray, group_size: int):
    """([..., N, K] i8, scales) -> ([..., K/16, 4N] u32 holding q + 128, Marlin scales) (marlin.rs:582-626,639-760)."""
    lead = q.shape[:-2]
    n, k = q.shape[-2:]
    b = (q.astype(np.int16) + 128).astype(np.uint8)
    kn = np.swapaxes(b, -1, -2)
    t = np.moveaxis(kn.reshape(lead + (k // 16, 16, n // 16, 16)), -3, -2)
    blk = t.reshape(lead + (k // 16, n * 16 // 1024, 1024))[..., _weight_perm_int8()]
    u = blk.reshape(lead + (k // 16, 4 * n, 4)).astype(np.uint32)
    shifts = np.arange(4, dtype=np.uint32) * np.uint32(8)
    return np.bitwise_or.reduce(u << shifts, axis=-1).astype(np.uint32), _scales_to_marlin(scales, lead, k, n, group_size)


class MarlinCacheFile:
    def __init__(self, path: str, config_json: Optional[bytes] = None):
        self.path = path
        self._f = open(path, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        if len(self._mm) < CACHE_HEADER_SIZE or self._mm[:4] != CACHE_MAGIC:
            raise ValueError(f"{path}: not a Krasis expert cache")
        version, = struct.unpack_from("<I", self._mm, 4)
        if version != CACHE_VERSION_MARLIN:
            raise ValueError(f"{path}: cache version {version}, expected {CACHE_VERSION_MARLIN} (Marlin GPU cache)")
        (self.hidden_size, self.moe_intermediate_size, self.n_routed_experts, self.num_moe_layers, self.group_size,
         self.config_hash, self.n_shared_experts) = struct.unpack_from("<7Q", self._mm, 8)
        if config_json is not None and fnv1a(config_json) != self.config_hash:
            raise ValueError("cache was built for a different config.json (FNV-1a hash mismatch)")
        name = path.rsplit("/", 1)[-1]
        self.bits = 8 if "_int8_" in name else 4
        self._sizes = marlin_expert_byte_sizes(self.hidden_size, self.moe_intermediate_size, self.group_size, self.bits)
        per_routed = sum(self._sizes) * self.n_routed_experts * self.num_moe_layers
        h, gs, div = self.hidden_size, self.group_size, (8 if self.bits == 4 else 4)
        sm = self.n_shared_experts * self.moe_intermediate_size
        shared = self.num_moe_layers * ((h // div) * 2 * sm * 4 + (h // gs) * 2 * sm * 2 + (sm // div) * h * 4 + (sm // gs) * h * 2) \
            if self.n_shared_experts else 0
        if len(self._mm) != CACHE_HEADER_SIZE + per_routed + shared:          # mod.rs:1082-1105,2433-2442
            raise ValueError(f"cache size mismatch: expected {CACHE_HEADER_SIZE + per_routed + shared} bytes, got {len(self._mm)}")

    def layer_marlin_arrays(self, layer: int, e0: int, e1: int):
        """Zero-copy views of experts [e0, e1) of one layer, still in Marlin order."""
        h, m, gs = self.hidden_size, self.moe_intermediate_size, self.group_size
        a, b, c, d = self._sizes
        per = a + b + c + d
        base = CACHE_HEADER_SIZE + (layer * self.n_routed_experts + e0) * per
        raw = np.frombuffer(self._mm, np.uint8, (e1 - e0) * per, base).reshape(e1 - e0, per)
        hw2 = marlin_w2_padded_n(h, m)
        f = 2 if self.bits == 4 else 4                       # u32 words per (k-tile, n): [K/16, 2N] for INT4, [K/16, 4N] for INT8
        return (raw[:, :a].view(np.uint32).reshape(-1, h // 16, f * 2 * m), raw[:, a:a + b].view(np.uint16).reshape(-1, h // gs, 2 * m),
                raw[:, a + b:a + b + c].view(np.uint32).reshape(-1, m // 16, f * hw2), raw[:, a + b + c:].view(np.uint16).reshape(-1, m // gs, hw2))

    def layer_quantiser_arrays(self, layer: int, e0: int, e1: int):
        """(w13_q [E,2I,H/8] u32, w13_s [E,2I,H/gs] u16, w2_q [E,H,I/8] u32, w2_s [E,H,I/gs] u16): what quantize_int4 emitted."""
        if marlin_w2_padded_n(self.hidden_size, self.moe_intermediate_size) != self.hidden_size:
            raise NotImplementedError("padded w2 (hidden == intermediate, not a multiple of 256)")
        p13, s13, p2, s2 = self.layer_marlin_arrays(layer, e0, e1)
        f = marlin_to_rowmajor_int4 if self.bits == 4 else marlin_to_rowmajor_int8
        return f(p13, s13, self.group_size) + f(p2, s2, self.group_size)


def load_experts_from_marlin_cache(engine, path: str, config_json: Optional[bytes] = None, start_layer: int = 0):
    """KrasisEngine.load()'s cache fast path (src/weights/mod.rs:2367-2520) for the B200 engine."""
    from .engine import QuantizedExperts
    c = MarlinCacheFile(path, config_json)
    if (c.hidden_size, c.moe_intermediate_size, c.n_routed_experts) != (engine.hidden_size(), engine.intermediate_size(), engine.num_experts()):
        raise ValueError("cache geometry does not match the engine")          # mod.rs:2402-2416
    if c.bits != engine.gpu_num_bits() or c.group_size != engine.group_size():
        raise ValueError("cache bits / group size do not match the engine")
    if start_layer + engine.num_moe_layers() > c.num_moe_layers:
        raise ValueError("layer range exceeds the cache")                      # mod.rs:2425-2430
    for i in range(engine.num_moe_layers()):
        engine.load_quantized_layer(i, QuantizedExperts(*c.layer_quantiser_arrays(start_layer + i, engine.expert_start, engine.expert_end)))
