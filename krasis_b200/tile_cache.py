"""B200-layout expert cache: quantise / re-tile once, then start from disk.

The reference caches its GPU experts as `~/.krasis/cache/<model>/experts_marlin_int{4,8}_g{gs}.bin` with a 64-byte header
(magic "KRAS", version, geometry, FNV-1a hash of config.json, n_shared_experts — src/weights/mod.rs:857-893,4117-4144) that is
validated before the body is trusted (:2382-2423) and an exact total-size check (:1082-1108).  This module keeps that
discipline for the KB2 tile layout the tcgen05 kernels consume (krasis_b200/csrc/moe_common.cuh), so a restart neither
re-quantises nor re-tiles the 37 GiB of Qwen3-Coder-Next experts:

    header (64 B)  "KRAS" | u32 version = 6 | u64 hidden | u64 moe_intermediate | u64 n_routed_experts | u64 num_moe_layers
                   | u64 group_size | u64 fnv1a(config.json) | u64 n_shared_experts | (4 B: num_bits)
    body           for each MoE layer: w13 tiles [E][...] | w13 scale tiles [E][...] | w2 tiles [E][...] | w2 scale tiles [E][...]
                   (the four buffers of kb2_tiled_bytes for ALL experts; a rank's experts are a contiguous slice of each)

Versions 3 / 4 / 5 are the reference's Marlin / CPU / GGUF-sourced caches (mod.rs:4117-4214); 6 is new and only this
library reads it.  File name: `experts_kb2_int{bits}_g{gs}.bin`, next to the reference's files.
"""
import mmap
import os
import struct
from typing import Optional

import numpy as np

from .marlin_cache import CACHE_HEADER_SIZE, CACHE_MAGIC, fnv1a

CACHE_VERSION_KB2 = 6


def cache_file_name(num_bits: int, group_size: int = 128) -> str:
    return f"experts_kb2_int{num_bits}_g{group_size}.bin"


def pack_header(hidden: int, inter: int, n_experts: int, n_layers: int, group_size: int, config_hash: int, n_shared: int,
                num_bits: int) -> bytes:
    """64 bytes in the reference's field order (mod.rs:4117-4144); num_bits rides in the high half of the version word
    (the reference's CPU cache likewise encodes num_bits in a spare header word, mod.rs:4146-4179)."""
    h = struct.pack("<4sI7Q", CACHE_MAGIC, CACHE_VERSION_KB2 | (num_bits << 16), hidden, inter, n_experts, n_layers, group_size,
                    config_hash, n_shared)
    assert len(h) == CACHE_HEADER_SIZE
    return h


def unpack_header(raw: bytes) -> dict:
    if len(raw) < CACHE_HEADER_SIZE:
        raise ValueError("cache file shorter than its header")
    magic, ver, hidden, inter, n_exp, n_layers, gs, chash, n_shared = struct.unpack("<4sI7Q", raw[:CACHE_HEADER_SIZE])
    if magic != CACHE_MAGIC:
        raise ValueError(f"bad cache magic {magic!r}")
    if ver & 0xFFFF != CACHE_VERSION_KB2:
        raise ValueError(f"cache version {ver & 0xFFFF} is not the KB2 tile cache (version {CACHE_VERSION_KB2})")
    return dict(hidden=hidden, inter=inter, n_experts=n_exp, n_layers=n_layers, group_size=gs, config_hash=chash,
                n_shared=n_shared, num_bits=ver >> 16)


def _global_bytes(engine, which: int) -> int:
    e_loc = engine.expert_end - engine.expert_start
    return engine.tiled_bytes(which) // e_loc * engine.num_experts()


def expected_size(engine) -> int:
    return CACHE_HEADER_SIZE + engine.num_moe_layers() * sum(_global_bytes(engine, w) for w in range(4))


def write_tile_cache(engine, path: str, config_json: bytes, n_shared_experts: int = 0) -> int:
    """Dump every layer's tiled experts of a single-rank engine (which holds all experts); returns the file size.
    Written to `path + ".tmp"` and renamed, so a crash never leaves a truncated file that passes the header check."""
    import ctypes as C
    from . import capi
    if engine.expert_start != 0 or engine.expert_end != engine.num_experts():
        raise ValueError("write the cache from a single-rank engine (it must hold all experts)")
    hdr = pack_header(engine.hidden_size(), engine.intermediate_size(), engine.num_experts(), engine.num_moe_layers(),
                      engine.group_size(), fnv1a(config_json), n_shared_experts, engine.gpu_num_bits())
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(hdr)
        for layer in range(engine.num_moe_layers()):
            for which in range(4):
                n = engine.tiled_bytes(which)
                buf = np.empty(n, np.uint8)
                capi.check(engine._lib.kb2_export_experts_tiled_host(engine._h, layer, which, buf.ctypes.data, n))
                f.write(buf.tobytes())
    size = os.path.getsize(tmp)
    if size != expected_size(engine):
        os.remove(tmp)
        raise RuntimeError(f"cache size {size} != expected {expected_size(engine)}")
    os.replace(tmp, path)
    return size


def load_tile_cache(engine, path: str, config_json: Optional[bytes] = None) -> None:
    """Validate header + total size against the engine's configuration (and config.json's hash when given), then upload
    this rank's expert slice of every layer.  Raises ValueError on any mismatch, like the reference's validation path
    (mod.rs:2382-2423) which then rebuilds the cache."""
    from . import capi
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    h = unpack_header(mm[:CACHE_HEADER_SIZE])
    want = dict(hidden=engine.hidden_size(), inter=engine.intermediate_size(), n_experts=engine.num_experts(),
                n_layers=engine.num_moe_layers(), group_size=engine.group_size(), num_bits=engine.gpu_num_bits())
    for k, v in want.items():
        if h[k] != v:
            raise ValueError(f"cache {k} = {h[k]} does not match the engine ({v})")
    if config_json is not None and h["config_hash"] != fnv1a(config_json):
        raise ValueError("cache was built from a different config.json (FNV-1a hash mismatch)")
    if len(mm) != expected_size(engine):
        raise ValueError(f"cache size {len(mm)} != expected {expected_size(engine)}")
    e_loc, E = engine.expert_end - engine.expert_start, engine.num_experts()
    off = CACHE_HEADER_SIZE
    for layer in range(engine.num_moe_layers()):
        bufs = []
        for which in range(4):
            g, n = _global_bytes(engine, which), engine.tiled_bytes(which)
            per = g // E
            a = np.frombuffer(mm, dtype=np.uint8, count=n, offset=off + per * engine.expert_start) if n else np.empty(0, np.uint8)
            bufs.append(np.ascontiguousarray(a))
            off += g
        capi.check(engine._lib.kb2_load_experts_tiled_host(engine._h, layer, *[b.ctypes.data if b.size else None for b in bufs]))
