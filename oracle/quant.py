"""Krasis symmetric INT4 / INT8 group quantiser (oracle; test infrastructure only).

This is the *weight-value contract* of the reference: the GPU
path sees exactly the integers and BF16 scales these functions produce; the
Marlin repack that follows is a pure permutation.

Restates:
  quantize_int4      src/weights/marlin.rs:145-207
  dequantize_int4    src/weights/marlin.rs:210-234
  quantize_int8      src/weights/marlin.rs:65-114
  dequantize_int8    src/weights/marlin.rs:117-133
  marlin perm tables src/weights/marlin.rs:256-321
  marlin_repack      src/weights/marlin.rs:330-491   (INT4)
  dequantize_marlin  src/weights/marlin.rs:496-576   (inverse, used to read caches)
All arithmetic is done in float32 exactly as the Rust code does it
(f32 division, f32 reciprocal, f32 multiply, f32::round = half away from zero).
"""
import numpy as np

from .bf16 import bf16_bits_to_f32, f32_to_bf16_bits

GROUP_SIZE = 128  # marlin.rs:12
PACK_FACTOR = 8   # marlin.rs:15


def _round_half_away(x32: np.ndarray) -> np.ndarray:
    """Rust f32::round: half away from zero (exact; done in f64 so +0.5 cannot round)."""
    x = x32.astype(np.float64)
    return np.copysign(np.floor(np.abs(x) + 0.5), x)


def _group_scales(w: np.ndarray, group_size: int, qmax: float) -> np.ndarray:
    rows, cols = w.shape
    g = w.reshape(rows, cols // group_size, group_size)
    amax = np.max(np.abs(g), axis=2).astype(np.float32)
    scale = np.where(amax == 0.0, np.float32(1.0), amax / np.float32(qmax)).astype(np.float32)
    return f32_to_bf16_bits(scale)  # [rows, groups] uint16


def quantize_int4(weight_bf16_bits: np.ndarray, group_size: int = GROUP_SIZE):
    """marlin.rs:145-207.  weight [rows, cols] raw BF16 (uint16).

    Returns (packed [rows, cols/8] uint32 — nibble j of a word = column 8c+j,
    value q+8; scales [rows, cols/group_size] raw BF16 uint16).
    """
    wb = np.asarray(weight_bf16_bits, dtype=np.uint16)
    rows, cols = wb.shape
    assert cols % group_size == 0 and cols % PACK_FACTOR == 0
    w = bf16_bits_to_f32(wb)
    scales = _group_scales(w, group_size, 7.0)
    s = bf16_bits_to_f32(scales)
    inv = np.where(s == 0.0, np.float32(0.0), np.float32(1.0) / s).astype(np.float32)
    prod = (w.reshape(rows, -1, group_size) * inv[:, :, None]).astype(np.float32)
    q = np.clip(_round_half_away(prod), -8.0, 7.0).astype(np.int32).reshape(rows, cols)
    u4 = ((q + 8) & 0xF).astype(np.uint32).reshape(rows, cols // PACK_FACTOR, PACK_FACTOR)
    shifts = (np.arange(PACK_FACTOR, dtype=np.uint32) * np.uint32(4))
    packed = np.bitwise_or.reduce(u4 << shifts[None, None, :], axis=2).astype(np.uint32)
    return packed, scales


def unpack_int4(packed: np.ndarray) -> np.ndarray:
    """[rows, cols/8] uint32 -> signed int8 [rows, cols] in [-8, 7]."""
    p = np.asarray(packed, dtype=np.uint32)
    shifts = (np.arange(PACK_FACTOR, dtype=np.uint32) * np.uint32(4))
    u4 = (p[:, :, None] >> shifts[None, None, :]) & np.uint32(0xF)
    return (u4.astype(np.int16) - 8).astype(np.int8).reshape(p.shape[0], -1)


def dequantize_int4(packed: np.ndarray, scales: np.ndarray, group_size: int = GROUP_SIZE) -> np.ndarray:
    """marlin.rs:210-234: (nib - 8) * scale in f32."""
    q = unpack_int4(packed).astype(np.float32)
    rows, cols = q.shape
    s = bf16_bits_to_f32(scales)
    return (q.reshape(rows, -1, group_size) * s[:, :, None]).astype(np.float32).reshape(rows, cols)


def quantize_int8(weight_bf16_bits: np.ndarray, group_size: int = GROUP_SIZE):
    """marlin.rs:65-114.  Returns (data int8 [rows, cols], scales raw BF16 [rows, groups])."""
    wb = np.asarray(weight_bf16_bits, dtype=np.uint16)
    rows, cols = wb.shape
    assert cols % group_size == 0
    w = bf16_bits_to_f32(wb)
    scales = _group_scales(w, group_size, 127.0)
    s = bf16_bits_to_f32(scales)
    inv = np.where(s == 0.0, np.float32(0.0), np.float32(1.0) / s).astype(np.float32)
    prod = (w.reshape(rows, -1, group_size) * inv[:, :, None]).astype(np.float32)
    q = np.clip(_round_half_away(prod), -128.0, 127.0).astype(np.int8).reshape(rows, cols)
    return q, scales


def dequantize_int8(data: np.ndarray, scales: np.ndarray, group_size: int = GROUP_SIZE) -> np.ndarray:
    """marlin.rs:117-133."""
    rows, cols = data.shape
    s = bf16_bits_to_f32(scales)
    return (data.astype(np.float32).reshape(rows, -1, group_size) * s[:, :, None]).astype(
        np.float32).reshape(rows, cols)


# ---------------------------------------------------------------------------
# Marlin GPU layout (needed only to exchange bytes with existing Krasis caches /
# the KrasisEngine.get_expert_* hand-off; the B200 kernels use their own tiling).
# ---------------------------------------------------------------------------

def marlin_weight_perm_int4() -> np.ndarray:
    """marlin.rs:256-295 (== vLLM get_weight_perm(num_bits=4)); dest -> src, 1024 entries."""
    perm = []
    for i in range(32):
        col = i // 4
        perm1 = []
        for block in (0, 1):
            for row in (2 * (i % 4), 2 * (i % 4) + 1, 2 * (i % 4 + 4), 2 * (i % 4 + 4) + 1):
                perm1.append(16 * row + col + 8 * block)
        for j in range(4):
            perm.extend(p + 256 * j for p in perm1)
    perm = np.asarray(perm, dtype=np.int64)
    interleave = np.array([0, 2, 4, 6, 1, 3, 5, 7])
    return perm.reshape(-1, 8)[:, interleave].reshape(-1)


def marlin_scale_perms():
    """marlin.rs:302-321."""
    scale_perm = np.array([i + 8 * j for i in range(8) for j in range(8)], dtype=np.int64)
    offs = [0, 1, 8, 9, 16, 17, 24, 25]
    single = np.array([2 * i + o for i in range(4) for o in offs], dtype=np.int64)
    return scale_perm, single


def marlin_repack_int4(packed: np.ndarray, scales: np.ndarray, group_size: int = GROUP_SIZE):
    """marlin.rs:330-491.  ([N, K/8] u32, [N, K/gs] bf16) -> ([K/16, 2N] u32, [K/gs, N] bf16)."""
    n = packed.shape[0]
    k = packed.shape[1] * 8
    assert k % 16 == 0 and n % 64 == 0
    u = (unpack_int4(packed).astype(np.int16) + 8).astype(np.uint8)   # [N, K] 0..15
    kn = u.T                                                          # [K, N]
    t = kn.reshape(k // 16, 16, n // 16, 16).transpose(0, 2, 1, 3).reshape(k // 16, n * 16)
    perm = marlin_weight_perm_int4()
    t = t.reshape(k // 16, -1, 1024)[:, :, perm].reshape(k // 16, n * 16)
    v = t.reshape(k // 16, 2 * n, 8).astype(np.uint32)
    shifts = (np.arange(8, dtype=np.uint32) * np.uint32(4))
    out = np.bitwise_or.reduce(v << shifts[None, None, :], axis=2).astype(np.uint32)
    st = np.ascontiguousarray(np.asarray(scales, dtype=np.uint16).T)  # [K/gs, N]
    sp, sp1 = marlin_scale_perms()
    p = sp if group_size < k else sp1
    st = st.reshape(-1, len(p))[:, p].reshape(k // group_size, n)
    return out, st


def marlin_unpack_int4(mpacked: np.ndarray, mscales: np.ndarray, group_size: int = GROUP_SIZE):
    """Inverse of marlin_repack_int4 (marlin.rs:496-576): back to ([N, K/8] u32, [N, K/gs] bf16)."""
    k16, n2 = mpacked.shape
    k, n = k16 * 16, n2 // 2
    shifts = (np.arange(8, dtype=np.uint32) * np.uint32(4))
    t = ((np.asarray(mpacked, np.uint32)[:, :, None] >> shifts) & np.uint32(0xF)).astype(np.uint8)
    t = t.reshape(k16, -1, 1024)
    perm = marlin_weight_perm_int4()
    inv = np.empty_like(t)
    inv[:, :, perm] = t
    kn = inv.reshape(k16, n // 16, 16, 16).transpose(0, 2, 1, 3).reshape(k, n)
    u = np.ascontiguousarray(kn.T).astype(np.uint32).reshape(n, k // 8, 8)
    packed = np.bitwise_or.reduce(u << shifts[None, None, :], axis=2).astype(np.uint32)
    sp, sp1 = marlin_scale_perms()
    p = sp if group_size < k else sp1
    s = np.asarray(mscales, np.uint16).reshape(-1, len(p))
    sinv = np.empty_like(s)
    sinv[:, p] = s
    scales = np.ascontiguousarray(sinv.reshape(k // group_size, n).T)
    return packed, scales


def marlin_weight_perm_int8() -> np.ndarray:
    """marlin.rs:586-625 (== vLLM get_weight_perm(num_bits=8)): the INT4 base permutation with the [0,2,1,3] interleave."""
    perm = []
    for i in range(32):
        col = i // 4
        perm1 = []
        for block in (0, 1):
            for row in (2 * (i % 4), 2 * (i % 4) + 1, 2 * (i % 4 + 4), 2 * (i % 4 + 4) + 1):
                perm1.append(16 * row + col + 8 * block)
        for j in range(4):
            perm.extend(p + 256 * j for p in perm1)
    perm = np.asarray(perm, dtype=np.int64)
    return perm.reshape(-1, 4)[:, [0, 2, 1, 3]].reshape(-1)


def marlin_repack_int8(data: np.ndarray, scales: np.ndarray, group_size: int = GROUP_SIZE):
    """marlin.rs:639-760.  ([N, K] i8, [N, K/gs] bf16) -> ([K/16, 4N] u32, [K/gs, N] bf16)."""
    n, k = data.shape
    assert k % 16 == 0 and n % 64 == 0
    u = (data.astype(np.int16) + 128).astype(np.uint8)                 # [N, K] 0..255
    t = u.T.reshape(k // 16, 16, n // 16, 16).transpose(0, 2, 1, 3).reshape(k // 16, n * 16)
    t = t.reshape(k // 16, -1, 1024)[:, :, marlin_weight_perm_int8()].reshape(k // 16, n * 16)
    v = t.reshape(k // 16, 4 * n, 4).astype(np.uint32)
    shifts = np.arange(4, dtype=np.uint32) * np.uint32(8)
    out = np.bitwise_or.reduce(v << shifts[None, None, :], axis=2).astype(np.uint32)
    st = np.ascontiguousarray(np.asarray(scales, dtype=np.uint16).T)
    sp, sp1 = marlin_scale_perms()
    p = sp if group_size < k else sp1
    st = st.reshape(-1, len(p))[:, p].reshape(k // group_size, n)
    return out, st

