"""GGUF block formats: sizes, dequantisation, synthetic block generators
(oracle; test infrastructure only).

Restates the reference's element formulas:
  block sizes / bytes        src/gguf.rs:56-85
  get_scale_min_k4           src/gguf.rs:666-674 == src/gguf_kernels.rs:640-648
  Q8_0                       src/gguf.rs:574-593      d:f16 | i8[32]                w = d*q
  Q4_0                       src/gguf.rs:635-664      d:f16 | qs[16]                el 0-15 low nib, 16-31 high nib, (nib-8)*d
  Q5_0                       src/gguf.rs:599-633      d:f16 | qh:u32 | qs[16]       ((nib | bit<<4) - 16)*d
  Q4_K                       src/gguf.rs:681-738      d,dmin:f16 | sc[12] | qs[128] w = (d*sc)*q - (dmin*mn)
  Q5_K                       src/gguf.rs:744-805
  Q6_K                       src/gguf_kernels.rs:594-635 (authoritative, == ggml)
                             src/gguf.rs:813-866 uses sc[0/2/4/6] for all l — a suspected
                             reference bug; exposed here as dequant_q6_k_gguf_rs.
All products are evaluated in float32 in the order the Rust code uses.
"""
import numpy as np

# ggml type ids (src/gguf.rs:15-31)
GGML_F32, GGML_F16, GGML_Q4_0, GGML_Q5_0, GGML_Q8_0 = 0, 1, 2, 6, 8
GGML_Q4_K, GGML_Q5_K, GGML_Q6_K, GGML_BF16 = 12, 13, 14, 30

# (block elements, block bytes)  — src/gguf.rs:56-85
BLOCK = {
    GGML_F32: (1, 4), GGML_F16: (1, 2), GGML_BF16: (1, 2),
    GGML_Q4_0: (32, 18), 3: (32, 20), GGML_Q5_0: (32, 22), 7: (32, 24),
    GGML_Q8_0: (32, 34), 9: (32, 40),
    10: (256, 84), 11: (256, 110), GGML_Q4_K: (256, 144), GGML_Q5_K: (256, 176),
    GGML_Q6_K: (256, 210), 15: (256, 276),
}
NAMES = {GGML_Q4_0: "Q4_0", GGML_Q5_0: "Q5_0", GGML_Q8_0: "Q8_0", GGML_Q4_K: "Q4_K",
         GGML_Q5_K: "Q5_K", GGML_Q6_K: "Q6_K"}


def _f16(b: np.ndarray) -> np.ndarray:
    """[..., 2] uint8 little-endian -> float32."""
    return np.ascontiguousarray(b).view(np.float16).astype(np.float32)[..., 0]


def get_scale_min_k4(j: int, scales: np.ndarray):
    """src/gguf.rs:666-674.  scales [..., 12] uint8 -> (sc, mn) uint8 arrays."""
    s = scales
    if j < 4:
        return s[..., j] & 63, s[..., j + 4] & 63
    sc = (s[..., j + 4] & 0xF) | ((s[..., j - 4] >> 6) << 4)
    mn = (s[..., j + 4] >> 4) | ((s[..., j] >> 6) << 4)
    return sc, mn


def q4k_scales_mins(scales12: np.ndarray):
    """All eight (sc, mn) pairs: returns two [..., 8] uint8 arrays."""
    sc = np.stack([get_scale_min_k4(j, scales12)[0] for j in range(8)], axis=-1)
    mn = np.stack([get_scale_min_k4(j, scales12)[1] for j in range(8)], axis=-1)
    return sc.astype(np.uint8), mn.astype(np.uint8)


def dequant_q8_0(data: np.ndarray, n: int) -> np.ndarray:
    b = np.asarray(data, np.uint8)[: n // 32 * 34].reshape(-1, 34)
    d = _f16(b[:, 0:2])
    q = b[:, 2:34].view(np.int8).astype(np.float32)
    return (d[:, None] * q).astype(np.float32).reshape(-1)


def dequant_q4_0(data: np.ndarray, n: int) -> np.ndarray:
    b = np.asarray(data, np.uint8)[: n // 32 * 18].reshape(-1, 18)
    d = _f16(b[:, 0:2])
    qs = b[:, 2:18]
    nib = np.concatenate([qs & 0xF, qs >> 4], axis=1).astype(np.int32) - 8
    return (d[:, None] * nib.astype(np.float32)).astype(np.float32).reshape(-1)


def dequant_q5_0(data: np.ndarray, n: int) -> np.ndarray:
    b = np.asarray(data, np.uint8)[: n // 32 * 22].reshape(-1, 22)
    d = _f16(b[:, 0:2])
    qh = np.ascontiguousarray(b[:, 2:6]).view(np.uint32)[:, 0]
    qs = b[:, 6:22]
    q4 = np.concatenate([qs & 0xF, qs >> 4], axis=1).astype(np.int32)
    bit = ((qh[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.int32)
    q = (q4 | (bit << 4)) - 16
    return (d[:, None] * q.astype(np.float32)).astype(np.float32).reshape(-1)


def dequant_q4_k(data: np.ndarray, n: int) -> np.ndarray:
    b = np.asarray(data, np.uint8)[: n // 256 * 144].reshape(-1, 144)
    d, dmin = _f16(b[:, 0:2]), _f16(b[:, 2:4])
    sc, mn = q4k_scales_mins(b[:, 4:16])
    qs = b[:, 16:144].reshape(-1, 4, 32)
    q = np.stack([qs & 0xF, qs >> 4], axis=2).reshape(-1, 8, 32).astype(np.float32)
    d1 = (d[:, None] * sc.astype(np.float32)).astype(np.float32)
    m1 = (dmin[:, None] * mn.astype(np.float32)).astype(np.float32)
    return ((d1[:, :, None] * q).astype(np.float32) - m1[:, :, None]).astype(np.float32).reshape(-1)


def dequant_q5_k(data: np.ndarray, n: int) -> np.ndarray:
    b = np.asarray(data, np.uint8)[: n // 256 * 176].reshape(-1, 176)
    d, dmin = _f16(b[:, 0:2]), _f16(b[:, 2:4])
    sc, mn = q4k_scales_mins(b[:, 4:16])
    qh = b[:, 16:48]
    qs = b[:, 48:176].reshape(-1, 4, 32)
    q4 = np.stack([qs & 0xF, qs >> 4], axis=2).reshape(-1, 8, 32).astype(np.int32)
    hb = ((qh[:, None, :] >> np.arange(8, dtype=np.uint8)[None, :, None]) & 1).astype(np.int32)
    q = (q4 + 16 * hb).astype(np.float32)
    d1 = (d[:, None] * sc.astype(np.float32)).astype(np.float32)
    m1 = (dmin[:, None] * mn.astype(np.float32)).astype(np.float32)
    return ((d1[:, :, None] * q).astype(np.float32) - m1[:, :, None]).astype(np.float32).reshape(-1)


def _q6k_fields(data: np.ndarray, n: int):
    b = np.asarray(data, np.uint8)[: n // 256 * 210].reshape(-1, 210)
    ql = b[:, 0:128].reshape(-1, 2, 64)
    qh = b[:, 128:192].reshape(-1, 2, 32)
    sc = b[:, 192:208].view(np.int8).reshape(-1, 2, 8).astype(np.float32)
    d = _f16(b[:, 208:210])
    q0 = (ql[:, :, 0:32] & 0xF) | (((qh >> 0) & 3) << 4)
    q1 = (ql[:, :, 32:64] & 0xF) | (((qh >> 2) & 3) << 4)
    q2 = (ql[:, :, 0:32] >> 4) | (((qh >> 4) & 3) << 4)
    q3 = (ql[:, :, 32:64] >> 4) | (((qh >> 6) & 3) << 4)
    q = np.stack([q0, q1, q2, q3], axis=2).astype(np.int32) - 32     # [nb, 2, 4, 32]
    return d, sc, q


def dequant_q6_k(data: np.ndarray, n: int) -> np.ndarray:
    """ggml-correct Q6_K (src/gguf_kernels.rs:594-635): scale index = is + 2*sub, is = l/16."""
    d, sc, q = _q6k_fields(data, n)
    l = np.arange(32)
    idx = (l // 16)[None, :] + 2 * np.arange(4)[:, None]            # [4, 32]
    s = sc[:, :, idx]                                               # [nb, 2, 4, 32]
    ds = (d[:, None, None, None] * s).astype(np.float32)
    return (ds * q.astype(np.float32)).astype(np.float32).reshape(-1)


def dequant_q6_k_gguf_rs(data: np.ndarray, n: int) -> np.ndarray:
    """Reference src/gguf.rs:813-866 verbatim semantics (sc[0/2/4/6] for every l)."""
    d, sc, q = _q6k_fields(data, n)
    s = sc[:, :, [0, 2, 4, 6]][:, :, :, None]
    ds = (d[:, None, None, None] * s).astype(np.float32)
    return (ds * q.astype(np.float32)).astype(np.float32).reshape(-1)


_DEQ = {GGML_Q8_0: dequant_q8_0, GGML_Q4_0: dequant_q4_0, GGML_Q5_0: dequant_q5_0,
        GGML_Q4_K: dequant_q4_k, GGML_Q5_K: dequant_q5_k, GGML_Q6_K: dequant_q6_k}


def dequantize(ggml_type: int, data: np.ndarray, n_elements: int) -> np.ndarray:
    """src/gguf.rs:871-886 dispatch (Q6_K uses the ggml-correct variant)."""
    if ggml_type == GGML_F32:
        return np.asarray(data, np.uint8)[: n_elements * 4].view(np.float32).copy()
    if ggml_type == GGML_F16:
        return np.asarray(data, np.uint8)[: n_elements * 2].view(np.float16).astype(np.float32)
    if ggml_type == GGML_BF16:
        u = np.asarray(data, np.uint8)[: n_elements * 2].view(np.uint16).astype(np.uint32) << 16
        return u.view(np.float32)
    return _DEQ[ggml_type](data, n_elements)


def row_bytes(ggml_type: int, k: int) -> int:
    be, bb = BLOCK[ggml_type]
    assert k % be == 0, f"K={k} not a multiple of block size {be}"
    return k // be * bb


# ---------------------------------------------------------------------------
# Synthetic tensors (gguf-py cannot *quantise* K-quants, so
# Q4_K test tensors are random blocks with sane fp16 d/dmin).
# ---------------------------------------------------------------------------

def random_q4_k(rng: np.random.Generator, rows: int, k: int, d_range=(1e-3, 1e-2)) -> np.ndarray:
    nb = rows * k // 256
    b = np.empty((nb, 144), np.uint8)
    d = rng.uniform(*d_range, nb).astype(np.float16)
    dm = rng.uniform(*d_range, nb).astype(np.float16)
    b[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
    b[:, 2:4] = dm.view(np.uint8).reshape(nb, 2)
    b[:, 4:144] = rng.integers(0, 256, (nb, 140), dtype=np.uint8)
    return b.reshape(rows, -1)


def quantize_q8_0(w: np.ndarray) -> np.ndarray:
    """ggml reference Q8_0 quantiser (d = amax/127, q = round(w/d)); rows of f32 -> bytes."""
    rows, k = w.shape
    g = np.asarray(w, np.float32).reshape(rows, k // 32, 32)
    amax = np.abs(g).max(axis=2)
    d = (amax / np.float32(127.0)).astype(np.float32)
    inv = np.where(d == 0, np.float32(0), np.float32(1) / d).astype(np.float32)
    q = np.rint(g * inv[:, :, None]).clip(-127, 127).astype(np.int8)
    out = np.empty((rows, k // 32, 34), np.uint8)
    out[:, :, 0:2] = d.astype(np.float16).view(np.uint8).reshape(rows, k // 32, 2)
    out[:, :, 2:] = q.view(np.uint8)
    return out.reshape(rows, -1)


def random_blocks(rng: np.random.Generator, ggml_type: int, rows: int, k: int) -> np.ndarray:
    """Random but well-formed blocks for any supported type: [rows, row_bytes] uint8."""
    if ggml_type == GGML_Q4_K:
        return random_q4_k(rng, rows, k)
    be, bb = BLOCK[ggml_type]
    nb = rows * k // be
    b = rng.integers(0, 256, (nb, bb), dtype=np.uint8)
    d = rng.uniform(1e-3, 1e-2, nb).astype(np.float16).view(np.uint8).reshape(nb, 2)
    if ggml_type == GGML_Q6_K:
        b[:, 208:210] = d
        b[:, 192:208] = rng.integers(-64, 64, (nb, 16)).astype(np.int8).view(np.uint8)
    elif ggml_type == GGML_Q5_K:
        b[:, 0:2] = d
        b[:, 2:4] = rng.uniform(1e-3, 1e-2, nb).astype(np.float16).view(np.uint8).reshape(nb, 2)
    else:
        b[:, 0:2] = d
    return b.reshape(rows, -1)
