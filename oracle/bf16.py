"""BF16 <-> F32 helpers (oracle; test infrastructure only).

Restates reference src/weights/marlin.rs:19-30:
  bf16_to_f32(v) = f32::from_bits(v << 16)
  f32_to_bf16(v) = (bits + 0x7FFF + ((bits >> 16) & 1)) >> 16      (round-nearest-even)
"""
import numpy as np


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    """uint16 raw BF16 -> float32 (marlin.rs:19-21)."""
    b = np.asarray(bits, dtype=np.uint16).astype(np.uint32) << np.uint32(16)
    return b.view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> uint16 raw BF16, round to nearest even (marlin.rs:25-30).

    Like the reference, NaN handling is whatever the integer add produces.
    """
    bits = np.ascontiguousarray(np.asarray(x, dtype=np.float32)).view(np.uint32)
    rnd = bits + (np.uint32(0x7FFF) + ((bits >> np.uint32(16)) & np.uint32(1)))
    return (rnd >> np.uint32(16)).astype(np.uint16)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round float32/float64 values to the nearest BF16, returned as float32."""
    return bf16_bits_to_f32(f32_to_bf16_bits(np.asarray(x, dtype=np.float32)))


def f16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return np.asarray(bits, dtype=np.uint16).view(np.float16).astype(np.float32)
