"""Host-side (numpy) statement of the KB2 tile layout (krasis_b200/csrc/moe_common.cuh) and its inverse.

The device does the re-tiling at load time (repack_*_kernel); this module exists for the hand-off in the OTHER direction:
serving the reference's `get_expert_*` / `write_experts_*_into` calls (src/moe.rs:1972-2481), which want the quantiser's
arrays (and from there the Marlin order, krasis_b200/marlin_cache.py) back out of an engine that only keeps tiles.  Pure byte
permutations; tests check them bit for bit against the device kernels.
"""
import numpy as np

_NIB_ORDER = np.array([0, 2, 4, 6, 1, 3, 5, 7])


def untile_int4(wq: np.ndarray, ws: np.ndarray, n_experts: int, N: int, K: int):
    """tiles (uint8 buffers of n_experts experts) -> (packed [E, N, K/8] u32, scales [E, N, K/128] u16) as quantize_int4 emits."""
    q = wq.view(np.uint32).reshape(n_experts, N // 128, K // 64, 2, 128, 4)             # [E][tile][kb][half][row][word]
    q = q.transpose(0, 1, 4, 2, 3, 5).reshape(n_experts, N, K // 8)                      # [E][row][kb, half, word]
    nib = (q[..., None] >> (np.arange(8, dtype=np.uint32) * 4)) & 0xF
    inv = np.argsort(_NIB_ORDER)
    packed = np.bitwise_or.reduce(nib[..., inv] << (np.arange(8, dtype=np.uint32) * 4), axis=-1).astype(np.uint32)
    s = ws.view(np.uint16).reshape(n_experts, N // 128, K // 128, 128).transpose(0, 1, 3, 2).reshape(n_experts, N, K // 128)
    return packed, np.ascontiguousarray(s)


def untile_int8(wq: np.ndarray, ws: np.ndarray, n_experts: int, N: int, K: int):
    """INT8 tiles: blob (128 rows x 64 K) = [quarter 0..3][row][16 B] -> ([E, N, K] i8, scales [E, N, K/128] u16)."""
    q = wq.view(np.int8).reshape(n_experts, N // 128, K // 64, 4, 128, 16)               # [E][tile][kb][quarter][row][16]
    q = q.transpose(0, 1, 4, 2, 3, 5).reshape(n_experts, N, K)
    s = ws.view(np.uint16).reshape(n_experts, N // 128, K // 128, 128).transpose(0, 1, 3, 2).reshape(n_experts, N, K // 128)
    return np.ascontiguousarray(q), np.ascontiguousarray(s)
