"""ctypes wrapper of oracle/cpu_moe.c (the C/AVX2 port of the reference's CPU expert path).
Test infrastructure + timed CPU baseline only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libkrasis_cpu_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.kcpu_moe_forward_int4.restype = C.c_int
        _lib.kcpu_moe_forward_int4.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] * 2 + [
            C.c_void_p, C.c_int]
        _lib.kcpu_moe_forward_gguf.restype = C.c_int
        _lib.kcpu_moe_forward_gguf.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p] * 3 + [C.c_int] * 2 + [C.c_void_p, C.c_int]
        _lib.kcpu_num_threads.restype = C.c_int
    return _lib


def to_unified(packed, scales):
    """[E][N][K/8] u32, [E][N][K/gs] bf16 (quantiser rows) -> CPU unified transposed
    [E][K/8][N], [E][K/gs][N] (src/weights/mod.rs:329-397)."""
    return (np.ascontiguousarray(np.transpose(packed, (0, 2, 1))),
            np.ascontiguousarray(np.transpose(scales, (0, 2, 1))))


def moe_forward_int4(w13_u, s13_u, w2_u, s2_u, x_bf16_bits, ids, wts, gs=128, nthreads=0):
    lib = load()
    E, H8, I2 = w13_u.shape
    H, I = H8 * 8, I2 // 2
    M, k = ids.shape
    out = np.empty((M, H), np.float32)
    ids = np.ascontiguousarray(ids, np.int32)
    wts = np.ascontiguousarray(wts, np.float32)
    x = np.ascontiguousarray(x_bf16_bits, np.uint16)
    rc = lib.kcpu_moe_forward_int4(w13_u.ctypes.data, s13_u.ctypes.data, w2_u.ctypes.data, s2_u.ctypes.data,
                                   E, H, I, gs, x.ctypes.data, ids.ctypes.data, wts.ctypes.data, M, k,
                                   out.ctypes.data, nthreads)
    assert rc == 0
    return out


def moe_forward_gguf(gate, up, down, gate_up_type, down_type, H, I, x_bf16_bits, ids, wts, nthreads=0):
    """gate/up uint8 [E, I, row_bytes(H)], down uint8 [E, H, row_bytes(I)] raw GGUF blocks (ggml type ids 8 = Q8_0, 12 = Q4_K)."""
    lib = load()
    M, k = ids.shape
    out = np.empty((M, H), np.float32)
    gate, up, down = (np.ascontiguousarray(a, np.uint8) for a in (gate, up, down))
    ids = np.ascontiguousarray(ids, np.int32)
    wts = np.ascontiguousarray(wts, np.float32)
    x = np.ascontiguousarray(x_bf16_bits, np.uint16)
    rc = lib.kcpu_moe_forward_gguf(gate.ctypes.data, up.ctypes.data, down.ctypes.data, int(gate_up_type), int(down_type),
                                   gate.shape[0], H, I, x.ctypes.data, ids.ctypes.data, wts.ctypes.data, M, k, out.ctypes.data, nthreads)
    if rc != 0:
        raise ValueError("unsupported GGUF type for the CPU INT16 path")
    return out


def num_threads():
    return load().kcpu_num_threads()
