"""Dense / elementwise building blocks of the prefill layer loop behind the reference's function names
(python/krasis/weight_loader.py: quantize_to_int8, int8_linear; python/krasis/layer.py: norms, shared expert).
Every function launches libkrasis_b200 kernels; torch only owns the buffers."""
from typing import Optional, Tuple

import torch

from . import capi


def _s(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _dev(t):
    return t.device.index or 0


def _chk_bf16(t, name):
    if not t.is_cuda or t.dtype != torch.bfloat16 or t.dim() != 2 or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous 2-D CUDA bf16 tensor (krasis_b200 has no CPU path)")


def rmsnorm(x: torch.Tensor, weight_f32: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """flashinfer.norm.rmsnorm(x, w, eps) (layer.py:163-165)."""
    _chk_bf16(x, "x")
    out = torch.empty_like(x) if out is None else out
    capi.check(capi.load().kb2_rmsnorm(x.data_ptr(), None, weight_f32.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                                       float(eps), _dev(x), _s(x)))
    return out


def fused_add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight_f32: torch.Tensor, eps: float) -> None:
    """flashinfer.norm.fused_add_rmsnorm — IN PLACE: residual += x; x = rmsnorm(residual) * w (layer.py:283-285)."""
    _chk_bf16(x, "x")
    _chk_bf16(residual, "residual")
    capi.check(capi.load().kb2_rmsnorm(x.data_ptr(), residual.data_ptr(), weight_f32.data_ptr(), x.data_ptr(), x.shape[0],
                                       x.shape[1], float(eps), _dev(x), _s(x)))


def fused_add_rmsnorm_q8(x: torch.Tensor, residual: torch.Tensor, weight_f32: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """fused_add_rmsnorm that also returns the INT8 row quantisation (q int8 [M,H], scale f32 [M]) of the normed rows — exactly what
    int8_linear (weight_loader.py:46-99) would compute from them, so the shared expert's first GEMM skips its quantisation pass."""
    _chk_bf16(x, "x")
    _chk_bf16(residual, "residual")
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    qs = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    capi.check(capi.load().kb2_rmsnorm_q8(x.data_ptr(), residual.data_ptr(), weight_f32.data_ptr(), x.data_ptr(), q.data_ptr(), qs.data_ptr(),
                                          x.shape[0], x.shape[1], float(eps), _dev(x), _s(x)))
    return q, qs


def sum_slots(slots_ptr: int, world: int, rows: int, hidden: int, device: torch.device) -> torch.Tensor:
    """bf16(sum over the `world` partial rows of every token) from a peer-filled receive buffer [rows][world][hidden] bf16."""
    out = torch.empty((rows, hidden), dtype=torch.bfloat16, device=device)
    capi.check(capi.load().kb2_sum_slots_bf16(slots_ptr, world, out.data_ptr(), rows, hidden, device.index or 0,
                                              torch.cuda.current_stream(device).cuda_stream))
    return out


def quantize_to_int8(weight_bf16: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """weight_loader.py:25-43: per-row symmetric INT8; returns (int8 [N,K], scale bf16 [N])."""
    _chk_bf16(weight_bf16, "weight")
    n, k = weight_bf16.shape
    q = torch.empty((n, k), dtype=torch.int8, device=weight_bf16.device)
    s = torch.empty(n, dtype=torch.bfloat16, device=weight_bf16.device)
    capi.check(capi.load().kb2_quantize_rows_int8(weight_bf16.data_ptr(), q.data_ptr(), None, s.data_ptr(), n, k,
                                                  _dev(weight_bf16), _s(weight_bf16)))
    return q, s


def int8_linear(x: torch.Tensor, weight_int8: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """weight_loader.py:46-99 (W8A8, bias-free)."""
    _chk_bf16(x, "x")
    m, k = x.shape
    n = weight_int8.shape[0]
    out = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
    xq = torch.empty((m, k), dtype=torch.int8, device=x.device)
    xs = torch.empty(m, dtype=torch.float32, device=x.device)
    capi.check(capi.load().kb2_int8_linear(x.data_ptr(), weight_int8.data_ptr(), scale.data_ptr(), out.data_ptr(),
                                           xq.data_ptr(), xs.data_ptr(), m, n, k, _dev(x), _s(x)))
    return out


def int8_linear_q8(xq: torch.Tensor, xs: torch.Tensor, weight_int8: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """int8_linear on an activation that is already row-quantised (xq int8 [M,K], xs f32 [M])."""
    m, k = xq.shape
    n = weight_int8.shape[0]
    out = torch.empty((m, n), dtype=torch.bfloat16, device=xq.device)
    capi.check(capi.load().kb2_int8_linear_q8(xq.data_ptr(), xs.data_ptr(), weight_int8.data_ptr(), scale.data_ptr(), out.data_ptr(),
                                              m, n, k, _dev(xq), _s(xq)))
    return out


def silu_mul_int8_linear(gate_up: torch.Tensor, weight_int8: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """int8_linear(silu(gate) * up, W) with the activation quantised straight from the [M, 2K] gate|up rows (bit-identical to
    silu_and_mul followed by int8_linear)."""
    _chk_bf16(gate_up, "gate_up")
    m, k = gate_up.shape[0], gate_up.shape[1] // 2
    n = weight_int8.shape[0]
    out = torch.empty((m, n), dtype=torch.bfloat16, device=gate_up.device)
    xq = torch.empty((m, k), dtype=torch.int8, device=gate_up.device)
    xs = torch.empty(m, dtype=torch.float32, device=gate_up.device)
    act = None if k in (256, 512, 1024, 2048) else torch.empty((m, k), dtype=torch.bfloat16, device=gate_up.device)
    capi.check(capi.load().kb2_silu_mul_int8_linear(gate_up.data_ptr(), weight_int8.data_ptr(), scale.data_ptr(), out.data_ptr(),
                                                    act.data_ptr() if act is not None else None, xq.data_ptr(), xs.data_ptr(), m, n, k,
                                                    _dev(gate_up), _s(gate_up)))
    return out


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    _chk_bf16(x, "x")
    out = torch.empty((x.shape[0], x.shape[1] // 2), dtype=torch.bfloat16, device=x.device)
    capi.check(capi.load().kb2_silu_and_mul(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1] // 2, _dev(x), _s(x)))
    return out


class SharedExpert:
    """TransformerLayer._shared_expert_forward (layer.py:508-524): INT8 gate_up / down + optional sigmoid gate."""

    def __init__(self, gate_up_proj_bf16: torch.Tensor, down_proj_bf16: torch.Tensor, shared_expert_gate: Optional[torch.Tensor] = None):
        self.gate_up = quantize_to_int8(gate_up_proj_bf16.contiguous())
        self.down = quantize_to_int8(down_proj_bf16.contiguous())
        self.gate = shared_expert_gate.reshape(-1).contiguous() if shared_expert_gate is not None else None

    def forward(self, hidden: torch.Tensor, hidden_q8=None) -> torch.Tensor:
        """hidden_q8 = (q, scale) from fused_add_rmsnorm_q8 when the caller's norm already quantised these rows."""
        gu = int8_linear_q8(hidden_q8[0], hidden_q8[1], *self.gate_up) if hidden_q8 is not None else int8_linear(hidden, *self.gate_up)
        out = silu_mul_int8_linear(gu, *self.down)
        if self.gate is not None:
            capi.check(capi.load().kb2_sigmoid_gate_mul(hidden.data_ptr(), self.gate.data_ptr(), out.data_ptr(), hidden.shape[0],
                                                         hidden.shape[1], out.shape[1], _dev(hidden), _s(hidden)))
        return out
