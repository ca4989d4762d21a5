"""Attention blocks behind the reference's call surface (python/krasis/linear_attention.py, attention.py).

  GatedDeltaNetAttention(cfg, layer_idx, weights, device).forward(hidden, is_decode) / reset_state()
      mirrors python/krasis/linear_attention.py:118-214,393-470 for the prefill path (M > 1, chunked).
  linear(x, weight)  — torch.nn.functional.linear for BF16 weights on tcgen05 (attention.py:526-529,672).
All arithmetic runs inside libkrasis_b200.so; torch supplies tensors and streams.
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import capi


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _bf16_host(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        return t.detach().to(torch.bfloat16).cpu().contiguous().view(torch.int16).numpy().view(np.uint16)
    return np.ascontiguousarray(t, dtype=np.uint16)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16):
    """out = x @ weight.T (+ bias): x [M,K] bf16 cuda, weight [N,K] bf16 cuda."""
    if not (x.is_cuda and weight.is_cuda) or x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("linear: x and weight must be CUDA bf16 tensors (krasis_b200 has no CPU path)")
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1] or not x.is_contiguous() or not weight.is_contiguous():
        raise ValueError(f"linear: shape mismatch {tuple(x.shape)} x {tuple(weight.shape)}")
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    b = bias.float().contiguous() if bias is not None else None
    capi.check(capi.load().kb2_linear_bf16(x.data_ptr(), weight.data_ptr(), b.data_ptr() if b is not None else None,
                                           out.data_ptr(), M, N, K, int(out_dtype == torch.float32),
                                           x.device.index or 0, _stream(x.device)))
    return out


class GatedDeltaNetAttention:
    """Drop-in for python/krasis/linear_attention.py:GatedDeltaNetAttention on the prefill path.

    `cfg` needs: hidden_size, linear_num_key_heads, linear_num_value_heads, linear_key_head_dim,
    linear_value_head_dim, linear_conv_kernel_dim, rms_norm_eps (python/krasis/linear_attention.py:148-158).
    `weights` keys as in the reference (:167-176): in_proj_qkvz, in_proj_ba, out_proj, conv1d_weight, A_log,
    dt_bias, norm_weight — BF16 tensors (INT8 attention weights are disabled in the reference, config.py:209)."""

    def __init__(self, cfg, layer_idx: int, weights: dict, device, max_tokens: int = 8192):
        self.cfg, self.layer_idx = cfg, layer_idx
        self.device = torch.device(device)
        self._lib = capi.load()
        for k, v in weights.items():
            if isinstance(v, tuple):
                raise ValueError(f"{k}: INT8 attention weights are not supported (the reference disables them, config.py:209)")
        c = capi.GdnConfig(cfg.hidden_size, cfg.linear_num_key_heads, cfg.linear_num_value_heads,
                           cfg.linear_key_head_dim, cfg.linear_value_head_dim, cfg.linear_conv_kernel_dim,
                           float(cfg.rms_norm_eps), max_tokens, 1, self.device.index or 0)
        self._h = C.c_void_p()
        capi.check(self._lib.kb2_gdn_create(C.byref(c), C.byref(self._h)))
        arrs = [_bf16_host(weights[k]) for k in ("in_proj_qkvz", "in_proj_ba", "conv1d_weight", "A_log", "dt_bias",
                                                 "norm_weight", "out_proj")]
        capi.check(self._lib.kb2_gdn_set_weights_host(self._h, 0, *[a.ctypes.data for a in arrs]))
        self._c = c

    def __del__(self):
        try:
            if self._h:
                self._lib.kb2_gdn_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def reset_state(self):
        capi.check(self._lib.kb2_gdn_reset_state(self._h, 0, _stream(self.device)))

    def forward(self, hidden: torch.Tensor, is_decode: bool = False) -> torch.Tensor:
        if is_decode:
            raise NotImplementedError("M=1 recurrent decode is out of scope (SURVEY.md §8: prefill path only)")
        if not hidden.is_cuda or hidden.dtype != torch.bfloat16 or hidden.dim() != 2 or not hidden.is_contiguous() \
                or hidden.shape[1] != self._c.hidden_size:
            raise ValueError(f"hidden: expected contiguous CUDA bf16 [M, {self._c.hidden_size}]")
        out = torch.empty_like(hidden)
        capi.check(self._lib.kb2_gdn_forward(self._h, 0, hidden.data_ptr(), out.data_ptr(), hidden.shape[0],
                                             _stream(hidden.device)))
        return out

    def state(self):
        """(conv_state [C,K] bf16-as-float32, recurrent_state [nv,dk,dv] float32) on host — tests only."""
        c = self._c
        Cc = 2 * c.num_k_heads * c.k_head_dim + c.num_v_heads * c.v_head_dim
        conv = np.zeros((Cc, c.conv_kernel), np.uint16)
        rec = np.zeros((c.num_v_heads, c.k_head_dim, c.v_head_dim), np.float32)
        capi.check(self._lib.kb2_gdn_get_state_host(self._h, 0, conv.ctypes.data, rec.ctypes.data))
        return (conv.astype(np.uint32) << 16).view(np.float32), rec
