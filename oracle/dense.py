"""Oracle restatements of the dense / elementwise pieces around the hot path (torch CPU; test infrastructure only).

  quantize_to_int8, int8_linear   python/krasis/weight_loader.py:25-43, 46-99 (executed literally; torch._int_mm is
                                   replaced by an exact int64 matmul, which is what INT32 accumulation computes)
  rmsnorm / fused_add_rmsnorm     flashinfer.norm semantics as called from python/krasis/layer.py:163-183,283-308
  silu_and_mul                    flashinfer.activation.silu_and_mul (layer.py:502,512)
  shared_expert_forward           python/krasis/layer.py:508-524
"""
import torch


def quantize_to_int8(weight_bf16):
    w = weight_bf16.float()
    amax = w.abs().amax(dim=1).clamp(min=1e-10)
    scale = amax / 127.0
    w_int8 = (w / scale.unsqueeze(1)).round().clamp(-128, 127).to(torch.int8)
    return w_int8, scale.to(torch.bfloat16)


def int8_linear(x, weight_int8, scale):
    x_float = x.float()
    x_amax = x_float.abs().amax(dim=1, keepdim=True).clamp(min=1e-10)
    x_scale = x_amax / 127.0
    x_int8 = (x_float / x_scale).round().clamp(-128, 127).to(torch.int8)
    out_int32 = (x_int8.to(torch.int64) @ weight_int8.to(torch.int64).t()).to(torch.int32)   # exact, == torch._int_mm
    out = out_int32.float() * (x_scale * scale.float().unsqueeze(0))
    return out.to(torch.bfloat16)


def rmsnorm(x, weight, eps):
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * weight.float()).to(x.dtype)


def fused_add_rmsnorm(x, residual, weight, eps):
    """Returns (normed, new_residual): s = x + residual in fp32; residual <- bf16(s); normed <- bf16(rmsnorm(s) * w)."""
    s = x.float() + residual.float()
    normed = (s * torch.rsqrt(s.pow(2).mean(-1, keepdim=True) + eps) * weight.float()).to(x.dtype)
    return normed, s.to(x.dtype)


def silu_and_mul(x):
    n = x.shape[-1] // 2
    g, u = x[..., :n].float(), x[..., n:].float()
    return (g / (1.0 + torch.exp(-g)) * u).to(x.dtype)


def shared_expert_forward(hidden, gate_up, down, gate_w=None):
    """layer.py:508-524 with INT8 weights: gate_up/down = (int8 [N,K], scale bf16 [N]); gate_w [1,H] bf16 or None."""
    act = silu_and_mul(int8_linear(hidden, *gate_up))
    out = int8_linear(act, *down)
    if gate_w is not None:
        out = torch.sigmoid(torch.nn.functional.linear(hidden, gate_w)) * out
    return out
