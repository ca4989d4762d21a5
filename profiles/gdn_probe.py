"""One Gated-DeltaNet layer at Qwen3-Coder-Next geometry (H=2048, nk=16, nv=32, dk=dv=128, conv 4), M=8192 prefill.
Usage: python profiles/gdn_probe.py [iters]   — prints CUDA-event ms per forward; run under ncu for the launch list."""
import sys
import types

import torch

sys.path.insert(0, ".")
from krasis_b200.attention import GatedDeltaNetAttention  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(0)
nk, nv, dk, dv, H, K, M = 16, 32, 128, 128, 2048, 4, 8192
kd, vd = nk * dk, nv * dv
bf = torch.bfloat16
w = dict(in_proj_qkvz=(torch.randn(2 * kd + 2 * vd, H) * 0.02).to(bf), in_proj_ba=(torch.randn(2 * nv, H) * 0.02).to(bf),
         out_proj=(torch.randn(H, vd) * 0.02).to(bf), conv1d_weight=(torch.randn(2 * kd + vd, 1, K) * 0.5).to(bf),
         A_log=(torch.randn(nv) * 0.5).to(bf), dt_bias=(torch.randn(nv) * 0.5).to(bf), norm_weight=torch.ones(dv).to(bf))
cfg = types.SimpleNamespace(hidden_size=H, linear_num_key_heads=nk, linear_num_value_heads=nv, linear_key_head_dim=dk,
                            linear_value_head_dim=dv, linear_conv_kernel_dim=K, rms_norm_eps=1e-6)
lay = GatedDeltaNetAttention(cfg, 0, w, "cuda:0", max_tokens=M)
x = torch.randn(M, H, device="cuda").to(bf)
lay.forward(x, is_decode=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    lay.reset_state()
    y = lay.forward(x, is_decode=False)
e1.record()
torch.cuda.synchronize()
print(f"gdn layer forward: {e0.elapsed_time(e1) / iters:.3f} ms  (finite={bool(torch.isfinite(y.float()).all())})")
