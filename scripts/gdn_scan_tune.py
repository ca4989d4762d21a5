"""Tuning harness for the tcgen05 GDN chunk scan (GPU): per-kernel times of one QCN-geometry GDN layer at 8192 tokens for
a clock64 timeline of one CTA (KB2_GDN_SCAN_TRACE) and of the chunk-prepare kernels (KB2_GDN_PREPARE_TRACE).
Scan slots: MMA warp 0 s_ready seen, 1 G1 issued, 2 v_ready seen, 3 G2+G3 issued; core thread 0: 4 g1_done seen, 6 g2_done seen,
7 S tiles written, 8 g3_done seen, 9 epilogue done; thread 64 (a VP row): 10 accumulators loaded, 11 v parked, 12 after the
256-thread barrier, 13 v tiles stored, 5 v_ready arrive done."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from krasis_b200 import capi  # noqa: E402
from krasis_b200.attention import GatedDeltaNetAttention  # noqa: E402

torch.manual_seed(0)
nk, nv, dk, dv, H, K, M = 16, 32, 128, 128, 256, 4, 8192
kd, vd = nk * dk, nv * dv
bf = torch.bfloat16
w = dict(in_proj_qkvz=(torch.randn(2 * kd + 2 * vd, H) * 0.15).to(bf), in_proj_ba=(torch.randn(2 * nv, H) * 0.15).to(bf),
         out_proj=(torch.randn(H, vd) * 0.05).to(bf), conv1d_weight=(torch.randn(2 * kd + vd, 1, K) * 0.5).to(bf),
         A_log=(torch.randn(nv) * 0.5).to(bf), dt_bias=(torch.randn(nv) * 0.5).to(bf), norm_weight=(1 + 0.1 * torch.randn(dv)).to(bf))
cfg = types.SimpleNamespace(hidden_size=H, linear_num_key_heads=nk, linear_num_value_heads=nv, linear_key_head_dim=dk,
                            linear_value_head_dim=dv, linear_conv_kernel_dim=K, rms_norm_eps=1e-6)
lay = GatedDeltaNetAttention(cfg, 0, w, "cuda:0", max_tokens=M)
x = torch.randn(M, H).to(bf).cuda()
for _ in range(3):
    lay.reset_state()
    lay.forward(x)
capi.kernel_profile(True)
for _ in range(5):
    lay.reset_state()
    y = lay.forward(x)
prof = capi.kernel_profile_collect()
capi.kernel_profile(False)
print("GDN layer: " + ", ".join(f"{n} {t / c * 1e3:.1f}us" for n, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]) if n.startswith("gdn")))
trace = torch.zeros(8 * 16, dtype=torch.int64, device="cuda")
os.environ["KB2_GDN_SCAN_TRACE"] = str(trace.data_ptr())
lay.reset_state()
lay.forward(x)
torch.cuda.synchronize()
del os.environ["KB2_GDN_SCAN_TRACE"]
t = trace.cpu().view(8, 16)
base = t[:, 0:1]
print("  scan timeline (cycles since the MMA thread saw s_ready), chunks 8..15:")
for r in (t - base).tolist():
    print("   ", r[:14])
print("  chunk period (cycles):", (t[1:, 0] - t[:-1, 0]).tolist())

# ---- tcgen05 chunk-prepare: time + per-phase timeline of CTA 0 (slots: MMA thread 0 inputs landed, 1 MMA-A issued, 2 images seen,
# 3 MMA-B/C issued; core thread 0: 4 gates scanned, 5 k.k^T seen, 6 A^T built, 7 T solved, 8 images written, 9 vcorr/kcd products seen,
# 10 outputs stored; version 2 only: 11 T zeroed, 12 diagonal 16x16 blocks inverted, 13 level 1 done)
for mode in ("1", "0"):
    os.environ["KB2_GDN_PREPARE_MMA_SYNC"] = mode
    for _ in range(2):
        lay.reset_state()
        lay.forward(x)
    capi.kernel_profile(True)
    for _ in range(5):
        lay.reset_state()
        lay.forward(x)
    prof = capi.kernel_profile_collect()
    capi.kernel_profile(False)
    print(("mma.sync" if mode == "1" else "tcgen05") + " prepare: " + ", ".join(f"{n} {t / c * 1e3:.1f}us" for n, (t, c) in prof.items() if "prepare" in n))
    if mode == "1":
        continue
    trace = torch.zeros(8 * 16, dtype=torch.int64, device="cuda")
    os.environ["KB2_GDN_PREPARE_TRACE"] = str(trace.data_ptr())
    lay.reset_state()
    lay.forward(x)
    torch.cuda.synchronize()
    del os.environ["KB2_GDN_PREPARE_TRACE"]
    t = trace.cpu().view(8, 16)[:, :14]
    print("  prepare timeline (cycles since inputs landed), loop iterations 2..9:")
    for r in (t - t[:, 0:1]).tolist()[:4]:
        print("   ", r)
    print("  unit period (cycles):", (t[1:, 0] - t[:-1, 0]).tolist())
