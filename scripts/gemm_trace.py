"""Per-work-item timeline of the grouped expert GEMMs (tuning aid; run on the GPU box).

For CTA 0 and its first 16 work items prints (cycles): tmem wait, K loop length, waits of the MMA thread on the A (dequantised
weights) and B (token rows) operands, epilogue phase 1 / phase 2 lengths, and the waits of the dequantiser / producers.
Slots written by grouped_gemm.cu: 0 item start (MMA thread), 1 accumulators free, 2 last MMA committed, 3 sum wait a_full, 4 sum wait
b_full, 5 epilogue saw tmem_full, 6 TMEM drained, 7 stores done, 8 dequant sum wait w_full, 9 sum wait a_empty, 10 sum st+wait,
11 dequant item end, 12 weight producer sum wait w_empty, 13 token producer sum wait b_empty, 14 n_tok.
"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from krasis_b200 import KrasisEngine, capi

H, I, E, k, M = 2048, 512, 512, 10, 8192
bits = int(os.environ.get("BITS", "4"))
g = torch.Generator(device="cuda").manual_seed(1)
eng = KrasisEngine(hidden_size=H, moe_intermediate_size=I, n_routed_experts=E, num_experts_per_tok=k, num_moe_layers=1, num_bits=bits, max_tokens=M)
def rq(*shape):
    if bits == 4:
        return torch.randint(-2 ** 31, 2 ** 31 - 1, shape[:-1] + (shape[-1] // 8,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    return torch.randint(-128, 128, shape, dtype=torch.int64, device="cuda", generator=g).to(torch.int8)
def rs(*shape):
    return ((torch.rand(shape, device="cuda", generator=g) * 0.004 + 0.002).to(torch.bfloat16)).view(torch.int16)
eng.load_quantized_layer_dev(0, rq(E, 2 * I, H), rs(E, 2 * I, H // 128), rq(E, H, I), rs(E, H, I // 128))
rng = np.random.default_rng(0)
ids = torch.from_numpy(np.stack([rng.choice(E, k, replace=False) for _ in range(M)]).astype(np.int32)).cuda()
wts = torch.from_numpy(rng.dirichlet(np.ones(k), M).astype(np.float32)).cuda()
x = torch.randn(M, H, device="cuda", generator=g).to(torch.bfloat16)
for _ in range(3):
    eng.moe_forward(0, x, ids, wts, routed_only=True)
capi.kernel_profile(True)
for _ in range(5):
    eng.moe_forward(0, x, ids, wts, routed_only=True)
prof = capi.kernel_profile_collect(); capi.kernel_profile(False)
print(", ".join(f"{n} {t / c * 1e3:.1f}us" for n, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])))
for which in ("1", "2"):
    tr = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
    os.environ[f"KB2_GEMM{which}_TRACE"] = str(tr.data_ptr())
    eng.moe_forward(0, x, ids, wts, routed_only=True)
    torch.cuda.synchronize()
    del os.environ[f"KB2_GEMM{which}_TRACE"]
    t = tr.cpu().view(16, 16).numpy()
    print(f"GEMM{which}: item  n_tok | period  tmem_wait  k_loop  wait_A  wait_B | epi_lag  phase1  phase2 | dq_wait_w  dq_wait_a_empty  dq_st | prodW_wait prodB_wait")
    for i in range(1, 15):
        r = t[i]
        print(f"   {i:3d} {r[14]:5d} | {t[i + 1][0] - r[0]:6d} {r[1] - r[0]:9d} {r[2] - r[1]:7d} {r[3]:7d} {r[4]:7d} | {r[5] - r[2]:6d} {r[6] - r[5]:7d} {r[7] - r[6]:7d} |"
              f" {r[8]:8d} {r[9]:12d} {r[10]:8d} | {r[12]:8d} {r[13]:8d}")
eng.close()
