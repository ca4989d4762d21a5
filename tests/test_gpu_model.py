"""End-to-end prefill of a small hybrid (Gated DeltaNet + GQA) MoE model through krasis_b200.model.KrasisModel
against a forward assembled from the oracle pieces.  Bar (SURVEY.md A.6 / tests/test_prefill_vs_decode.py:124-131):
last-token logits cosine >= 0.999 and greedy match."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import attention as A, dense as D, moe as omoe, router  # noqa: E402
from tests.test_gpu_moe import untile_expert_int4  # noqa: E402


def test_small_hybrid_model_prefill_matches_oracle():
    from krasis_b200.model import HybridMoEConfig, KrasisModel
    cfg = HybridMoEConfig(hidden_size=256, num_hidden_layers=4, full_attention_interval=4, vocab_size=512,
                          n_routed_experts=8, num_experts_per_tok=2, moe_intermediate_size=128,
                          shared_expert_intermediate_size=128, num_attention_heads=4, num_key_value_heads=2,
                          gqa_head_dim=128, partial_rotary_factor=0.5, rope_theta=10000.0,
                          linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=32, linear_value_head_dim=32,
                              synthetic_router_std=1.0)   # decisive router: near-ties (legitimately implementation-dependent) become rare
    M = 200
    model = KrasisModel(cfg, device=0, max_tokens=M, keep_weights=True)
    assert model.layer_types == ["linear_attention"] * 3 + ["full_attention"]
    assert model._kv_layer_offsets == [-1, -1, -1, 0]
    g = torch.Generator().manual_seed(9)
    tok = torch.randint(0, cfg.vocab_size, (M,), generator=g)
    pos = torch.arange(M)
    logits = model.forward(tok.cuda(), pos.cuda(), model.new_sequence()).cpu()
    all_logits = model.forward(tok.cuda(), pos.cuda(), model.new_sequence(), return_all_logits=True).cpu()
    assert logits.shape == (1, cfg.vocab_size) and all_logits.shape == (M, cfg.vocab_size)
    assert torch.equal(all_logits[-1:], logits)          # reset_state + fresh sequence => identical rerun

    # ---- oracle forward from the very same weights
    cpu = lambda t: t.detach().cpu()
    hidden = cpu(model.embedding)[tok]
    residual = None
    eps = cfg.rms_norm_eps
    H, I, E, k = cfg.hidden_size, cfg.moe_intermediate_size, cfg.n_routed_experts, cfg.num_experts_per_tok
    for i, lay in enumerate(model.layers):
        wn = cpu(lay.input_norm).to(torch.bfloat16)
        if residual is None:
            residual, hidden = hidden, D.rmsnorm(hidden, wn, eps)
        else:
            hidden, residual = D.fused_add_rmsnorm(hidden, residual, wn, eps)
        w = {kk: cpu(v) for kk, v in lay._w.items()}
        if lay.layer_type == "linear_attention":
            attn, _, _ = A.gdn_layer_prefill(hidden, w, dict(nk=cfg.linear_num_key_heads, nv=cfg.linear_num_value_heads,
                                                             dk=cfg.linear_key_head_dim, dv=cfg.linear_value_head_dim, eps=eps))
        else:
            attn, _, _ = A.gqa_layer_prefill(hidden, w, dict(nh=cfg.num_attention_heads, nkv=cfg.num_key_value_heads,
                                                             d=cfg.gqa_head_dim, rotary_dim=cfg.rotary_dim,
                                                             theta=cfg.rope_theta, eps=eps), pos)
        h, residual = D.fused_add_rmsnorm(attn, residual, cpu(lay.post_attn_norm).to(torch.bfloat16), eps)
        ts, gate = lay._experts, lay._gate
        wq13, ws13 = ts[0].cpu().numpy(), ts[1].view(torch.int16).cpu().numpy().view(np.uint8)
        wq2, ws2 = ts[2].cpu().numpy(), ts[3].view(torch.int16).cpu().numpy().view(np.uint8)
        q13, s13, q2, s2 = [], [], [], []
        for e in range(E):
            a, b = untile_expert_int4(wq13, ws13, e, 2 * I, H)
            c, d_ = untile_expert_int4(wq2, ws2, e, H, I)
            q13.append(a); s13.append(b); q2.append(c); s2.append(d_)
        layer = omoe.Int4Layer(np.stack(q13), np.stack(s13), np.stack(q2), np.stack(s2))
        hn = h.float().numpy()
        ids, wts = router.compute_routing(hn, cpu(gate).float().numpy(), k, norm_topk_prob=True)
        routed = omoe.moe_forward_gpu_path(layer, hn, ids, wts)
        gu, dn, gw = [cpu(t) if t is not None else None for t in lay._shared_w]
        shared = D.shared_expert_forward(h, D.quantize_to_int8(gu), D.quantize_to_int8(dn), gw)
        hidden = torch.from_numpy(omoe.finish_gpu_path(routed, 1.0, shared.float().numpy())).to(torch.bfloat16)
    hidden, _ = D.fused_add_rmsnorm(hidden, residual, cpu(model.final_norm).to(torch.bfloat16), eps)
    want = D.int8_linear(hidden, *D.quantize_to_int8(cpu(model._lm_head_bf16))).float()
    cos_last = torch.nn.functional.cosine_similarity(logits, want[-1:]).item()
    assert cos_last >= 0.999, cos_last
    # the GPU's top token must be the oracle's top token or tie with it within 4 BF16 ulp of the largest logit
    # (logits are BF16: two candidates closer than the rounding step can swap)
    assert float(want[-1][int(logits.argmax())]) >= float(want[-1].max()) - 4 * 2 ** -8 * float(want[-1].abs().max())
    cos_all = torch.nn.functional.cosine_similarity(all_logits, want, dim=1)
    # a router near-tie that resolves differently (fp32 summation order) re-routes that one token: allow a few rows
    # (and, through the linear-attention state of the next layers, the tokens right after it): allow a minority of rows
    top_ok = want.gather(1, all_logits.argmax(1)[:, None]).squeeze(1) >= want.max(1).values - 4 * 2 ** -8 * want.abs().max(1).values
    assert (cos_all >= 0.995).float().mean().item() > 0.9 and top_ok.float().mean().item() > 0.9
    assert cos_all.median().item() >= 0.999


def test_graph_replayed_prefill_is_bit_identical_to_eager_launches():
    """KrasisModel.forward_graphed replays the whole prefill step from a CUDA graph (same kernels, same order, static input
    buffers): logits must equal the eager path bit for bit, for the captured inputs and for different token ids replayed later."""
    from krasis_b200.model import HybridMoEConfig, KrasisModel
    cfg = HybridMoEConfig(hidden_size=256, num_hidden_layers=4, full_attention_interval=4, vocab_size=512,
                          n_routed_experts=8, num_experts_per_tok=2, moe_intermediate_size=128,
                          shared_expert_intermediate_size=128, num_attention_heads=4, num_key_value_heads=2,
                          gqa_head_dim=128, partial_rotary_factor=0.5, rope_theta=10000.0,
                          linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=32, linear_value_head_dim=32)
    M = 192
    model = KrasisModel(cfg, device=0, max_tokens=M)
    g = torch.Generator().manual_seed(3)
    pos = torch.arange(M).cuda()
    for trial in range(3):
        tok = torch.randint(0, cfg.vocab_size, (M,), generator=g).cuda()
        eager = model.forward(tok, pos, model.new_sequence()).clone()
        replay = model.forward_graphed(tok, pos).clone()
        assert model._graphs.get(M) not in (None, False), "capture fell back to eager launches"
        assert torch.equal(eager, replay), f"trial {trial}"
