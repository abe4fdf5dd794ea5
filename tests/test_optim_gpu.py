"""Fused clip + AdamW over the flat buffers vs torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW on the same
parameters (the reference's optimizer step, train.py:86-89).  fp32 buffer: 1e-6 relative.  bf16 buffer: torch
rounds every intermediate to bf16 while the kernel keeps fp32 inside and rounds once -> compare both with an fp32
AdamW on the same values: |kernel - fp32| <= |torch_bf16 - fp32| + 1 bf16 ulp."""
import sys
import types
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = Path(__file__).resolve().parent / "golden"


def test_flat_adamw_matches_torch(cuda_dev):
    from navillm_b200.optim import FlatAdamW
    from tests.test_navmodel_gpu import build_model
    g = torch.load(GOLD / "nav_amp_bf16.pt", weights_only=False)
    model, _ = build_model(g, cuda_dev)
    model._ensure()
    gen = torch.Generator(device="cpu").manual_seed(0)
    params = [p for p in model.parameters()]
    for p in params:                                      # synthetic gradients, large enough to trigger clipping
        p.grad.copy_((torch.randn(p.shape, generator=gen) * 3).to(p.dtype))
    ref_p = [p.detach().clone() for p in params]
    ref_g = [p.grad.detach().clone() for p in params]
    opt = FlatAdamW(model, lr=1e-2)
    for it in range(3):
        opt.step(max_grad_norm=40.0)
    torch.cuda.synchronize()
    # torch reference on copies: bf16 params as torch would (bf16 states) and an fp32 shadow
    tp = [torch.nn.Parameter(x.clone()) for x in ref_p]
    tf = [torch.nn.Parameter(x.float().clone()) for x in ref_p]
    o1 = torch.optim.AdamW(tp, lr=1e-2)
    o2 = torch.optim.AdamW(tf, lr=1e-2)
    for it in range(3):
        for q, gq in zip(tp, ref_g):
            q.grad = gq.clone()
        for q, gq in zip(tf, ref_g):
            q.grad = gq.float().clone()
        n1 = torch.nn.utils.clip_grad_norm_(tp, 40.0)
        n2 = torch.nn.utils.clip_grad_norm_(tf, 40.0)
        o1.step(); o2.step()
    assert abs(float(opt.grad_norm()) - float(n2)) <= 2e-3 * float(n2)
    for p, a, b in zip(params, tp, tf):
        mine, t_same, t32 = p.detach().float(), a.detach().float(), b.detach()
        if p.dtype == torch.float32:
            assert torch.allclose(mine, t32, rtol=2e-5, atol=1e-6)
        else:
            e_ref = (t_same - t32).abs().max().item()
            e_mine = (mine - t32).abs().max().item()
            assert e_mine <= e_ref + 2.0 ** -8 * t32.abs().max().item(), (e_mine, e_ref)
    sd = opt.state_dict()
    assert len(sd["state"]) == len(params) and sd["state"][0]["exp_avg"].shape == params[0].shape


def test_optimizer_state_interop_with_torch_adamw_indexing(cuda_dev):
    """State written by torch.optim.AdamW over ``model.parameters()`` (what the reference saves, tools/optims.py:43,69-71)
    loads by index; a state dict whose per-index shapes do not match this model's parameter order (e.g. written with
    up_proj / down_proj swapped) is rejected with a clear error instead of being copied or silently mis-assigned; a lazy
    zero_grad followed by an optimizer step with NO backward in between applies zero gradients, not stale ones."""
    from navillm_b200.optim import FlatAdamW
    from tests.test_navmodel_gpu import build_model
    g = torch.load(GOLD / "nav_amp_bf16.pt", weights_only=False)
    model, _ = build_model(g, cuda_dev)
    model._ensure()
    params = [p for p in model.parameters() if p.requires_grad]
    # a torch AdamW over the same parameter list, one step -> its state dict uses parameter INDICES
    for p in params:
        p.grad.copy_(torch.ones_like(p) * 0.01)
    topt = torch.optim.AdamW(params, lr=0.0)
    topt.step()
    tsd = topt.state_dict()
    opt = FlatAdamW(model, lr=1e-3)
    opt.load_state_dict(tsd)
    back = opt.state_dict()
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    for i in (0, names.index("lang_model.model.layers.0.mlp.down_proj.weight"), names.index("lang_model.model.layers.0.mlp.up_proj.weight"), len(params) - 1):
        assert torch.equal(back["state"][i]["exp_avg"].cpu(), tsd["state"][i]["exp_avg"].cpu()), names[i]
    # swapped down/up entries (the order of an implementation that registers gate, up, down): rejected
    i_dn, i_up = names.index("lang_model.model.layers.0.mlp.down_proj.weight"), names.index("lang_model.model.layers.0.mlp.up_proj.weight")
    assert i_dn < i_up                                   # reference order: gate_proj, down_proj, up_proj (transformers 4.28)
    if params[i_dn].shape != params[i_up].shape:
        bad = {"state": dict(tsd["state"]), "param_groups": tsd["param_groups"]}
        bad["state"][i_dn], bad["state"][i_up] = tsd["state"][i_up], tsd["state"][i_dn]
        with pytest.raises(ValueError):
            opt.load_state_dict(bad)
    with pytest.raises(ValueError):
        opt.load_state_dict({"state": {}, "param_groups": [{"params": list(range(len(params) + 1))}]})
    # lazy zero + step without a backward: per-layer gradients count as zero
    q = model.lang_model.model.layers[0].self_attn.q_proj.weight
    q.grad.fill_(5.0)
    model.zero_grad(lazy=True)
    opt2 = FlatAdamW(model, lr=1e-2, weight_decay=0.0)
    before = q.detach().float().clone()
    opt2.step(max_grad_norm=40.0)
    torch.cuda.synchronize()
    assert float(q.grad.float().abs().max()) == 0.0 and torch.equal(q.detach().float(), before)
