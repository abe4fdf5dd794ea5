"""Panorama encoder on the fp32 kernels (navillm_b200/image_embedding.py) vs the CPU oracle restatement of
the reference's ImageEmbeddings (oracle.forward_panorama), forward and parameter gradients.

Two numerical modes (ops.set_pano_precision):
  fp32  exact CUDA-core GEMMs: both sides fp32, only summation order differs -> 1e-4 (forward) / 2e-4 (gradients)
        relative to the tensor's max.
  tf32  tcgen05 kind::tf32 GEMMs (the product default; what the reference's pinned torch 1.10 did on Ampere+ with its
        default allow_tf32 = True): operands carry a 10-bit mantissa (unit roundoff 2^-10 under truncation), errors
        accumulate over 2 encoder layers + projections -> 4e-3 (forward) / 1e-2 (gradients) relative to the max.
"""
import sys
import types
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def rel_err(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


@pytest.mark.parametrize("B,N,lens,with_obj", [(2, 12, [12, 9], True), (3, 36, [36, 36, 36], False), (2, 70, [70, 33], False)])
def test_pano_encoder_matches_oracle(cuda_dev, B, N, lens, with_obj, pano_precision):
    precision = pano_precision
    from oracle import navillm_oracle as O
    from navillm_b200.image_embedding import ImageEmbeddings
    tol_f, tol_g = (1e-4, 2e-4) if precision == "fp32" else (4e-3, 1e-2)
    cfg = O.OracleConfig(hidden=256, n_layers=1, n_heads=2, inter=256, vocab=70, image_feat_size=72, obj_feat_size=40,
                         pano_hidden=128, pano_heads=2, pano_inter=256, num_pano_layers=2)
    sd = O.init_state_dict(cfg, seed=1)
    vis_cfg = types.SimpleNamespace(hidden_size=128, num_attention_heads=2, intermediate_size=256, hidden_dropout_prob=0.1,
                                    image_feat_size=72, angle_feat_size=4, obj_feat_size=40, output_size=256, num_pano_layers=2)
    mod = ImageEmbeddings(vis_cfg, use_obj=True).eval()
    mod.load_state_dict({k[len("img_embeddings."):]: v for k, v in sd.items() if k.startswith("img_embeddings.")})
    mod = mod.to(cuda_dev)
    g = torch.Generator().manual_seed(B * 100 + N)
    view = torch.randn(B, N, 72, generator=g)
    loc = torch.randn(B, N, 7, generator=g)
    types_ = torch.randint(0, 2, (B, N), generator=g)
    lens_t = torch.tensor(lens)
    kw = {}
    if with_obj:
        kw = dict(obj_img_fts=torch.randn(B, 5, 40, generator=g), obj_lens=torch.tensor([5, 3][:B]),
                  obj_loc_fts=torch.randn(B, 5, 7, generator=g))
    # oracle with autograd
    sda = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("img_embeddings.")}
    ref = O.forward_panorama(sda, cfg, view, lens_t, loc, types_, **kw)
    Gp = torch.randn(ref["pano_embeds"].shape, generator=g)
    loss = (ref["pano_embeds"] * Gp).sum()
    if with_obj:
        Go = torch.randn(ref["obj_embeds"].shape, generator=g)
        loss = loss + (ref["obj_embeds"] * Go).sum()
    loss.backward()
    # CUDA
    out = mod.forward_panorama_per_step(view.to(cuda_dev), lens_t.to(cuda_dev), loc.to(cuda_dev), types_.to(cuda_dev),
                                        **{k: v.to(cuda_dev) for k, v in kw.items()})
    assert rel_err(out["pano_embeds"].detach().cpu(), ref["pano_embeds"].detach()) < tol_f
    assert torch.equal(out["pano_masks"].cpu(), ref["pano_masks"])
    l2 = (out["pano_embeds"] * Gp.to(cuda_dev)).sum()
    if with_obj:
        assert rel_err(out["obj_embeds"].detach().cpu(), ref["obj_embeds"].detach()) < tol_f
        l2 = l2 + (out["obj_embeds"] * Go.to(cuda_dev)).sum()
    l2.backward()
    torch.cuda.synchronize()
    for name, p in mod.named_parameters():
        r = sda["img_embeddings." + name].grad
        if r is None:
            continue
        assert p.grad is not None, name
        assert rel_err(p.grad.cpu(), r) < tol_g, f"{name}: {rel_err(p.grad.cpu(), r)}"
