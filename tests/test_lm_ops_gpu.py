"""Row-wise LM kernels (csrc/lm_ops.cu) against plain PyTorch fp32 references of the same ops.

Tolerances: each kernel rounds at the same points as the reference's bf16 eager op, so forward results
agree to 1 bf16 ulp (rtol 8e-3); backward kernels keep fp32 inside where eager autograd rounds every
intermediate to bf16, so they are compared with an fp32 reference at 2 bf16 ulp (rtol 1.6e-2).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def close(a, b, rtol, atol):
    a, b = a.float(), b.float()
    bad = (a - b).abs() > (atol + rtol * b.abs())
    assert not bool(bad.any()), f"max abs diff {(a - b).abs().max().item()} ({int(bad.sum())} bad of {bad.numel()})"


def test_rmsnorm_fwd_bwd(cuda_dev):
    from navillm_b200 import ops
    torch.manual_seed(0)
    T, D = 333, 4096
    x = torch.randn(T, D, device=cuda_dev).to(bf16)
    w = (1 + 0.1 * torch.randn(D, device=cuda_dev)).to(bf16)
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-6)
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
    ref = (w.float() * (xf * r).to(bf16).float()).to(bf16)
    close(y, ref, 8e-3, 1e-3)
    close(rstd, r.squeeze(-1), 1e-5, 1e-6)
    # backward vs autograd (fp32 math)
    dy = torch.randn(T, D, device=cuda_dev).to(bf16)
    dres = torch.randn(T, D, device=cuda_dev).to(bf16)
    dw = torch.zeros(D, device=cuda_dev, dtype=bf16)
    dx = ops.rmsnorm_bwd(x, w, rstd, dy, dres=dres, dw=dw)
    xa = x.float().requires_grad_(True)
    wa = w.float().requires_grad_(True)
    ya = wa * (xa * torch.rsqrt(xa.pow(2).mean(-1, keepdim=True) + 1e-6))
    ya.backward(dy.float())
    close(dx, xa.grad + dres.float(), 1.6e-2, 2e-2)
    close(dw, wa.grad, 1.6e-2, 0.15)


def _rope_tables(max_pos, hd, dev):
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], -1)
    return emb.cos().to(bf16).to(dev), emb.sin().to(bf16).to(dev)


def test_rope(cuda_dev):
    from navillm_b200 import ops
    torch.manual_seed(1)
    T, H, hd = 77, 4, 128
    cos_t, sin_t = _rope_tables(256, hd, cuda_dev)
    x = torch.randn(T, 3 * H * hd, device=cuda_dev).to(bf16)
    pos = torch.randint(0, 256, (T,), device=cuda_dev, dtype=torch.int32)
    y = x.clone()
    ops.rope_(y, pos, cos_t, sin_t, 2 * H)          # q and k blocks
    xv = x[:, :2 * H * hd].view(T, 2 * H, hd)
    c, s = cos_t[pos.long()][:, None, :], sin_t[pos.long()][:, None, :]
    rot = torch.cat([-xv[..., hd // 2:], xv[..., :hd // 2]], -1)
    ref = (xv * c) + (rot * s)                      # bf16 eager arithmetic, like HF apply_rotary_pos_emb
    assert torch.equal(y[:, :2 * H * hd].view(T, 2 * H, hd), ref)
    assert torch.equal(y[:, 2 * H * hd:], x[:, 2 * H * hd:])
    # backward = rotation by -theta
    g = torch.randn(T, 2 * H * hd, device=cuda_dev).to(bf16)
    gb = g.clone()
    ops.rope_(gb, pos, cos_t, sin_t, 2 * H, backward=True)
    gv = g.view(T, 2 * H, hd).float()
    rot_t = torch.cat([gv[..., hd // 2:], -gv[..., :hd // 2]], -1)
    refb = gv * c.float() + rot_t * s.float()
    close(gb.view(T, 2 * H, hd), refb, 1.6e-2, 1e-2)


def test_swiglu(cuda_dev):
    from navillm_b200 import ops
    torch.manual_seed(2)
    T, F = 130, 11008
    gu = torch.randn(T, 2 * F, device=cuda_dev).to(bf16)
    h = ops.swiglu_fwd(gu)
    g, u = gu[:, :F], gu[:, F:]
    ref = torch.nn.functional.silu(g) * u
    close(h, ref, 8e-3, 1e-3)
    dh = torch.randn(T, F, device=cuda_dev).to(bf16)
    dgu = ops.swiglu_bwd(gu, dh)
    ga, ua = g.float().requires_grad_(True), u.float().requires_grad_(True)
    (torch.nn.functional.silu(ga) * ua).backward(dh.float())
    close(dgu[:, :F], ga.grad, 1.6e-2, 1e-2)
    close(dgu[:, F:], ua.grad, 1.6e-2, 1e-2)


def test_embed_and_scatter(cuda_dev):
    from navillm_b200 import ops
    torch.manual_seed(3)
    V, D, T = 500, 256, 90
    E = torch.randn(V, D, device=cuda_dev).to(bf16)
    ids = torch.randint(0, V, (T,), device=cuda_dev, dtype=torch.int32)
    vis_src = torch.full((T,), -1, device=cuda_dev, dtype=torch.int32)
    sel = torch.tensor([3, 10, 11, 50, 89], device=cuda_dev)
    vis_src[sel] = torch.arange(5, device=cuda_dev, dtype=torch.int32)
    vis = torch.randn(5, D, device=cuda_dev)
    out = ops.embed_fwd(ids, E, vis_src, vis)
    ref = E[ids.long()].clone()
    ref[sel] = (ref[sel].float() + vis).to(bf16)     # bf16 += fp32 -> fp32 add, rounded to bf16
    assert torch.equal(out, ref)
    dx = torch.randn(T, D, device=cuda_dev).to(bf16)
    dvis = ops.embed_bwd_vis(dx, vis_src, 5)
    assert torch.equal(dvis, dx[sel].float())
    dE = torch.zeros(V, D, device=cuda_dev, dtype=bf16)
    ops.embed_bwd_weight_(dx, ids, dE)
    refE = torch.zeros(V, D, device=cuda_dev).index_add_(0, ids.long(), dx.float())
    close(dE, refE, 8e-3, 1e-2)


def test_head_and_rows(cuda_dev):
    from navillm_b200 import ops
    torch.manual_seed(4)
    T, D, O, R = 64, 4096, 100, 5
    h = torch.randn(T, D, device=cuda_dev).to(bf16)
    rows = torch.tensor([63, 2, 17, 40, 5], device=cuda_dev, dtype=torch.int32)
    x = ops.gather_rows(h, rows)
    assert torch.equal(x, h[rows.long()])
    W = (torch.randn(O, D, device=cuda_dev) * 0.02).to(bf16)
    b = torch.randn(O, device=cuda_dev).to(bf16)
    y = ops.head_fwd(x, W, b)
    ref = x.float() @ W.float().t() + b.float()
    close(y, ref, 8e-3, 1e-2)
    dy = torch.randn(R, O, device=cuda_dev).to(bf16)
    dW = torch.zeros(O, D, device=cuda_dev, dtype=bf16)
    db = torch.zeros(O, device=cuda_dev, dtype=bf16)
    dx = ops.head_bwd(dy, x, W, dW=dW, db=db)
    close(dx, dy.float() @ W.float(), 1.6e-2, 1e-2)
    close(dW, dy.float().t() @ x.float(), 1.6e-2, 2e-2)
    close(db, dy.float().sum(0), 1.6e-2, 2e-2)
    dst = torch.zeros(T, D, device=cuda_dev, dtype=bf16)
    ops.scatter_rows_(dx, rows, dst)
    assert torch.equal(dst[rows.long()], dx) and float(dst.float().abs().sum()) == float(dx.float().abs().sum())


def test_ce(cuda_dev):
    from navillm_b200 import ops
    torch.manual_seed(5)
    N, V = 37, 32006
    logits = (torch.randn(N, V, device=cuda_dev) * 2).to(bf16)
    labels = torch.randint(0, 32000, (N,), device=cuda_dev, dtype=torch.int32)
    labels[::5] = -100
    special = torch.tensor([32000, 32001, 32002, 32003, 32004], device=cuda_dev, dtype=torch.int32)
    n_act = int((labels >= 0).sum())
    row_loss, dl = ops.ce_fwd_bwd(logits, labels, special, grad_scale=1.0 / n_act)
    la = logits.float().clone()
    la[:, special.long()] = float("-inf")
    la.requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(la, labels.long(), ignore_index=-100)
    loss.backward()
    close(row_loss.sum() / n_act, loss.detach(), 1e-4, 1e-4)
    close(dl, la.grad, 1.6e-2, 1e-6)


@pytest.mark.parametrize("cols", [32006, 33, 8])
def test_scale_in_place_by_a_device_scalar(cuda_dev, cols):
    """nv_scale_bf16 (the upstream gradient of the scalar LM loss folded into the stored dlogits): equals the fp32 product
    rounded once to bf16, leaves the padding columns of the strided buffer alone."""
    from navillm_b200 import ops
    ld = (cols + 7) // 8 * 8 + 8
    g = torch.Generator(device="cpu").manual_seed(cols)
    buf = torch.randn(5, ld, generator=g).to(cuda_dev, torch.bfloat16)
    keep = buf.clone()
    x = buf[:, :cols]
    s = torch.tensor([0.37], device=cuda_dev, dtype=torch.float32)
    ops.scale_(x, s)
    assert torch.equal(x, (keep[:, :cols].float() * 0.37).to(torch.bfloat16))
    assert torch.equal(buf[:, cols:], keep[:, cols:])
