"""Size-independent properties of the CUDA path at BASELINE.json's FULL widths (hidden 4096, 32 heads x 128,
FFN 11008; sequences up to 2048 tokens; 2 decoder layers to keep the run short -- every layer is the same code).
The oracle cannot run these sizes in seconds, so parity here is structural:

  * packing invariance  -- a sequence's hidden states do not depend on which other sequences share the batch
                           (bit-exact: the same tiles compute the same numbers);
  * causality           -- perturbing token t leaves hidden states of tokens < t bit-exact;
  * position semantics  -- left padding shifts positions exactly like the reference's arange(S) (RoPE uses the
                           padded column index), checked against an explicitly shifted run;
  * gradient linearity  -- backward(2*g) == 2*backward(g) up to bf16 rounding, and weight gradients accumulate;
  * GEMM identities at full size -- (A B^T) via the K-major path equals the MN-major path on transposed storage;
                           dgrad/wgrad shapes of the 7B model run through the CTA-pair kernel.
"""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def stack(cuda_dev):
    from navillm_b200 import llama
    dims = llama.LlamaDims(hidden=4096, n_layers=2, n_heads=32, inter=11008, vocab=1024)
    with torch.device(cuda_dev):
        model = llama.LlamaModelParams(dims)
    g = torch.Generator(device=cuda_dev).manual_seed(0)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g, device=cuda_dev))
            else:
                p.normal_(0, 0.02, generator=g)
    flat = llama.FlatParams(model.flat_order(), cuda_dev)
    return dims, model, flat, llama.LlamaCore(dims, model, flat)


def run(core, x, seqlens, pos=None, save=False):
    dev = x.device
    cu = torch.tensor([0] + list(torch.tensor(seqlens).cumsum(0)), dtype=torch.int32, device=dev)
    if pos is None:
        pos = torch.cat([torch.arange(L, dtype=torch.int32) for L in seqlens]).to(dev)
    return core.forward(x, pos, cu, seqlens, save=save)


def test_packing_invariance_and_causality_full_width(stack, cuda_dev):
    dims, model, flat, core = stack
    g = torch.Generator(device=cuda_dev).manual_seed(1)
    lens = [2048, 700, 1024, 333]
    xs = [(torch.randn(L, dims.hidden, generator=g, device=cuda_dev) * 0.5).to(bf16) for L in lens]
    full, _ = run(core, torch.cat(xs), lens)
    off = 0
    for i, L in enumerate(lens):
        alone, _ = run(core, xs[i].clone(), [L])
        assert torch.equal(alone, full[off:off + L]), f"sequence {i} depends on its batch neighbours"
        off += L
    # causality: perturb the token at 1500 of sequence 0
    x2 = xs[0].clone()
    x2[1500] += 1.0
    pert, _ = run(core, x2, [2048])
    assert torch.equal(pert[:1500], full[:1500]) and not torch.equal(pert[1500:], full[1500:2048])
    assert bool(torch.isfinite(full.float()).all())


def test_left_padding_positions_full_width(stack, cuda_dev):
    """Reference semantics: a left-padded row of length L in a batch padded to S uses positions S-L..S-1."""
    dims, model, flat, core = stack
    g = torch.Generator(device=cuda_dev).manual_seed(2)
    L, S = 600, 1024
    x = (torch.randn(L, dims.hidden, generator=g, device=cuda_dev) * 0.5).to(bf16)
    shifted = torch.arange(S - L, S, dtype=torch.int32, device=cuda_dev)
    a, _ = run(core, x.clone(), [L], pos=shifted)
    b, _ = run(core, x.clone(), [L])
    assert not torch.equal(a, b)                     # positions matter (RoPE is relative only inside attention scores...
    # ...but the attention pattern is shift-invariant: scores depend on position differences only, so the outputs
    # agree up to the bf16 rounding of the rotated q/k (different absolute angles): a loose numeric check
    assert (a.float() - b.float()).abs().max().item() <= 0.08 * b.float().abs().max().item()


def test_backward_linearity_and_accumulation_full_width(stack, cuda_dev):
    dims, model, flat, core = stack
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    lens = [900, 1100]
    x = (torch.randn(sum(lens), dims.hidden, generator=g, device=cuda_dev) * 0.5).to(bf16)
    dy = torch.randn(sum(lens), dims.hidden, generator=g, device=cuda_dev).to(bf16)
    w = model.layers[0].mlp.down_proj.weight

    def bwd(scale):
        flat.flat_grad.zero_()
        _, tape = run(core, x.clone(), lens, save=True)
        dx = core.backward((dy.float() * scale).to(bf16), tape)
        return dx.float(), w.grad.float().clone()

    dx1, gw1 = bwd(1.0)
    dx2, gw2 = bwd(2.0)
    assert (dx2 - 2 * dx1).abs().max().item() <= 2e-2 * dx2.abs().max().item()
    assert (gw2 - 2 * gw1).abs().max().item() <= 2e-2 * gw2.abs().max().item()
    # accumulation: a second backward on top of the first doubles the weight gradient
    _, tape = run(core, x.clone(), lens, save=True)
    core.backward(dy.clone(), tape)
    gw_acc = w.grad.float()
    bwd(1.0)
    _, tape = run(core, x.clone(), lens, save=True)
    core.backward(dy.clone(), tape)
    assert (w.grad.float() - 2 * gw1).abs().max().item() <= 2e-2 * (2 * gw1).abs().max().item()
    assert bool(torch.isfinite(gw_acc).all())


def test_gemm_major_forms_agree_at_7b_shapes(cuda_dev):
    from navillm_b200 import ops
    g = torch.Generator(device=cuda_dev).manual_seed(4)
    T, D, F = 4096, 4096, 11008
    x = torch.randn(T, D, generator=g, device=cuda_dev).to(bf16)
    w = (torch.randn(2 * F, D, generator=g, device=cuda_dev) * 0.02).to(bf16)
    y_k = ops.gemm(x, w)                                     # K-major A, K-major B (forward form), CTA-pair kernel
    y_mn = ops.gemm(x.t().contiguous(), w.t().contiguous(), a_mn=True, b_mn=True)   # same product, MN-major storage
    y_1 = ops.gemm(x, w, block_n=256)                        # single-CTA kernel
    assert torch.equal(y_k, y_mn) and torch.equal(y_k, y_1)  # same k order, same fp32 accumulation -> bit-exact
    dy = torch.randn(T, 2 * F, generator=g, device=cuda_dev).to(bf16)
    dx = ops.gemm(dy, w, b_mn=True)                          # dgrad
    dw = ops.gemm(dy, x, a_mn=True, b_mn=True)               # wgrad
    ref_dx = (dy[:64].float() @ w.float())
    assert (dx[:64].float() - ref_dx).abs().max().item() <= 2e-2 * ref_dx.abs().max().item()
    ref_dw = dy[:, :64].float().t() @ x.float()
    assert (dw[:64].float() - ref_dw).abs().max().item() <= 2e-2 * ref_dw.abs().max().item()
