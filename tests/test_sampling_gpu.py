"""Sampled decoding (nv_sample_topk) against the oracle's restatement of HF ``sample`` (oracle.sampling_probs, pinned
against transformers' own warpers in test_sampling_cpu.py).  The reference draws with torch.multinomial from the torch CUDA
generator, so token ids cannot be compared draw by draw; what is compared is (i) the DISTRIBUTION the kernel draws from,
element by element, (ii) that a draw is the inverse CDF of that distribution at the supplied uniform number, (iii) empirical
frequencies over many draws, (iv) through ``generate``: every sampled token lies in the oracle's top-k support of its step
(teacher-forced oracle forward over the generated ids), runs are reproducible under ``torch.manual_seed``."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

V = 32006
SPECIAL = [32000, 32001, 32002, 32003, 32004, 32005]


def _logits(B, seed, scale=2.5):
    g = torch.Generator().manual_seed(seed)
    lg = (torch.randn(B, V, generator=g) * scale).to(torch.bfloat16)
    lg[0, 5000:5100] = lg[0].float().topk(60)[0][-1]               # a run of ties straddling the k-th value
    return lg


def _run(lg, u, temperature, top_k, dev, finished=None, eos=2, pad=0, stop=True):
    from navillm_b200 import ops
    B = lg.shape[0]
    special = torch.tensor(SPECIAL, dtype=torch.int32, device=dev)
    fin = torch.zeros(B, dtype=torch.int32, device=dev) if finished is None else finished.to(dev)
    nxt = torch.full((B,), -1, dtype=torch.int32, device=dev)
    probs = torch.empty(B, V, dtype=torch.float32, device=dev)
    ops.sample_topk(lg.to(dev), special, fin, eos, pad, stop, temperature, top_k, u.to(dev), nxt, probs_out=probs)
    return nxt.cpu(), probs.cpu(), fin.cpu()


@pytest.mark.parametrize("temperature,top_k", [(1.0, 50), (0.7, 50), (1.5, 8), (1.0, 0), (0.3, 1)])
def test_distribution_and_inverse_cdf(cuda_dev, temperature, top_k):
    from oracle import navillm_oracle as O
    B = 8
    lg = _logits(B, seed=11)
    u = torch.tensor([0.0, 0.999999, 0.5, 0.25, 0.75, 0.1, 0.9, 0.333], dtype=torch.float32)
    nxt, probs, _ = _run(lg, u, temperature, top_k, cuda_dev)
    scores = lg.clone()
    scores[:, SPECIAL] = float("-inf")                             # what ModifiedLM.forward hands to generate (modified_lm.py:122-124)
    ref = O.sampling_probs(scores, temperature, top_k).float()
    # (i) the distribution: same support, values equal up to one bf16 ulp (expf / division differ in the last fp32 bit)
    assert torch.equal(probs > 0, ref > 0), f"support differs: {(probs > 0).sum(-1).tolist()} vs {(ref > 0).sum(-1).tolist()}"
    ulp = torch.where(ref > 0, 2.0 ** (torch.floor(torch.log2(ref.clamp_min(1e-38))) - 7), torch.zeros_like(ref))
    assert bool(((probs - ref).abs() <= ulp).all())
    frac_exact = float((probs == ref)[ref > 0].float().mean())
    assert frac_exact > 0.98, frac_exact
    assert bool((probs[:, SPECIAL] == 0).all())
    # (ii) the draw: inverse CDF (token order) of the kernel's distribution at u, checked in float64
    cdf = probs.double().cumsum(-1)
    for b in range(B):
        t = int(nxt[b])
        assert probs[b, t] > 0
        target = float(u[b]) * float(cdf[b, -1])
        lo = float(cdf[b, t - 1]) if t > 0 else 0.0
        tol = 1e-5 * float(cdf[b, -1])
        assert lo - tol <= target <= float(cdf[b, t]) + tol, (b, t, lo, target, float(cdf[b, t]))
    if top_k == 1:
        assert torch.equal(nxt.long(), scores.float().argmax(-1))


def test_empirical_frequencies(cuda_dev):
    """20k draws of one row: observed frequencies of the kept tokens agree with the oracle's probabilities (chi-square)."""
    from oracle import navillm_oracle as O
    lg = _logits(1, seed=3, scale=1.0)
    n = 20000
    scores = lg.clone()
    scores[:, SPECIAL] = float("-inf")
    ref = O.sampling_probs(scores, 0.9, 20).float()[0]
    ref = ref / ref.sum()
    g = torch.Generator().manual_seed(0)
    u = torch.rand(n, generator=g)
    nxt, _, _ = _run(lg.expand(n, V).contiguous(), u, 0.9, 20, cuda_dev)
    counts = torch.bincount(nxt.long(), minlength=V).float()
    assert float(counts[ref == 0].sum()) == 0
    kept = ref > 0
    exp = ref[kept] * n
    chi2 = float((((counts[kept] - exp) ** 2) / exp).sum())
    dof = int(kept.sum()) - 1
    assert chi2 < dof + 6 * (2 * dof) ** 0.5, (chi2, dof)


def test_finished_rows_and_eos(cuda_dev):
    B = 4
    lg = _logits(B, seed=5)
    lg[1, :] = -30.0
    lg[1, 2] = 30.0                                                # row 1 must draw EOS (id 2)
    fin = torch.tensor([0, 0, 1, 0], dtype=torch.int32)
    nxt, _, fin2 = _run(lg, torch.full((B,), 0.5), 1.0, 50, cuda_dev, finished=fin, eos=2, pad=0)
    assert int(nxt[2]) == 0 and int(nxt[1]) == 2
    assert fin2.tolist() == [0, 1, 1, 0]
    nxt, _, fin3 = _run(lg, torch.full((B,), 0.5), 1.0, 50, cuda_dev, finished=fin, eos=2, pad=0, stop=False)
    assert fin3.tolist() == [0, 0, 1, 0]


def test_generate_do_sample_support_and_reproducibility(cuda_dev):
    from oracle import navillm_oracle as O
    from tests.test_navmodel_gpu import build_model
    from tests.test_oracle_golden import load
    g, cfg, tok = load("amp_bf16")
    model, _ = build_model(g, cuda_dev)
    sd = g["state_dict"]
    qa = g["qa_in"]
    feats = qa["features"]
    lens = torch.tensor([f.shape[0] for f in feats])
    view = torch.stack([torch.cat([f, f.new_zeros(int(lens.max()) - f.shape[0], f.shape[1])], 0) for f in feats], 0)
    pano = O.forward_panorama(sd, cfg, view, lens)
    pe = pano["pano_embeds"] + O._pos_embed(torch.zeros(pano["pano_embeds"].shape[:2] + (14,)), sd, "vp_pos_embeddings")
    pe = pe + sd["token_type_embeddings.weight"][0]
    cand = pe[pano["pano_masks"]]
    text = tok(qa["prompts"])
    S0 = text["input_ids"].shape[1]
    n_new, T, K = 10, 0.8, 12

    def gen(seed):
        torch.manual_seed(seed)
        return model.lang_model.generate(input_ids=text["input_ids"], attention_mask=text["attention_mask"], cand_vis=cand.to(cuda_dev),
                                         max_new_tokens=n_new, stop_on_eos=False, do_sample=True, temperature=T, top_k=K).cpu()
    a, b, c = gen(1), gen(1), gen(2)
    assert a.shape == (text["input_ids"].shape[0], S0 + n_new)
    assert torch.equal(a, b), "same torch seed must reproduce the sampled ids"
    assert not torch.equal(a, c), "different seeds should give different samples"
    assert not bool(torch.isin(a[:, S0:], torch.tensor(model.lang_model.special_token_ids)).any())
    # teacher-forced oracle over the sampled sequence: the token drawn at step t must be in the oracle's top-k support of
    # that step (or within the bf16 noise floor of its k-th score: the two stacks round differently)
    mask = torch.cat([text["attention_mask"], torch.ones(a.shape[0], n_new, dtype=text["attention_mask"].dtype)], 1)
    pos = (mask.long().cumsum(-1) - 1).masked_fill(mask == 0, 1)
    out = O.modified_lm_forward(sd, cfg, a, mask, cand_vis=cand, position_ids=pos)
    logits = out["logits"].float()
    outside = 0
    for t in range(n_new):
        sc = logits[:, S0 + t - 1, :]
        kth = sc.topk(K)[0][:, -1]
        tokv = sc.gather(1, a[:, S0 + t:S0 + t + 1]).squeeze(1)
        ulp = 2.0 ** (torch.floor(torch.log2(kth.abs().clamp_min(1e-30))) - 7)
        assert bool((tokv >= kth - 3 * ulp).all()), (t, tokv.tolist(), kth.tolist())
        outside += int((tokv < kth).sum())
    print(f"\n[do_sample] tokens inside the oracle's top-{K} support: {a.shape[0] * n_new - outside} of {a.shape[0] * n_new}")
