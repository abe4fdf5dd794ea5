"""Greedy generation (prefill + KV-cache decode on the kernels, CUDA-graph replay) against the CPU oracle's
restated HF greedy search.  Token ids must be bit-exact wherever the oracle's decision is not a numerical
near-tie: at each step the oracle's top-1/top-2 logit margin is compared with the bf16 noise floor
(2 ulp of the top logit); a step inside the noise floor may legitimately pick either candidate, after which
the sequences diverge and the comparison for that row stops."""
import sys
import types
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = Path(__file__).resolve().parent / "golden"


def test_greedy_generate_matches_oracle(cuda_dev):
    from oracle import navillm_oracle as O
    from tests.test_navmodel_gpu import build_model, to_dev
    from tests.test_oracle_golden import load
    g, cfg, tok = load("amp_bf16")
    model, _ = build_model(g, cuda_dev)
    sd = g["state_dict"]
    qa = g["qa_in"]
    n_new = 12
    # oracle (bf16, like the reference): visual tokens as in forward_3dqa, then greedy with logits recorded
    feats = qa["features"]
    lens = torch.tensor([f.shape[0] for f in feats])
    view = torch.stack([torch.cat([f, f.new_zeros(int(lens.max()) - f.shape[0], f.shape[1])], 0) for f in feats], 0)
    pano = O.forward_panorama(sd, cfg, view, lens)
    pe = pano["pano_embeds"] + O._pos_embed(torch.zeros(pano["pano_embeds"].shape[:2] + (14,)), sd, "vp_pos_embeddings")
    pe = pe + sd["token_type_embeddings.weight"][0]
    cand = pe[pano["pano_masks"]]
    text = tok(qa["prompts"])
    ref_ids, ref_logits = O.greedy_generate(sd, cfg, text["input_ids"], text["attention_mask"], cand_vis=cand,
                                            max_new_tokens=n_new, stop_on_eos=False, return_logits=True)
    S0 = text["input_ids"].shape[1]
    for graph in (False, True):
        ids = model.lang_model.generate(input_ids=text["input_ids"], attention_mask=text["attention_mask"],
                                        cand_vis=cand.to(cuda_dev), max_new_tokens=n_new, stop_on_eos=False,
                                        use_cuda_graph=graph).cpu()
        assert ids.shape == ref_ids.shape and torch.equal(ids[:, :S0], text["input_ids"])
        exact = 0
        for b in range(ids.shape[0]):
            for t in range(n_new):
                lg = ref_logits[t][b]
                top2 = torch.topk(lg, 2).values
                noise = 2 * 2.0 ** -8 * top2[0].abs().item()
                if ids[b, S0 + t] == ref_ids[b, S0 + t]:
                    exact += 1
                    continue
                assert (top2[0] - top2[1]).item() <= noise, \
                    f"graph={graph} row {b} step {t}: token {ids[b, S0 + t]} != oracle {ref_ids[b, S0 + t]} with margin {(top2[0] - top2[1]).item():.4g} > noise {noise:.4g}"
                break                                             # legitimate near-tie: sequences diverge from here
        assert exact >= n_new, f"too few exactly matching tokens ({exact})"


def test_3dqa_generate_mode_runs_and_decodes(cuda_dev):
    from tests.test_navmodel_gpu import build_model, to_dev
    g = torch.load(GOLD / "nav_amp_bf16.pt", weights_only=False)
    model, tok = build_model(g, cuda_dev)
    out = model("3dqa", to_dev(dict(g["qa_in"]), cuda_dev), training=False, max_new_tokens=6, do_sample=False)
    assert len(out["generated_sentences"]) == 2 and all(isinstance(s, str) for s in out["generated_sentences"])
    out2 = model("3dqa", to_dev(dict(g["qa_in"]), cuda_dev), training=False, max_new_tokens=6, do_sample=True, temperature=0.7)
    assert len(out2["generated_sentences"]) == 2
