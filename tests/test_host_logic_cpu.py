"""Host-side logic that runs without a GPU: tokenizer framing (left pad / left truncation / pair type ids),
prompt packing (positions, visual scatter order, label rows), and the oracle's own invariants."""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from navillm_b200.tokenizer import SyntheticTokenizer  # noqa: E402


def test_tokenizer_framing():
    tok = SyntheticTokenizer(base_vocab=1000)
    out = tok(["a b <cand> <cls_1>", ["long prompt here <hist>", "answer</s>"]])
    ids, m, tt = out["input_ids"], out["attention_mask"], out["token_type_ids"]
    assert ids.shape == m.shape == tt.shape
    assert m[0, 0] == 0 and ids[0, 0] == tok.pad_token_id          # left padding
    assert ids[0, m[0].bool()][0] == tok.bos_token_id
    assert ids[0, -1] == tok.special["<cls_1>"] and ids[0, -2] == tok.special["<cand>"]
    row = ids[1]
    assert int((row == tok.bos_token_id).sum()) == 2                 # BOS per segment
    assert row[-1] == tok.eos_token_id and tt[1, -1] == 1 and tt[1, 0] == 0
    long = tok(["w " * 3000], max_length=1024)
    assert long["input_ids"].shape[1] == 1024 and long["input_ids"][0, 0] != tok.bos_token_id   # left truncation
    assert len(tok) == 1006 and tok.pad_token_id == 1005


def test_packed_prompt_indices():
    from navillm_b200.modified_lm import PackedPrompt
    tok = SyntheticTokenizer(base_vocab=1000)
    lm = SimpleNamespace(cand_token_id=[tok.special["<cand>"]], hist_token_id=[tok.special["<hist>"]],
                         obj_token_id=[tok.special["<obj>"]], cls_token_id=[tok.special["<cls_1>"], tok.special["<cls_2>"]])
    text = tok([["x <hist> <cand> <cand> y", "ans one</s>"], ["p <cand> q r s t u", "b</s>"]])
    labels = text["input_ids"].clone()
    labels[text["token_type_ids"] == 0] = -100
    pp = PackedPrompt(text["input_ids"], text["attention_mask"], lm, torch.device("cpu"), labels=labels)
    assert pp.T == int(text["attention_mask"].sum()) and pp.seqlens == text["attention_mask"].sum(1).tolist()
    assert pp.cu.tolist() == [0, pp.seqlens[0], pp.T]
    # plain-forward positions = column index in the padded row (pads counted), like HF arange(S)
    S = text["input_ids"].shape[1]
    assert pp.pos[:pp.seqlens[0]].tolist() == list(range(S - pp.seqlens[0], S))
    # visual rows: cand rows first (row-major over the batch), then hist rows
    vs = pp.vis_src.numpy()
    ids = pp.ids.numpy()
    assert vs[ids == tok.special["<cand>"]].tolist() == [0, 1, 2]
    assert vs[ids == tok.special["<hist>"]].tolist() == [3]
    assert (vs[(ids != tok.special["<cand>"]) & (ids != tok.special["<hist>"])] == -1).all()
    # loss rows predict the NEXT token and only answer tokens are targets
    tgt = pp.loss_tgt.numpy()
    rows = pp.loss_rows.numpy()
    assert (ids[rows + 1] == tgt).all() and pp.n_loss == int((labels[:, 1:] != -100).sum())
    gen = PackedPrompt(text["input_ids"], text["attention_mask"], lm, torch.device("cpu"), generate_positions=True)
    assert gen.pos[:3].tolist() == [0, 1, 2]                          # generate: cumsum(mask) - 1
    # token order for the deterministic embedding gradient: the stable argsort torch.sort would give, made on the host
    ref_sorted, ref_order = torch.sort(pp.ids.to(torch.int64), stable=True)
    assert pp.tok_order.tolist() == ref_order.tolist() and pp.tok_sorted.tolist() == ref_sorted.tolist()
    with torch.no_grad():                                             # inference: no backward, no sort, nothing uploaded
        assert PackedPrompt(text["input_ids"], text["attention_mask"], lm, torch.device("cpu")).tok_order.numel() == 0


def test_oracle_rope_tables_match_product_tables():
    from navillm_b200.llama import LlamaDims, rope_tables
    from oracle import navillm_oracle as O
    d = LlamaDims(hidden=256, n_heads=2, n_layers=1, inter=256, vocab=64, max_pos=64)
    cos, sin = rope_tables(d, torch.device("cpu"))
    cfg = O.OracleConfig(hidden=256, n_heads=2)
    c2, s2 = O.rope_tables(cfg, torch.arange(64)[None], torch.bfloat16)
    assert torch.equal(cos, c2[0]) and torch.equal(sin, s2[0])


# ---------------------------------------------------------------------------------------------------------
# cross-step prefix-KV reuse: host-side planning (navillm_b200.modified_lm.plan_prefix_reuse)
# ---------------------------------------------------------------------------------------------------------
def _left_pad(rows, pad=0):
    import numpy as np
    S = max(len(r) for r in rows)
    ids = np.full((len(rows), S), pad, dtype=np.int64)
    msk = np.zeros((len(rows), S), dtype=bool)
    for i, r in enumerate(rows):
        ids[i, S - len(r):] = r
        msk[i, S - len(r):] = True
    return ids, msk


def test_prefix_reuse_plan_over_a_three_step_rollout():
    import types
    import numpy as np
    import pytest
    from navillm_b200.modified_lm import plan_prefix_reuse
    CAND, HIST, CLS = 900, 901, 902
    cache = types.SimpleNamespace(B=2, max_len=64, ids=[np.zeros(0, dtype=np.int64) for _ in range(2)], off=[None, None])
    instr = [[1, 5, 6, 7, 8], [1, 9, 10]]

    def prompt(b, t, n_cand):
        p = list(instr[b]) + [20]
        for i in range(t):
            p += [30 + i, HIST]
        p += [21] + [x for j in range(n_cand) for x in (40 + j, CAND)] + [22, CLS]
        return p

    # step 0: nothing cached, everything is encoded; positions = left-pad offset + index (HF arange over the padded row)
    rows = [prompt(0, 0, 2), prompt(1, 0, 3)]
    ids, msk = _left_pad(rows)
    tok, pos, vis, ql, cached, kvl, cls = plan_prefix_reuse(ids, msk, cache, [0, 0], 5, CAND, HIST, CLS)
    S0 = ids.shape[1]
    assert cached == [0, 0] and ql == [len(rows[0]), len(rows[1])] and kvl == ql
    assert cache.off == [S0 - len(rows[0]), S0 - len(rows[1])]
    assert pos[0].tolist() == list(range(cache.off[0], cache.off[0] + len(rows[0])))
    assert [int(v) for v in vis[0] if v >= 0] == [0, 1] and [int(v) for v in vis[1] if v >= 0] == [2, 3, 4]     # row-major <cand> order
    assert cls == [len(rows[0]) - 1, len(rows[0]) + len(rows[1]) - 1]
    off = list(cache.off)

    # step 1: one <hist> per row; the prefix up to the end of the old history section is reused
    rows = [prompt(0, 1, 3), prompt(1, 1, 1)]
    ids, msk = _left_pad(rows)
    tok, pos, vis, ql, cached, kvl, cls = plan_prefix_reuse(ids, msk, cache, [1, 1], 4, CAND, HIST, CLS)
    assert cached == [len(instr[0]) + 1, len(instr[1]) + 1]                   # instruction + the "history" marker token 20
    assert cache.off == off                                                   # rows keep their first-step rotary offset
    assert pos[0][0] == off[0] + cached[0] and kvl == [len(rows[0]), len(rows[1])]
    # visual sources: candidates first (row-major over the batch), then hist rows flattened sample-major
    assert [int(v) for v in vis[0] if v >= 0] == [4 + 0, 0, 1, 2] and [int(v) for v in vis[1] if v >= 0] == [4 + 1, 3]

    # step 2: the first <hist> is now part of the reusable prefix; the second one is new
    rows = [prompt(0, 2, 1), prompt(1, 2, 1)]
    ids, msk = _left_pad(rows)
    tok, pos, vis, ql, cached, kvl, cls = plan_prefix_reuse(ids, msk, cache, [2, 2], 2, CAND, HIST, CLS)
    assert cached == [len(instr[0]) + 1 + 2, len(instr[1]) + 1 + 2]
    assert [int(v) for v in vis[0] if v >= 0] == [2 + 0 + 1, 0] and [int(v) for v in vis[1] if v >= 0] == [2 + 2 + 1, 1]
    assert tok[0][0] == 31 and tok[0][-1] == CLS

    # a different instruction in row 1 (new episode without reset): nothing of that row is reused
    instr[1] = [1, 11, 12]
    rows = [prompt(0, 2, 1), prompt(1, 2, 1)]
    ids, msk = _left_pad(rows)
    _, _, _, _, cached, _, _ = plan_prefix_reuse(ids, msk, cache, [2, 2], 2, CAND, HIST, CLS)
    assert cached[0] == rows[0].index(CAND) and cached[1] == 1                # row 0: everything before its first <cand>

    # contract violations fail loudly
    with pytest.raises(RuntimeError, match="hist_vis rows"):
        plan_prefix_reuse(ids, msk, cache, [2, 1], 2, CAND, HIST, CLS)
    with pytest.raises(RuntimeError, match="cand_vis rows"):
        plan_prefix_reuse(ids, msk, cache, [2, 2], 3, CAND, HIST, CLS)
    cache.max_len = 8
    with pytest.raises(ValueError, match="exceeds the prefix cache length"):
        plan_prefix_reuse(ids, msk, cache, [2, 2], 2, CAND, HIST, CLS)


def test_tokenizer_fallback_is_explicit():
    """A missing / unusable tokenizer directory must not silently turn into hash-derived token ids (pretrained weights would
    run on garbage): the synthetic stand-in is used only when asked for (from_scratch / model_config.tokenizer)."""
    import warnings

    import pytest
    from navillm_b200.modified_lm import ModifiedLlamaForCausalLM
    cfg = SimpleNamespace(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=1, vocab_size=64, rms_norm_eps=1e-6)
    lm = ModifiedLlamaForCausalLM(cfg, SimpleNamespace(precision="amp_bf16"))
    with pytest.raises(RuntimeError, match="tokenizer"):
        lm.init_tokenizer("/nonexistent/vicuna-7b")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        lm.init_tokenizer("/nonexistent/vicuna-7b", allow_synthetic=True)
    assert any("SyntheticTokenizer" in str(x.message) for x in w)
    assert lm.cls_token == ["<cls_1>", "<cls_2>"] and len(lm.special_token_ids) == 5


def test_grad_sync_skips_clean_segments_and_tapers_chunks():
    """GradSync bookkeeping without a process group: the overlap schedule (groups of `chunk` layers, the last group cut into
    single layers) and the complement of the clean (all-zero) gradient segments."""
    from navillm_b200.parallel import GradSync
    gs = GradSync()
    assert gs.layer_hook(None, [], 8) is None and gs.exchange() == 0          # no process group: nothing armed, nothing sent
    # the range complement used by exchange(): [lo, hi) minus clean segments
    def complement(lo, hi, clean):
        ranges, cur = [], lo
        for a, b in sorted(clean):
            a, b = max(a, lo), min(b, hi)
            if a >= b:
                continue
            if a > cur:
                ranges.append((cur, a))
            cur = max(cur, b)
        if cur < hi:
            ranges.append((cur, hi))
        return ranges
    assert complement(100, 1000, [(400, 700)]) == [(100, 400), (700, 1000)]
    assert complement(100, 1000, [(0, 150), (900, 2000)]) == [(150, 900)]
    assert complement(100, 1000, []) == [(100, 1000)]


def test_fuse_obj_row_maps_reproduce_the_reference_layout():
    """`--fuse_obj` (models/image_embedding.py:81-93): the encoder input row b is [views[b,:vl] ; objs[b,:ol]] zero-padded to
    max(vl + ol); the view rows are read back afterwards.  The index maps the kernels gather / scatter with must reproduce
    exactly that layout (built here with torch indexing) and be inverse to each other."""
    import types
    import torch
    from navillm_b200.image_embedding import ImageEmbeddings
    cfg = types.SimpleNamespace(hidden_size=32, num_attention_heads=2, intermediate_size=64, hidden_dropout_prob=0.1,
                                image_feat_size=16, angle_feat_size=4, obj_feat_size=12, output_size=48, num_pano_layers=1)
    mod = ImageEmbeddings(cfg, use_obj=True, fuse_obj=True)
    assert [k for k in mod.state_dict() if k.startswith("obj_linear")] == ["obj_linear.0.weight", "obj_linear.0.bias",
                                                                          "obj_linear.1.weight", "obj_linear.1.bias"]
    B, N, O = 3, 5, 4
    vl, ol = torch.tensor([5, 2, 4]), torch.tensor([3, 0, 4])
    obj = torch.randn(B, O, 12)
    m = mod._fuse_maps(vl, obj, ol, torch.randn(B, O, 7), B, N, "cpu")
    Nf = int((vl + ol).max())
    assert m["Nf"] == Nf and m["lens_f"].tolist() == (vl + ol).tolist() and m["twos"].tolist() == [2] * (B * O)
    views = torch.arange(B * N, dtype=torch.float32).view(B, N) + 1          # view row ids 1..
    objs = -(torch.arange(B * O, dtype=torch.float32).view(B, O) + 1)        # object row ids -1..
    want = torch.zeros(B, Nf)
    for b in range(B):
        want[b, :vl[b]] = views[b, :vl[b]]
        want[b, vl[b]:vl[b] + ol[b]] = objs[b, :ol[b]]

    def gather(src, idx):
        out = torch.zeros(idx.numel())
        ok = idx >= 0
        out[ok] = src.reshape(-1)[idx[ok].long()]
        return out
    fused = gather(views, m["view_src"]) + gather(objs, m["obj_src"])
    assert torch.equal(fused.view(B, Nf), want)
    back = gather(fused, m["view_back"]).view(B, N)
    for b in range(B):
        assert torch.equal(back[b, :vl[b]], views[b, :vl[b]]) and bool((back[b, vl[b]:] == 0).all())
    oback = gather(fused, m["obj_back"]).view(B, O)
    for b in range(B):
        assert torch.equal(oback[b, :ol[b]], objs[b, :ol[b]]) and bool((oback[b, ol[b]:] == 0).all())
    # a model built without objects cannot fuse them
    import pytest
    plain = ImageEmbeddings(cfg, use_obj=False, fuse_obj=True)
    with pytest.raises(RuntimeError):
        plain._fuse_maps(vl, obj, ol, torch.randn(B, O, 7), B, N, "cpu")
