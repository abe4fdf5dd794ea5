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


def test_oracle_rope_tables_match_product_tables():
    from navillm_b200.llama import LlamaDims, rope_tables
    from oracle import navillm_oracle as O
    d = LlamaDims(hidden=256, n_heads=2, n_layers=1, inter=256, vocab=64, max_pos=64)
    cos, sin = rope_tables(d, torch.device("cpu"))
    cfg = O.OracleConfig(hidden=256, n_heads=2)
    c2, s2 = O.rope_tables(cfg, torch.arange(64)[None], torch.bfloat16)
    assert torch.equal(cos, c2[0]) and torch.equal(sin, s2[0])
