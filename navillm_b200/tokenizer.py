"""Tokenizer front-end of the modified LM (host side; mirrors models/modified_lm.py:56-87).

The reference loads ``LlamaTokenizer.from_pretrained(vicuna)``, registers the special tokens
``<cand> <hist> <obj> <cls_1> <cls_2>`` (+ ``<PAD>``), and tokenises with left padding, LEFT truncation
at 1024 and ``return_token_type_ids=True`` (pair inputs ``[prompt, answer]`` get type ids 0/1).

* ``HFTokenizerAdapter`` does exactly that when the Vicuna tokenizer files are available locally.
* ``SyntheticTokenizer`` is a deterministic stand-in for boxes without any tokenizer files (no network
  here): whitespace/special-token splitting and a stable word hash into the base vocabulary.  It
  reproduces the *framing* the model depends on -- BOS per segment, pair type ids, left padding, left
  truncation, special-token ids 32000.. -- which is all the hot path observes.
"""
from __future__ import annotations

import re
import zlib
from typing import List, Sequence, Union

import torch

SPECIAL_TOKENS = ["<cand>", "<hist>", "<obj>", "<cls_1>", "<cls_2>"]


class TokenBatch(dict):
    """dict with ``.to(device)`` like transformers.BatchEncoding."""

    def to(self, device):
        return TokenBatch({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class SyntheticTokenizer:
    bos_token, eos_token, unk_token, pad_token = "<s>", "</s>", "<unk>", "<PAD>"
    unk_token_id, bos_token_id, eos_token_id = 0, 1, 2
    padding_side = "left"
    truncation_side = "left"

    def __init__(self, base_vocab: int = 32000):
        self.base_vocab = base_vocab
        self.special = {tok: base_vocab + i for i, tok in enumerate(SPECIAL_TOKENS)}
        self.pad_token_id = base_vocab + len(SPECIAL_TOKENS)
        self._split = re.compile("(" + "|".join(re.escape(t) for t in SPECIAL_TOKENS + [self.eos_token]) + ")")
        self._id2word: dict[int, str] = {}

    def __len__(self) -> int:
        return self.base_vocab + len(SPECIAL_TOKENS) + 1

    def _word_id(self, w: str) -> int:
        i = 3 + zlib.crc32(w.encode("utf-8")) % (self.base_vocab - 3)
        self._id2word.setdefault(i, w)
        return i

    def encode(self, text: str, add_special_tokens: bool = True) -> List[int]:
        ids: List[int] = [self.bos_token_id] if add_special_tokens else []
        for piece in self._split.split(text):
            if not piece:
                continue
            if piece in self.special:
                ids.append(self.special[piece])
            elif piece == self.eos_token:
                ids.append(self.eos_token_id)
            else:
                ids.extend(self._word_id(w) for w in piece.split())
        return ids

    def __call__(self, text: Union[str, Sequence], max_length: int = 1024, padding: bool = True, truncation: bool = True,
                 return_tensors: str = "pt", add_special_tokens: bool = True, return_token_type_ids: bool = True):
        if isinstance(text, str):
            text = [text]
        rows, types = [], []
        for item in text:
            if isinstance(item, (list, tuple)):          # pair: [prompt, answer]
                a = self.encode(item[0], add_special_tokens)
                b = self.encode(item[1], add_special_tokens)
                ids, tt = a + b, [0] * len(a) + [1] * len(b)
            else:
                ids = self.encode(item, add_special_tokens)
                tt = [0] * len(ids)
            if truncation and len(ids) > max_length:     # truncation_side = 'left'
                ids, tt = ids[-max_length:], tt[-max_length:]
            rows.append(ids)
            types.append(tt)
        S = max(len(r) for r in rows)
        B = len(rows)
        input_ids = torch.full((B, S), self.pad_token_id, dtype=torch.long)
        attn = torch.zeros((B, S), dtype=torch.long)
        tti = torch.zeros((B, S), dtype=torch.long)
        for b, (ids, tt) in enumerate(zip(rows, types)):  # padding_side = 'left'
            n = len(ids)
            input_ids[b, S - n:] = torch.tensor(ids, dtype=torch.long)
            attn[b, S - n:] = 1
            tti[b, S - n:] = torch.tensor(tt, dtype=torch.long)
        return TokenBatch(input_ids=input_ids, attention_mask=attn, token_type_ids=tti)

    def batch_decode(self, seqs, skip_special_tokens: bool = True, clean_up_tokenization_spaces: bool = False):
        out = []
        drop = {self.bos_token_id, self.eos_token_id, self.unk_token_id, self.pad_token_id, *self.special.values()}
        for s in seqs:
            words = []
            for i in s:
                i = int(i)
                if skip_special_tokens and i in drop:
                    continue
                words.append(self._id2word.get(i, f"<{i}>"))
            out.append(" ".join(words))
        return out


class HFTokenizerAdapter:
    """The reference's own construction (models/modified_lm.py:56-75) for a local tokenizer directory."""

    def __init__(self, path: str):
        # the slow sentencepiece LlamaTokenizer, like the reference (models/modified_lm.py:57): the fast tokenizer can
        # split text around added special tokens differently from what released checkpoints were trained with
        from transformers import LlamaTokenizer
        self.tok = LlamaTokenizer.from_pretrained(path, padding_side="left", truncation_side="left")
        self.tok.add_special_tokens({"additional_special_tokens": list(SPECIAL_TOKENS)})
        if self.tok.pad_token is None:
            self.tok.add_special_tokens({"pad_token": "<PAD>"})
        self.special = {t: self.tok.encode(t, add_special_tokens=False)[0] for t in SPECIAL_TOKENS}

    def __len__(self):
        return len(self.tok)

    def __getattr__(self, k):
        return getattr(self.tok, k)

    def __call__(self, *a, **kw):
        return self.tok(*a, **kw)
