"""The oracle's sampling distribution (oracle.sampling_probs) against the installed transformers' own logits warpers: the
reference samples through HF GenerationMixin.sample (models/nav_model.py:388-396 with do_sample / temperature from
tasks/agents/llava.py:58-62), whose warper list under the generation defaults is [Temperature, TopK(50)]."""
import pytest
import torch

from oracle import navillm_oracle as O


def _hf_probs(scores, temperature, top_k):
    lp = pytest.importorskip("transformers.generation.logits_process")
    s = scores
    if temperature != 1.0:
        s = lp.TemperatureLogitsWarper(temperature)(None, s)
    if top_k:
        s = lp.TopKLogitsWarper(top_k=top_k, min_tokens_to_keep=1)(None, s)
    return torch.softmax(s, dim=-1)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("temperature,top_k", [(1.0, 50), (0.7, 50), (1.3, 5), (0.2, 0), (1.0, 1)])
def test_sampling_probs_equal_hf_warpers(dtype, temperature, top_k):
    g = torch.Generator().manual_seed(5)
    scores = (torch.randn(6, 1000, generator=g) * 3).to(dtype)
    scores[:, [3, 17, 999]] = float("-inf")                       # special tokens (models/modified_lm.py:122-124)
    scores[2, 100:140] = scores[2, 100]                           # a run of ties
    got = O.sampling_probs(scores, temperature, top_k)
    ref = _hf_probs(scores.clone(), temperature, top_k)
    assert got.dtype == dtype
    assert torch.equal(got, ref)
    kept = (got > 0).sum(-1)
    if top_k:
        assert bool((kept >= min(top_k, 1)).all())
        assert bool((kept[[0, 1, 3, 4, 5]] <= max(top_k, 1)).all())   # rows without ties keep exactly <= k tokens
    assert bool((got[:, [3, 17, 999]] == 0).all())


def test_generation_default_top_k_is_50():
    """The reference never passes top_k, so HF's generation default applies to its do_sample runs (transformers 4.28:
    GenerationConfig(top_k=50, top_p=1.0); 5.x keeps the same values in its table of global defaults)."""
    tr = pytest.importorskip("transformers")
    g = tr.GenerationConfig()
    if g.top_k is None:                                           # 5.x: filled in from the global defaults at generate()
        d = tr.GenerationConfig._get_default_generation_params()
        assert d["top_k"] == 50 and d["top_p"] == 1.0 and d["temperature"] == 1.0
    else:
        assert g.top_k == 50 and g.top_p == 1.0
