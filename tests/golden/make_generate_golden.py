"""Generate tests/golden/generate_{fp32,amp_bf16}.pt: token ids produced by the REFERENCE's generation branch
(models/nav_model.py:386-402 -> HF GenerationMixin.generate -> models/modified_lm.py:89-146) on the golden 3dqa batch.

    python tests/golden/make_generate_golden.py

The reference pins transformers==4.28.0; under the installed 5.5.0 its prepare_inputs_for_generation
(models/modified_lm.py:184-199) fails only because it forwards its arguments POSITIONALLY to the parent method, whose
signature changed.  The adapter below is that method with the same body, keyword arguments, and the emptiness test of
the cache written for 5.x cache objects ("not past_key_values" upstream).  Everything else - the model, its forward,
the visual-token injection, HF's greedy search loop - runs unmodified.  tests/test_oracle_golden.py checks the oracle's
restated greedy loop against these ids; tests/test_generate_gpu.py checks the CUDA path against the oracle.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_golden as MG  # noqa: E402


def main():
    for precision in ("fp32", "amp_bf16"):
        model, tfv = MG.build_reference_model(precision)
        from transformers import LlamaForCausalLM
        import models.modified_lm as ml

        def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None,
                                          cand_vis=None, hist_vis=None, obj_vis=None, **kwargs):
            model_inputs = LlamaForCausalLM.prepare_inputs_for_generation(
                self, input_ids, past_key_values=past_key_values, attention_mask=attention_mask, inputs_embeds=inputs_embeds, **kwargs)
            empty = past_key_values is None or (hasattr(past_key_values, "get_seq_length") and past_key_values.get_seq_length() == 0)
            if empty:
                model_inputs["cand_vis"], model_inputs["hist_vis"], model_inputs["obj_vis"] = cand_vis, hist_vis, obj_vis
            return model_inputs

        ml.ModifiedLlamaForCausalLM.prepare_inputs_for_generation = prepare_inputs_for_generation
        gold = torch.load(Path(__file__).resolve().parent / f"nav_{precision}.pt", weights_only=False)
        model.load_state_dict(gold["state_dict"], strict=True)
        model.eval()
        captured = {}
        orig = model.lang_model.generate

        def spy(*a, **k):
            ids = orig(*a, **k)
            captured["ids"] = ids.clone()
            captured["prompt_len"] = k["input_ids"].shape[1]
            return ids
        model.lang_model.generate = spy
        out = {"meta": {"transformers": tfv, "precision": precision, "max_new_tokens": 10}}
        with torch.no_grad():
            sent = model("3dqa", dict(gold["qa_in"]), training=False, max_new_tokens=10, do_sample=False)["generated_sentences"]
        out.update(ids=captured["ids"], prompt_len=captured["prompt_len"], sentences=sent)
        # summarization generation branch (models/nav_model.py:320-341): <hist> + <cand> visual tokens, max_new_tokens=50
        # fixed by the reference, greedy; once free and once constrained by a Trie (tools/trie.py; caller
        # tasks/agents/mp3d_agent.py:545-556 builds it from the candidate answers)
        from tools.trie import Trie
        tok = model.lang_model.tokenizer
        with torch.no_grad():
            s_free = model("summarization", dict(gold["sum_in"]), training=False)["generated_sentences"]
        out.update(sum_ids=captured["ids"], sum_prompt_len=captured["prompt_len"], sum_sentences=s_free)
        words = [[21, 22, 23], [21, 22, 30, 31], [21, 40], [50, 51, 52]]
        trie = Trie(tok.bos_token_id, tok.eos_token_id)
        for w in words:
            trie.insert(w)
        with torch.no_grad():
            s_trie = model("summarization", dict(gold["sum_in"]), training=False, trie=trie)["generated_sentences"]
        out.update(trie_words=words, trie_ids=captured["ids"], trie_sentences=s_trie)
        path = Path(__file__).resolve().parent / f"generate_{precision}.pt"
        torch.save(out, path)
        print("wrote", path, out["ids"][:, out["prompt_len"]:].tolist(), "| sum:", out["sum_ids"][:, out["sum_prompt_len"]:][:, :8].tolist(),
              "| trie:", out["trie_ids"][:, out["sum_prompt_len"]:].tolist())


if __name__ == "__main__":
    main()
