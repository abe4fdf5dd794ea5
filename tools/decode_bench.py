"""C3 (BASELINE.json configs[2]): ScanQA-shaped greedy generation, B=8, 256 visual <cand> tokens + 64 text tokens
(S0 = 320), 128 new tokens, Vicuna-7B random init.  Reports prefill ms, ms/token of the CUDA-graph decode step and
the HBM roofline fraction (weights 13.21 GB + KV read per token)."""
import json
import sys
import time
import types
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    model = bench.build_model(dev)
    model.eval()
    lm = model.lang_model
    lm._ensure() if lm.core is not None else model._ensure()
    B, NV, NT, NEW = 8, 256, 64, 128
    tok = lm.tokenizer
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(3, 31999, (B, NV + NT), generator=g)
    ids[:, 0] = tok.bos_token_id
    ids[:, 8:8 + NV] = lm.cand_token_id[0]
    mask = torch.ones_like(ids)
    cand = torch.randn(B * NV, 4096, generator=g).to(dev)
    for graph in (True,):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = lm.generate(input_ids=ids, attention_mask=mask, cand_vis=cand, max_new_tokens=NEW, stop_on_eos=False,
                          use_cuda_graph=graph)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        out = lm.generate(input_ids=ids, attention_mask=mask, cand_vis=cand, max_new_tokens=NEW, stop_on_eos=False,
                          use_cuda_graph=graph)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        # prefill alone
        out1 = lm.generate(input_ids=ids, attention_mask=mask, cand_vis=cand, max_new_tokens=1, stop_on_eos=False)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        total, prefill = t2 - t1, t3 - t2
        per_tok = (total - prefill) / (NEW - 1)
        bytes_tok = 2 * (6.476e9 + 0.131e9) + B * (NV + NT + NEW / 2) * 524288
        pk = bench.peaks()
        print(json.dumps({"config": "C3 generate B=8 S0=320 new=128 Vicuna-7B", "cuda_graph": graph, "total_s": total,
                          "prefill_s": prefill, "ms_per_token": per_tok * 1e3, "tokens_per_s": B / per_tok,
                          "hbm_GBps_achieved": bytes_tok / per_tok / 1e9, "hbm_peak_GBps": pk["hbm_gbs"],
                          "hbm_frac": bytes_tok / per_tok / 1e9 / pk["hbm_gbs"], "first_call_s": t1 - t0,
                          "out_shape": list(out.shape)}), flush=True)


if __name__ == "__main__":
    main()
