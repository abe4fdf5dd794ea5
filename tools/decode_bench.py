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
    def run(n_new):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = lm.generate(input_ids=ids, attention_mask=mask, cand_vis=cand, max_new_tokens=n_new, stop_on_eos=False,
                          use_cuda_graph=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, out

    run(NEW)                                                   # warm-up (allocator, tensor maps)
    # per-token time from the difference of two lengths: prefill, graph capture and instantiation cancel
    t_long = min(run(NEW)[0] for _ in range(3))
    t_short = min(run(NEW // 2)[0] for _ in range(3))
    t_one, out1 = run(1)
    _, out = run(NEW)
    per_tok = (t_long - t_short) / (NEW - NEW // 2)
    bytes_tok = 2 * (6.476e9 + 0.131e9) + B * (NV + NT + 0.75 * NEW) * 524288     # weights + mean KV read of the second half
    pk = bench.peaks()
    print(json.dumps({"config": "C3 generate B=8 S0=320 new=128 Vicuna-7B", "cuda_graph": True, "total_s": t_long,
                      "prefill_plus_first_token_s": t_one, "ms_per_token": per_tok * 1e3, "tokens_per_s": B / per_tok,
                      "hbm_GBps_achieved": bytes_tok / per_tok / 1e9, "hbm_peak_GBps": pk["hbm_gbs"],
                      "hbm_frac": bytes_tok / per_tok / 1e9 / pk["hbm_gbs"],
                      "method": "(t[128 new] - t[64 new]) / 64, best of 3 each: prefill and graph capture cancel",
                      "out_shape": list(out.shape)}), flush=True)


if __name__ == "__main__":
    main()
