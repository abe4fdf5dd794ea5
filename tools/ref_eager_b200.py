"""Context bar: what a plain PyTorch user of the REFERENCE gets on the same B200.

NOT the product and not a target: the oracle port of the reference LM (oracle/navillm_oracle.py::llama_model = HF LLaMA eager
attention, models/modified_lm.py:112-116) run on `cuda` in bf16 eager -- cuBLAS GEMMs, materialised [B,32,S,S] scores -- for
the C2 training step (fwd + bwd, 32 full-width layers), on a quarter batch (B=4 rows of C2's length distribution padded to
the longest, like the reference) so that the eager attention's saved activations stay far below HBM capacity, scaled to
nav-steps/s.  Beside it: library attention (torch SDPA / flash-attn varlen, fwd+bwd) at C2's attention shapes, i.e. what the
hand-written tcgen05 attention kernels are up against.

    python tools/ref_eager_b200.py > gpurun_out/ref_eager.json
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from oracle import navillm_oracle as O  # noqa: E402  (tools/ may use the checker; the product never does)


def timeit(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


def main():
    dev = torch.device("cuda:0")
    bf16 = torch.bfloat16
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    B = 4
    lens = np.random.RandomState(1234).randint(bench.LEN_LO, bench.LEN_HI + 1, size=16)[:B]
    S = int(lens.max())
    cfg = O.OracleConfig(hidden=bench.D_MODEL, n_layers=bench.N_LAYERS, n_heads=bench.N_HEADS, inter=bench.D_FF, vocab=64, precision="amp_bf16")
    with torch.device(dev):
        g = torch.Generator(device=dev).manual_seed(0)
        sd = {}
        for l in range(cfg.n_layers):
            p = f"lang_model.model.layers.{l}"
            for nm, shp in (("self_attn.q_proj", (4096, 4096)), ("self_attn.k_proj", (4096, 4096)), ("self_attn.v_proj", (4096, 4096)),
                            ("self_attn.o_proj", (4096, 4096)), ("mlp.gate_proj", (11008, 4096)), ("mlp.up_proj", (11008, 4096)),
                            ("mlp.down_proj", (4096, 11008))):
                sd[f"{p}.{nm}.weight"] = torch.empty(shp, dtype=bf16).normal_(0, 0.02, generator=g).requires_grad_(True)
            sd[f"{p}.input_layernorm.weight"] = torch.ones(4096, dtype=bf16, requires_grad=True)
            sd[f"{p}.post_attention_layernorm.weight"] = torch.ones(4096, dtype=bf16, requires_grad=True)
        sd["lang_model.model.norm.weight"] = torch.ones(4096, dtype=bf16, requires_grad=True)
        emb = torch.randn(B, S, 4096, generator=g).to(bf16).requires_grad_(True)
        mask = torch.zeros(B, S, dtype=torch.long)
        for b, L in enumerate(lens):
            mask[b, S - int(L):] = 1

        def step():
            for v in sd.values():
                v.grad = None
            h = O.llama_model(sd, cfg, emb, mask)
            h[:, -1].float().sum().backward()

        ms = timeit(step)
    out["reference_eager_bf16"] = {"rows": B, "padded_len": S, "real_tokens": int(lens.sum()), "ms_per_step_of_4_rows": ms,
                                   "nav_steps_per_s": B / ms * 1e3,
                                   "note": "oracle port of the reference LM (HF eager attention) on cuda, bf16, fwd+bwd, 32 layers; "
                                           "pano encoder/heads (<0.1% of the FLOPs) not included; peak memory GB below",
                                   "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    del sd, emb
    torch.cuda.empty_cache()

    # library attention at C2's attention shapes (16 sequences U{256..1024}, 32 heads, hd 128), fwd + bwd
    seqlens = np.random.RandomState(1234).randint(256, 1025, size=16).tolist()
    T = sum(seqlens)
    flops_f = sum(2 * 4096 * float(s) * s for s in seqlens)
    lib = {}
    try:
        from flash_attn import flash_attn_varlen_func
        q, k, v = (torch.randn(T, 32, 128, device=dev, dtype=bf16, requires_grad=True) for _ in range(3))
        cu = torch.tensor([0] + list(np.cumsum(seqlens)), dtype=torch.int32, device=dev)
        do = torch.randn(T, 32, 128, device=dev, dtype=bf16)
        fwd = timeit(lambda: flash_attn_varlen_func(q, k, v, cu, cu, max(seqlens), max(seqlens), causal=True), iters=10, warm=3)

        def fb():
            o = flash_attn_varlen_func(q, k, v, cu, cu, max(seqlens), max(seqlens), causal=True)
            o.backward(do)
        tot = timeit(fb, iters=10, warm=3)
        lib["flash_attn_2.8_varlen_c2_ragged"] = {"fwd_ms": fwd, "fwd_tflops": flops_f / fwd / 1e9, "bwd_ms": tot - fwd,
                                                  "bwd_tflops_algorithmic": 2.5 * flops_f / (tot - fwd) / 1e9}
    except Exception as e:  # pragma: no cover
        lib["flash_attn_2.8_varlen_c2_ragged"] = {"error": f"{type(e).__name__}: {e}"}
    try:
        import torch.nn.functional as F
        q, k, v = (torch.randn(4, 32, 2048, 128, device=dev, dtype=bf16, requires_grad=True) for _ in range(3))
        do = torch.randn_like(q)
        fl = 4 * 2 * 4096 * 2048.0 * 2048
        fwd = timeit(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=True), iters=10, warm=3)

        def fb2():
            F.scaled_dot_product_attention(q, k, v, is_causal=True).backward(do)
        tot = timeit(fb2, iters=10, warm=3)
        lib["torch_sdpa_dense_2048x4"] = {"fwd_ms": fwd, "fwd_tflops": fl / fwd / 1e9, "bwd_ms": tot - fwd,
                                          "bwd_tflops_algorithmic": 2.5 * fl / (tot - fwd) / 1e9}
    except Exception as e:  # pragma: no cover
        lib["torch_sdpa_dense_2048x4"] = {"error": f"{type(e).__name__}: {e}"}
    out["library_attention"] = lib
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
