"""Time attention forward/backward kernels at the C2 shape (16 ragged sequences U{256..1024}, 32 heads)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from navillm_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
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
    H, HD = 32, 128
    for name, seqlens in (("c2_ragged", np.random.RandomState(1234).randint(256, 1025, size=16).tolist()),
                          ("dense_1024x16", [1024] * 16), ("dense_2048x4", [2048] * 4), ("short_128x128", [128] * 128),
                          ("short_256x64", [256] * 64)):
        T = sum(seqlens)
        qkv = torch.randn(T, 3 * H * HD, device=dev, dtype=torch.bfloat16)
        do = torch.randn(T, H * HD, device=dev, dtype=torch.bfloat16)
        cu = torch.tensor([0] + list(np.cumsum(seqlens)), dtype=torch.int32, device=dev)
        o, lse = ops.attn_fwd(qkv, cu, seqlens, H)
        dqkv = torch.empty_like(qkv)
        fwd = timeit(lambda: ops.attn_fwd(qkv, cu, seqlens, H, out=o, lse=lse))
        bwd = timeit(lambda: ops.attn_bwd(qkv, o, do, lse, cu, seqlens, H, dqkv=dqkv))
        flops_f = sum(2 * 4096 * float(s) * s for s in seqlens)          # causal QK^T + PV, all heads
        print(f"{name}: T={T} fwd {fwd:.3f} ms ({flops_f / fwd / 1e9:.0f} TFLOP/s)  bwd {bwd:.3f} ms "
              f"({2.5 * flops_f / bwd / 1e9:.0f} TFLOP/s algorithmic)", flush=True)


if __name__ == "__main__":
    main()
