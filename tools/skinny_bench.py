"""Time the M <= 128 (decode) GEMM shapes of Vicuna-7B and report the weight-streaming bandwidth.
Usage: python tools/skinny_bench.py [pkg_root]   (pkg_root: another checkout to compare against)"""
import sys
from pathlib import Path

import torch

root = Path(sys.argv[1]).resolve() if len(sys.argv) > 1 else Path(__file__).resolve().parents[1]
sys.path.insert(0, str(root))
from navillm_b200 import ops  # noqa: E402


def timeit(fn, iters=30, warmup=5):
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
    print("package:", root)
    # rotate over 8 weight copies so that the weights come from HBM, not L2
    for name, N, K in (("qkv", 12288, 4096), ("o", 4096, 4096), ("gateup", 22016, 4096), ("down", 4096, 11008)):
        ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) for _ in range(8)]
        x = torch.randn(8, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(8, N, device=dev, dtype=torch.bfloat16)
        for bn in (128, 32):
            i = [0]

            def f():
                ops.gemm(x, ws[i[0] % 8], out=out, block_n=bn)
                i[0] += 1
            try:
                ms = timeit(f)
            except Exception as e:  # older checkouts have no 32-column variant
                print(f"  {name} bn={bn}: {e}")
                continue
            print(f"  {name:7s} M=8 N={N} K={K} bn={bn}: {ms * 1e3:7.1f} us  {N * K * 2 / ms / 1e6:7.0f} GB/s")
        if hasattr(ops, "gemm_skinny"):
            i = [0]

            def f2():
                ops.gemm_skinny(x, ws[i[0] % 8], out=out)
                i[0] += 1
            ms = timeit(f2)
            print(f"  {name:7s} M=8 N={N} K={K} swap-AB: {ms * 1e3:7.1f} us  {N * K * 2 / ms / 1e6:7.0f} GB/s")
        if name == "gateup" and hasattr(ops, "gemm_skinny_swiglu"):
            i = [0]
            h = torch.empty(8, N // 2, device=dev, dtype=torch.bfloat16)

            def f3():
                ops.gemm_skinny_swiglu(x, ws[i[0] % 8], out=h)
                i[0] += 1
            ms = timeit(f3)
            print(f"  {name:7s} M=8 N={N} K={K} swap-AB + SwiGLU: {ms * 1e3:7.1f} us  {N * K * 2 / ms / 1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
