"""Time the HBM-bound row kernels of the LLaMA stack at the C2 token count (T = 10 425, D = 4096) and report achieved GB/s
against the measured HBM peak.  Tensors rotate through 6 copies (> L2) so that reads come from HBM."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from navillm_b200 import ops  # noqa: E402


def timeit(fn, iters=30, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for i in range(iters):
        fn(i)
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


def main():
    dev = torch.device("cuda:0")
    T, D, NC = 10425, 4096, 6
    pk = bench.peaks()["hbm_gbs"]
    xs = [torch.randn(T, D, device=dev, dtype=torch.bfloat16) for _ in range(NC)]
    dys = [torch.randn(T, D, device=dev, dtype=torch.bfloat16) for _ in range(NC)]
    drs = [torch.randn(T, D, device=dev, dtype=torch.bfloat16) for _ in range(NC)]
    w = torch.randn(D, device=dev, dtype=torch.bfloat16)
    out = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(D, device=dev, dtype=torch.bfloat16)
    _, rstd = ops.rmsnorm_fwd(xs[0], w, 1e-6)
    row = T * D * 2
    ms = timeit(lambda i: ops.rmsnorm_fwd(xs[i % NC], w, 1e-6, out=out))
    print(f"rmsnorm_fwd        {ms * 1e3:7.1f} us  {2 * row / ms / 1e6:6.0f} GB/s  ({2 * row / ms / 1e6 / pk:.0%} of {pk:.0f})")
    ms = timeit(lambda i: ops.rmsnorm_bwd(xs[i % NC], w, rstd, dys[i % NC], dres=drs[i % NC], dx=out, dw=dw))
    print(f"rmsnorm_bwd+dres   {ms * 1e3:7.1f} us  {4 * row / ms / 1e6:6.0f} GB/s  ({4 * row / ms / 1e6 / pk:.0%})  [x, dy, dres read; dx written; + colsum]")
    ms = timeit(lambda i: ops.rmsnorm_bwd(xs[i % NC], w, rstd, dys[i % NC], dx=out, dw=dw))
    print(f"rmsnorm_bwd        {ms * 1e3:7.1f} us  {3 * row / ms / 1e6:6.0f} GB/s  ({3 * row / ms / 1e6 / pk:.0%})")


if __name__ == "__main__":
    main()
