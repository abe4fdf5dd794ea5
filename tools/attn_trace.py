"""Developer tool: in-kernel phase trace of the attention-backward dq pass.

Build the library with the trace compiled in, run on the GPU box, print per-CTA timelines (SM clocks):

    NV_NVCC_EXTRA=-DNV_ATTN_TRACE python -m navillm_b200.build
    python tools/attn_trace.py > gpurun_out/attn_trace.txt
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from navillm_b200 import ops  # noqa: E402
from navillm_b200._lib import load  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    H, HD = 32, 128
    seqlens = np.random.RandomState(1234).randint(256, 1025, size=16).tolist()
    T = sum(seqlens)
    qkv = torch.randn(T, 3 * H * HD, device=dev, dtype=torch.bfloat16)
    do = torch.randn(T, H * HD, device=dev, dtype=torch.bfloat16)
    cu = torch.tensor([0] + list(np.cumsum(seqlens)), dtype=torch.int32, device=dev)
    o, lse = ops.attn_fwd(qkv, cu, seqlens, H)
    dqkv = torch.empty_like(qkv)
    for _ in range(3):
        ops.attn_bwd(qkv, o, do, lse, cu, seqlens, H, dqkv=dqkv)
    torch.cuda.synchronize()
    L = load()
    L.nv_debug_attn_trace.restype = ctypes.c_int
    L.nv_debug_attn_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    buf = np.zeros(256 * 64, dtype=np.uint64)
    kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 1           # 1 = dq pass (default), 0 = dk/dv pass
    n = L.nv_debug_attn_trace(kernel, buf.ctypes.data, buf.size)
    if n == 0:
        print("library built without -DNV_ATTN_TRACE")
        return
    print("kernel", "dq" if kernel == 1 else "dkv")
    tr = buf.reshape(256, 64).astype(np.int64)
    nblk = sum((s + 127) // 128 for s in seqlens)
    g0 = min(tr[c, 5] for c in range(nblk) if tr[c, 0])
    print("seqlens", seqlens, "blocks", nblk)
    print("cta n_sub sm  start_us | setup res_wait | total | done->epi epi->exit | per-sub events (rel. to entry): "
          "[mma_sdp_issued, cw_sdp_ready, cw_ds_written, mma_ds_seen]")
    for c in range(nblk):
        t = tr[c]
        if t[0] == 0:
            continue
        e = t[0]
        n_sub = int(t[3])
        subs = []
        for k in range(min(n_sub, 8)):
            subs.append((int(t[10 + 2 * k] - e), int(t[30 + 2 * k] - e), int(t[31 + 2 * k] - e), int(t[11 + 2 * k] - e)))
        print(f"{c:3d} {n_sub:2d} {int(t[4]):3d} {(t[5] - g0) / 1e3:8.1f} | {int(t[1] - e):5d} {int(t[2] - e):6d} | "
              f"{int(t[52] - e):6d} | {int(t[51] - t[50]):5d} {int(t[52] - t[51]):5d} | {subs}")


if __name__ == "__main__":
    main()
