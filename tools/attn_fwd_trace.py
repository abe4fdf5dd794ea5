"""Developer tool: in-kernel phase trace of the attention FORWARD kernel (csrc/attn_fwd.cu).

    NV_NVCC_EXTRA=-DNV_ATTN_TRACE python -m navillm_b200.build --force
    python tools/attn_fwd_trace.py [dense] > gpurun_out/attn_fwd_trace.txt
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
    dense = len(sys.argv) > 1 and sys.argv[1] == "dense"
    seqlens = [2048] * 4 if dense else np.random.RandomState(1234).randint(256, 1025, size=16).tolist()
    T = sum(seqlens)
    qkv = torch.randn(T, 3 * H * HD, device=dev, dtype=torch.bfloat16)
    cu = torch.tensor([0] + list(np.cumsum(seqlens)), dtype=torch.int32, device=dev)
    for _ in range(3):
        ops.attn_fwd(qkv, cu, seqlens, H)
    torch.cuda.synchronize()
    L = load()
    L.nv_debug_attn_trace.restype = ctypes.c_int
    L.nv_debug_attn_trace.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    buf = np.zeros(256 * 64, dtype=np.uint64)
    n = L.nv_debug_attn_trace(2, buf.ctypes.data, buf.size)
    if n == 0:
        print("library built without -DNV_ATTN_TRACE")
        return
    tr = buf.reshape(256, 64).astype(np.int64)
    ncta = sum((s + 255) // 256 for s in seqlens)
    g0 = min(tr[c, 5] for c in range(ncta) if tr[c, 0])
    print("seqlens", seqlens, "CTAs per head", ncta)
    print("cta nblk sm start_us | setup q+k0 | total | lastPV->epi_end | MMA per block j: (S(j+1) issued, PV0(j), PV1(j)) ; softmax tile0 per block: "
          "(s_full seen, exp done, pv_done waited, p_ready arrived)")
    for c in range(ncta):
        t = tr[c]
        if t[0] == 0:
            continue
        e = t[0]
        nb = int(t[3])
        mma = [tuple(int(t[10 + 3 * j + k] - e) if t[10 + 3 * j + k] else -1 for k in range(3)) for j in range(min(nb, 8))]
        sm = [tuple(int(t[34 + 4 * j + k] - e) if t[34 + 4 * j + k] else -1 for k in range(4)) for j in range(min(nb, 6))]
        print(f"{c:3d} {nb:2d} {int(t[4]):3d} {(t[5] - g0) / 1e3:7.1f} | {int(t[1] - e):5d} {int(t[2] - e):6d} | {int(t[60] - e):6d} | "
              f"{int(t[59] - t[58]) if t[59] else -1:5d} | {mma} ; {sm}")


if __name__ == "__main__":
    main()
