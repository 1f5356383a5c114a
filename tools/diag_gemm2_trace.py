"""Phase timeline inside the 2-CTA tcgen05 GEMM (rg_probe_gemm_trace with rg_set_gemm_kernel mode 2):
globaltimer stamps (ns) of the last two launches of a 6-launch PDL chain, medians over CTAs.
python tools/diag_gemm2_trace.py [persist_tiles]"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_gesture_b200 import _lib, ops

lib = _lib.load()
torch.zeros(1, device="cuda")
pt = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ops.set_gemm_kernel(2, 0, pt)
names = ["entry", "prologue", "pdl-wait", "loads issued", "1st stage", "last commit", "epi start", "epi issued", "stores done", "exit"]
for (M, N, K) in [(1376, 512, 512), (6880, 512, 512), (6880, 1536, 512), (6880, 512, 2048)]:
    for epi in (0, 1, 2):
        ctas10 = (N // 128) * ((M + 127) // 128) * 10
        buf = np.zeros(2 * ctas10, dtype=np.int64)
        _lib.check(lib.rg_probe_gemm_trace(M, N, K, 0, epi, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), buf.size,
                                           _lib.stream_ptr()))
        tiles = (N // 256) * ((M + 255) // 256)
        grid = 2 * (tiles if tiles < pt else min(tiles, 74))
        a = buf[:ctas10][: grid * 16].reshape(grid, 16)
        b = buf[ctas10:][: grid * 16].reshape(grid, 16)
        lead = b[0::2]                                  # leader CTAs have the MMA stamps
        t0 = b[:, 0].min()
        def med(col, arr=b):
            v = arr[:, col]
            v = v[v > 0]
            return int(np.median(v) - t0) if len(v) else -1
        line = " ".join(f"{n}={med(i, lead if i in (4, 5) else b)}" for i, n in enumerate(names))
        print(f"M={M} N={N} K={K} epi={epi} grid={grid} sms={len(set(b[:, 10]))}: {line} | "
              f"prev launch: first entry {int(a[:, 0].min() - t0)} last exit {int(a[:, 9].max() - t0)}; this: last exit {int(b[:, 9].max() - t0)}",
              flush=True)
ops.set_gemm_kernel(0, 0, 296)
