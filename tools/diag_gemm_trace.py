"""Phase timeline inside the tcgen05 GEMM (rg_probe_gemm_trace): medians over CTAs, in SM cycles."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_gesture_b200 import _lib

lib = _lib.load()
torch.zeros(1, device="cuda")
names = ["prologue", "pdl-wait", "1st stage", "mainloop", "acc->epi", "tmem->smem", "stores", "TOTAL"]
shapes = [(2752, 512, 512), (2752, 1536, 512), (2752, 1024, 512), (2752, 512, 2048)]
for split in (0,):
    for (M, N, K) in shapes:
        for pred in (0, 1, 2):
            ctas = (N // 128) * ((M + 127) // 128)
            buf = np.zeros(ctas * 10, dtype=np.int64)
            _lib.check(lib.rg_probe_gemm_trace(M, N, K, split, pred, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), buf.size,
                                               _lib.stream_ptr()))
            t = buf.reshape(ctas, 10)
            d = np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4],
                          t[:, 6] - t[:, 5], t[:, 7] - t[:, 6], t[:, 7] - t[:, 0]], 1)
            gt = t[:, 8] - t[:, 8].min()
            med = np.median(d, 0).astype(int)
            print(f"split={split} M={M} N={N} K={K} epi={pred} ctas={ctas} sms={len(set(t[:, 9]))}: " +
                  " ".join(f"{n}={v}" for n, v in zip(names, med)) +
                  f" | total max {d[:, 7].max()} | CTA start spread ns: med {int(np.median(gt))} max {gt.max()}")
