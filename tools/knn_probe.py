"""kNN sweep outside bench.py: exact scan vs tensor-core path at N x 768, k=8 (timing + equality check).

    python tools/knn_probe.py [N] [Q ...]
"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_gesture_b200 import _lib  # noqa: E402
from rag_gesture_b200.parallel import KnnIndex, knn_topk  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
QS = [int(a) for a in sys.argv[2:]] or [64, 256, 4096]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
db = torch.nn.functional.normalize(torch.randn(N, 768, device=dev, generator=g), dim=1)
t0 = time.perf_counter()
index = KnnIndex(db)
torch.cuda.synchronize()
print(f"index build {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


for Q in QS:
    q = torch.nn.functional.normalize(torch.randn(Q, 768, device=dev, generator=g), dim=1)
    ms_tc = timed(lambda: knn_topk(db, q, 8, index=index))
    i_tc, s_tc = knn_topk(db, q, 8, index=index)
    unc = index.last_uncertified
    line = f"Q={Q}: tc {ms_tc:.3f} ms = {Q / ms_tc * 1e3:.0f} q/s, {2.0 * Q * N * 768 / ms_tc / 1e9:.1f} TF/s, uncertified {unc}"
    if Q <= 512 or N <= 200_000:
        ms_ex = timed(lambda: knn_topk(db, q, 8), reps=3)
        i_ex, s_ex = knn_topk(db, q, 8)
        line += f"; exact {ms_ex:.3f} ms; identical idx {torch.equal(i_tc, i_ex)} scores {torch.equal(s_tc, s_ex)}"
    ms = ctypes.c_float()
    _lib.check(_lib.load().rg_probe_knn_tc(index.handle, _lib.ptr(q), Q, 5, _lib.ptr(flush), flush.numel(),
                                           ctypes.byref(ms), _lib.stream_ptr()))
    line += f"; knn_tc_kernel alone {ms.value:.3f} ms = {2.0 * Q * N * 768 / ms.value / 1e9:.1f} TF/s"
    print(line, flush=True)
