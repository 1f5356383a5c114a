"""Host-side profile of MotionDiffusion.prepare / forward on the bench batch (cProfile)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rag_gesture_b200 as R  # noqa: E402
from rag_gesture_b200 import _lib, config as C, synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
cfg = C.model_cfg()
cfg["use_retrieval_for_test"] = True
cfg["model"]["precision"] = _lib.PREC_BF16
arch = R.build_architecture(cfg, database=S.SyntheticGestureDataset(bench.N_DB, seed=7))
arch.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
arch = arch.to(dev).eval()
batch = bench.make_batch(0, 64)


def once():
    db = arch.model.database
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    t0 = time.perf_counter()
    gb = arch.prepare(**dict(batch, inference_kwargs=bench.infer_kwargs()))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = arch.run_prepared(gb)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    arch.finish(gb, out)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


for _ in range(2):
    once()
print("prepare %.1f ms, run_prepared %.1f ms, finish %.1f ms" % tuple(1e3 * t for t in once()))
pr = cProfile.Profile()
pr.enable()
once()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
