"""Sequential forward vs GuidedPipeline on the same seeded host batches: per-batch max |diff|."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_gesture_b200 as R  # noqa: E402
from rag_gesture_b200 import config as C, synthetic as S  # noqa: E402
from rag_gesture_b200.architecture import GuidedPipeline  # noqa: E402

dev = torch.device("cuda:0")
cfg = C.model_cfg()
cfg["use_retrieval_for_test"] = True
arch = R.build_architecture(cfg, database=S.SyntheticGestureDataset(1200, seed=7))
arch.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
arch = arch.to(dev).eval()
qs = S.SyntheticGestureDataset(48, seed=8)
KEYS = ("prev_latentout", "pred_upper")


def batches():
    for ids in ([1, 2, 4], [7, 8], [10, 11, 12, 13], [20]):
        b = S.collate([qs[i] for i in ids])
        b["retrieval_method"] = "discourse"
        b["inference_kwargs"] = dict(use_inversion=True, outpaint=False, inversion_start_time=-1,
                                     insertion_guidance=True, guidance_iters=[0] * 25 + list(range(25)), guidance_lr=0.1)
        yield b


def reset():
    db = arch.model.database
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    torch.manual_seed(5)
    torch.cuda.manual_seed(6)


def grab(r):
    out = {k: r[k].cpu() for k in KEYS}
    out["n_ex"] = sum(len(x) for x in r["retrieval_dict"]["retr_startends"])
    return out


reset(); s1 = [grab(arch(**b)) for b in batches()]
reset(); s2 = [grab(arch(**b)) for b in batches()]
reset(); p1 = [grab(r) for r in GuidedPipeline(arch).run(batches())]
reset(); p2 = [grab(r) for r in GuidedPipeline(arch).run(batches())]
for name, a, b in (("seq vs seq", s1, s2), ("seq vs pipe", s1, p1), ("pipe vs pipe", p1, p2)):
    for i, (x, y) in enumerate(zip(a, b)):
        print(name, "batch", i, "n_ex", x["n_ex"], y["n_ex"],
              {k: float((x[k] - y[k]).abs().max()) for k in KEYS}, flush=True)
