"""Wall-clock split of MotionDiffusion.prepare on the bench batch (no profiler overhead)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import rag_gesture_b200 as R
from rag_gesture_b200 import _lib, config as C, synthetic as S

dev = torch.device("cuda:0")
cfg = C.model_cfg(); cfg["use_retrieval_for_test"] = True; cfg["model"]["precision"] = _lib.PREC_BF16
arch = R.build_architecture(cfg, database=S.SyntheticGestureDataset(bench.N_DB, seed=7))
arch.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
arch = arch.to(dev).eval()
batch = bench.make_batch(0, 64)
db = arch.model.database
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3

orig_retrieve, orig_stack = db.retrieve, torch.stack
def timed_retrieve(*a, **k):
    t0 = time.perf_counter(); r = orig_retrieve(*a, **k); T["retrieve(64x)"] = T.get("retrieve(64x)", 0.0) + (time.perf_counter() - t0) * 1e3; return r
db.retrieve = timed_retrieve
orig_rank = None
for it in range(3):
    T.clear()
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds): d.clear()
    t0 = time.perf_counter(); kw = arch._scatter(dict(batch)); tick("scatter H2D", t0)
    t0 = time.perf_counter()
    gb = arch.prepare(**dict(batch, inference_kwargs=bench.infer_kwargs())); tick("prepare total (incl. scatter again)", t0)
    idx = db.text_index(dev)
    t0 = time.perf_counter()
    for _ in range(77): idx.rank(batch["text_features"][0].to(dev), list(range(20)), 20)
    tick("77 x rank(20 rows)", t0)
    smps_t0 = time.perf_counter(); smps = [db.dataset[n] for _, _, n in [(0, 0, db.dataset.names[i]) for i in range(96)]]; tick("96 dataset fetches", smps_t0)
    t0 = time.perf_counter()
    keys = ("motion_upper", "motion_lower", "motion_face", "motion_hands", "trans", "facial", "contact", "motion_mask", "word", "audio", "speaker_id", "motion")
    ex = {k: torch.stack([s_[k] for s_ in smps], 0).to(dev, non_blocking=True) for k in keys}; tick("stack + H2D of 96 exemplars", t0)
print({k: round(v, 1) for k, v in T.items()})
