"""Host-side timeline of GuidedPipeline on the bench workload (64 clips, 4096-entry DB): per batch, how long the
main thread waits for stage 1, spends in the deferred condition encode, in run_pass (enqueue) and in finish, and how
long stage 1 itself takes on the worker.  python tools/diag_e2e.py [n_batches]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rag_gesture_b200 as R  # noqa: E402
from rag_gesture_b200 import _lib, config as C, synthetic as S  # noqa: E402
from rag_gesture_b200.architecture import GuidedPipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prio = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vae = len(sys.argv) > 3 and sys.argv[3] == "vae"
dev = torch.device("cuda:0")
cfg = C.model_cfg()
cfg["use_retrieval_for_test"] = True
cfg["model"]["precision"] = _lib.PREC_BF16
if vae:
    import tempfile
    with tempfile.TemporaryDirectory() as root:
        cfg["model"]["vae_cfg"] = S.write_vae_files(root, latent_dim=C.LATENT_DIM, num_heads=4, ff_size=1024, num_layers=4)
        arch = R.build_architecture(cfg, database=S.SyntheticGestureDataset(bench.N_DB, seed=7))
else:
    arch = R.build_architecture(cfg, database=S.SyntheticGestureDataset(bench.N_DB, seed=7))
arch.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
arch = arch.to(dev).eval()
batch = bench.make_batch(0, bench.B_PER_GPU)
db = arch.model.database
T = {"stage1": [], "  h2d": [], "  encode": [], "  retrieve": [], "  encode_many": [], "  db.forward": [], "  precond": [], "wait": [],
     "cond": [], "pass": [], "finish": []}


def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T[name].append(time.perf_counter() - t0)
        return r
    return w


pipe = GuidedPipeline(arch, side_priority=prio)
print("side stream priority", prio)
arch._scatter = timed("  h2d", arch._scatter)
codec = arch.model.gesture_rep_encoder
codec.encode = timed("  encode", codec.encode)
if hasattr(codec, "encode_many"):
    codec.encode_many = timed("  encode_many", codec.encode_many)
codec.decode = timed("  decode", codec.decode)
T["  decode"] = []
db.retrieve_many = timed("  retrieve", db.retrieve_many)
db.forward = timed("  db.forward", db.forward)
arch.model.get_precompute_condition = timed("  precond", arch.model.get_precompute_condition)
pipe._stage1 = timed("stage1", pipe._stage1)
arch.encode_clip_conditions = timed("cond", arch.encode_clip_conditions)
arch.run_pass = timed("pass", arch.run_pass)
arch.finish = timed("finish", arch.finish)


MARKS = []


def mark(name, fn):
    def w(*a, **k):
        MARKS.append((name + ">", time.perf_counter()))
        r = fn(*a, **k)
        MARKS.append((name + "<", time.perf_counter()))
        return r
    return w


db.forward = mark("forward", db.forward)
db.retrieve_many = mark("retrieve", db.retrieve_many)
db.exemplar_corpus = mark("corpus", db.exemplar_corpus)
if hasattr(codec, "encode_many"):
    codec.encode_many = mark("encode_many", codec.encode_many)
db.place_window = mark("place", db.place_window)


def gen(k):
    for _ in range(k):
        for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
            d.clear()
        yield dict(batch, inference_kwargs=bench.infer_kwargs())


for _ in pipe.run(gen(3)):
    pass
torch.cuda.synchronize()
for v in T.values():
    v.clear()
t0 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
marks = []
for res in pipe.run(gen(n)):
    marks.append(time.perf_counter() - t0)
b.record()
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"{n} batches in {tot * 1e3:.1f} ms host / {a.elapsed_time(b):.1f} ms device = {tot / n * 1e3:.1f} ms per batch")
for k, v in T.items():
    print(f"  {k:12s} n={len(v):2d}  mean {1e3 * sum(v) / max(1, len(v)):7.2f} ms   " + " ".join(f"{1e3 * x:6.1f}" for x in v[:10]))
# phases inside the LAST db.forward call
last = max(i for i, (n_, _) in enumerate(MARKS) if n_ == "forward>")
seq, seen = [], set()
for n_, t in MARKS[last:]:
    if n_.startswith("place") and n_ in seen:
        continue
    seen.add(n_)
    seq.append((n_, t))
place_end = max(t for n_, t in MARKS[last:] if n_ == "place<")
print("  last db.forward: " + "  ".join(f"{n_} +{1e3 * (t - seq[0][1]):.1f}" for n_, t in seq) + f"  last place< +{1e3 * (place_end - seq[0][1]):.1f}")
print("  yields at (ms): " + " ".join(f"{1e3 * m:.0f}" for m in marks))
