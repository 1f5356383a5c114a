"""Diagnostics: CUDA-event timing of one rg_denoise call and of the K6 precompute at the bench's
batch sizes.  python tools/diag_step.py [precision] [reps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_gesture_b200 as R  # noqa: E402
from rag_gesture_b200 import _lib, config as C, synthetic as S  # noqa: E402

prec = {"bf16": _lib.PREC_BF16, "bf16x3": _lib.PREC_BF16X3, "fp32": _lib.PREC_FP32}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
model = R.build_submodule(dict(C.denoiser_cfg(), precision=prec), database=None, use_retrieval_for_test=False)
model.load_state_dict(S.synthetic_state_dict(0), strict=False)
model = model.to(dev).eval()
diff = R.build_diffusion(C.diffusion_test_cfg())
eng = model.rg_engine(diff)


def ev_time(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    t_launch = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, t_launch


for B in [int(b) for b in os.environ.get('DIAG_B', '64,96,160').split(',')]:
    cond = S.synthetic_conditions(B, seed=5)
    xf = eng.encode_conditions(cond["word"].to(dev), cond["audio"].to(dev), cond["speaker_ids"].to(dev))
    t_state, _ = ev_time(lambda: eng.precompute_state(xf), 3)
    state = eng.precompute_state(xf)
    x = S.synthetic_latents(B, seed=6).to(dev)
    sm = S.motion_mask(B).to(dev)
    qm = torch.stack([S.query_masks(B)[c] for c in C.CONDS], 0).to(dev).contiguous()
    out = torch.empty_like(x)
    for lanes in [int(v) for v in os.environ.get("DIAG_LANES", "1,2,3,4").split(",")]:
        eng.set_lanes(lanes)
        t_dn, t_cpu = ev_time(lambda: eng.denoise(x, sm, qm, state, step_idx=10, out=out), reps)
        print(f"B={B:4d} lanes={lanes}: rg_denoise {t_dn:7.3f} ms GPU ({t_cpu:6.3f} ms host enqueue)")
    eng.set_lanes(0)
    t_dn, t_cpu = ev_time(lambda: eng.denoise(x, sm, qm, state, step_idx=10, out=out), reps)
    t_up, _ = ev_time(lambda: eng.ddim_update(x, out, 10, -1, out=out), reps)
    print(f"B={B:4d}: rg_denoise {t_dn:7.3f} ms GPU ({t_cpu:6.3f} ms host enqueue) = "
          f"{B * 3.348 / t_dn:7.1f} TFLOP/s algorithmic; K6 state {t_state:7.2f} ms; ddim_update {t_up * 1e3:6.1f} us")
