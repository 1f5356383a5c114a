"""Diagnostics: one rg_denoise evaluation (full, and GEMM-only chain) per GEMM kernel choice
(1 = 128x128 one-tile-per-CTA, 2 = persistent 2-CTA) at several batch sizes.
python tools/diag_gemm2.py [precision] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_gesture_b200 as R  # noqa: E402
from rag_gesture_b200 import _lib, config as C, ops, synthetic as S  # noqa: E402

prec = {"bf16": _lib.PREC_BF16, "bf16x3": _lib.PREC_BF16X3}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
model = R.build_submodule(dict(C.denoiser_cfg(), precision=prec), database=None, use_retrieval_for_test=False)
model.load_state_dict(S.synthetic_state_dict(0), strict=False)
model = model.to(dev).eval()
diff = R.build_diffusion(C.diffusion_test_cfg())
eng = model.rg_engine(diff)
lib = _lib.load()


def ev_time(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for B in [int(b) for b in os.environ.get("DIAG_B", "32,64,96,160,256").split(",")]:
    cond = S.synthetic_conditions(B, seed=5)
    xf = eng.encode_conditions(cond["word"].to(dev), cond["audio"].to(dev), cond["speaker_ids"].to(dev))
    state = eng.precompute_state(xf)
    x = S.synthetic_latents(B, seed=6).to(dev)
    sm = S.motion_mask(B).to(dev)
    qm = torch.stack([S.query_masks(B)[c] for c in C.CONDS], 0).to(dev).contiguous()
    out = torch.empty_like(x)
    step = lambda: eng.denoise(x, sm, qm, state, step_idx=10, out=out)
    res = {}
    for mode, pt in ((1, 0), (3, 0), (2, 1), (2, 10 ** 6)):
        ops.set_gemm_kernel(mode, 0, pt)
        t_full = ev_time(step, reps)
        res[(mode, pt)] = out.clone()
        _lib.check(lib.rg_probe_gemm_only(eng._h, 1))
        t_gemm = ev_time(step, reps)
        _lib.check(lib.rg_probe_gemm_only(eng._h, 0))
        name = {(1, 0): "128x128      ", (3, 0): "pair128      ", (2, 1): "2cta persist ", (2, 10 ** 6): "2cta one-tile"}[(mode, pt)]
        print(f"B={B:4d} M={B * 43:6d} kernel {name}: evaluation {t_full:7.3f} ms, GEMM-only chain {t_gemm:7.3f} ms = "
              f"{B * 3.291 / t_gemm:7.1f} TFLOP/s algorithmic", flush=True)
    ops.set_gemm_kernel(0, 0, 296)
    vals = list(res.values())
    print(f"B={B:4d}: outputs bit-identical across kernels: {all(torch.equal(vals[0], v) for v in vals[1:])}", flush=True)
