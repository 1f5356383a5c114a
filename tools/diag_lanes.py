import os, sys, torch
sys.path.insert(0, os.getcwd())
import rag_gesture_b200 as R
from rag_gesture_b200 import _lib, config as C, synthetic as S
dev = torch.device("cuda:0")
model = R.build_submodule(dict(C.denoiser_cfg(), precision=_lib.PREC_BF16), database=None, use_retrieval_for_test=False)
model.load_state_dict(S.synthetic_state_dict(0), strict=False)
model = model.to(dev).eval()
diff = R.build_diffusion(C.diffusion_test_cfg())
eng = model.rg_engine(diff)
def ev(fn, n=20):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for B in (64, 96, 128, 160, 224):
    cond = S.synthetic_conditions(B, seed=5)
    xf = eng.encode_conditions(cond["word"].to(dev), cond["audio"].to(dev), cond["speaker_ids"].to(dev))
    state = eng.precompute_state(xf)
    x = S.synthetic_latents(B, seed=6).to(dev); sm = S.motion_mask(B).to(dev)
    qm = torch.stack([S.query_masks(B)[c] for c in C.CONDS], 0).to(dev).contiguous()
    out = torch.empty_like(x)
    res = []
    for lanes in (1, 2, 3):
        eng.set_lanes(lanes)
        res.append(ev(lambda: eng.denoise_groups(x, sm, qm, state, [(B * 2 // 5, 30), (B - B * 2 // 5, 12)], out=out)))
    eng.set_lanes(0)
    print(f"B={B}: graph-replayed grouped evaluation, lanes 1/2/3: " + " / ".join(f"{t:.3f}" for t in res) + " ms", flush=True)
