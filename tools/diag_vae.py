"""Device time of the TransformerVAE codec at the bench shape (64 clips + 96 exemplars encoded, 64 clips decoded),
per GEMM tier, and the kernels it spends that time in (torch profiler).  python tools/diag_vae.py"""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_gesture_b200 import config as C, synthetic as S  # noqa: E402
from rag_gesture_b200.vae import GestureRepEncoder  # noqa: E402

dev = torch.device("cuda:0")
with tempfile.TemporaryDirectory() as root:
    cfg = S.write_vae_files(root, latent_dim=C.LATENT_DIM, num_heads=4, ff_size=1024, num_layers=4)
    enc = GestureRepEncoder(cfg, "time").to(dev).eval()
g = torch.Generator().manual_seed(1)
F = C.MAX_SEQ_LEN


def inputs(B):
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    return dict(motion_upper=0.4 * r(B, F, 39), motion_lower=0.4 * r(B, F, 27), motion_face=0.2 * r(B, F, 3),
                motion_hands=0.3 * r(B, F, 90), motion_transl=0.5 * r(B, F, 3), motion_facial=0.5 * r(B, F, 100),
                motion_contact=(r(B, F, 4) > 0).float(), motion_mask=torch.ones(B, F, device=dev))


clips, ex = inputs(64), inputs(96)


def step():
    m, _ = enc.encode(**{k: v.clone() for k, v in clips.items()})
    enc.encode_many(**{k: v.clone() for k, v in ex.items()})
    return enc.decode(m)


for tier, graphs in ((None, False), ("bf16x3", False), ("bf16", False), ("bf16x3", True), (None, True)):
    enc.set_gemm_tier(tier)
    enc.enable_graphs(graphs)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    a.record()
    for _ in range(5):
        step()
    b.record()
    host = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    print(f"tier {tier} graphs {graphs}: device {a.elapsed_time(b) / 5:.2f} ms per step, host enqueue {1e3 * host:.2f} ms")
    if tier == "bf16x3" and not graphs:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
