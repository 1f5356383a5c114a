"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
through refshim.py) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py denoiser   # one group

The reference ships no tests or golden vectors, so these files are the parity pin for
oracle/ and, through it, for the CUDA path.  Inputs are never stored, only seeds: both sides
regenerate them with rag_gesture_b200.synthetic; `in_digest` guards against RNG drift.

Documented adjustments to the reference, none of which touch its arithmetic:
  * scale_func_cfg=None (the shipped value crashes in forward_test, raggesture.py:1102);
  * weights are loaded with load_state_dict from synthetic_state_dict (zero-initialised layers
    would make the model the zero function);
  * GestureRepEncoder (needs VAE yaml + checkpoints that are not in the repo) is replaced by a
    shape-only stand-in; the denoiser never calls it.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refshim  # noqa: E402
from rag_gesture_b200 import config as C  # noqa: E402
from rag_gesture_b200 import synthetic as S  # noqa: E402


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


class ShapeOnlyCodec(torch.nn.Module):
    def __init__(self, vae_cfg, body_part_cat_axis="time"):
        super().__init__()
        self.vae_latent_dim = vae_cfg["latent_dim"]
        self.body_part_cat_axis = body_part_cat_axis


def build_reference_denoiser(ns, sd):
    ns.dt.GestureRepEncoder = ShapeOnlyCodec
    cfg = C.denoiser_cfg()
    cfg.pop("type")
    model = ns.rg.ReGestureTransformer(**cfg, database=None, use_retrieval_for_test=False)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.eval()


def build_reference_diffusion(ns):
    return ns.arch.build_diffusion(C.diffusion_test_cfg())


def model_kwargs_for(model, cond, B):
    pc = model.get_precompute_condition(device="cpu", text=cond["word"], audio=cond["audio"],
                                        speaker_ids=cond["speaker_ids"], re_dict=1)
    return dict(xf_out=pc["xf_out"], re_dict=None, query_mask=S.query_masks(B),
                motion_mask=S.motion_mask(B), sample_idx=None)


def gen_schedule(ns):
    d = build_reference_diffusion(ns)
    save("schedule", timestep_map=np.array(d.timestep_map, dtype=np.int64),
         alphas_cumprod=d.alphas_cumprod, alphas_cumprod_prev=d.alphas_cumprod_prev,
         alphas_cumprod_next=d.alphas_cumprod_next,
         sqrt_recip_alphas_cumprod=d.sqrt_recip_alphas_cumprod,
         sqrt_recipm1_alphas_cumprod=d.sqrt_recipm1_alphas_cumprod,
         sqrt_alphas_cumprod=d.sqrt_alphas_cumprod,
         sqrt_one_minus_alphas_cumprod=d.sqrt_one_minus_alphas_cumprod)


def gen_denoiser(ns):
    """One denoiser evaluation (a clip-step without the DDIM update), B=2, two timesteps."""
    sd = S.synthetic_state_dict(0)
    model = build_reference_denoiser(ns, sd)
    B = 2
    cond = S.synthetic_conditions(B, seed=11)
    x = S.synthetic_latents(B, seed=12)
    kw = model_kwargs_for(model, cond, B)
    outs = {}
    with torch.no_grad():
        for tau in (14, 514, 999):
            ts = torch.full((B,), tau, dtype=torch.int64)
            outs[f"x0_t{tau}"] = model(x, ts, **{k: (dict(v) if isinstance(v, dict) else v)
                                                 for k, v in kw.items()})
    save("denoiser_step", in_digest=np.array(digest(x, cond["word"], cond["audio"], sd["out.weight"])),
         xf_text=kw["xf_out"]["xf_text"][:, :4], **outs)

    # normal-scale cross-attention values: rows 20/30 are NOT collapsed (see synthetic.py header)
    sdn = S.synthetic_state_dict(1, normal_scale=True)
    model.load_state_dict(sdn, strict=False)
    kw = model_kwargs_for(model, cond, B)
    with torch.no_grad():
        ts = torch.full((B,), 514, dtype=torch.int64)
        o = model(x, ts, **kw)
    save("denoiser_step_normal_scale", in_digest=np.array(digest(x, sdn["out.weight"])), x0_t514=o)


def gen_loops(ns):
    """Config 1 (plain 50-step DDIM, B=1), the reverse loop, and a guided loop (B=2)."""
    sd = S.synthetic_state_dict(0)
    model = build_reference_denoiser(ns, sd)
    diff = build_reference_diffusion(ns)
    T, D = C.N_TOKENS, C.LATENT_DIM

    # plain DDIM, B=1 (BASELINE.json configs[0]); noise from the global generator like the reference
    cond = S.synthetic_conditions(1, seed=21)
    kw = model_kwargs_for(model, cond, 1)
    torch.manual_seed(1234)
    traj = []
    with torch.no_grad():
        final = None
        for out in diff.ddim_sample_loop_progressive(model, (1, T, D), clip_denoised=False,
                                                     model_kwargs=kw, eta=0):
            traj.append(out["sample"])
            final = out["sample"]
    save("ddim_plain_b1", final=final, step49=traj[0], step25=traj[24], step1=traj[48])

    # inversion of one exemplar latent (B=1), all 50 levels
    xs = S.synthetic_latents(1, seed=22, scale=0.5)
    with torch.no_grad():
        inv = diff.ddim_reverse_sample_loop(model, start_img=xs, clip_denoised=False,
                                            model_kwargs=kw, eta=0, return_all_timesteps=True)
    inv = torch.cat(inv, dim=0)                                   # [50,43,512]
    save("ddim_reverse_b1", inv0=inv[0], inv24=inv[24], inv49=inv[49])

    # guided loop, B=2: exemplar windows (upper+hands rows) of the inverted latent inserted at two
    # different query windows; clip 1 additionally exercises an empty row set at one level
    B = 2
    cond2 = S.synthetic_conditions(B, seed=23)
    kw2 = model_kwargs_for(model, cond2, B)
    inv_list = torch.zeros(50, B, T, D)
    n = C.N_CHUNKS
    for b, (q0, q1, r0, r1) in enumerate([(2, 5, 4, 7), (6, 10, 0, 4)]):
        inv_list[:, b, q0:q1] = inv[:, r0:r1]
        inv_list[:, b, n + 1 + q0:n + 1 + q1] = inv[:, n + 1 + r0:n + 1 + r1]
    torch.manual_seed(4321)
    start = torch.randn(B, T, D)
    start[inv_list[49] != 0] = inv_list[49][inv_list[49] != 0]
    g_iters = [0] * 25 + list(range(25))                            # decreasing_till_25, visualize.py:90-91
    with torch.inference_mode(False):
        final_g = diff.ddim_guided_sample_loop(model, (B, T, D), noise=start.clone(),
                                               clip_denoised=False, model_kwargs=kw2, eta=0,
                                               in_seq=None, guidance_iters=g_iters,
                                               inverted_latent_list=inv_list, guidance_lr=0.1)
    torch.manual_seed(4321)
    start2 = torch.randn(B, T, D)
    start2[inv_list[49] != 0] = inv_list[49][inv_list[49] != 0]
    final_g0 = diff.ddim_guided_sample_loop(model, (B, T, D), noise=start2.clone(),
                                            clip_denoised=False, model_kwargs=kw2, eta=0,
                                            in_seq=None, guidance_iters=[0] * 50,
                                            inverted_latent_list=inv_list, guidance_lr=0.1)
    print("guided(g) == guided(0):", bool((final_g == final_g0).all()))
    # long-form mode: prev_latent as in_seq on every step of the plain loop (diffusion_architecture.py:449-462)
    prev = torch.zeros(1, T, D)
    prev[:, [0, n + 1, 2 * n + 2, 3 * n + 3]] = S.synthetic_latents(1, seed=24)[:, [9, 20, 31, 42]]
    torch.manual_seed(999)
    final_prev = diff.ddim_sample_loop(model, (1, T, D), clip_denoised=False, model_kwargs=kw,
                                       eta=0, in_seq=prev)
    save("ddim_guided_b2", final=final_g, dead_guidance_bit_identical=np.array(bool((final_g == final_g0).all())),
         final_prev_latent_b1=final_prev)


GROUPS = {"schedule": gen_schedule, "denoiser": gen_denoiser, "loops": gen_loops}

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    ns = refshim.load()
    which = sys.argv[1:] or list(GROUPS)
    for g in which:
        GROUPS[g](ns)
