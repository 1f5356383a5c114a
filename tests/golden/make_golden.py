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


def make_ref_db_class(ns):
    """The reference RetrievalDatabase with only __init__ replaced: the in-RAM dicts it normally
    loads from its LMDB cache (raggesture.py:244-303) are filled from the synthetic dataset, in LMDB
    cursor order; retrieve() / forward() are the reference's own code."""
    class RefDB(ns.rg.RetrievalDatabase):
        def __init__(self, dataset=None, num_retrieval=1, topk=2, latent_dim=512, text_latent_dim=768,
                     max_seq_len=150, motion_fps=15, motion_framechunksize=15, **kw):
            torch.nn.Module.__init__(self)
            self.num_retrieval, self.topk, self.latent_dim = num_retrieval, topk, latent_dim
            self.text_latent_dim, self.max_seq_len = text_latent_dim, max_seq_len
            self.motion_fps, self.motion_framechunksize = motion_fps, motion_framechunksize
            self.retrieval_method = {"discourse": ns.discourse.discourse_retrieval,
                                     "gesture_type": ns.gesture_type.gesture_type_retrieval}
            self.train_indexes, self.test_indexes = {}, {}
            self.train_dbounds, self.test_dbounds = {}, {}
            self.train_qbounds, self.test_qbounds = {}, {}
            self.dataset = dataset
            d = {k: {} for k in ("text", "sense", "discbounds", "gesture_labels", "prominence", "gestprom")}
            for i in sorted(range(len(dataset)), key=lambda i: dataset.names[i].encode("ascii")):
                spk, disc, prom, gest, _ = dataset.annotations(i)
                name = dataset.names[i]
                d["text"][name] = [dataset.text_feature(i), spk]
                d["sense"][name] = [spk] + [(x[1], x[0]) for x in disc]
                d["discbounds"][name] = [(x[1], x[0], x[4], x[5], x[6], x[7]) for x in disc]
                d["gesture_labels"][name] = [spk] + list(gest)
                d["prominence"][name] = ns.rag_utils.map_conns_to_prominence([x[0] for x in disc], prom)
                d["gestprom"][name] = ns.rag_utils.map_conns_to_prominence([g["word"] for g in gest], prom)
            for k, v in d.items():
                setattr(self, "idx_2_" + k, v)
    return RefDB


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    return o


N_DB, N_QUERY = 1200, 48


def gen_retrieval(ns):
    """discourse_retrieval + sort_sidx_by_textsimilarity + RetrievalDatabase.forward window placement
    of the reference on the synthetic annotated dataset (queries = samples from a second dataset)."""
    import json
    from rag_gesture_b200.codec import SyntheticGestureCodec
    ds = S.SyntheticGestureDataset(N_DB, seed=7)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    db = make_ref_db_class(ns)(dataset=ds, **C.retrieval_cfg()).eval()
    out = {"queries": []}
    for i in range(N_QUERY):
        spk, disc, prom, gest, _ = qs.annotations(i)
        idx, bounds, qb = ns.discourse.discourse_retrieval(
            text="", discourse=disc, prominence=prom, speaker_id=spk, db_idx_2_sense=db.idx_2_sense,
            db_idx_2_discbounds=db.idx_2_discbounds, db_idx_2_prominence=db.idx_2_prominence,
            encoded_text=qs.text_feature(i), text_feat_cache=db.idx_2_text)
        out["queries"].append(_jsonable({"idx": idx, "bounds": bounds, "qbounds": qb}))
    names = list(db.idx_2_text.keys())[:200]
    out["sim_order"] = ns.rag_utils.sort_sidx_by_textsimilarity(names, "", qs.text_feature(0), db.idx_2_text)
    # window placement through the reference forward (codec: synthetic; RNG seeded)
    codec = SyntheticGestureCodec(C.denoiser_cfg()["vae_cfg"])
    batch = S.collate([qs[i] for i in range(N_QUERY)])
    cond = dict(text=batch["raw_word"], audio=batch["raw_audio"], text_enc=batch["word"],
                text_features=batch["text_features"], audio_enc=batch["audio"], discourse=batch["discourse"],
                prominence=batch["prominence"], speaker_ids=batch["speaker_ids"],
                gesture_labels=batch["gesture_labels"], text_times=batch["text_segments"])
    torch.manual_seed(5)
    with torch.no_grad():
        re = db(cond, batch["motion_length"], "cpu", idx=batch["sample_name"], retrieval_method="discourse",
                gesture_rep_encoder=codec)
    out["retr_startends"] = _jsonable(re["retr_startends"])
    out["query_startends"] = _jsonable(re["query_startends"])
    out["raw_sample_names"] = _jsonable(re["raw_sample_names"])
    out["re_mask_sum"] = re["re_mask"].sum(1).tolist()
    with open(os.path.join(HERE, "retrieval.json"), "w") as f:
        json.dump(out, f)
    save("retrieval_latents", raw_motion_latents=re["raw_motion_latents"][:4])
    print("wrote retrieval.json; exemplars placed:", sum(len(x) for x in re["retr_startends"]))


def gen_gesture_type(ns):
    """gesture_type_retrieval of the reference (with `fuzz.partial_ratio` restated, oracle/fuzz_ratio.py: the shipped
    word-similarity path, rag/utils.py:269-270) on the synthetic gesture labels, and RetrievalDatabase.forward with
    retrieval_method="gesture_type" (exemplar padding -0.2/+0.1 s above 0.9 s, raggesture.py:617-627)."""
    import json
    from rag_gesture_b200.codec import SyntheticGestureCodec
    ds = S.SyntheticGestureDataset(N_DB, seed=7)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    db = make_ref_db_class(ns)(dataset=ds, **C.retrieval_cfg()).eval()
    out = {"queries": []}
    for i in range(N_QUERY):
        spk, _, _, gest, _ = qs.annotations(i)
        idx, bounds, qb = ns.gesture_type.gesture_type_retrieval(
            text="", gesture_labels=gest, speaker_id=spk, db_idx_2_gesture_labels=db.idx_2_gesture_labels,
            encoded_text=qs.text_feature(i), text_feat_cache=db.idx_2_text)
        out["queries"].append(_jsonable({"idx": idx, "bounds": bounds, "qbounds": qb}))
    codec = SyntheticGestureCodec(C.denoiser_cfg()["vae_cfg"])
    batch = S.collate([qs[i] for i in range(N_QUERY)])
    cond = dict(text=batch["raw_word"], audio=batch["raw_audio"], text_enc=batch["word"],
                text_features=batch["text_features"], audio_enc=batch["audio"], discourse=batch["discourse"],
                prominence=batch["prominence"], speaker_ids=batch["speaker_ids"],
                gesture_labels=batch["gesture_labels"], text_times=batch["text_segments"])
    torch.manual_seed(5)
    with torch.no_grad():
        re = db(cond, batch["motion_length"], "cpu", idx=batch["sample_name"], retrieval_method="gesture_type",
                gesture_rep_encoder=codec)
    out["retr_startends"] = _jsonable(re["retr_startends"])
    out["query_startends"] = _jsonable(re["query_startends"])
    out["raw_sample_names"] = _jsonable(re["raw_sample_names"])
    out["raw_type2words"] = _jsonable(re["raw_type2words"])
    out["re_mask_sum"] = re["re_mask"].sum(1).tolist()
    out["latents_digest"] = digest(re["raw_motion_latents"])
    with open(os.path.join(HERE, "gesture_type.json"), "w") as f:
        json.dump(out, f)
    save("gesture_type_latents", raw_motion_latents=re["raw_motion_latents"][:4])
    print("wrote gesture_type.json; query labels:", sum(len(q["idx"]) for q in out["queries"]),
          "exemplars placed:", sum(len(x) for x in re["retr_startends"]))


def build_reference_architecture(ns, ds, sd):
    from rag_gesture_b200.codec import SyntheticGestureCodec
    ns.dt.GestureRepEncoder = SyntheticGestureCodec
    ns.rg.RetrievalDatabase = make_ref_db_class(ns)
    cfg = C.model_cfg()
    cfg["use_retrieval_for_test"] = True
    arch = ns.builder.build_architecture(cfg, database=ds)
    missing, unexpected = arch.model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("gesture_rep_encoder.") for k in missing)
    return arch.eval()


def gen_pipeline(ns):
    """MotionDiffusion.forward of the reference, B=3: discourse retrieval -> per-exemplar inversion ->
    insertion-guided sampling (decreasing_till_25) -> decode; then a 2-window prev_latent chain."""
    ds = S.SyntheticGestureDataset(N_DB, seed=7)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    arch = build_reference_architecture(ns, ds, S.synthetic_state_dict(0))
    pick = [1, 2, 4]
    batch = S.collate([qs[i] for i in pick])
    batch["retrieval_method"] = "discourse"
    batch["inference_kwargs"] = dict(use_inversion=True, outpaint=False, inversion_start_time=-1,
                                     insertion_guidance=True, guidance_iters=[0] * 25 + list(range(25)),
                                     guidance_lr=0.1)
    torch.manual_seed(2024)
    with torch.no_grad():
        res = arch(**batch)
    n_ex = sum(len(x) for x in res["retrieval_dict"]["retr_startends"])
    print("pipeline exemplars:", n_ex, res["retrieval_dict"]["raw_sample_names"])
    # long-form chaining: window 2 starts from window 1's last tokens
    # (a different clip per window: a repeated sample_name would hit the reference's broken
    #  retrieval-cache branch, raggesture.py:365)
    outs = []
    prev = None
    torch.manual_seed(77)
    for w in range(2):
        bw = S.collate([qs[5 + w]])
        bw["retrieval_method"] = "discourse"
        bw["inference_kwargs"] = dict(use_inversion=True, insertion_guidance=True,
                                      guidance_iters=[0] * 25 + list(range(25)), guidance_lr=0.1,
                                      use_prev_latent=True, prev_latent=prev)
        with torch.no_grad():
            r = arch(**bw)
        prev = r["prev_latentout"]
        outs.append(prev)
    save("pipeline_b3", prev_latentout=res["prev_latentout"], pred_upper=res["pred_upper"][:, ::10],
         pred_hands=res["pred_hands"][:, ::10], n_exemplars=np.array(n_ex),
         chain_w0=outs[0], chain_w1=outs[1])


def gen_pipeline_gesture_type(ns):
    """MotionDiffusion.forward of the reference with retrieval_method="gesture_type", B=2: exemplars chosen by the
    semantic-gesture rules, inverted and inserted with guidance exactly as in the discourse run."""
    ds = S.SyntheticGestureDataset(N_DB, seed=7)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    arch = build_reference_architecture(ns, ds, S.synthetic_state_dict(0))
    batch = S.collate([qs[i] for i in [0, 2]])
    batch["retrieval_method"] = "gesture_type"
    batch["inference_kwargs"] = dict(use_inversion=True, outpaint=False, inversion_start_time=-1,
                                     insertion_guidance=True, guidance_iters=[0] * 25 + list(range(25)),
                                     guidance_lr=0.1)
    torch.manual_seed(515)
    with torch.no_grad():
        res = arch(**batch)
    n_ex = sum(len(x) for x in res["retrieval_dict"]["retr_startends"])
    print("gesture_type pipeline exemplars:", n_ex, res["retrieval_dict"]["raw_sample_names"])
    save("pipeline_gesture_type_b2", prev_latentout=res["prev_latentout"], pred_upper=res["pred_upper"][:, ::10],
         pred_hands=res["pred_hands"][:, ::10], n_exemplars=np.array(n_ex))


def gen_outpaint(ns):
    """The outpaint branch of MotionDiffusion.forward (diffusion_architecture.py:279-283,472): no inversion, no
    guidance; the retrieved exemplar latents placed at their query windows (`raw_motion_latents`) are blended into
    the sample on every step of the plain DDIM loop."""
    ds = S.SyntheticGestureDataset(N_DB, seed=7)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    arch = build_reference_architecture(ns, ds, S.synthetic_state_dict(0))
    batch = S.collate([qs[i] for i in [9, 14]])
    batch["retrieval_method"] = "discourse"
    batch["inference_kwargs"] = dict(outpaint=True, use_inversion=False, insertion_guidance=False)
    torch.manual_seed(4242)
    with torch.no_grad():
        res = arch(**batch)
    seq = res["retrieval_dict"]["raw_motion_latents"]
    print("outpaint: rows with an exemplar latent:", int((seq != 0).any(-1).sum()))
    save("pipeline_outpaint_b2", prev_latentout=res["prev_latentout"], pred_upper=res["pred_upper"][:, ::10],
         n_rows=np.array(int((seq != 0).any(-1).sum())))


TWO_BRANCH = dict(scale_func_cfg=dict(coarse_scale=6.5, both_coef=0.52351, text_coef=-0.28419, retr_coef=2.39872),
                  per_joint_scale=dict(upper=1.2, hands=0.9, face=1.0, lowertransl=1.1))


def gen_two_branch(ns):
    """The 2-branch mode of forward_test (raggesture.py:925-954,1041-1111): text branch + "none" branch (keys - 1e6,
    values of a zeroed condition) mixed with scale_func_retr's coefficients and the per-body-part joint scale.  The
    shipped config sets scale_func_cfg but not per_joint_scale and crashes (AttributeError :1102); with
    per_joint_scale given the reference runs.  random.randint picks the coefficient set above t = 100: Python's
    `random` is seeded before every call / loop."""
    import random
    ns.dt.GestureRepEncoder = ShapeOnlyCodec
    cfg = C.denoiser_cfg()
    cfg.pop("type")
    cfg.update(TWO_BRANCH)
    sd = S.synthetic_state_dict(0)
    model = ns.rg.ReGestureTransformer(**cfg, database=None, use_retrieval_for_test=False)
    model.load_state_dict(sd, strict=False)
    model.eval()
    diff = build_reference_diffusion(ns)
    B, T, D = 2, C.N_TOKENS, C.LATENT_DIM
    cond, x = S.synthetic_conditions(B, seed=11), S.synthetic_latents(B, seed=12)
    kw = model_kwargs_for(model, cond, B)
    outs = {}
    with torch.no_grad():
        for tau in (50, 514, 999):
            random.seed(7 + tau)
            outs[f"x0_t{tau}"] = model(x, torch.full((B,), tau, dtype=torch.int64),
                                       **{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
        cond1 = S.synthetic_conditions(1, seed=21)
        kw1 = model_kwargs_for(model, cond1, 1)
        torch.manual_seed(31)
        random.seed(32)
        outs["plain_final"] = diff.ddim_sample_loop(model, (1, T, D), clip_denoised=False, model_kwargs=kw1, eta=0)
        random.seed(33)
        inv = diff.ddim_reverse_sample_loop(model, start_img=S.synthetic_latents(1, seed=22, scale=0.5),
                                            clip_denoised=False, model_kwargs=kw1, eta=0, return_all_timesteps=True)
        outs["inv49"] = inv[-1]
    save("denoiser_two_branch", **outs)


def gen_rotation(ns):
    """6D cross-fade of tools/longform_synthesis.py:449-471 with the reference's rotation_conversions."""
    import importlib
    rc = importlib.import_module("mogen.models.utils.rotation_conversions")
    g = torch.Generator().manual_seed(3)
    bs, F, J = 2, 15, 7
    prev = 0.8 * torch.randn(bs, F, J * 3, generator=g)
    new = 0.8 * torch.randn(bs, F, J * 3, generator=g)
    to6 = lambda a: rc.matrix_to_rotation_6d(rc.axis_angle_to_matrix(a.reshape(bs, F, J, 3))).reshape(bs, F, J * 6)
    w = torch.linspace(0, 1, F).unsqueeze(0).unsqueeze(-1)
    blended = to6(prev) * (1 - w) + to6(new) * w
    out = rc.matrix_to_axis_angle(rc.rotation_6d_to_matrix(blended.reshape(bs, F, J, 6))).reshape(bs, F, J * 3)
    save("rotation_crossfade", prev=prev, new=new, out=out, six=to6(prev))


def gen_postprocess(ns):
    """tools/visualize.py:204-291 restated with the reference's rotation_conversions: part recomposition by the
    dataset's 0/1 masks and the 15 -> 30 fps interpolation in 6D space."""
    import importlib
    rc = importlib.import_module("mogen.models.utils.rotation_conversions")
    g = torch.Generator().manual_seed(12)
    bs, n, J = 2, 20, 55
    perm = torch.randperm(J, generator=g)
    masks = {}
    for name, cnt, off in (("upper", 13, 0), ("lower", 9, 13), ("hands", 30, 22), ("face", 1, 52)):
        m = np.zeros(J * 3)
        for j in perm[off:off + cnt].tolist():
            m[3 * j:3 * j + 3] = 1
        masks[name] = m
    parts = {k: 0.6 * torch.randn(bs, n, int(m.sum()), generator=g) for k, m in masks.items()}
    facial, trans = torch.randn(bs, n, 100, generator=g), torch.randn(bs, n, 3, generator=g)
    motion = torch.zeros(bs, n, J * 3)
    for k in ("upper", "lower", "hands", "face"):
        motion[..., masks[k].astype(bool)] = parts[k]
    up = lambda x: torch.nn.functional.interpolate(x.permute(0, 2, 1), scale_factor=30 / 15, mode="linear").permute(0, 2, 1)
    six = rc.matrix_to_rotation_6d(rc.axis_angle_to_matrix(motion.reshape(bs, n, J, 3))).reshape(bs, n, J * 6)
    six = up(six)
    aa = rc.matrix_to_axis_angle(rc.rotation_6d_to_matrix(six.reshape(bs, 2 * n, J, 6))).reshape(bs, 2 * n, J * 3)
    save("postprocess", motion=motion, aa30=aa, facial30=up(facial), trans30=up(trans), facial=facial, trans=trans,
         **{f"mask_{k}": v for k, v in masks.items()}, **{f"part_{k}": v for k, v in parts.items()})


VAE_VARIANTS = {
    "a": dict(arch="all_encoder", position_embedding="learned", vae_dist="normal", pre_norm=False, activation="gelu"),
    "b": dict(arch="encoder_decoder", position_embedding="sine", vae_dist="multivariate_normal", pre_norm=True,
              activation="relu", num_layers=2),
}


def codec_inputs(B, seed):
    """Synthetic SMPL-X style inputs of GestureRepEncoder.encode (axis-angle parts, translation, expressions,
    contacts), len150."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *shape, s=1.0: s * torch.randn(*shape, generator=g)
    F = C.MAX_SEQ_LEN
    return dict(motion_upper=r(B, F, 39, s=0.4), motion_lower=r(B, F, 27, s=0.4), motion_face=r(B, F, 3, s=0.2),
                motion_hands=r(B, F, 90, s=0.3), motion_transl=r(B, F, 3, s=0.5), motion_facial=r(B, F, 100, s=0.5),
                motion_contact=(r(B, F, 4) > 0).float(), motion_mask=torch.ones(B, F))


def write_vae_files(root, variant, seed0, shapes_of, latent_dim=64):
    """YAML + checkpoint per body part, laid out as load_vae expects (diffusion_transformer.py:151-167);
    every second checkpoint carries DataParallel's 'module.' prefix."""
    import yaml
    cfg = {"frame_chunk_size": C.FRAME_CHUNK, "latent_dim": latent_dim}
    for i, part in enumerate(("upper", "hands", "face", "lowertrans")):
        args = S.vae_args(part, latent_dim=latent_dim, **VAE_VARIANTS[variant])
        d = os.path.join(root, part)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "cfg.yaml"), "w") as f:
            yaml.safe_dump(args, f)
        sd = S.synthetic_vae_state_dict(shapes_of(args), seed0 + i)
        if i % 2:
            sd = {"module." + k: v for k, v in sd.items()}
        torch.save({"model_state": sd}, os.path.join(d, os.path.basename(args["test_ckpt"])))
        cfg[f"{part}_cfg"] = os.path.join(d, "cfg.yaml")
    return cfg


def gen_codec(ns):
    """TransformerVAE (both decoder architectures / position embeddings / distributions) and GestureRepEncoder
    of the unmodified reference on synthetic weights (rag_gesture_b200.synthetic.synthetic_vae_state_dict)."""
    import importlib
    import json
    import tempfile
    from argparse import Namespace
    gv = importlib.import_module("mogen.models.transformers.gesture_vae")
    out = {}
    for name, var in VAE_VARIANTS.items():
        args = S.vae_args("upper", **var)
        ref = gv.TransformerVAE(Namespace(**args)).eval()
        shapes = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        ref.load_state_dict(S.synthetic_vae_state_dict(shapes, 100))
        x = 0.5 * torch.randn(3, C.MAX_SEQ_LEN, args["nfeats"], generator=torch.Generator().manual_seed(5))
        with torch.no_grad():
            torch.manual_seed(21)
            z, _ = ref.encode_to_dist(x)
            rec = ref.decode(z)
            torch.manual_seed(22)
            lengths = [150, 120, 45]
            z_r, _ = ref.encode_to_dist(x, lengths)
            rec_r = ref.decode(z_r, lengths)
        out.update({f"{name}_keys": json.dumps({k: list(v) for k, v in shapes.items()}), f"{name}_z": z,
                    f"{name}_rec": rec[:, ::5], f"{name}_z_ragged": z_r, f"{name}_rec_ragged": rec_r[:, ::5],
                    f"{name}_in_digest": digest(x)})
    with tempfile.TemporaryDirectory() as root:
        def shapes_of(args):
            return {k: tuple(v.shape) for k, v in gv.TransformerVAE(Namespace(**args)).state_dict().items()}
        cfg = write_vae_files(root, "a", 200, shapes_of)
        enc = ns.dt.GestureRepEncoder(cfg, "time").eval()
        inp = codec_inputs(2, 31)
        with torch.no_grad():
            torch.manual_seed(41)
            motion, mask = enc.encode(**{k: v.clone() for k, v in inp.items()})
            dec = enc.decode(motion)
            torch.manual_seed(42)                                  # exemplars are encoded one by one at B=1
            singles = [enc.encode(**{k: v[e:e + 1].clone() for k, v in inp.items()})[0] for e in range(2)]
        out.update(enc_motion=motion, enc_mask=mask, enc_single=torch.cat(singles, 0), enc_in_digest=digest(*inp.values()))
        for k, v in zip(("upper", "lower", "face", "hands", "transl", "exps", "contact"), dec):
            out[f"dec_{k}"] = v[:, ::10]
    save("codec_vae", **out)


GROUPS = {"rotation": gen_rotation, "schedule": gen_schedule, "denoiser": gen_denoiser, "loops": gen_loops,
          "retrieval": gen_retrieval, "gesture_type": gen_gesture_type, "pipeline": gen_pipeline,
          "pipeline_gesture_type": gen_pipeline_gesture_type, "outpaint": gen_outpaint, "two_branch": gen_two_branch,
          "codec": gen_codec,
          "postprocess": gen_postprocess}

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    ns = refshim.load()
    which = sys.argv[1:] or list(GROUPS)
    for g in which:
        GROUPS[g](ns)
