"""End-to-end parity of the CUDA path (fused engine + samplers, through the reference-shaped API)
against the golden vectors of the unmodified reference and against the oracle.  Needs a GPU.

Tolerances (fp32 tier of BASELINE.json north_star: rel-L2 <= 1e-3 on denoised latents)."""
import pytest
import torch

from conftest import rel_l2
from rag_gesture_b200 import config as C
from rag_gesture_b200 import synthetic as S

pytestmark = pytest.mark.gpu
TOL_STEP = 1e-4      # one denoiser evaluation
TOL_LOOP = 1e-3      # 50-step trajectories (north_star fp32 tier)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model(dev, sd0):
    from rag_gesture_b200 import mogen_api as M
    cfg = C.denoiser_cfg()
    m = M.build_submodule(cfg, database=None, use_retrieval_for_test=False)
    missing, unexpected = m.load_state_dict(sd0, strict=False)
    assert not unexpected and all(k.startswith("gesture_rep_encoder.") for k in missing), (missing, unexpected)
    return m.to(dev).eval()


@pytest.fixture(scope="module")
def diffusion():
    from rag_gesture_b200.diffusion import build_diffusion
    return build_diffusion(C.diffusion_test_cfg())


def _kw(model, cond, B, dev):
    pc = model.get_precompute_condition(device=dev, text=cond["word"], audio=cond["audio"],
                                        speaker_ids=cond["speaker_ids"], re_dict=1)
    return dict(xf_out=pc["xf_out"], re_dict=None, sample_idx=None,
                query_mask={k: v.to(dev) for k, v in S.query_masks(B).items()},
                motion_mask=S.motion_mask(B).to(dev))


def test_state_dict_keys_match_reference(model):
    keys = [k for k in model.state_dict().keys() if not k.startswith("gesture_rep_encoder.")]
    assert keys and set(keys) == set(S.denoiser_param_shapes().keys())
    for k, shp in S.denoiser_param_shapes().items():
        assert tuple(model.state_dict()[k].shape) == tuple(shp), k


def test_condition_encoding(model, golden, dev):
    g = golden("denoiser_step")
    cond = S.synthetic_conditions(2, seed=11)
    kw = _kw(model, cond, 2, dev)
    assert rel_l2(kw["xf_out"]["xf_text"][:, :4].cpu(), torch.from_numpy(g["xf_text"])) < 1e-5


def test_denoiser_step_fused_vs_golden(model, golden, dev):
    g = golden("denoiser_step")
    B = 2
    kw = _kw(model, S.synthetic_conditions(B, seed=11), B, dev)
    x = S.synthetic_latents(B, seed=12).to(dev)
    with torch.no_grad():
        for tau in (14, 514, 999):
            out = model(x, torch.full((B,), tau, device=dev), **kw).cpu()
            assert rel_l2(out, torch.from_numpy(g[f"x0_t{tau}"])) < TOL_STEP, tau
        # off-schedule timestep through the on-the-fly table row, against the oracle
        from oracle import denoiser as OD
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    cond = S.synthetic_conditions(B, seed=11)
    xf = OD.encode_conditions(sd, cond["word"], cond["audio"], cond["speaker_ids"])
    ref = OD.denoiser_forward(sd, x.cpu(), torch.full((B,), 777), S.motion_mask(B), xf, S.query_masks(B))
    with torch.no_grad():
        out = model(x, torch.full((B,), 777, device=dev), **kw).cpu()
    assert rel_l2(out, ref) < TOL_STEP


def test_denoiser_module_path_matches_fused(model, dev):
    """forward_test over the per-module kernels (K/V recomputed per call, arbitrary emb) == fused."""
    from oracle import denoiser as OD
    B = 2
    cond = S.synthetic_conditions(B, seed=31)
    kw = _kw(model, cond, B, dev)
    x = S.synthetic_latents(B, seed=32).to(dev)
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ts = torch.full((B,), 428)
    emb = OD._lin(sd, "time_embed.2", torch.nn.functional.silu(OD._lin(sd, "time_embed.0", OD.timestep_embedding(ts, 512))))
    h = OD.embed_latents(sd, x.cpu())
    with torch.no_grad():
        out_m = model.forward_test(h=h.to(dev), src_mask=S.motion_mask(B).unsqueeze(-1).to(dev), emb=emb.to(dev),
                                   xf_out=kw["xf_out"], query_mask=kw["query_mask"], timesteps=ts.to(dev)).cpu()
        out_f = model(x, ts.to(dev), **kw).cpu()
    assert rel_l2(out_m, out_f) < 2e-5


@pytest.mark.parametrize("tier", ["fp32", "bf16x3", "bf16"])
def test_normal_scale_row_groups(tier, golden, dev):
    """Cross-attention values at ordinary scale (the regime of a trained checkpoint): rows 20/30 get
    y - 1e6 rounded to a 1/16 grid (efficient_attention.py:98), a discontinuous function; every other
    row group must still agree.  Rows 20/30 and the other rows are reported separately per tier."""
    from rag_gesture_b200 import _lib as L
    from rag_gesture_b200 import mogen_api as M
    prec = {"fp32": L.PREC_FP32, "bf16x3": L.PREC_BF16X3, "bf16": L.PREC_BF16}[tier]
    g = golden("denoiser_step_normal_scale")
    sd = S.synthetic_state_dict(1, normal_scale=True)
    m = M.build_submodule(dict(C.denoiser_cfg(), precision=prec), database=None, use_retrieval_for_test=False)
    m.load_state_dict(sd, strict=False)
    m = m.to(dev).eval()
    B = 2
    kw = _kw(m, S.synthetic_conditions(B, seed=11), B, dev)
    x = S.synthetic_latents(B, seed=12).to(dev)
    with torch.no_grad():
        out = m(x, torch.full((B,), 514, device=dev), **kw).cpu()
    ref = torch.from_numpy(g["x0_t514"])
    rows = [r for r in range(C.N_TOKENS) if r not in (20, 30)]
    e_other, e_2030 = rel_l2(out[:, rows], ref[:, rows]), rel_l2(out[:, [20, 30]], ref[:, [20, 30]])
    print("normal-scale (%s tier) rel-L2: other rows %.3g, rows 20/30 %.3g" % (tier, e_other, e_2030))
    # measured on B200 (fp32 tier): 1.0e-2 / 4.8e-2 -- a handful of 1/16-grid roundings flip between any two
    # fp32 implementations (also CPU vs GPU runs of the reference itself); see DESIGN.md.  The bf16 tier rounds
    # the GEMM operands to 2^-9, which moves more elements across a grid line: its bound is wider.
    assert e_other < (1e-1 if tier == "bf16" else 5e-2)
    assert rel_l2(out, ref) < (3e-1 if tier == "bf16" else 2e-1)


def test_plain_ddim_loop_config1(model, diffusion, golden, dev):
    """BASELINE.json configs[0]: B=1, 50-step plain DDIM, same noise as the reference run."""
    g = golden("ddim_plain_b1")
    kw = _kw(model, S.synthetic_conditions(1, seed=21), 1, dev)
    tape = S.NoiseTape(1234)
    diffusion.noise_fn = tape.randn
    traj = [o["sample"].cpu() for o in diffusion.ddim_sample_loop_progressive(
        model, (1, C.N_TOKENS, C.LATENT_DIM), clip_denoised=False, model_kwargs=kw, eta=0)]
    diffusion.noise_fn = None
    assert tape.n == 51
    assert rel_l2(traj[0], torch.from_numpy(g["step49"])) < TOL_STEP
    assert rel_l2(traj[24], torch.from_numpy(g["step25"])) < TOL_LOOP
    assert rel_l2(traj[49], torch.from_numpy(g["final"])) < TOL_LOOP


def _inverted(model, diffusion, dev):
    kw = _kw(model, S.synthetic_conditions(1, seed=21), 1, dev)
    xs = S.synthetic_latents(1, seed=22, scale=0.5).to(dev)
    inv = diffusion.ddim_reverse_sample_loop(model, start_img=xs, clip_denoised=False, model_kwargs=kw,
                                             eta=0, return_all_timesteps=True)
    return torch.cat(inv, 0), kw


def test_reverse_loop(model, diffusion, golden, dev):
    g = golden("ddim_reverse_b1")
    inv, _ = _inverted(model, diffusion, dev)
    inv = inv.cpu()
    assert rel_l2(inv[0], torch.from_numpy(g["inv0"])) < TOL_STEP
    assert rel_l2(inv[24], torch.from_numpy(g["inv24"])) < TOL_LOOP
    assert rel_l2(inv[49], torch.from_numpy(g["inv49"])) < TOL_LOOP


def test_reverse_loop_batched_equals_single(model, diffusion, dev):
    """Batching the B=1 inversions of the reference changes nothing: no op couples clips."""
    cond = S.synthetic_conditions(3, seed=41)
    kw3 = _kw(model, cond, 3, dev)
    xs = S.synthetic_latents(3, seed=42, scale=0.5).to(dev)
    all3 = diffusion.ddim_reverse_sample_loop(model, start_img=xs, clip_denoised=False, model_kwargs=kw3, eta=0)
    one = {k: (v[1:2] if torch.is_tensor(v) else v) for k, v in S.synthetic_conditions(3, seed=41).items()}
    kw1 = _kw(model, one, 1, dev)
    single = diffusion.ddim_reverse_sample_loop(model, start_img=xs[1:2].contiguous(), clip_denoised=False,
                                                model_kwargs=kw1, eta=0)
    assert torch.equal(all3[1:2], single)


def _guided_loop_check(model, diffusion, golden, dev, tol_loop, tier):
    """Guided loop with insertion guidance (gaussian_diffusion.py:1233-1395) and the long-form
    prev-latent blend vs the unmodified reference's golden trajectory, for one precision tier."""
    gg = golden("ddim_guided_b2")
    inv, kw1 = _inverted(model, diffusion, dev)
    B, T, D, n = 2, C.N_TOKENS, C.LATENT_DIM, C.N_CHUNKS
    kw2 = _kw(model, S.synthetic_conditions(B, seed=23), B, dev)
    inv_list = torch.zeros(50, B, T, D, device=dev)
    for b, (q0, q1, r0, r1) in enumerate([(2, 5, 4, 7), (6, 10, 0, 4)]):
        inv_list[:, b, q0:q1] = inv[:, r0:r1]
        inv_list[:, b, n + 1 + q0:n + 1 + q1] = inv[:, n + 1 + r0:n + 1 + r1]
    finals = []
    for skip in (True, False):
        tape = S.NoiseTape(4321)
        start = tape.randn((B, T, D), dev)
        nz = inv_list[49] != 0
        start[nz] = inv_list[49][nz]
        diffusion.noise_fn, diffusion.skip_dead_guidance = tape.randn, skip
        finals.append(diffusion.ddim_guided_sample_loop(
            model, (B, T, D), noise=start, clip_denoised=False, model_kwargs=kw2, eta=0, in_seq=None,
            guidance_iters=[0] * 25 + list(range(25)), inverted_latent_list=inv_list, guidance_lr=0.1).cpu())
    diffusion.noise_fn, diffusion.skip_dead_guidance = None, True
    assert torch.equal(finals[0], finals[1])            # executing the gradient steps changes nothing
    err = rel_l2(finals[0], torch.from_numpy(gg["final"]))
    print(f"guided loop ({tier} tier) rel-L2 vs reference: {err:.3g}")
    assert err < tol_loop

    # long-form mode: prev_latent blended on every step of the plain loop
    prev = torch.zeros(1, T, D)
    prev[:, [0, n + 1, 2 * n + 2, 3 * n + 3]] = S.synthetic_latents(1, seed=24)[:, [9, 20, 31, 42]]
    tape = S.NoiseTape(999)
    diffusion.noise_fn = tape.randn
    fp = diffusion.ddim_sample_loop(model, (1, T, D), clip_denoised=False, model_kwargs=kw1, eta=0,
                                    in_seq=prev.to(dev)).cpu()
    diffusion.noise_fn = None
    err_p = rel_l2(fp, torch.from_numpy(gg["final_prev_latent_b1"]))
    print(f"prev-latent loop ({tier} tier) rel-L2 vs reference: {err_p:.3g}")
    assert err_p < tol_loop


def test_guided_loop_and_dead_guidance(model, diffusion, golden, dev):
    _guided_loop_check(model, diffusion, golden, dev, TOL_LOOP, "fp32")


def test_batch_sizes_and_ragged(model, diffusion, dev):
    """B = 1, 3, 64 give per-clip identical answers (no cross-clip coupling, tile edges)."""
    B = 64
    cond = S.synthetic_conditions(B, seed=51)
    kw = _kw(model, cond, B, dev)
    x = S.synthetic_latents(B, seed=52).to(dev)
    with torch.no_grad():
        full = model(x, torch.full((B,), 600, device=dev), **kw)
        for sl in (slice(0, 1), slice(5, 8), slice(63, 64)):
            sub = {k: v[sl] for k, v in cond.items()}
            n = sl.stop - sl.start
            out = model(x[sl].contiguous(), torch.full((n,), 600, device=dev), **_kw(model, sub, n, dev))
            assert torch.equal(out, full[sl])


# ---- tensor-core precisions (tcgen05): bf16x3 = fp32 tier (<= 1e-3), bf16 = reduced tier (<= 2e-2) ----------
from rag_gesture_b200 import _lib as _L  # noqa: E402

TC_TIERS = [(_L.PREC_BF16X3, 1e-4, 1e-3), (_L.PREC_BF16, 1e-2, 2e-2)]


@pytest.fixture(scope="module", params=TC_TIERS, ids=["bf16x3", "bf16"])
def tc_model(request, dev, sd0):
    from rag_gesture_b200 import mogen_api as M
    prec, tol_step, tol_loop = request.param
    m = M.build_submodule(dict(C.denoiser_cfg(), precision=prec), database=None, use_retrieval_for_test=False)
    m.load_state_dict(sd0, strict=False)
    return m.to(dev).eval(), tol_step, tol_loop


def test_tc_denoiser_step_vs_golden(tc_model, golden, dev):
    model, tol_step, _ = tc_model
    g = golden("denoiser_step")
    B = 2
    kw = _kw(model, S.synthetic_conditions(B, seed=11), B, dev)
    x = S.synthetic_latents(B, seed=12).to(dev)
    with torch.no_grad():
        for tau in (14, 514, 999):
            out = model(x, torch.full((B,), tau, device=dev), **kw).cpu()
            err = rel_l2(out, torch.from_numpy(g[f"x0_t{tau}"]))
            print(f"precision {model.precision} tau {tau}: rel-L2 {err:.3g}")
            assert err < tol_step, tau


def test_tc_loops_vs_golden(tc_model, diffusion, golden, dev):
    model, _, tol_loop = tc_model
    g = golden("ddim_plain_b1")
    kw = _kw(model, S.synthetic_conditions(1, seed=21), 1, dev)
    tape = S.NoiseTape(1234)
    diffusion.noise_fn = tape.randn
    final = diffusion.ddim_sample_loop(model, (1, C.N_TOKENS, C.LATENT_DIM), clip_denoised=False,
                                       model_kwargs=kw, eta=0).cpu()
    diffusion.noise_fn = None
    err = rel_l2(final, torch.from_numpy(g["final"]))
    gr = golden("ddim_reverse_b1")
    xs = S.synthetic_latents(1, seed=22, scale=0.5).to(dev)
    inv = diffusion.ddim_reverse_sample_loop(model, start_img=xs, clip_denoised=False, model_kwargs=kw, eta=0)
    err_r = rel_l2(inv.cpu(), torch.from_numpy(gr["inv49"]))
    print(f"precision {model.precision}: plain loop rel-L2 {err:.3g}, reverse loop {err_r:.3g}")
    assert err < tol_loop and err_r < tol_loop


def test_tc_guided_loop_vs_golden(tc_model, diffusion, golden, dev):
    """The guided loop (insertion guidance, dead gradient steps) and the prev-latent loop on the tensor-core
    tiers against the reference's golden trajectory: bf16x3 <= 1e-3, bf16 <= 2e-2 (north_star)."""
    model, _, tol_loop = tc_model
    _guided_loop_check(model, diffusion, golden, dev, tol_loop, {2: "bf16x3", 1: "bf16"}.get(model.precision, "tc"))


def test_tc_bench_shape_groups_vs_oracle(tc_model, diffusion, dev):
    """The benched shape: B = 64 clips at guided level i together with E = 96 exemplars at inversion level j
    through rg_denoise_groups (one kernel chain, M = 6880 rows: the persistent 2-CTA GEMM path), a sample of
    clips from both groups against the CPU oracle (reference algorithm as written, fp32)."""
    from oracle import denoiser as OD
    model, tol_step, _ = tc_model
    eng = model.rg_engine(diffusion)
    B, E = 64, 96
    cond = S.synthetic_conditions(B + E, seed=101)
    kw = _kw(model, cond, B + E, dev)
    prep = model.prepare_batch(kw, B + E)
    x = S.synthetic_latents(B + E, seed=102).to(dev)
    li, lj = 37, 12
    out = eng.denoise_groups(x, prep.src_mask, prep.query_mask, prep.state, [(B, li), (E, lj)]).cpu()
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    tmap = diffusion.timestep_map
    for b in (0, 63, 64, 131, 159):
        sub = {k: v[b:b + 1] for k, v in cond.items()}
        xf = OD.encode_conditions(sd, sub["word"], sub["audio"], sub["speaker_ids"])
        tau = int(tmap[li if b < B else lj])
        ref = OD.denoiser_forward(sd, x[b:b + 1].cpu(), torch.full((1,), tau), S.motion_mask(1), xf, S.query_masks(1))
        err = rel_l2(out[b:b + 1], ref)
        print(f"precision {model.precision} bench shape clip {b} (tau {tau}): rel-L2 {err:.3g}")
        assert err < tol_step, (b, err)


def test_tc_batch_independence(tc_model, dev):
    """Ragged M (TMA zero-fill, row guards): a clip's answer does not depend on its batch."""
    model, _, _ = tc_model
    B = 7
    cond = S.synthetic_conditions(B, seed=61)
    kw = _kw(model, cond, B, dev)
    x = S.synthetic_latents(B, seed=62).to(dev)
    with torch.no_grad():
        full = model(x, torch.full((B,), 300, device=dev), **kw)
        sub = {k: v[3:4] for k, v in cond.items()}
        one = model(x[3:4].contiguous(), torch.full((1,), 300, device=dev), **_kw(model, sub, 1, dev))
    assert torch.equal(one, full[3:4])


def test_state_cache_is_not_reused_across_batches(model, dev):
    """Regression: the K6 state cache was keyed on data_ptr/_version, so a new batch whose condition
    tensors landed on recycled storage silently reused the previous batch's cross-attention state."""
    B = 2
    x = S.synthetic_latents(B, seed=82).to(dev)
    outs = []
    for seed in (81, 83, 81):
        kw = _kw(model, S.synthetic_conditions(B, seed=seed), B, dev)
        with torch.no_grad():
            outs.append(model(x, torch.full((B,), 300, device=dev), **kw).clone())
        del kw                                   # storage goes back to the caching allocator
    model._state_cache = (None, None)
    kw = _kw(model, S.synthetic_conditions(B, seed=83), B, dev)
    with torch.no_grad():
        fresh = model(x, torch.full((B,), 300, device=dev), **kw)
    assert torch.equal(outs[1], fresh) and torch.equal(outs[0], outs[2]) and not torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("which", ["fp32", "tc"])
def test_lanes_do_not_change_results(which, model, tc_model, diffusion, dev):
    """rg_set_lanes: cutting the batch into 1..4 concurrent clip-range chains is bit-identical
    (ragged split 7 = 2+2+3 / 1+2+2+2 included), also across consecutive steps reusing the workspaces."""
    mdl = model if which == "fp32" else tc_model[0]
    eng = mdl.rg_engine(diffusion)
    B = 7
    kw = _kw(mdl, S.synthetic_conditions(B, seed=91), B, dev)
    x = S.synthetic_latents(B, seed=92).to(dev)
    outs = []
    try:
        for lanes in (1, 2, 3, 4, 0):
            eng.set_lanes(lanes)
            with torch.no_grad():
                y = mdl(x, torch.full((B,), 300, device=dev), **kw)
                y = mdl(y, torch.full((B,), 280, device=dev), **kw)
            outs.append(y.clone())
    finally:
        eng.set_lanes(0)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    with pytest.raises(RuntimeError):
        eng.set_lanes(9)


@pytest.mark.parametrize("which", ["fp32", "tc"])
def test_denoise_groups_equals_per_level_calls(which, model, tc_model, diffusion, dev):
    """rg_denoise_groups: a batch whose clip ranges sit at DIFFERENT schedule levels gives every clip exactly
    what a single-level rg_denoise gives it (ragged groups 3 + 1 + 4, levels 49 / 0 / 17), for the fp32 and the
    tensor-core tiers; bad group tables are refused."""
    mdl = model if which == "fp32" else tc_model[0]
    eng = mdl.rg_engine(diffusion)
    B, groups = 8, [(3, 49), (1, 0), (4, 17)]
    kw = _kw(mdl, S.synthetic_conditions(B, seed=61), B, dev)
    prep = mdl.prepare_batch(kw, B)
    x = S.synthetic_latents(B, seed=62).to(dev)
    joint = eng.denoise_groups(x, prep.src_mask, prep.query_mask, prep.state, groups)
    b0 = 0
    for n, step in groups:
        sl = slice(b0, b0 + n)
        qm = prep.query_mask[:, sl].contiguous()
        ref = eng.denoise(x[sl].contiguous(), prep.src_mask[sl].contiguous(), qm, prep.state[sl].contiguous(), step_idx=step)
        assert torch.equal(joint[sl], ref), (n, step)
        b0 += n
    with pytest.raises(RuntimeError):
        eng.denoise_groups(x, prep.src_mask, prep.query_mask, prep.state, [(4, 1), (4, 50)])    # level outside the schedule


def test_fused_guided_and_reverse_loops_bit_identical(model, diffusion, dev):
    """ddim_guided_and_reverse_loops (guided sampling of one batch + inversion of another batch's exemplars, one
    kernel chain per level) == ddim_guided_sample_loop and ddim_reverse_sample_loop run separately, bit for bit,
    with the same noise tape."""
    B, E, T, D, n = 2, 3, C.N_TOKENS, C.LATENT_DIM, C.N_CHUNKS
    kwg = _kw(model, S.synthetic_conditions(B, seed=71), B, dev)
    kwr = _kw(model, S.synthetic_conditions(E, seed=72), E, dev)
    start_img = S.synthetic_latents(E, seed=73, scale=0.5).to(dev)
    inv_list = torch.zeros(50, B, T, D, device=dev)
    inv_list[:, 0, 2:5] = S.synthetic_latents(50, seed=74)[:, 2:5].to(dev)
    inv_list[:, 1, n + 3:n + 8] = S.synthetic_latents(50, seed=75)[:, 3:8].to(dev)
    iters = [0] * 25 + list(range(25))

    def start_noise(tape):
        s = tape.randn((B, T, D), dev)
        nz = inv_list[49] != 0
        s[nz] = inv_list[49][nz]
        return s

    tape = S.NoiseTape(55)
    diffusion.noise_fn = tape.randn
    ref_g = diffusion.ddim_guided_sample_loop(model, (B, T, D), noise=start_noise(tape), clip_denoised=False,
                                              model_kwargs=kwg, eta=0, in_seq=None, guidance_iters=iters,
                                              inverted_latent_list=inv_list, guidance_lr=0.1)
    ref_r = diffusion.ddim_reverse_sample_loop(model, start_img=start_img, clip_denoised=False, model_kwargs=kwr,
                                               eta=0, return_all_timesteps=True)
    tape = S.NoiseTape(55)
    diffusion.noise_fn = tape.randn
    out_g, out_r = diffusion.ddim_guided_and_reverse_loops(
        model, guided=dict(shape=(B, T, D), noise=start_noise(tape), model_kwargs=kwg, in_seq=None,
                           guidance_iters=iters, inverted_latent_list=inv_list, guidance_lr=0.1),
        reverse=dict(start_img=start_img, model_kwargs=kwr))
    diffusion.noise_fn = None
    assert torch.equal(out_g, ref_g)
    assert len(out_r) == len(ref_r) == 50
    for a, b in zip(out_r, ref_r):
        assert torch.equal(a, b)


def test_run_levels_equals_step_loops(model, diffusion, dev):
    """SpacedDiffusion.run_levels (all 50 levels inside one rg_run_levels call) == the step-by-step loops with the
    reference's names, bit for bit and with the same noise-tape consumption: inversion only, plain sampling
    with and without an in_seq blended at every level, guided sampling with the dead gradient steps executed."""
    B, E, T, D, n = 2, 2, C.N_TOKENS, C.LATENT_DIM, C.N_CHUNKS
    kwg = _kw(model, S.synthetic_conditions(B, seed=81), B, dev)
    kwr = _kw(model, S.synthetic_conditions(E, seed=82), E, dev)
    start_img = S.synthetic_latents(E, seed=83, scale=0.5).to(dev)
    ref_r = diffusion.ddim_reverse_sample_loop(model, start_img=start_img, clip_denoised=False, model_kwargs=kwr,
                                               eta=0, return_all_timesteps=True)
    _, out_r = diffusion.run_levels(model, reverse=dict(start_img=start_img, model_kwargs=kwr))
    assert all(torch.equal(a, b) for a, b in zip(out_r, ref_r)) and len(out_r) == 50

    prev = torch.zeros(B, T, D)
    prev[:, [0, n + 1, 2 * n + 2, 3 * n + 3]] = S.synthetic_latents(B, seed=84)[:, [9, 20, 31, 42]]
    for in_seq in (None, prev.to(dev)):
        tape = S.NoiseTape(7)
        diffusion.noise_fn = tape.randn
        ref = diffusion.ddim_sample_loop(model, (B, T, D), clip_denoised=False, model_kwargs=kwg, eta=0, in_seq=in_seq)
        tail_ref = tape.randn((4,), dev)
        tape = S.NoiseTape(7)
        diffusion.noise_fn = tape.randn
        out, _ = diffusion.run_levels(model, guided=dict(shape=(B, T, D), noise=None, model_kwargs=kwg, in_seq=in_seq))
        assert torch.equal(out, ref)
        assert torch.equal(tape.randn((4,), dev), tail_ref)         # the tape was consumed identically

    inv_list = torch.zeros(50, B, T, D, device=dev)
    inv_list[:, 0, 2:5] = S.synthetic_latents(50, seed=85)[:, 2:5].to(dev)
    iters = [0] * 25 + list(range(25))
    try:
        diffusion.skip_dead_guidance = False
        tape = S.NoiseTape(8)
        diffusion.noise_fn = tape.randn
        ref = diffusion.ddim_guided_sample_loop(model, (B, T, D), noise=None, clip_denoised=False, model_kwargs=kwg,
                                                eta=0, in_seq=prev.to(dev), guidance_iters=iters,
                                                inverted_latent_list=inv_list, guidance_lr=0.1)
        tape = S.NoiseTape(8)
        diffusion.noise_fn = tape.randn
        out, _ = diffusion.run_levels(model, guided=dict(shape=(B, T, D), noise=None, model_kwargs=kwg,
                                                         in_seq=prev.to(dev), guidance_iters=iters,
                                                         inverted_latent_list=inv_list, guidance_lr=0.1))
        assert torch.equal(out, ref)
    finally:
        diffusion.noise_fn, diffusion.skip_dead_guidance = None, True


@pytest.mark.parametrize("which", ["fp32", "tc"])
def test_graph_replay_is_bit_identical(which, model, tc_model, diffusion, dev):
    """rg_set_graphs: the evaluation chain replayed from a CUDA graph (second and later calls on the same
    buffers) == direct launches, for single-level and grouped evaluations, across levels (the level's table
    row is re-staged outside the graph), after the workspace grew (graphs are dropped), and over a whole
    run_levels loop."""
    mdl = model if which == "fp32" else tc_model[0]
    eng = mdl.rg_engine(diffusion)
    B = 6
    kw = _kw(mdl, S.synthetic_conditions(B, seed=111), B, dev)
    prep = mdl.prepare_batch(kw, B)
    x = S.synthetic_latents(B, seed=112).to(dev)
    out = torch.empty_like(x)
    try:
        eng.set_graphs(False)
        ref = {lvl: eng.denoise(x, prep.src_mask, prep.query_mask, prep.state, step_idx=lvl).clone() for lvl in (3, 30, 49)}
        ref_g = eng.denoise_groups(x, prep.src_mask, prep.query_mask, prep.state, [(2, 5), (4, 44)]).clone()
        eng.set_graphs(True)
        for rep in range(3):                        # direct, capture + replay, replay
            for lvl in (3, 30, 49):
                eng.denoise(x, prep.src_mask, prep.query_mask, prep.state, step_idx=lvl, out=out)
                assert torch.equal(out, ref[lvl]), (rep, lvl)
            eng.denoise_groups(x, prep.src_mask, prep.query_mask, prep.state, [(2, 5), (4, 44)], out=out)
            assert torch.equal(out, ref_g), rep
        # two lanes (concurrent clip-range chains, forked and joined inside the captured graph; ragged split 6 = 3 + 3
        # for the level call, and a group boundary inside a lane for the grouped call)
        eng.set_lanes(2)
        for rep in range(3):
            eng.denoise(x, prep.src_mask, prep.query_mask, prep.state, step_idx=30, out=out)
            assert torch.equal(out, ref[30]), ("lanes", rep)
            eng.denoise_groups(x, prep.src_mask, prep.query_mask, prep.state, [(2, 5), (4, 44)], out=out)
            assert torch.equal(out, ref_g), ("lanes", rep)
        eng.set_lanes(0)
        # a larger batch grows the workspace: captured graphs must not be replayed on the moved buffers
        B2 = 40
        kw2 = _kw(mdl, S.synthetic_conditions(B2, seed=113), B2, dev)
        prep2 = mdl.prepare_batch(kw2, B2)
        x2 = S.synthetic_latents(B2, seed=114).to(dev)
        big = [eng.denoise(x2, prep2.src_mask, prep2.query_mask, prep2.state, step_idx=7).clone() for _ in range(3)]
        assert torch.equal(big[0], big[1]) and torch.equal(big[0], big[2])
        for rep in range(2):
            eng.denoise(x, prep.src_mask, prep.query_mask, prep.state, step_idx=30, out=out)
            assert torch.equal(out, ref[30])
        # whole loops: 50 levels on the same buffers (level 0 direct, level 1 captures, 48 replays)
        start = S.synthetic_latents(B, seed=115, scale=0.5).to(dev)
        kwr = dict(kw)
        _, inv_g = diffusion.run_levels(mdl, reverse=dict(start_img=start, model_kwargs=kwr))
        eng.set_graphs(False)
        _, inv_d = diffusion.run_levels(mdl, reverse=dict(start_img=start, model_kwargs=kwr))
        assert all(torch.equal(a, b) for a, b in zip(inv_g, inv_d))
    finally:
        eng.set_graphs(True)
        eng.set_lanes(0)


def test_engine_follows_weight_changes(dev, sd0):
    """the packed device engine is rebuilt when ANY parameter changes -- an in-place edit of a middle
    layer, and a checkpoint loaded through a PARENT module (whose load_state_dict never calls the child's
    override) -- not only when the first / last tensor does."""
    import rag_gesture_b200 as R
    arch = R.build_architecture(C.model_cfg(), database=None)
    arch.model.load_state_dict(sd0, strict=False)
    arch = arch.to(dev).eval()
    m = arch.model
    B = 2
    kw = _kw(m, S.synthetic_conditions(B, seed=121), B, dev)
    x = S.synthetic_latents(B, seed=122).to(dev)
    t = torch.full((B,), 300, device=dev)
    with torch.no_grad():
        base = m(x, t, **kw).clone()
        m.temporal_decoder_blocks[3].ffn.linear1.weight.mul_(1.5)          # in-place edit of a middle layer
        edited = m(x, t, **kw).clone()
    assert not torch.equal(base, edited)
    sd1 = S.synthetic_state_dict(5)
    arch.load_state_dict({"model." + k: v for k, v in sd1.items()}, strict=False)   # through the parent
    with torch.no_grad():
        m._state_cache = (None, None)
        kw = _kw(m, S.synthetic_conditions(B, seed=121), B, dev)
        other = m(x, t, **kw).clone()
    ref = R.build_submodule(C.denoiser_cfg(), database=None, use_retrieval_for_test=False)
    ref.load_state_dict(sd1, strict=False)
    ref = ref.to(dev).eval()
    with torch.no_grad():
        want = ref(x, t, **_kw(ref, S.synthetic_conditions(B, seed=121), B, dev))
    assert torch.equal(other, want)


def test_single_speaker_config(dev, sd0):
    """num_speakers == 1 (diffusion_transformer.py:545-546): the speaker condition is zeros [B, B, latent] -- the
    reference's shape, quirk included -- and the evaluation matches the oracle run on exactly that tensor."""
    from oracle import denoiser as OD
    from rag_gesture_b200 import mogen_api as M
    cfg = dict(C.denoiser_cfg(), speaker_embedding=dict(num_speakers=1))
    sd = dict(sd0)
    sd["speaker_embedding.weight"] = sd0["speaker_embedding.weight"][:1].clone()
    m = M.build_submodule(cfg, database=None, use_retrieval_for_test=False)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    m = m.to(dev).eval()
    B = 3
    cond = S.synthetic_conditions(B, seed=131)
    kw = _kw(m, cond, B, dev)
    assert tuple(kw["xf_out"]["xf_spk"].shape) == (B, B, C.LATENT_DIM) and not bool(kw["xf_out"]["xf_spk"].any())
    x = S.synthetic_latents(B, seed=132).to(dev)
    with torch.no_grad():
        out = m(x, torch.full((B,), 514, device=dev), **kw).cpu()
    xf = OD.encode_conditions(sd0, cond["word"], cond["audio"], cond["speaker_ids"])
    xf["xf_spk"] = torch.zeros(B, B, C.LATENT_DIM)
    ref = OD.denoiser_forward(sd, x.cpu(), torch.full((B,), 514), S.motion_mask(B), xf, S.query_masks(B))
    assert rel_l2(out, ref) < TOL_STEP


TWO_BRANCH = dict(scale_func_cfg=dict(coarse_scale=6.5, both_coef=0.52351, text_coef=-0.28419, retr_coef=2.39872),
                  per_joint_scale=dict(upper=1.2, hands=0.9, face=1.0, lowertransl=1.1))


@pytest.mark.parametrize("tier", ["fp32", "bf16x3", "bf16"])
def test_two_branch_mode_vs_reference(tier, golden, dev, sd0, diffusion):
    """forward_test with scale_func_cfg (raggesture.py:925-954,1041-1111) -- the batch evaluated as text
    branch + "none" branch in one 2B-clip chain and mixed by rg_mix_branches -- against the unmodified reference run
    with per_joint_scale set (make_golden.py two_branch): single evaluations below and above t = 100 (the random
    coefficient set, Python's `random` seeded like the reference run), a plain 50-step loop and an inversion loop."""
    import random
    from rag_gesture_b200 import _lib as L
    from rag_gesture_b200 import mogen_api as M
    # bf16: the mix extrapolates between the branches (w * text + (1 - w) * none with w up to 7.5), which amplifies the
    # operand rounding of a single evaluation ~5x relative to the single-branch mode
    prec, tol_step, tol_loop = {"fp32": (L.PREC_FP32, 1e-4, 1e-3), "bf16x3": (L.PREC_BF16X3, 2e-4, 1e-3),
                                "bf16": (L.PREC_BF16, 2e-2, 2e-2)}[tier]
    g = golden("denoiser_two_branch")
    m = M.build_submodule(dict(C.denoiser_cfg(), precision=prec, **TWO_BRANCH), database=None, use_retrieval_for_test=False)
    m.load_state_dict(sd0, strict=False)
    m = m.to(dev).eval()
    B = 2
    kw = _kw(m, S.synthetic_conditions(B, seed=11), B, dev)
    x = S.synthetic_latents(B, seed=12).to(dev)
    with torch.no_grad():
        for tau in (50, 514, 999):
            random.seed(7 + tau)
            out = m(x, torch.full((B,), tau, device=dev), **kw).cpu()
            err = rel_l2(out, torch.from_numpy(g[f"x0_t{tau}"]))
            print(f"two-branch ({tier} tier) tau {tau}: rel-L2 {err:.3g}")
            assert err < tol_step, tau
    kw1 = _kw(m, S.synthetic_conditions(1, seed=21), 1, dev)
    diffusion.noise_fn = lambda shape, device: torch.randn(*shape).to(device)      # global CPU generator, like the reference
    try:
        torch.manual_seed(31)
        random.seed(32)
        final = diffusion.ddim_sample_loop(m, (1, C.N_TOKENS, C.LATENT_DIM), clip_denoised=False, model_kwargs=kw1, eta=0).cpu()
        random.seed(33)
        inv = diffusion.ddim_reverse_sample_loop(m, start_img=S.synthetic_latents(1, seed=22, scale=0.5).to(dev),
                                                 clip_denoised=False, model_kwargs=kw1, eta=0, return_all_timesteps=True)
        # the whole-loop entry point falls back to the step-by-step loops in this mode: same draws, same result
        random.seed(33)
        _, inv2 = diffusion.run_levels(m, reverse=dict(start_img=S.synthetic_latents(1, seed=22, scale=0.5).to(dev), model_kwargs=kw1))
    finally:
        diffusion.noise_fn = None
    e1, e2 = rel_l2(final, torch.from_numpy(g["plain_final"])), rel_l2(inv[-1].cpu(), torch.from_numpy(g["inv49"]))
    print(f"two-branch ({tier} tier): plain loop rel-L2 {e1:.3g}, reverse loop {e2:.3g}")
    assert e1 < tol_loop and e2 < tol_loop
    assert torch.equal(inv2[-1], inv[-1])
    # scale_func_cfg without per_joint_scale: the reference's AttributeError (raggesture.py:1102)
    bad = M.build_submodule(dict(C.denoiser_cfg(), scale_func_cfg=TWO_BRANCH["scale_func_cfg"]), database=None,
                            use_retrieval_for_test=False)
    bad.load_state_dict(sd0, strict=False)
    bad = bad.to(dev).eval()
    with pytest.raises(AttributeError), torch.no_grad():
        bad(x, torch.full((B,), 50, device=dev), **_kw(bad, S.synthetic_conditions(B, seed=11), B, dev))
