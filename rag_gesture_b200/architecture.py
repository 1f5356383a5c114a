"""MotionDiffusion: the orchestration of one guided batch, same constructor / forward(**batch) /
result keys as mogen/models/architectures/diffusion_architecture.py:64-582 (inference branch).

What changes is HOW the batch is executed on a B200:
  * all exemplars of all clips are DDIM-inverted in ONE batched reverse loop (the reference runs one
    50-step loop per exemplar at B=1 inside a double Python loop, :323-354); no denoiser op couples
    clips, so the latents are the same (tests/test_gpu_denoiser.py::test_reverse_loop_batched_equals_single);
  * cross-attention K/V state and the timestep table are computed once (engine K6/K7);
  * the guided loop skips the provably dead gradient steps.
The global-RNG draw order of the reference is preserved: GT encode, exemplar encodes
(clip-major), start_noise, then per step randn_like(in_seq), randn_like(x).
Training branch, DDPM sampling and the `visualize_inversion` debug outputs are out of scope.
"""
import copy

import torch
import torch.nn as nn

from . import config as CFG
from .diffusion import build_diffusion
from .mogen_api import ARCHITECTURES, LOSSES, build_loss, build_submodule


@LOSSES.register_module()
class MSELoss(nn.Module):
    """Constructor-compatible placeholder (mogen/models/losses/mse_loss.py:33-70): MotionDiffusion
    builds `loss_recon` even for inference (diffusion_architecture.py:86); training is out of scope."""

    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert reduction in (None, "none", "mean", "sum")
        self.reduction, self.loss_weight = reduction or "none", loss_weight

    def forward(self, *a, **k):
        raise NotImplementedError("training losses are out of scope of rg_b200")


@ARCHITECTURES.register_module()
class MotionDiffusion(nn.Module):
    def __init__(self, model=None, loss_recon=None, loss_gen=None, loss_contact=None, loss_laplace=None,
                 diffusion_train=None, diffusion_test=None, init_cfg=None, inference_type="ddpm",
                 genloss_acceleration_weight=True, genloss_hands_weight=2, genloss_smooth=True,
                 body_part_lossweights=None, **kwargs):
        super().__init__()
        self.loss_recon = build_loss(loss_recon)
        if loss_contact is not None or loss_gen is not None or loss_laplace is not None:
            raise NotImplementedError("auxiliary training losses are out of scope of rg_b200")
        self.model = build_submodule(copy.deepcopy(model), **kwargs)
        self.diffusion_train = build_diffusion(diffusion_train) if diffusion_train is not None else None
        self.diffusion_test = build_diffusion(diffusion_test)
        self.inference_type = inference_type
        self.body_part_lossweights = body_part_lossweights

    @staticmethod
    def _row_sets(T):
        n = (T - 3) // 4
        return (list(range(0, n)), list(range(n + 1, 2 * n + 1)), list(range(2 * n + 2, 3 * n + 2)),
                list(range(3 * n + 3, T)))

    def _scatter(self, kwargs):
        """Host -> device move of the batch (what MMDataParallel.scatter does for the reference,
        tools/visualize.py:146): tensors and lists of tensors, asynchronously from pinned memory."""
        dev = self.model.out.weight.device

        def mv(v):
            if torch.is_tensor(v):
                return v.to(dev, non_blocking=True)
            if isinstance(v, list) and v and torch.is_tensor(v[0]):
                return [t.to(dev, non_blocking=True) for t in v]
            return v
        return {k: mv(v) for k, v in kwargs.items()}

    def forward(self, **kwargs):
        """batch dict (beatx_collate_fn schema, host or device tensors) -> the same dict plus
        retrieval_dict, prev_latentout [B,43,512] and the decoded pred_* poses (:479,572-577)."""
        gb = self.prepare(**kwargs)
        output = self.run_prepared(gb)
        return self.finish(gb, output)

    # ---- stage 1: host logic + codec + retrieval (everything that is not the denoising loops) ----
    def prepare(self, defer_conditions=False, **kwargs):
        """defer_conditions=True leaves out the one engine call of this stage (the text/audio/speaker
        pre-projection of the B clips); `encode_clip_conditions(gb)` adds it later.  GuidedPipeline uses
        this to run the stage on a worker thread without touching the denoiser handle."""
        if self.training:
            raise NotImplementedError("rg_b200 is inference-only; call model.eval()")
        kwargs = self._scatter(kwargs)
        codec = self.model.gesture_rep_encoder
        enc_args = (kwargs["motion_upper"], kwargs["motion_lower"], kwargs["motion_face"], kwargs["motion_hands"],
                    kwargs["trans"], kwargs["facial"], kwargs["contact"], kwargs["motion_mask"].float())
        device = enc_args[0].device
        # The reference encodes the clips first and retrieves afterwards (diffusion_architecture.py:230-262).  A codec
        # that hands out its Gaussian draws (vae.GestureRepEncoder.draw_encode_eps) gets them taken HERE, at that place
        # in the random stream, and runs its passes AFTER the retrieval stage: the stage's host-side waits (speaker
        # ids, text-similarity ranks) then queue behind the input copies only, not behind the clips' encode pass on a
        # GPU that is busy with the previous batch's loops.  Same values, different order on the stream.
        late_eps = None
        if hasattr(codec, "draw_encode_eps") and device.type == "cuda":
            late_eps = codec.draw_encode_eps(enc_args[0])
        else:
            with torch.no_grad():
                motion, motion_mask = codec.encode(*enc_args)

        ik = kwargs.get("inference_kwargs", {})
        gb = GuidedBatch()
        gb.use_outpaint = ik.pop("outpaint", False)
        gb.use_inversion = ik.pop("use_inversion", False)
        gb.inversion_start_time = ik.pop("inversion_start_time", -1)
        if ik.pop("visualize_inversion", False):
            raise NotImplementedError("visualize_inversion debug outputs are out of scope of rg_b200")
        gb.use_guidance = ik.pop("insertion_guidance", False)
        gb.guidance_iters = ik.pop("guidance_iters", [10] * 50)
        gb.guidance_lr = ik.pop("guidance_lr", 0.1)
        gb.use_prev = ik.pop("use_prev_latent", False)
        gb.prev_latent = ik.pop("prev_latent", None)
        gb.extra = ik
        if gb.use_prev:
            assert not gb.use_outpaint
        if gb.use_outpaint:
            assert not gb.use_inversion and not gb.use_guidance
        if gb.use_guidance:
            assert not gb.use_outpaint and gb.use_inversion
        if self.inference_type != "ddim":
            raise NotImplementedError("rg_b200 implements inference_type='ddim' only")

        kwargs.update({"text": kwargs["word"], "raw_text": kwargs["raw_word"], "text_times": kwargs["text_segments"]})
        with torch.no_grad():
            if defer_conditions:
                model_kwargs = self.model.get_precompute_condition(device=device, xf_out={}, **kwargs)
                gb.cond_inputs = (kwargs["text"], kwargs["audio"], kwargs["speaker_ids"])
            else:
                model_kwargs = self.model.get_precompute_condition(device=device, **kwargs)
            if late_eps is not None:
                motion, motion_mask = codec.encode(*enc_args, eps=late_eps)
        kwargs["motion_mask"] = motion_mask
        B, T = motion.shape[:2]
        # cross-attention is skipped on rows {(T-3)//4 * k}: NOT the separator rows
        qmask = torch.ones_like(motion_mask)
        for col in ((T - 3) // 4, 2 * (T - 3) // 4, 3 * (T - 3) // 4):     # one column at a time: a list index would be
            qmask[:, col] = 0                                              # uploaded with a blocking pageable copy
        query_masks = {c: qmask for c in CFG.CONDS}
        model_kwargs["query_mask"] = query_masks
        model_kwargs["motion_mask"] = motion_mask
        model_kwargs["sample_idx"] = kwargs.get("sample_idx", None)
        retrieval_dict = model_kwargs["re_dict"]
        kwargs["retrieval_dict"] = retrieval_dict     # the reference deep-copies (:275); nothing mutates it here
        gb.results, gb.model_kwargs, gb.shape, gb.device = kwargs, model_kwargs, (B, T, codec.vae_latent_dim), device

        if gb.use_outpaint:
            seq = retrieval_dict["raw_motion_latents"]
            assert seq.shape[1] == 1
            gb.outpaint_seq = seq.squeeze(1)
        if gb.use_prev and gb.prev_latent is not None:
            gb.prev_latent = self.mask_prev_latent(gb.prev_latent.to(device))

        # all exemplars of all clips, in the reference's visiting order (clip-major, dict order)
        if gb.use_inversion:
            lat = retrieval_dict["retr_uncropped_latents"]
            gb.jobs = [(b, q) for b, lats in enumerate(lat) for q in lats.keys()]
            if gb.jobs:
                cat = lambda key: torch.cat([lat[b][q][key] for b, q in gb.jobs], 0).to(device)
                gb.ex = {k: cat(k) for k in ("retr_text", "retr_audio", "retr_spkid", "retr_motion_mask",
                                             "retr_motion_latent")}
                clip_of = torch.tensor([b for b, _ in gb.jobs])
                if device.type == "cuda":              # pinned + asynchronous: a pageable upload would block the host
                    clip_of = clip_of.pin_memory()     # behind the encode passes enqueued above
                clip_of = clip_of.to(device, non_blocking=True)
                gb.ex_query_mask = {c: m[clip_of] for c, m in query_masks.items()}
                gb.windows = [(retrieval_dict["retr_startends"][b][q], retrieval_dict["query_startends"][b][q])
                              for b, q in gb.jobs]
        return gb

    def encode_clip_conditions(self, gb):
        """The deferred part of prepare(defer_conditions=True): xf_text / xf_audio / xf_spk of the B clips."""
        if gb.cond_inputs is not None:
            text, audio, spk = gb.cond_inputs
            with torch.no_grad():
                gb.model_kwargs["xf_out"] = self.model.encode_all_conditions(text, audio, spk, gb.device)
            gb.cond_inputs = None
        return gb

    def mask_prev_latent(self, prev):
        """Long-form chaining (:286-297): only the first token of each body part is kept, and it takes the
        LAST token of that part of the previous window; None stays None."""
        if prev is None:
            return None
        masked = torch.zeros_like(prev)
        for rows in self._row_sets(prev.shape[1]):
            masked[:, rows[0]] = prev[:, rows[-1]]
        return masked

    def invert_many(self, gbs):
        """DDIM-invert the exemplars of one or several prepared batches in ONE batched 50-step reverse loop
        (the reference: one loop per exemplar at B=1, :323-354).  Sets gb.inv = [steps, E_gb, T, D]."""
        todo = [gb for gb in gbs if gb.use_inversion and gb.jobs and gb.inv is None]
        if not todo:
            return
        diff, device = self.diffusion_test, todo[0].device
        cat = lambda key: torch.cat([gb.ex[key] for gb in todo], 0)
        self.model._state_cache = (None, None)
        with torch.no_grad():
            ex_kwargs = self.model.get_precompute_condition(device=device, text=cat("retr_text"), audio=cat("retr_audio"),
                                                            speaker_ids=cat("retr_spkid"), re_dict=1)
        ex_kwargs["query_mask"] = {c: torch.cat([gb.ex_query_mask[c] for gb in todo], 0) for c in CFG.CONDS}
        ex_kwargs["motion_mask"] = cat("retr_motion_mask")
        if todo[0].extra:       # unknown sampler keywords: let the step-by-step loop accept or refuse them
            inv = diff.ddim_reverse_sample_loop(self.model, start_img=cat("retr_motion_latent"), clip_denoised=False,
                                                progress=False, model_kwargs=ex_kwargs, eta=0,
                                                return_all_timesteps=True, **todo[0].extra)
        else:                   # all levels in one library call
            _, inv = diff.run_levels(self.model, reverse=dict(start_img=cat("retr_motion_latent"), model_kwargs=ex_kwargs))
        inv = torch.stack(inv, 0)                       # [steps, E_total, T, D], clean -> noisy
        off = 0
        for gb in todo:
            gb.inv = inv[:, off:off + len(gb.jobs)]
            off += len(gb.jobs)

    # ---- stage 2: the hot path proper, device-resident inputs -> output latents ---------------------
    def insertion_targets(self, gb):
        """The deterministic part of the guided loop's inputs for a batch whose exemplars are inverted (gb.inv):
        upper-body and hands windows of the inverted latents placed at the query windows (:394-407).  Returns
        (start_rows [B,T,D], start_mask [B,T,D] bool, inv_per_t [S,B,T,D] or None): start_noise takes
        start_rows where start_mask is set; inv_per_t are the per-level insertion targets.  No RNG is drawn
        here, so a rank that owns the window can compute them and ship them to the rank that samples
        (LongformSynthesizer)."""
        diff, (B, T, D), device = self.diffusion_test, gb.shape, gb.device
        n = (T - 3) // 4
        start_rows = torch.zeros(B, T, D, device=device)
        start_mask = torch.zeros(B, T, D, dtype=torch.bool, device=device)
        inv_per_t = torch.zeros(diff.num_timesteps, B, T, D, device=device) if gb.use_guidance else None
        if gb.jobs:
            inv = gb.inv
            src, dst = [], []                           # (exemplar, row) -> (clip, row), in the reference's order
            for e, ((b, _), ((r0, r1), (q0, q1))) in enumerate(zip(gb.jobs, gb.windows)):
                assert r1 - r0 == q1 - q0
                for o in (0, n + 1):                    # upper body and hands only
                    src += [(e, o + r) for r in range(r0, r1)]
                    dst += [(b, o + q) for q in range(q0, q1)]
            if len(set(dst)) == len(dst):               # disjoint windows: one gather/scatter per tensor
                if dst:
                    # one pinned, asynchronous copy of the four index lists: torch.tensor(list, device=cuda) is a
                    # blocking pageable copy that makes the host wait for everything enqueued before it (the whole
                    # previous pass), so the host could never run ahead of the device
                    idx = torch.tensor([*zip(*src), *zip(*dst)], dtype=torch.int64)
                    if device.type == "cuda":
                        idx = idx.pin_memory()
                    ei, ri, bi, qi = idx.to(device, non_blocking=True).unbind(0)
                    start_rows[bi, qi] = inv[gb.inversion_start_time][ei, ri]
                    start_mask[bi, qi] = True
                    if gb.use_guidance:
                        inv_per_t[:, bi, qi] = inv[:, ei, ri]
            else:                                       # overlapping windows: later exemplars overwrite earlier ones
                for e, ((b, _), ((r0, r1), (q0, q1))) in enumerate(zip(gb.jobs, gb.windows)):
                    for o in (0, n + 1):
                        start_rows[b, o + q0:o + q1] = inv[gb.inversion_start_time, e, o + r0:o + r1]
                        start_mask[b, o + q0:o + q1] = True
                        if gb.use_guidance:
                            inv_per_t[:, b, o + q0:o + q1] = inv[:, e, o + r0:o + r1]
        return start_rows, start_mask, inv_per_t

    def _guided_inputs(self, gb):
        """start_noise and the per-level insertion targets of a batch: the one RNG draw of this stage
        (start_noise, :300) plus insertion_targets (computed here, or shipped in gb.targets)."""
        diff, (B, T, D), device = self.diffusion_test, gb.shape, gb.device
        n = (T - 3) // 4
        start_noise = diff._randn((B, T, D), device)
        start_rows, start_mask, inv_per_t = gb.targets if gb.targets is not None else self.insertion_targets(gb)
        start_noise = torch.where(start_mask, start_rows, start_noise)
        if gb.use_guidance and gb.use_prev and gb.prev_latent is not None:
            inv_per_t = inv_per_t.clone() if gb.targets is not None else inv_per_t
            inv_per_t[:, :, [0, n + 1, 2 * n + 2, 3 * n + 3], :] = 0
        return start_noise, inv_per_t

    def run_prepared(self, gb, keep_inversion=False):
        """K6 state for the B clips and E exemplars, ONE batched 50-step inversion of the exemplars,
        window insertion, 50 guided (or plain) sampling steps: 50 * (B + E) clip-steps."""
        diff, (B, T, D) = self.diffusion_test, gb.shape
        start_noise, inv_per_t = None, None
        if gb.use_inversion:
            if gb.jobs and gb.targets is None:
                # the reference draws start_noise before it inverts (:300); inversion draws nothing, so the
                # order of the two does not change any generator's sequence
                self.invert_many([gb])
            start_noise, inv_per_t = self._guided_inputs(gb)
            if not keep_inversion:
                gb.inv = None                           # a GuidedBatch re-run (bench) inverts again
        self.model._state_cache = (None, None)
        noise = start_noise if gb.use_inversion else None
        if not gb.extra:        # all levels in one library call (rg_run_levels)
            if gb.use_guidance:
                spec = dict(in_seq=gb.prev_latent if gb.use_prev else None, guidance_iters=gb.guidance_iters,
                            inverted_latent_list=inv_per_t, guidance_lr=gb.guidance_lr)
            else:
                spec = dict(in_seq=gb.prev_latent if gb.use_prev else (gb.outpaint_seq if gb.use_outpaint else None))
            output, _ = diff.run_levels(self.model, guided=dict(shape=(B, T, D), noise=noise,
                                                                model_kwargs=gb.model_kwargs, **spec))
        elif gb.use_guidance:
            output = diff.ddim_guided_sample_loop(
                self.model, (B, T, D), noise=start_noise if gb.use_inversion else None, clip_denoised=False,
                progress=False, model_kwargs=gb.model_kwargs, eta=0,
                in_seq=gb.prev_latent if gb.use_prev else None, guidance_iters=gb.guidance_iters,
                inverted_latent_list=inv_per_t, guidance_lr=gb.guidance_lr, **gb.extra)
        else:
            in_seq = gb.prev_latent if gb.use_prev else (gb.outpaint_seq if gb.use_outpaint else None)
            output = diff.ddim_sample_loop(
                self.model, (B, T, D), noise=start_noise if gb.use_inversion else None, clip_denoised=False,
                progress=False, model_kwargs=gb.model_kwargs, eta=0, in_seq=in_seq, **gb.extra)
        if getattr(self.model, "post_process") is not None:
            output = self.model.post_process(output)
        return output

    def run_pass(self, gb_guided=None, gb_invert=None):
        """One pass of the two-batch software pipeline: the guided sampling of `gb_guided` (exemplars already
        inverted by the previous pass) and the DDIM inversion of `gb_invert`'s exemplars as ONE kernel chain
        per level (SpacedDiffusion.ddim_guided_and_reverse_loops).  Either may be None (first / last pass).
        Falls back to the separate loops whenever the pair is not an insertion-guided pair.  Returns the
        output latents of gb_guided (or None); sets gb_invert.inv."""
        g, v = gb_guided, gb_invert
        fusable = (g is not None and v is not None and g is not v and g.use_guidance and g.use_inversion
                   and v.use_inversion and len(v.jobs) > 0 and v.inv is None and not g.extra and not v.extra
                   and (not g.jobs or g.inv is not None))
        if not fusable:
            if v is not None:
                self.invert_many([v])
            return self.run_prepared(g) if g is not None else None
        diff, (B, T, D), device = self.diffusion_test, g.shape, g.device
        start_noise, inv_per_t = self._guided_inputs(g)
        g.inv = None
        with torch.no_grad():
            ex_kwargs = self.model.get_precompute_condition(device=device, text=v.ex["retr_text"], audio=v.ex["retr_audio"],
                                                            speaker_ids=v.ex["retr_spkid"], re_dict=1)
        ex_kwargs["query_mask"] = v.ex_query_mask
        ex_kwargs["motion_mask"] = v.ex["retr_motion_mask"]
        output, inv = diff.ddim_guided_and_reverse_loops(
            self.model,
            guided=dict(shape=(B, T, D), noise=start_noise, model_kwargs=g.model_kwargs,
                        in_seq=g.prev_latent if g.use_prev else None, guidance_iters=g.guidance_iters,
                        inverted_latent_list=inv_per_t, guidance_lr=g.guidance_lr),
            reverse=dict(start_img=v.ex["retr_motion_latent"], model_kwargs=ex_kwargs))
        v.inv = torch.stack(inv, 0)
        if getattr(self.model, "post_process") is not None:
            output = self.model.post_process(output)
        return output

    def run_many(self, gbs):
        """Outputs of a list of prepared batches through the two-batch pipeline: pass k runs the guided loop
        of batch k-1 together with the inversion of batch k."""
        outs, prev = [], None
        for gb in list(gbs) + [None]:
            out = self.run_pass(prev, gb)
            if prev is not None:
                outs.append(out)
            prev = gb
        return outs

    # ---- stage 3: decode ---------------------------------------------------------------------------------
    def finish(self, gb, output):
        results = gb.results
        results["prev_latentout"] = output
        with torch.no_grad():
            up, lo, fa, ha, tr, ex, _ = self.model.gesture_rep_encoder.decode(output)
        results.update(pred_upper=up, pred_lower=lo, pred_facepose=fa, pred_hands=ha, pred_transl=tr, pred_exps=ex)
        return results


class GuidedBatch:
    """Device-resident inputs of one guided batch between MotionDiffusion.prepare and run_prepared."""
    jobs, ex, ex_query_mask, windows, outpaint_seq, prev_latent, inv = (), None, None, (), None, None, None
    cond_inputs = None
    targets = None          # (start_rows, start_mask, inv_per_t) computed elsewhere (insertion_targets on another rank)

    def clip_steps(self, num_timesteps):
        """Work of the batch in the metric's unit: 50 * (B + E)."""
        return num_timesteps * (self.shape[0] + len(self.jobs))


def _record_streams(obj, stream, _seen=None):
    """Tell the caching allocator that every CUDA tensor reachable from `obj` is also used on `stream`."""
    _seen = set() if _seen is None else _seen
    if id(obj) in _seen:
        return
    _seen.add(id(obj))
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_streams(v, stream, _seen)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_streams(v, stream, _seen)
    elif isinstance(obj, GuidedBatch):
        _record_streams(vars(obj), stream, _seen)


class GuidedPipeline:
    """Throughput form of `for batch in loader: model(**batch)` (tools/visualize.py:189-200): while the
    denoising loops of batch i run, a worker thread executes stage 1 of batch i+1 -- H2D of the pinned host
    batch, codec encode, discourse retrieval with its text-similarity ranking, exemplar fetch and encode -- on
    a side stream, and the inversion loop of batch i+1's exemplars is fused level by level with the guided loop
    of batch i (MotionDiffusion.run_pass).  The worker never touches the denoiser handle (the clip
    pre-projection is deferred to the main thread).

    Random streams.  The sampler noise is drawn by the main thread from the default CUDA generator.  The codec's
    rsample noise is drawn by the worker: a codec that draws from the CPU generator (SyntheticGestureCodec) is
    the only user of that generator, so results are bit-identical to sequential forward() calls.  The
    TransformerVAE codec (vae.GestureRepEncoder) draws ON THE DEVICE; sharing the default CUDA generator between
    the two threads would make the interleaving depend on thread timing, so for the duration of run() the codec
    gets its own device generator (`codec.generator`, seeded from torch.cuda.initial_seed() unless the caller
    installed one).  Results are then reproducible run to run, and equal to sequential forward() calls made with
    the same `codec.generator` installed -- but NOT to sequential calls that draw codec and sampler noise from
    the one default generator, as the reference does.  In the tensor-core tiers the codec's passes are replayed as
    CUDA graphs inside run() (`codec_graphs`): still bit-reproducible run to run, and equal to the eager sequential
    calls up to the rounding of the few cuBLAS calls left in those passes (tests/test_gpu_codec.py).

        for results in GuidedPipeline(model).run(loader): ...
    """

    def __init__(self, arch, side_priority=0, codec_graphs=None):
        self.arch = arch
        self.device = arch.model.out.weight.device
        if self.device.type != "cuda":
            raise RuntimeError("rg_b200: GuidedPipeline needs the model on a CUDA device (no CPU fallback)")
        # Stage 1's kernels (codec, text-similarity ranks, exemplar gathers) share the GPU with the loops.  A
        # high-priority side stream (-1) was measured and is WORSE (tools/diag_e2e.py 16 -1: 121 ms per batch against
        # 87 at priority 0, with 150-300 ms stalls in the stage's host-to-device copies), so the default stays 0.
        self.side = torch.cuda.Stream(self.device, priority=side_priority)
        # A codec that can replay its passes as CUDA graphs (vae.GestureRepEncoder.enable_graphs) does so inside run():
        # by default whenever its GEMMs are on the tensor-core path, where the passes are bound by the host's eager
        # launches; the fp32 tier keeps the eager passes (bit-equal to sequential forward() calls).
        codec = arch.model.gesture_rep_encoder
        self.codec_graphs = (getattr(codec, "gemm_tier", None) is not None) if codec_graphs is None else bool(codec_graphs)
        # The main thread re-acquires the GIL after every blocking call; while the worker runs Python it may
        # wait one switch interval each time (default 5 ms).  The loops are one C call per pass
        # (rg_run_levels), so a moderate interval is enough; very short ones (50 us) make both threads thrash.
        self.switch_interval = 5e-4

    @staticmethod
    def codec_generator(device, seed=None):
        """The device generator GuidedPipeline gives a device-drawing codec (see the class docstring)."""
        g = torch.Generator(device=device)
        g.manual_seed((torch.cuda.initial_seed() if seed is None else int(seed)) ^ 0x5DEECE66D)
        return g

    def _stage1(self, kwargs, main):
        # No wait on the main stream here: the consumer may be a whole pass ahead of the device, and the
        # host-side waits of this stage (text-similarity ranks) would then stall behind that pass.  Memory
        # handed to the main stream is protected by record_stream below.
        with torch.cuda.device(self.device), torch.cuda.stream(self.side):
            gb = self.arch.prepare(defer_conditions=True, **kwargs)
            ready = self.side.record_event()
        _record_streams(gb, main)
        return gb, ready

    def run(self, batches):
        """batches: iterable of forward() keyword dicts (host or device tensors).  Yields the result dict of
        each batch in order; at most one batch is prepared ahead."""
        import sys
        main = torch.cuda.current_stream(self.device)
        it = iter(batches)
        try:
            first = next(it)
        except StopIteration:
            return
        old_interval = sys.getswitchinterval()
        sys.setswitchinterval(min(old_interval, self.switch_interval))
        codec = self.arch.model.gesture_rep_encoder
        own_gen = getattr(codec, "draws_on_device", False) and getattr(codec, "generator", None) is None
        if own_gen:
            codec.generator = self.codec_generator(self.device)
        graphs_were = getattr(codec, "use_graphs", None)
        if self.codec_graphs and graphs_were is False:
            codec.use_graphs = True                 # captured passes are kept on the codec between runs
        try:
            yield from self._run(it, first, main)
        finally:
            sys.setswitchinterval(old_interval)
            if own_gen:
                codec.generator = None
            if graphs_were is not None:
                codec.use_graphs = graphs_were

    def _run(self, it, first, main):
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1, thread_name_prefix="rg-stage1") as pool:
            fut = pool.submit(self._stage1, dict(first), main)
            cur = None                              # prepared + inverted, waiting for its guided pass
            while fut is not None or cur is not None:
                gb = None
                if fut is not None:
                    gb, ready = fut.result()
                    nxt = next(it, None)
                    fut = pool.submit(self._stage1, dict(nxt), main) if nxt is not None else None
                    main.wait_event(ready)
                    self.arch.encode_clip_conditions(gb)
                out = self.arch.run_pass(cur, gb)   # guided loop of `cur` fused with the inversion of `gb`
                if cur is not None:
                    yield self.arch.finish(cur, out)
                cur = gb
