"""Host side of the DDIM sampler: schedule tables and the three loops, same entry points as the
reference's `diffusion_test` object (mogen/models/utils/gaussian_diffusion.py):

    ddim_sample_loop            (:1042)   ddim_reverse_sample_loop   (:1137)
    ddim_guided_sample_loop     (:1233)   ddim_sample / ddim_reverse_sample (:910, :1003)
    q_sample (:459)  num_timesteps  timestep_map  alphas_cumprod*  (attributes callers read)

Every arithmetic step (denoiser, eps / x_{t+-1} update, in_seq blend, guidance step) is a call into
librg_b200.so; this file only sequences them and draws the Gaussian noise in the reference's order
 so that same-seed runs line up.  Supported configuration = what
MotionDiffusion.forward uses at inference: model_mean_type="start_x", eta=0, clip_denoised=False,
no cond_fn / denoised_fn / pre_seq, classifier_free_guidance_scale=0.  Anything else raises.
DDPM sampling (p_sample_loop), training losses and VLB terms are out of scope.
"""
import math

import numpy as np
import torch


def get_named_beta_schedule(name, n):
    """float64 beta tables (gaussian_diffusion.py:229-268)."""
    if name == "linear":
        scale = 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if name == "scaled_linear":
        return np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, n, dtype=np.float64) ** 2
    if name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return np.array([min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)], dtype=np.float64)
    raise NotImplementedError(f"unknown beta schedule: {name}")


def space_timesteps(num_timesteps, section_counts, num_inference_timesteps=None):
    """Which of the original timesteps the respaced process keeps (gaussian_diffusion.py:1629-1711)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {want} steps with an integer stride")
        if section_counts == "leading":
            ratio = num_timesteps // num_inference_timesteps
            return set((np.arange(0, num_inference_timesteps) * ratio).round().tolist())
        if section_counts == "trailing":
            ratio = num_timesteps / num_inference_timesteps
            t = np.round(np.arange(num_timesteps, 0, -ratio)).astype(np.int64) - 1
            return set(np.append(t, 0).tolist())
        section_counts = [int(x) for x in section_counts.split(",")]
        if num_inference_timesteps is not None:
            assert sum(section_counts) == num_inference_timesteps
    per, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for i, count in enumerate(section_counts):
        size = per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion:
    """Respaced Gaussian diffusion tables + DDIM loops driving the CUDA engine."""

    def __init__(self, use_timesteps, betas, model_mean_type="start_x", model_var_type="fixed_large",
                 classifier_free_guidance_scale=0, **_unused):
        if model_mean_type != "start_x":
            raise NotImplementedError("rg_b200 implements the START_X parameterisation only")
        if classifier_free_guidance_scale:
            raise NotImplementedError("classifier-free guidance (2-branch mode) is out of scope (SURVEY 8f.4)")
        self.model_mean_type, self.model_var_type = model_mean_type, model_var_type
        self.classifier_free_guidance_scale = 0
        base = np.cumprod(1.0 - np.asarray(betas, dtype=np.float64), axis=0)
        self.use_timesteps = set(int(t) for t in use_timesteps)
        self.original_num_steps = len(base)
        self.timestep_map, new_betas, last = [], [], 1.0
        for i, ac in enumerate(base):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        self.betas = np.array(new_betas, dtype=np.float64)
        self.num_timesteps = len(self.betas)
        ac = np.cumprod(1.0 - self.betas, axis=0)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.noise_fn = None          # tests install a tape here; default = torch.randn on device
        self.skip_dead_guidance = True

    # fp32 coefficients exactly as the reference forms them at use (float64 table -> .float(),
    # then th.sqrt / 1 - x in fp32 on the tensor; gaussian_diffusion.py:1623, 994-995, 1035-1036)
    def coef_table(self):
        f32 = np.float32
        abp = self.alphas_cumprod_prev.astype(f32)
        abn = self.alphas_cumprod_next.astype(f32)
        c = np.empty((self.num_timesteps, 8), dtype=f32)
        c[:, 0] = self.sqrt_recip_alphas_cumprod.astype(f32)
        c[:, 1] = self.sqrt_recipm1_alphas_cumprod.astype(f32)
        c[:, 2] = np.sqrt(abp)
        c[:, 3] = np.sqrt(f32(1) - abp)
        c[:, 4] = np.sqrt(abn)
        c[:, 5] = np.sqrt(f32(1) - abn)
        c[:, 6] = self.sqrt_alphas_cumprod.astype(f32)
        c[:, 7] = self.sqrt_one_minus_alphas_cumprod.astype(f32)
        return c

    # -- helpers ------------------------------------------------------------------------------------
    def _randn(self, shape, device):
        if self.noise_fn is not None:
            return self.noise_fn(tuple(shape), device)
        return torch.randn(*shape, device=device)

    def _engine(self, model):
        eng = getattr(model, "rg_engine", None)
        if eng is None:
            raise RuntimeError("rg_b200 sampler needs a CUDA-backed denoiser (ReGestureTransformer of "
                               "rag_gesture_b200.mogen_api); there is no generic PyTorch path")
        return eng(self)

    @staticmethod
    def _check(clip_denoised, denoised_fn, cond_fn, eta, pre_seq):
        if clip_denoised or denoised_fn is not None or cond_fn is not None or eta != 0.0 or pre_seq is not None:
            raise NotImplementedError("rg_b200 sampler supports clip_denoised=False, eta=0, no "
                                      "denoised_fn/cond_fn/pre_seq (what MotionDiffusion.forward uses)")

    @staticmethod
    def _uniform_step(t):
        i = int(t[0])
        if not bool((t == i).all()):
            raise NotImplementedError("all clips of a batch must share one timestep")
        return i

    # -- single steps (API parity with :459, :910, :1003) -----------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        raise NotImplementedError("q_sample is fused into the in_seq blend kernel (rg_blend_in_seq); "
                                  "the stand-alone form is only used by training")

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None,
                    model_kwargs=None, eta=0.0, pre_seq=None, in_seq=None):
        self._check(clip_denoised, denoised_fn, cond_fn, eta, pre_seq)
        eng = self._engine(model)
        prep = model.prepare_batch(model_kwargs, x.shape[0])
        out = self._forward_step(eng, prep, x, self._uniform_step(t), in_seq)
        return {"sample": out[0], "pred_xstart": out[1]}

    def ddim_reverse_sample(self, model, x, t, clip_denoised=True, denoised_fn=None,
                            model_kwargs=None, eta=0.0, pre_seq=None):
        self._check(clip_denoised, denoised_fn, None, eta, pre_seq)
        eng = self._engine(model)
        prep = model.prepare_batch(model_kwargs, x.shape[0])
        i = self._uniform_step(t)
        x0 = eng.denoise(x, prep.src_mask, prep.query_mask, prep.state, step_idx=i)
        return {"sample": eng.ddim_update(x, x0, i, +1), "pred_xstart": x0}

    def _forward_step(self, eng, prep, x, i, in_seq):
        """blend (if in_seq) -> denoise -> x_{t-1}; draws noise in the order of :945 then :991."""
        if in_seq is not None:
            x = eng.blend_in_seq(x, in_seq.contiguous(), self._randn(in_seq.shape, x.device), i)
        x0 = eng.denoise(x, prep.src_mask, prep.query_mask, prep.state, step_idx=i)
        self._randn(x.shape, x.device)      # randn_like(x) of :991 -- sigma = 0, value unused
        return eng.ddim_update(x, x0, i, -1), x0

    # -- loops ----------------------------------------------------------------------------------------------
    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                         cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0,
                         pre_seq=None, in_seq=None):
        final = None
        for out in self.ddim_sample_loop_progressive(model, shape, noise, clip_denoised, denoised_fn,
                                                     cond_fn, model_kwargs, device, progress, eta,
                                                     pre_seq, in_seq):
            final = out
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True,
                                     denoised_fn=None, cond_fn=None, model_kwargs=None, device=None,
                                     progress=False, eta=0.0, pre_seq=None, in_seq=None):
        self._check(clip_denoised, denoised_fn, cond_fn, eta, pre_seq)
        eng = self._engine(model)
        device = device if device is not None else eng.device
        img = noise if noise is not None else self._randn(shape, device)
        img = img.float().contiguous()
        prep = model.prepare_batch(model_kwargs, shape[0])
        with torch.no_grad():
            for i in reversed(range(self.num_timesteps)):
                img, x0 = self._forward_step(eng, prep, img, i, in_seq)
                yield {"sample": img, "pred_xstart": x0}

    def ddim_reverse_sample_loop(self, model, start_img, clip_denoised=True, denoised_fn=None,
                                 model_kwargs=None, device=None, progress=False, pre_seq=None,
                                 eta=0.0, return_all_timesteps=False, num_inv_timesteps=None):
        """DDIM inversion, clean -> noisy.  start_img may hold ANY number of exemplars: the
        reference inverts them one by one at B=1 (diffusion_architecture.py:323-354); no op of the
        denoiser couples clips, so one batched loop gives the same latents."""
        self._check(clip_denoised, denoised_fn, None, eta, pre_seq)
        if num_inv_timesteps is not None:
            raise NotImplementedError("num_inv_timesteps not implemented yet")   # as the reference, :1200
        eng = self._engine(model)
        img = start_img.float().contiguous()
        prep = model.prepare_batch(model_kwargs, img.shape[0])
        samples = []
        with torch.no_grad():
            for i in range(self.num_timesteps):
                x0 = eng.denoise(img, prep.src_mask, prep.query_mask, prep.state, step_idx=i)
                img = eng.ddim_update(img, x0, i, +1)
                if return_all_timesteps:
                    samples.append(img)
        return samples if return_all_timesteps else img

    def ddim_guided_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0,
                                pre_seq=None, in_seq=None, guidance_iters=None,
                                inverted_latent_list=None, guidance_lr=0.1):
        final = None
        for out in self.ddim_guided_sample_loop_progressive(
                model, shape, noise, clip_denoised, denoised_fn, cond_fn, model_kwargs, device,
                progress, eta, pre_seq, in_seq, guidance_iters, inverted_latent_list, guidance_lr):
            final = out
        return final["sample"]

    def ddim_guided_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True,
                                            denoised_fn=None, cond_fn=None, model_kwargs=None,
                                            device=None, progress=False, eta=0.0, pre_seq=None,
                                            in_seq=None, guidance_iters=None,
                                            inverted_latent_list=None, guidance_fn=None,
                                            guidance_lr=0.1):
        """Insertion-guided sampling (:1233-1395).  For every level below the first, in_seq becomes
        inverted_latent_list[i]; `guidance_iters[i]` gradient steps pull the masked rows towards it,
        then ddim_sample's blend re-noises the inverted latent into exactly those rows.  Because the
        blend overwrites every row the gradient steps touched, they cannot change the result
        (pinned by tests); `skip_dead_guidance=False` executes them anyway."""
        self._check(clip_denoised, denoised_fn, cond_fn, eta, pre_seq)
        if guidance_iters is None:
            guidance_iters = [1] * self.num_timesteps
        if inverted_latent_list is None:
            raise ValueError("inverted_latent_list must be provided for guided sampling")
        assert len(guidance_iters) == len(inverted_latent_list)
        eng = self._engine(model)
        device = device if device is not None else eng.device
        img = noise if noise is not None else self._randn(shape, device)
        img = img.float().contiguous()
        prep = model.prepare_batch(model_kwargs, shape[0])
        first = self.num_timesteps - 1
        with torch.no_grad():
            for i in reversed(range(self.num_timesteps)):
                if i != first:
                    in_seq = inverted_latent_list[i]
                    g = int(guidance_iters[i])
                    if g > 0 and not self.skip_dead_guidance:
                        img = eng.guidance_steps(img.clone(), in_seq.contiguous(), g, guidance_lr)
                img, x0 = self._forward_step(eng, prep, img, i, in_seq)
                yield {"sample": img, "pred_xstart": x0}

    def ddim_guided_and_reverse_loops(self, model, guided, reverse):
        """The guided sampling loop of one batch and the DDIM inversion loop of ANOTHER batch's exemplars
        advanced together: pass j evaluates the clips at level S-1-j and the exemplars at level j in ONE
        denoiser call (rg_denoise_groups), so the per-kernel fixed costs of the 100-kernel chain are paid once
        for both.  Per clip the arithmetic, and the order of the noise draws (only the guided loop draws),
        are those of ddim_guided_sample_loop / ddim_reverse_sample_loop: results are bit-identical.

        guided  = dict(shape, noise, model_kwargs, in_seq, guidance_iters, inverted_latent_list, guidance_lr)
        reverse = dict(start_img, model_kwargs)
        -> (final guided sample [B,T,D], list of S inverted latents [E,T,D])"""
        S = self.num_timesteps
        eng = self._engine(model)
        B, (T, D) = guided["shape"][0], guided["shape"][1:]
        inv_list, g_iters = guided["inverted_latent_list"], guided["guidance_iters"]
        if inv_list is None:
            raise ValueError("inverted_latent_list must be provided for guided sampling")
        assert len(g_iters) == len(inv_list)
        start = reverse["start_img"].float().contiguous()
        E, device = start.shape[0], start.device
        # K6 state of both groups (each through the model's own cache logic), then one joint batch
        model._state_cache = (None, None)
        pg = model.prepare_batch(guided["model_kwargs"], B)
        model._state_cache = (None, None)
        pr = model.prepare_batch(reverse["model_kwargs"], E)
        model._state_cache = (None, None)
        src_mask = torch.cat([pg.src_mask, pr.src_mask], 0)
        qm = None
        if pg.query_mask is not None and pr.query_mask is not None:
            qm = torch.cat([pg.query_mask, pr.query_mask], 1).contiguous()
        elif pg.query_mask is not None or pr.query_mask is not None:
            raise NotImplementedError("both groups need a query_mask, or neither")
        state = torch.cat([pg.state, pr.state], 0)
        del pg, pr
        img = guided["noise"] if guided["noise"] is not None else self._randn(guided["shape"], device)
        xj = torch.empty(B + E, T, D, device=device)
        xj[:B].copy_(img.float())
        xj[B:].copy_(start)
        x0 = torch.empty_like(xj)
        samples = torch.empty(S, E, T, D, device=device)
        in_seq, first = guided.get("in_seq", None), S - 1
        with torch.no_grad():
            for j in range(S):
                i = S - 1 - j
                if i != first:
                    in_seq = inv_list[i]
                    g = int(g_iters[i])
                    if g > 0 and not self.skip_dead_guidance:
                        eng.guidance_steps(xj[:B], in_seq.contiguous(), g, guided.get("guidance_lr", 0.1))
                if in_seq is not None:
                    eng.blend_in_seq(xj[:B], in_seq.contiguous(), self._randn(in_seq.shape, device), i, out=xj[:B])
                eng.denoise_groups(xj, src_mask, qm, state, [(B, i), (E, j)], out=x0)
                self._randn((B, T, D), device)      # randn_like(x) of :991 -- sigma = 0, value unused
                eng.ddim_update(xj[:B], x0[:B], i, -1, out=xj[:B])
                eng.ddim_update(xj[B:], x0[B:], j, +1, out=samples[j])
                xj[B:].copy_(samples[j])
        return xj[:B].clone(), list(samples.unbind(0))

    def p_sample_loop(self, *a, **k):
        raise NotImplementedError("DDPM ancestral sampling is outside the rg_b200 hot path (inference_type='ddim')")


def build_diffusion(cfg):
    """Same config keys as diffusion_architecture.py:25-61."""
    betas = get_named_beta_schedule(cfg["beta_scheduler"], cfg["diffusion_steps"])
    respace = cfg.get("respace", None)
    if respace is None:
        use = range(cfg["diffusion_steps"])
    else:
        use = space_timesteps(cfg["diffusion_steps"], respace, cfg.get("num_inference_timesteps", None))
    return SpacedDiffusion(use_timesteps=use, betas=betas, model_mean_type=cfg["model_mean_type"],
                           model_var_type=cfg["model_var_type"],
                           classifier_free_guidance_scale=cfg.get("classifier_free_guidance_scale", 0))
