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
        out = self._forward_step(eng, prep, x, self._uniform_step(t), in_seq, model)
        return {"sample": out[0], "pred_xstart": out[1]}

    def ddim_reverse_sample(self, model, x, t, clip_denoised=True, denoised_fn=None,
                            model_kwargs=None, eta=0.0, pre_seq=None):
        self._check(clip_denoised, denoised_fn, None, eta, pre_seq)
        eng = self._engine(model)
        prep = model.prepare_batch(model_kwargs, x.shape[0])
        i = self._uniform_step(t)
        x0 = self._x0(model, eng, prep, x, i)
        return {"sample": eng.ddim_update(x, x0, i, +1), "pred_xstart": x0}

    @staticmethod
    def _x0(model, eng, prep, x, i, coefs=None):
        """One denoiser evaluation at schedule level i: the single-branch engine call, or -- when the model was built
        with scale_func_cfg -- both branches of forward_test and their mix (ReGestureTransformer.two_branch_x0)."""
        if getattr(model, "two_branch", False):
            return model.two_branch_x0(eng, prep, x, step_idx=i, coefs=coefs)
        return eng.denoise(x, prep.src_mask, prep.query_mask, prep.state, step_idx=i)

    def _forward_step(self, eng, prep, x, i, in_seq, model=None):
        """blend (if in_seq) -> denoise -> x_{t-1}; draws noise in the order of :945 then :991."""
        if in_seq is not None:
            x = eng.blend_in_seq(x, in_seq.contiguous(), self._randn(in_seq.shape, x.device), i)
        x0 = self._x0(model, eng, prep, x, i)
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
                img, x0 = self._forward_step(eng, prep, img, i, in_seq, model)
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
        coefs = None
        if getattr(model, "two_branch", False) and img.shape[0] > 1:
            # 2-branch mode draws a coefficient set per evaluation above t = 100 (Python's `random`).  The reference
            # inverts exemplar by exemplar (E loops at B = 1), so exemplar e's 50 draws come before exemplar e+1's:
            # the batched loop pre-draws the table in that order and hands every level its column.
            rows = []
            for _ in range(img.shape[0]):
                cs = [model.scale_func_retr(int(eng.timestep_map[i])) for i in range(self.num_timesteps)]
                rows.append([[c["both_coef"], c["text_coef"], c["retr_coef"], c["none_coef"]] for c in cs])
            coefs = torch.tensor(rows, dtype=torch.float32)            # [E, S, 4]
        with torch.no_grad():
            for i in range(self.num_timesteps):
                x0 = self._x0(model, eng, prep, img, i, None if coefs is None else coefs[:, i])
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
                img, x0 = self._forward_step(eng, prep, img, i, in_seq, model)
                yield {"sample": img, "pred_xstart": x0}

    def run_levels(self, model, guided=None, reverse=None):
        """All S levels of a sampling loop and/or an inversion loop in ONE call into the library
        (rg_run_levels), the form MotionDiffusion uses on the hot path:

        guided  = dict(shape, noise, model_kwargs, in_seq, guidance_iters=None, inverted_latent_list=None,
                       guidance_lr=0.1): ddim_guided_sample_loop when inverted_latent_list is given, else
                       ddim_sample_loop (in_seq blended at every level);
        reverse = dict(start_img, model_kwargs): ddim_reverse_sample_loop(return_all_timesteps=True) of ANOTHER
                  batch's exemplars; pass j evaluates the clips at level S-1-j and the exemplars at level j in one
                  kernel chain (rg_denoise_groups), so the fixed costs of the ~100-kernel chain are paid once.
        -> (final sample [B,T,D] or None, list of S inverted latents [E,T,D] or None)

        Per clip the arithmetic equals the step-by-step loops of this class.  All Gaussian draws of the loop are
        made up front in the loop's order (per level: randn_like(in_seq) if there is one, then the unused
        randn_like(x) of gaussian_diffusion.py:991), so any generator / noise tape sees the same sequence."""
        S = self.num_timesteps
        eng = self._engine(model)
        if getattr(model, "two_branch", False):
            # 2-branch mode mixes two evaluations per level on the host side of the ABI: the step-by-step loops
            out_g = out_r = None
            if reverse is not None:                 # the reference inverts first (diffusion_architecture.py:323-354)
                out_r = self.ddim_reverse_sample_loop(model, start_img=reverse["start_img"], clip_denoised=False,
                                                      model_kwargs=reverse["model_kwargs"], eta=0, return_all_timesteps=True)
            if guided is not None:
                kw = dict(noise=guided.get("noise", None), clip_denoised=False, model_kwargs=guided["model_kwargs"], eta=0,
                          in_seq=guided.get("in_seq", None))
                if guided.get("inverted_latent_list", None) is not None:
                    out_g = self.ddim_guided_sample_loop(model, guided["shape"], guidance_iters=guided.get("guidance_iters", None),
                                                         inverted_latent_list=guided["inverted_latent_list"],
                                                         guidance_lr=guided.get("guidance_lr", 0.1), **kw)
                else:
                    out_g = self.ddim_sample_loop(model, guided["shape"], **kw)
            return out_g, out_r
        B = E = 0
        parts = []
        if guided is not None:
            B = guided["shape"][0]
            model._state_cache = (None, None)
            parts.append(model.prepare_batch(guided["model_kwargs"], B))
        if reverse is not None:
            start = reverse["start_img"].float().contiguous()
            E = start.shape[0]
            model._state_cache = (None, None)
            parts.append(model.prepare_batch(reverse["model_kwargs"], E))
        model._state_cache = (None, None)
        assert parts, "run_levels needs a guided and/or a reverse part"
        device = parts[0].state.device
        if len(parts) == 2:
            if (parts[0].query_mask is None) != (parts[1].query_mask is None):
                raise NotImplementedError("both groups need a query_mask, or neither")
            src_mask = torch.cat([p.src_mask for p in parts], 0)
            qm = None if parts[0].query_mask is None else torch.cat([p.query_mask for p in parts], 1).contiguous()
            state = torch.cat([p.state for p in parts], 0)
        else:
            src_mask, qm, state = parts[0].src_mask, parts[0].query_mask, parts[0].state
        del parts
        T, D = eng.n_tokens, eng.latent_dim
        xj = torch.empty(B + E, T, D, device=device)
        in_seq0 = inv_list = noise = g_iters = samples = None
        lr = 0.1
        if guided is not None:
            img = guided.get("noise", None)
            xj[:B].copy_((img if img is not None else self._randn(guided["shape"], device)).float())
            in_seq0, inv_list = guided.get("in_seq", None), guided.get("inverted_latent_list", None)
            g_iters, lr = guided.get("guidance_iters", None), guided.get("guidance_lr", 0.1)
            if inv_list is not None:
                if g_iters is None:
                    g_iters = [1] * S
                assert len(g_iters) == len(inv_list) == S
                inv_list = inv_list.float().contiguous()
            if in_seq0 is not None:
                in_seq0 = in_seq0.float().contiguous()
            if in_seq0 is not None or inv_list is not None:
                noise = torch.empty(S, B, T, D, device=device)
            for j in range(S):
                if (inv_list is not None and j != 0) or in_seq0 is not None:
                    noise[j].copy_(self._randn((B, T, D), device))
                self._randn((B, T, D), device)      # randn_like(x) of :991 -- sigma = 0, value unused
        if reverse is not None:
            xj[B:].copy_(start)
            samples = torch.empty(S, E, T, D, device=device)
        with torch.no_grad():
            eng.run_levels(S, xj, B, E, src_mask, qm, state, in_seq0=in_seq0, inv_list=inv_list, noise=noise,
                           guidance_iters=g_iters, guidance_lr=lr, run_dead_guidance=not self.skip_dead_guidance,
                           samples_out=samples)
        return (xj[:B].clone() if guided is not None else None,
                list(samples.unbind(0)) if reverse is not None else None)

    def ddim_guided_and_reverse_loops(self, model, guided, reverse):
        """Guided sampling of one batch + DDIM inversion of another batch's exemplars, one kernel chain per
        level (see run_levels)."""
        if guided.get("inverted_latent_list", None) is None:
            raise ValueError("inverted_latent_list must be provided for guided sampling")
        return self.run_levels(model, guided=guided, reverse=reverse)

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
