"""ORACLE (test infrastructure, not product): CPU restatement of the reference DDIM sampler.

Follows mogen/models/utils/gaussian_diffusion.py for the START_X / FIXED_LARGE / eta=0 /
clip_denoised=False configuration that MotionDiffusion.forward uses
(diffusion_architecture.py:345-474): float64 numpy tables, cast to fp32 at use (:1623), fp32
tensor arithmetic in the reference's operation order, the same global-RNG draw order
(SURVEY App. B).  Parity pin: tests/golden (see oracle/denoiser.py header).
"""
import numpy as np
import torch


def scaled_linear_betas(n=1000, beta_start=0.00085, beta_end=0.012):
    """get_named_beta_schedule("scaled_linear") gaussian_diffusion.py:252-266."""
    return np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2


def space_timesteps(num_timesteps, section_counts):
    """gaussian_diffusion.py:1629-1711, comma-separated-counts form only (the shipped config
    uses "15,15,8,6,6")."""
    counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(counts))
    start, steps = 0, []
    for i, cnt in enumerate(counts):
        size = size_per + (1 if i < extra else 0)
        if size < cnt:
            raise ValueError(f"cannot divide section of {size} steps into {cnt}")
        stride = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


class Schedule:
    """SpacedDiffusion.__init__ (:1723-1738) + GaussianDiffusion.__init__ tables (:399-440)."""

    def __init__(self, diffusion_steps=1000, respace="15,15,8,6,6"):
        base = np.cumprod(1.0 - scaled_linear_betas(diffusion_steps), axis=0)
        use = space_timesteps(diffusion_steps, respace)
        last, betas, self.timestep_map = 1.0, [], []
        for i, ac in enumerate(base):
            if i in use:
                betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        betas = np.array(betas, dtype=np.float64)
        self.num_timesteps = len(betas)
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)


def _ext(arr, t, shape):
    """_extract_into_tensor :1614-1626: float64 table -> index -> .float() -> broadcast."""
    res = torch.from_numpy(arr)[t].float()
    while res.dim() < len(shape):
        res = res[..., None]
    return res.expand(shape)


class OracleDiffusion:
    def __init__(self, schedule=None, randn=None):
        self.s = schedule or Schedule()
        self.num_timesteps = self.s.num_timesteps
        # every Gaussian draw goes through here, in the reference's order
        self.randn = randn or (lambda shape, device="cpu": torch.randn(*shape))

    # -- pieces ------------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise):
        """:459-477."""
        return (_ext(self.s.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + _ext(self.s.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def _model_x0_eps(self, model, x, t, model_kwargs):
        """p_mean_variance START_X branch :593-606 via _WrappedModel :1761-1764."""
        ts = torch.tensor(self.s.timestep_map, dtype=t.dtype)[t]
        kw = dict(model_kwargs or {})
        kw["do_clf_guidance"] = False
        x0 = model(x, ts, **kw)
        eps = (_ext(self.s.sqrt_recip_alphas_cumprod, t, x.shape) * x - x0) \
            / _ext(self.s.sqrt_recipm1_alphas_cumprod, t, x.shape)
        return x0, eps

    def blend_in_seq(self, x, in_seq, t):
        """in_seq outpainting blend of ddim_sample :934-947."""
        nz = (in_seq != 0).any(dim=-1)
        zero_mask = (~nz).to(torch.int)
        nz = nz.to(torch.int)
        x = x * zero_mask.unsqueeze(-1).float()
        x_t = self.q_sample(in_seq, t, self.randn(in_seq.shape))
        return x + x_t * nz.unsqueeze(-1).float()

    def ddim_sample(self, model, x, t, model_kwargs=None, in_seq=None):
        """:910-1001 with eta=0, clip_denoised=False, cond_fn=None, pre_seq=None."""
        if in_seq is not None:
            x = self.blend_in_seq(x, in_seq, t)
        x0, eps = self._model_x0_eps(model, x, t, model_kwargs)
        alpha_bar = _ext(self.s.alphas_cumprod, t, x.shape)
        alpha_bar_prev = _ext(self.s.alphas_cumprod_prev, t, x.shape)
        sigma = 0.0 * torch.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar)) \
            * torch.sqrt(1 - alpha_bar / alpha_bar_prev)
        noise = self.randn(x.shape)                     # drawn although sigma == 0 (:991)
        mean_pred = x0 * torch.sqrt(alpha_bar_prev) + torch.sqrt(1 - alpha_bar_prev - sigma ** 2) * eps
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        return mean_pred + nonzero * sigma * noise, x0

    def ddim_reverse_sample(self, model, x, t, model_kwargs=None):
        """:1003-1040."""
        x0, eps = self._model_x0_eps(model, x, t, model_kwargs)
        alpha_bar_next = _ext(self.s.alphas_cumprod_next, t, x.shape)
        return x0 * torch.sqrt(alpha_bar_next) + torch.sqrt(1 - alpha_bar_next) * eps

    # -- loops -------------------------------------------------------------------------------
    def ddim_sample_loop(self, model, shape, noise=None, model_kwargs=None, in_seq=None,
                         trajectory=None):
        """:1042-1135."""
        img = noise if noise is not None else self.randn(shape)
        for i in reversed(range(self.num_timesteps)):
            t = torch.tensor([i] * shape[0])
            with torch.no_grad():
                img, _ = self.ddim_sample(model, img, t, model_kwargs, in_seq)
            if trajectory is not None:
                trajectory.append(img)
        return img

    def ddim_reverse_sample_loop(self, model, start_img, model_kwargs=None):
        """:1137-1230 with return_all_timesteps=True: list of 50 latents, clean -> noisy."""
        img, out = start_img, []
        for i in range(self.num_timesteps):
            t = torch.tensor([i] * img.shape[0])
            with torch.no_grad():
                img = self.ddim_reverse_sample(model, img, t, model_kwargs)
            out.append(img)
        return out

    def ddim_guided_sample_loop(self, model, shape, noise=None, model_kwargs=None, in_seq=None,
                                guidance_iters=None, inverted_latent_list=None, guidance_lr=0.1,
                                autograd_guidance=True, trajectory=None):
        """:1233-1395.  The gradient steps are executed (through autograd like the reference when
        `autograd_guidance`, else in closed form) although the blend that follows overwrites
        every row they touch."""
        if guidance_iters is None:
            guidance_iters = [1] * self.num_timesteps
        if inverted_latent_list is None:
            raise ValueError("inverted_latent_list must be provided for guided sampling")
        assert len(guidance_iters) == len(inverted_latent_list)
        img = noise if noise is not None else self.randn(shape)
        first = self.num_timesteps - 1
        for i in reversed(range(self.num_timesteps)):
            if i != first:
                in_seq = inverted_latent_list[i]
                mask = (in_seq != 0).any(dim=-1)
                if autograd_guidance:
                    with torch.enable_grad():
                        lat = img.clone().detach().requires_grad_(True)
                        for _ in range(guidance_iters[i]):
                            loss = torch.nn.functional.mse_loss(lat * mask.unsqueeze(-1).float(), in_seq)
                            (g,) = torch.autograd.grad(loss, [lat], retain_graph=True)
                            lat = lat - guidance_lr * g
                    img = lat.detach()
                else:
                    m = mask.unsqueeze(-1).float()
                    for _ in range(guidance_iters[i]):
                        img = img - guidance_lr * (2.0 / img.numel()) * m * (img * m - in_seq)
            t = torch.tensor([i] * shape[0])
            with torch.no_grad():
                img, _ = self.ddim_sample(model, img, t, model_kwargs, in_seq)
            if trajectory is not None:
                trajectory.append(img)
        return img
