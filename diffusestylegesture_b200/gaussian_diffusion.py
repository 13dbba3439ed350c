"""Host mirror of the reference sampler interface (reference main/diffusion/gaussian_diffusion.py and
main/diffusion/respace.py): same class names, constructor arguments, table attributes and
``p_sample_loop`` / ``ddim_sample_loop`` signatures, but the 1000-step loop itself runs inside libdsg
(one C call per segment: ``dsg_sample_loop``, include/dsg.h) instead of ~150 ATen launches per step.

Only what the sampling path needs is here: schedule tables (float64, host, once), option checking
and the hand-off to the engine: ``p_sample_loop`` (with ``const_noise`` / ``dump_steps``), ``ddim_sample_loop``
(eta = 0) and ``plms_sample_loop`` (order 2..4).  Training losses, learned-variance models and guidance hooks
are outside the hot path; the corresponding arguments raise ``NotImplementedError`` (no silent fallback).
"""
import enum
import math

import numpy as np
import torch


class ModelMeanType(enum.Enum):       # gaussian_diffusion.py:68-75
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):        # gaussian_diffusion.py:78-89
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):            # gaussian_diffusion.py:92-101
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """gaussian_diffusion.py:48-65."""
    out = np.empty(num_diffusion_timesteps, dtype=np.float64)
    for i in range(num_diffusion_timesteps):
        t1, t2 = i / num_diffusion_timesteps, (i + 1) / num_diffusion_timesteps
        out[i] = min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta)
    return out


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.):
    """gaussian_diffusion.py:21-45."""
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps,
                                   lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def _reject(**opts):
    for name, (value, allowed) in opts.items():
        if value is not allowed and value != allowed:
            raise NotImplementedError(
                f"{name}={value!r}: not implemented by the B200 engine (reference-only option; there is no "
                "PyTorch fallback on this path)")


class GaussianDiffusion:
    """Schedule tables of gaussian_diffusion.py:161-198 and the sampling entry points (:608-671, :889-1003)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False, **unused):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.timestep_map = list(range(self.num_timesteps))       # identity unless SpacedDiffusion overrides
        self._calls = 0

    # ---- fp32 coefficient rows handed to dsg_set_schedule (the `.float()` of _extract_into_tensor, :1617) ----
    def engine_tables(self, sampler):
        f32 = np.float32
        n = self.num_timesteps
        coef = np.zeros((n, 4), dtype=np.float32)
        if sampler == "ddpm":
            if self.model_var_type != ModelVarType.FIXED_SMALL:
                raise NotImplementedError("only ModelVarType.FIXED_SMALL (the reference's sigma_small=True)")
            coef[:, 0] = self.posterior_mean_coef1.astype(f32)
            coef[:, 1] = self.posterior_mean_coef2.astype(f32)
            coef[:, 2] = np.exp(f32(0.5) * self.posterior_log_variance_clipped.astype(f32))     # th.exp(0.5*logvar), :557
        else:
            abar_prev = self.alphas_cumprod_prev.astype(f32)
            coef[:, 0] = self.sqrt_recip_alphas_cumprod.astype(f32)
            coef[:, 1] = self.sqrt_recipm1_alphas_cumprod.astype(f32)
            coef[:, 2] = np.sqrt(abar_prev)
            coef[:, 3] = np.sqrt(f32(1) - abar_prev)
        qs = np.stack([self.sqrt_alphas_cumprod.astype(f32), self.sqrt_one_minus_alphas_cumprod.astype(f32)], axis=1)
        return coef, qs, np.asarray(self.timestep_map, dtype=np.int32)

    # ---- the loop ----
    def _run(self, sampler, model, shape, noise, model_kwargs, skip_timesteps, init_image, device, const_noise=False,
             dump_steps=None, order=0):
        if self.model_mean_type != ModelMeanType.START_X:
            raise NotImplementedError("only ModelMeanType.START_X (the reference always predicts x_start)")
        if not isinstance(shape, (tuple, list)):
            raise AssertionError("shape must be a tuple or list")
        if model_kwargs is None or "y" not in model_kwargs:
            raise ValueError("model_kwargs={'y': {...}} with style/seed/audio is required")
        B = int(shape[0])
        engine = model.get_engine(B)
        g = engine.g
        if tuple(shape) != (B, g.njoints, 1, g.n_poses):
            raise ValueError(f"shape {tuple(shape)} != (B, {g.njoints}, 1, {g.n_poses})")
        y = model_kwargs["y"]
        model.check_mask_local(y)
        coef, qs, tmap = self.engine_tables(sampler)
        engine.set_schedule(sampler, coef, qs, tmap)
        engine.set_conditioning(y["style"], y["seed"], y["audio"], y.get("seed_last", None))
        if noise is not None:
            x = noise.detach().to(device=engine.device, dtype=torch.float32).clone().contiguous()
            assert tuple(x.shape) == tuple(shape)
        else:
            x = torch.empty(tuple(shape), device=engine.device, dtype=torch.float32)
        seed = int(y.get("noise_seed", torch.initial_seed())) & 0xFFFFFFFFFFFFFFFF
        segment = y.get("segment", None)
        if segment is None:          # successive calls must not reuse x_T (the reference's RNG state advances)
            segment = self._calls
        self._calls += 1
        if dump_steps is not None:
            n_run = self.num_timesteps - int(skip_timesteps)
            dump_steps = [int(i) for i in dump_steps if 0 <= int(i) < n_run]      # `if i in dump_steps` (:664): others never match
        return engine.sample_loop(x, noise is not None, seed, clip_ids=y.get("clip_ids", None), segment=int(segment),
                                  skip_timesteps=int(skip_timesteps), init_image=init_image, const_noise=bool(const_noise),
                                  dump_steps=dump_steps, plms_order=int(order))

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                      randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """gaussian_diffusion.py:608-671 — same signature; returns the final sample [B, J, 1, T] on the GPU, or — as the
        reference does when ``dump_steps`` is given (:647-669) — the list of samples after those loop iterations.
        ``const_noise`` (:544-545): every clip is noised with clip 0's per-step noise."""
        _reject(clip_denoised=(clip_denoised, False), denoised_fn=(denoised_fn, None), cond_fn=(cond_fn, None),
                randomize_class=(randomize_class, False), cond_fn_with_grad=(cond_fn_with_grad, False))
        return self._run("ddpm", model, shape, noise, model_kwargs, skip_timesteps, init_image, device,
                         const_noise=const_noise, dump_steps=dump_steps)

    def plms_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, order=2):
        """gaussian_diffusion.py:1105-1134 — Pseudo Linear Multistep sampling (deterministic; eps re-derived from the
        predicted x_start, :1049).  order 2..4: order 1 dereferences ``old_out=None`` on the first step in the reference."""
        _reject(clip_denoised=(clip_denoised, False), denoised_fn=(denoised_fn, None), cond_fn=(cond_fn, None),
                randomize_class=(randomize_class, False), cond_fn_with_grad=(cond_fn_with_grad, False))
        if not int(order) or not 1 <= order <= 4:
            raise ValueError('order is invalid (should be int from 1-4).')           # :1023-1024
        if int(order) == 1:
            raise TypeError("order=1: the reference fails on its first step ('NoneType' object is not subscriptable, "
                            "gaussian_diffusion.py:1069)")
        return self._run("plms", model, shape, noise, model_kwargs, skip_timesteps, init_image, device, order=int(order))

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """gaussian_diffusion.py:889-935 — eta = 0 only (deterministic DDIM).  ``dump_steps`` / ``const_noise`` raise
        NotImplementedError exactly as in the reference (:913-916)."""
        if dump_steps is not None:
            raise NotImplementedError()
        if const_noise:
            raise NotImplementedError()
        _reject(clip_denoised=(clip_denoised, False), denoised_fn=(denoised_fn, None), cond_fn=(cond_fn, None),
                randomize_class=(randomize_class, False), cond_fn_with_grad=(cond_fn_with_grad, False), eta=(eta, 0.0))
        return self._run("ddim", model, shape, noise, model_kwargs, skip_timesteps, init_image, device)
