"""Factories — host mirror of reference main/utils/model_util.py (load_model_wo_clip :8-12,
create_gaussian_diffusion :59-100)."""
from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps


def load_model_wo_clip(model, state_dict):
    missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
    assert len(unexpected_keys) == 0, unexpected_keys
    assert all(k.startswith('clip_model.') for k in missing_keys), missing_keys


def create_gaussian_diffusion(timestep_respacing=''):
    """Cosine schedule, 1000 steps, x_start prediction, fixed-small variance — hard-coded by the reference.
    ``timestep_respacing`` (reference: always '') is exposed so DDIM-100 / 50-step configs share the factory."""
    steps = 1000
    betas = gd.get_named_beta_schedule('cosine', steps, 1.)
    if not timestep_respacing:
        timestep_respacing = [steps]
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing), betas=betas,
                           model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                           loss_type=gd.LossType.MSE, rescale_timesteps=False)
