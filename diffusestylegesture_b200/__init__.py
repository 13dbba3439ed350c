"""diffusestylegesture_b200 — B200 (sm_100a) sampling engine for the DiffuseStyleGesture hot path.

Host mirror of the reference interface for that path (``create_model_and_diffusion``, ``MDM``,
``SpacedDiffusion.p_sample_loop``, ``sample.inference``) over libdsg.so (include/dsg.h).
"""
from .config import ModelGeometry, ZEGGS, BEAT_PLUS, TWH_PLUS, PRESETS, state_dict_spec  # noqa: F401

__all__ = ["ModelGeometry", "ZEGGS", "BEAT_PLUS", "TWH_PLUS", "PRESETS", "state_dict_spec"]
