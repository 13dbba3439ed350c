"""Timestep respacing — host mirror of reference main/diffusion/respace.py (space_timesteps :8-61,
SpacedDiffusion :64-114).  The ``_WrappedModel`` index->original-timestep gather of :117-129 becomes
the ``timestep_map`` array uploaded once by ``dsg_set_schedule``."""
import numpy as np

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Which original steps to keep: "ddimN" = fixed integer stride; list/"a,b,c" = per-section counts."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                kept = range(0, num_timesteps, stride)
                if len(kept) == want:
                    return set(kept)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for sec, count in enumerate(section_counts):
        size = base + (1 if sec < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += frac
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process over a subset of the base process's steps (respace.py:64-114)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base_abar = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64))
        tmap, betas, last = [], [], 1.0
        for i, abar in enumerate(base_abar):
            if i in self.use_timesteps:
                betas.append(1 - abar / last)
                last = abar
                tmap.append(i)
        kwargs["betas"] = np.array(betas)
        super().__init__(**kwargs)
        self.timestep_map = tmap
