"""Deterministic synthetic weights and conditioning (no checkpoint ships with the reference).

Every tensor is drawn from its own generator keyed by (seed, crc32(name)), so the same
``state_dict`` can be rebuilt bit-identically on any box without shipping 36 MB of weights:
the golden-vector generator (oracle/gen_golden.py) loads it into the *reference* ``MDM``,
the tests and bench.py load it into the engine.  Scales follow torch's default
``nn.Linear`` init (uniform +-1/sqrt(fan_in)); LayerNorm gains are perturbed around 1 so
that a wrong gain/bias wiring cannot hide.
"""
import math
import zlib

import numpy as np
import torch

from .config import ModelGeometry, state_dict_spec


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def positional_encoding_table(d_model, max_len=5000):
    """``PositionalEncoding.pe`` buffer (reference main/model/mdm.py:377-384), shape [max_len, 1, d]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def synthetic_state_dict(g: ModelGeometry, seed=0, gain=1.0):
    """Reference-keyed ``state_dict`` (fp32 CPU tensors) for geometry ``g``."""
    sd = {}
    for name, shape in state_dict_spec(g):
        gen = _gen(seed, name)
        if ".norm" in name:
            if name.endswith("weight"):
                t = 1.0 + 0.1 * (2 * torch.rand(shape, generator=gen) - 1)
            else:
                t = 0.05 * (2 * torch.rand(shape, generator=gen) - 1)
        else:
            fan_in = shape[-1] if len(shape) == 2 else None
            if fan_in is None:  # bias: fan_in of the matching weight
                wname = name[:-4] + "weight" if name.endswith("bias") else name
                wshape = dict(state_dict_spec(g)).get(wname)
                fan_in = wshape[-1] if wshape is not None else shape[0]
            bound = gain / math.sqrt(fan_in)
            t = bound * (2 * torch.rand(shape, generator=gen) - 1)
        sd[name] = t.float().contiguous()
    sd["sequence_pos_encoder.pe"] = positional_encoding_table(g.latent_dim, g.pe_max_len)
    # the same buffer is registered a second time through TimestepEmbedder (mdm.py:102, 438)
    sd["embed_timestep.sequence_pos_encoder.pe"] = sd["sequence_pos_encoder.pe"]
    hd = g.latent_dim // g.local_heads
    sd["rel_pos.inv_freq"] = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
    return sd


def synthetic_conditioning(g: ModelGeometry, batch, segment=0, seed=1234, clip_offset=0):
    """Synthetic per-segment conditioning: WavLM-shaped features N(0,1),
    style one-hot ``clip mod style_in``, zero seed pose.  Returns CPU fp32 tensors keyed as ``y``."""
    audio = torch.empty(batch, g.audio_frames, g.audio_dim)
    style = torch.zeros(batch, g.style_in)
    for b in range(batch):
        clip = clip_offset + b
        gen = _gen(seed, f"audio/{clip}/{segment}")
        audio[b] = torch.randn(g.audio_frames, g.audio_dim, generator=gen)
        style[b, clip % g.style_in] = 1.0
    y = {
        "audio": audio,
        "style": style,
        "seed": torch.zeros(batch, g.njoints, 1, g.n_seed),
        "mask_local": torch.ones(1, g.n_poses, dtype=torch.bool),
    }
    return y
