"""WavLM-Large geometry and parameter inventory (the conditioning front-end of the ZEGGS path).

The hyper-parameters are NOT in the reference tree — they ship inside the external checkpoint's ``cfg`` dict
(reference main/mydiffusion_zeggs/sample.py:35-36); these are the public WavLM-Large values.
Names below are the reference ``state_dict`` keys of ``WavLM`` (WavLM/WavLM.py:220-318, 378-612; modules_WavLM.py:303-455),
so a real ``WavLM-Large.pt['model']`` loads unchanged.
"""
import math
import zlib
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class WavLMGeometry:
    conv_layers: tuple = ((512, 10, 5), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 2, 2), (512, 2, 2))
    embed_dim: int = 1024
    ffn_dim: int = 4096
    heads: int = 16
    layers: int = 24
    conv_pos: int = 128
    conv_pos_groups: int = 16
    num_buckets: int = 320
    max_distance: int = 800

    def frames(self, n_samples):
        n = n_samples
        for _, k, s in self.conv_layers:
            n = (n - k) // s + 1
        return n


WAVLM_LARGE = WavLMGeometry()

WAVLM_LARGE_CFG = dict(extractor_mode='layer_norm', encoder_layers=24, encoder_embed_dim=1024, encoder_ffn_embed_dim=4096,
                       encoder_attention_heads=16, layer_norm_first=True, conv_bias=False, normalize=True, conv_pos=128,
                       conv_pos_groups=16, relative_position_embedding=True, num_buckets=320, max_distance=800,
                       gru_rel_pos=True, activation_fn='gelu',
                       conv_feature_layers='[(512,10,5)] + [(512,3,2)] * 4 + [(512,2,2)] * 2', dropout=0.0,
                       attention_dropout=0.0, activation_dropout=0.0, encoder_layerdrop=0.0, dropout_input=0.0,
                       dropout_features=0.0)


def wavlm_state_dict_spec(g: WavLMGeometry = WAVLM_LARGE):
    """Ordered (name, shape) list = order of the ``weights`` pointer array of ``dsg_wavlm_create`` (include/dsg.h)."""
    E, Fd, H = g.embed_dim, g.ffn_dim, g.heads
    spec, cin = [], 1
    for i, (c, k, _) in enumerate(g.conv_layers):
        spec += [(f"feature_extractor.conv_layers.{i}.0.weight", (c, cin, k)),
                 (f"feature_extractor.conv_layers.{i}.2.1.weight", (c,)),
                 (f"feature_extractor.conv_layers.{i}.2.1.bias", (c,))]
        cin = c
    spec += [("layer_norm.weight", (cin,)), ("layer_norm.bias", (cin,)),
             ("post_extract_proj.weight", (E, cin)), ("post_extract_proj.bias", (E,)),
             ("encoder.pos_conv.0.bias", (E,)), ("encoder.pos_conv.0.weight_g", (1, 1, g.conv_pos)),
             ("encoder.pos_conv.0.weight_v", (E, E // g.conv_pos_groups, g.conv_pos)),
             ("encoder.layers.0.self_attn.relative_attention_bias.weight", (g.num_buckets, H))]
    for l in range(g.layers):
        p = f"encoder.layers.{l}."
        spec += [(p + "self_attn.q_proj.weight", (E, E)), (p + "self_attn.q_proj.bias", (E,)),
                 (p + "self_attn.k_proj.weight", (E, E)), (p + "self_attn.k_proj.bias", (E,)),
                 (p + "self_attn.v_proj.weight", (E, E)), (p + "self_attn.v_proj.bias", (E,)),
                 (p + "self_attn.out_proj.weight", (E, E)), (p + "self_attn.out_proj.bias", (E,)),
                 (p + "self_attn.grep_linear.weight", (8, E // H)), (p + "self_attn.grep_linear.bias", (8,)),
                 (p + "self_attn.grep_a", (1, H, 1, 1)),
                 (p + "self_attn_layer_norm.weight", (E,)), (p + "self_attn_layer_norm.bias", (E,)),
                 (p + "fc1.weight", (Fd, E)), (p + "fc1.bias", (Fd,)),
                 (p + "fc2.weight", (E, Fd)), (p + "fc2.bias", (E,)),
                 (p + "final_layer_norm.weight", (E,)), (p + "final_layer_norm.bias", (E,))]
    spec += [("encoder.layer_norm.weight", (E,)), ("encoder.layer_norm.bias", (E,))]
    return spec


def synthetic_wavlm_state_dict(g: WavLMGeometry = WAVLM_LARGE, seed=0):
    """Deterministic synthetic WavLM-Large weights (315 M parameters, ~1.26 GB fp32): per-tensor seeded generators, scales
    of the reference's initialisers, LayerNorm gains perturbed around 1."""
    sd = {}
    for name, shape in wavlm_state_dict_spec(g):
        gen = torch.Generator(device="cpu")
        gen.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        u = lambda: 2 * torch.rand(shape, generator=gen) - 1
        if "layer_norm" in name or ".2.1." in name:
            t = 1.0 + 0.1 * u() if name.endswith("weight") else 0.05 * u()
        elif name.endswith("weight_g"):
            t = 1.4 * (1.0 + 0.1 * u())
        elif name.endswith("weight_v"):
            t = math.sqrt(3.0) * math.sqrt(4.0 / (g.conv_pos * g.embed_dim)) * u()
        elif name.endswith("grep_a"):
            t = 1.0 + 0.2 * u()
        elif "relative_attention_bias" in name:
            t = 0.5 * u()
        elif "conv_layers" in name:
            fan_in = shape[1] * shape[2]
            t = math.sqrt(6.0 / fan_in) * u()
        else:
            fan_in = shape[-1] if len(shape) == 2 else None
            if fan_in is None:
                wshape = dict(wavlm_state_dict_spec(g)).get(name[:-4] + "weight")
                fan_in = wshape[-1] if wshape is not None else 1024
            t = u() / math.sqrt(fan_in)
        sd[name] = t.float().contiguous()
    sd["mask_emb"] = torch.zeros(g.embed_dim)
    return sd


def synthetic_wav(batch, n_samples, seed=1234, clip_offset=0):
    """N(0, 0.1^2) waveform per clip."""
    out = torch.empty(batch, n_samples)
    for b in range(batch):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(int(seed) * 7919 + clip_offset + b)
        out[b] = 0.1 * torch.randn(n_samples, generator=gen)
    return out
