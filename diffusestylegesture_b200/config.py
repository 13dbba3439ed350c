"""Model geometry for the DiffuseStyleGesture denoiser variants the engine supports.

The reference hard-codes these numbers in constructor calls rather than in the
YAML (reference main/mydiffusion_zeggs/sample.py:51-56 for ZEGGS;
BEAT-TWH-main/mydiffusion_beat_twh/sample.py:35-41 and :299-325 for the "+"
presets), so they live here as named presets.
"""
from dataclasses import dataclass, asdict

VARIANT_ZEGGS_ATTN3 = 3      # cond_mode 'cross_local_attention3_style1' (reference main/model/mdm.py:194-233)
VARIANT_BEAT_ATTN4 = 4       # cond_mode 'cross_local_attention4_style1' (BEAT-TWH-main/model/mdm.py:187-224)
VARIANT_BEAT_ATTN5 = 5       # cond_mode 'cross_local_attention5_style1' ("++", BEAT-TWH-main/model/mdm.py:226-264)


@dataclass(frozen=True)
class ModelGeometry:
    variant: int = VARIANT_ZEGGS_ATTN3
    njoints: int = 1141          # J  (njoints * nfeats, nfeats == 1)
    n_poses: int = 88            # T  frames per segment
    n_seed: int = 8              # seed frames carried between segments
    latent_dim: int = 256        # D
    ff_size: int = 1024          # F
    num_layers: int = 8          # L
    num_heads: int = 4           # global self-attention heads
    local_heads: int = 8         # MDM.num_head (reference mdm.py:58)
    local_window: int = 11       # LocalAttention window_size (mdm.py:133)
    audio_dim: int = 1024        # WavLM feature width fed to WavEncoder
    audio_latent: int = 64       # WavEncoder output (mdm.py:51, 548)
    style_in: int = 6            # style one-hot width
    style_latent: int = 64       # embed_style output (attn3); == latent_dim for attn4
    pe_max_len: int = 5000       # PositionalEncoding max_len (mdm.py:373)

    @property
    def seq_len(self):           # S = T + 1 (token prepended, mdm.py:219)
        return self.n_poses + 1

    @property
    def audio_frames(self):      # frames covered by y['audio']
        if self.variant == VARIANT_ZEGGS_ATTN3:
            return self.n_poses
        return self.n_poses - (2 if self.variant == VARIANT_BEAT_ATTN5 else 1) * self.n_seed

    def as_dict(self):
        return asdict(self)


ZEGGS = ModelGeometry()

# DiffuseStyleGesture+ presets (BEAT-TWH-main/mydiffusion_beat_twh/sample.py:307-323,
# configs/DiffuseStyleGesture.yml:4-13). local head dim = D / 8 (48 or 64); window 15.
BEAT_PLUS = ModelGeometry(variant=VARIANT_BEAT_ATTN4, njoints=2052, n_poses=150, n_seed=30,
                          latent_dim=384, local_window=15, audio_dim=1434, audio_latent=96,
                          style_in=2, style_latent=384)
TWH_PLUS = ModelGeometry(variant=VARIANT_BEAT_ATTN4, njoints=2232, n_poses=150, n_seed=30,
                         latent_dim=512, local_window=15, audio_dim=1435, audio_latent=128,
                         style_in=17, style_latent=512)

# DiffuseStyleGesture++ (cond_mode cross_local_attention5: the last n_seed frames are conditioned on y['seed_last'])
BEAT_PLUSPLUS = ModelGeometry(variant=VARIANT_BEAT_ATTN5, njoints=2052, n_poses=150, n_seed=30,
                              latent_dim=384, local_window=15, audio_dim=1434, audio_latent=96,
                              style_in=2, style_latent=384)

PRESETS = {"zeggs": ZEGGS, "beat+": BEAT_PLUS, "twh+": TWH_PLUS, "beat++": BEAT_PLUSPLUS}


def state_dict_spec(g: ModelGeometry):
    """Ordered (name, shape) list of every tensor the engine ingests.

    Names are the reference ``state_dict`` keys (SURVEY.md section 8(a) parameter
    inventory; verified against ``MDM(...).state_dict()`` in oracle/gen_golden.py).
    The order of this list IS the order of the ``weights`` pointer array passed
    to ``dsg_engine_create`` (include/dsg.h).
    """
    D, F, J, A = g.latent_dim, g.ff_size, g.njoints, g.audio_latent
    spec = [
        ("WavEncoder.audio_feature_map.weight", (A, g.audio_dim)),
        ("WavEncoder.audio_feature_map.bias", (A,)),
        ("input_process.poseEmbedding.weight", (D, J)),
        ("input_process.poseEmbedding.bias", (D,)),
        ("input_process2.weight", (D, 2 * D + A)),
        ("input_process2.bias", (D,)),
        ("embed_timestep.time_embed.0.weight", (D, D)),
        ("embed_timestep.time_embed.0.bias", (D,)),
        ("embed_timestep.time_embed.2.weight", (D, D)),
        ("embed_timestep.time_embed.2.bias", (D,)),
        ("embed_style.weight", (g.style_latent, g.style_in)),
        ("embed_style.bias", (g.style_latent,)),
    ]
    if g.variant == VARIANT_ZEGGS_ATTN3:
        spec += [("embed_text.weight", (D - g.style_latent, J * g.n_seed)),
                 ("embed_text.bias", (D - g.style_latent,))]
    else:
        spec += [("embed_text.weight", (A, J)), ("embed_text.bias", (A,))]
    spec += [("output_process.poseFinal.weight", (J, D)),
             ("output_process.poseFinal.bias", (J,))]
    for l in range(g.num_layers):
        p = f"seqTransEncoder.layers.{l}."
        spec += [
            (p + "self_attn.in_proj_weight", (3 * D, D)),
            (p + "self_attn.in_proj_bias", (3 * D,)),
            (p + "self_attn.out_proj.weight", (D, D)),
            (p + "self_attn.out_proj.bias", (D,)),
            (p + "linear1.weight", (F, D)),
            (p + "linear1.bias", (F,)),
            (p + "linear2.weight", (D, F)),
            (p + "linear2.bias", (D,)),
            (p + "norm1.weight", (D,)),
            (p + "norm1.bias", (D,)),
            (p + "norm2.weight", (D,)),
            (p + "norm2.bias", (D,)),
        ]
    if g.variant == VARIANT_BEAT_ATTN5:      # the "++" tensors come last (include/dsg.h: DSG_VARIANT_ATTN5)
        spec += [("embed_text_last.weight", (A, J)), ("embed_text_last.bias", (A,))]
    return spec
