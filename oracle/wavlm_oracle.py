"""ORACLE — test infrastructure only (see oracle/dsg_oracle.py header).  CPU restatement of the WavLM-Large conditioning
forward used by the ZEGGS path: ``wav2wavlm`` = ``WavLM.extract_features`` + linear interpolation to n_poses frames
(reference main/mydiffusion_zeggs/sample.py:44-48; WavLM/WavLM.py:323-375, 378-504, 507-742; modules_WavLM.py:303-563).
Pinned by oracle/gen_golden_wavlm.py against the reference ``WavLM`` class (public Large hyper-parameters, synthetic
weights: the real checkpoint is an external download).
"""
import math

import torch
import torch.nn.functional as F


def relative_position_bucket(rel, num_buckets=320, max_distance=800):
    """MultiheadAttention._relative_positions_bucket, bidirectional (modules_WavLM.py:417-442)."""
    nb = num_buckets // 2
    out = (rel > 0).to(torch.long) * nb
    rel = rel.abs()
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(is_small, rel, large)


def position_bias(W, L, g):
    """compute_bias (modules_WavLM.py:444-455): [heads, L, L]."""
    ctx = torch.arange(L)[:, None]
    mem = torch.arange(L)[None, :]
    bucket = relative_position_bucket(mem - ctx, g.num_buckets, g.max_distance)
    return W["encoder.layers.0.self_attn.relative_attention_bias.weight"][bucket].permute(2, 0, 1)


def conv_features(W, g, wav, taps=None):
    """ConvFeatureExtractionModel, mode 'layer_norm' (WavLM.py:391-422, 485-504): [B, N] -> [B, L, 512]."""
    x = wav.unsqueeze(1)
    for i, (c, k, s) in enumerate(g.conv_layers):
        x = F.conv1d(x, W[f"feature_extractor.conv_layers.{i}.0.weight"], stride=s)
        x = F.layer_norm(x.transpose(1, 2), (c,), W[f"feature_extractor.conv_layers.{i}.2.1.weight"],
                         W[f"feature_extractor.conv_layers.{i}.2.1.bias"], 1e-5).transpose(1, 2)
        x = F.gelu(x)
        if taps is not None:
            taps[f"conv{i}"] = x.transpose(1, 2)
    return x.transpose(1, 2)


def extract_features(W, g, wav, taps=None):
    """WavLM.extract_features(source)[0] (WavLM.py:323-375) for mask=False, no padding mask: [B, N] -> [B, L, 1024]."""
    E, H = g.embed_dim, g.heads
    hd = E // H
    feats = conv_features(W, g, wav, taps)
    feats = F.layer_norm(feats, (feats.shape[-1],), W["layer_norm.weight"], W["layer_norm.bias"], 1e-5)
    x = F.linear(feats, W["post_extract_proj.weight"], W["post_extract_proj.bias"])
    # TransformerEncoder.extract_features (WavLM.py:570-612): weight-normed grouped positional conv + SamePad + GELU
    v, gw = W["encoder.pos_conv.0.weight_v"], W["encoder.pos_conv.0.weight_g"]
    w = gw * v / v.norm(p=2, dim=(0, 1), keepdim=True)                       # nn.utils.weight_norm(dim=2)
    xc = F.conv1d(x.transpose(1, 2), w, W["encoder.pos_conv.0.bias"], padding=g.conv_pos // 2, groups=g.conv_pos_groups)
    xc = F.gelu(xc[:, :, :-1])                                                 # SamePad for an even kernel
    x = x + xc.transpose(1, 2)
    if taps is not None:
        taps["x_pos"] = x
    B, L, _ = x.shape
    pb = position_bias(W, L, g)                                                # layer 0 only; reused by every layer
    for l in range(g.layers):
        p = f"encoder.layers.{l}."
        res = x
        h = F.layer_norm(x, (E,), W[p + "self_attn_layer_norm.weight"], W[p + "self_attn_layer_norm.bias"], 1e-5)
        # gated relative position bias (modules_WavLM.py:517-535): gates come from the layer INPUT split into heads
        ql = h.view(B, L, H, hd).permute(0, 2, 1, 3)
        gg = torch.sigmoid(F.linear(ql, W[p + "self_attn.grep_linear.weight"], W[p + "self_attn.grep_linear.bias"])
                           .view(B, H, L, 2, 4).sum(-1))
        gate_a, gate_b = gg.chunk(2, dim=-1)
        gate = gate_a * (gate_b * W[p + "self_attn.grep_a"] - 1.0) + 2.0       # [B,H,L,1]
        mask = gate * pb[None]                                                 # [B,H,L,L]
        q = F.linear(h, W[p + "self_attn.q_proj.weight"], W[p + "self_attn.q_proj.bias"]).view(B, L, H, hd).transpose(1, 2)
        k = F.linear(h, W[p + "self_attn.k_proj.weight"], W[p + "self_attn.k_proj.bias"]).view(B, L, H, hd).transpose(1, 2)
        vv = F.linear(h, W[p + "self_attn.v_proj.weight"], W[p + "self_attn.v_proj.bias"]).view(B, L, H, hd).transpose(1, 2)
        att = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2) + mask, dim=-1) @ vv
        att = att.transpose(1, 2).reshape(B, L, E)
        x = res + F.linear(att, W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
        res = x
        h = F.layer_norm(x, (E,), W[p + "final_layer_norm.weight"], W[p + "final_layer_norm.bias"], 1e-5)
        x = res + F.linear(F.gelu(F.linear(h, W[p + "fc1.weight"], W[p + "fc1.bias"])), W[p + "fc2.weight"], W[p + "fc2.bias"])
        if taps is not None and l in (0, 11, 23):
            taps[f"layer{l}"] = x
    return F.layer_norm(x, (E,), W["encoder.layer_norm.weight"], W["encoder.layer_norm.bias"], 1e-5)


def wav2wavlm(W, g, wav, n_poses=88, taps=None):
    """sample.py:44-48 (ZEGGS: the waveform is NOT layer-normalised)."""
    rep = extract_features(W, g, wav, taps)
    return F.interpolate(rep.transpose(1, 2), size=n_poses, align_corners=True, mode='linear').transpose(1, 2)
