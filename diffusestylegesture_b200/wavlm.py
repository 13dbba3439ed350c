"""``WavLM`` — host mirror of the reference conditioning model (reference main/mydiffusion_zeggs/WavLM/WavLM.py:220-375).
A real ``torch.nn.Module`` with the reference ``state_dict`` keys (``checkpoint['model']`` of WavLM-Large.pt loads
unchanged) that holds parameters only; ``extract_features`` / ``wav2wavlm`` run in libdsg (dsg_wavlm_*, include/dsg.h).
No PyTorch compute path, no CPU fallback."""
import ctypes
import math

import numpy as np
import torch
from torch import nn

from .engine import load_library, _check, _ptr, _stream
from .wavlm_config import WAVLM_LARGE, WavLMGeometry, wavlm_state_dict_spec, synthetic_wavlm_state_dict


class WavLMConfig:
    """Attribute bag like the reference's (WavLM.py:162-217); only the Large architecture is implemented."""

    def __init__(self, cfg=None):
        from .wavlm_config import WAVLM_LARGE_CFG
        self.__dict__.update(WAVLM_LARGE_CFG)
        if cfg is not None:
            self.__dict__.update(cfg)


def _relative_position_bucket(rel, num_buckets, max_distance):
    """modules_WavLM.py:417-442 (bidirectional), same torch ops so that bucket boundaries agree bit for bit."""
    nb = num_buckets // 2
    out = (rel > 0).to(torch.long) * nb
    rel = torch.abs(rel)
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(is_small, rel, large)


class _Node(nn.Module):
    pass


class WavLM(nn.Module):
    def __init__(self, cfg=None, max_batch=16, n_samples=70400):
        super().__init__()
        cfg = cfg if isinstance(cfg, WavLMConfig) else WavLMConfig(cfg)
        g = WAVLM_LARGE
        if (cfg.encoder_layers, cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads) != \
                (g.layers, g.embed_dim, g.ffn_dim, g.heads) or cfg.extractor_mode != "layer_norm" or not cfg.layer_norm_first \
                or not cfg.gru_rel_pos or not cfg.relative_position_embedding or cfg.conv_bias:
            raise NotImplementedError("only the WavLM-Large architecture is implemented by the B200 engine")
        self.cfg, self.geometry = cfg, g
        self.max_batch, self.n_samples = int(max_batch), int(n_samples)
        init = synthetic_wavlm_state_dict(g, seed=0)
        for name, _ in wavlm_state_dict_spec(g):
            self._register(name, init[name])
        self._register("mask_emb", init["mask_emb"])
        self._h = None
        self._stale = True
        self.lib = None

    def _register(self, dotted, tensor):
        node = self
        parts = dotted.split('.')
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._stale = True
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._stale = True
        return super()._apply(fn, *a, **kw)

    def close(self):
        if self._h is not None:
            self.lib.dsg_wavlm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _engine(self, n_samples):
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("WavLM is on the CPU: the B200 engine has no CPU path — call model.to('cuda:N') first")
        if self._h is None or self._stale or n_samples != self.n_samples:
            self.close()
            lib = self.lib = load_library()
            lib.dsg_wavlm_create.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p),
                                             ctypes.c_int32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
            lib.dsg_wavlm_create.restype = ctypes.c_int
            lib.dsg_wavlm_forward.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
            lib.dsg_wavlm_forward.restype = ctypes.c_int
            lib.dsg_wavlm_frames.argtypes = [ctypes.c_void_p]
            lib.dsg_wavlm_frames.restype = ctypes.c_int32
            lib.dsg_wavlm_launch_count.argtypes = [ctypes.c_void_p]
            lib.dsg_wavlm_launch_count.restype = ctypes.c_int64
            lib.dsg_wavlm_destroy.argtypes = [ctypes.c_void_p]
            lib.dsg_wavlm_destroy.restype = None
            g = self.geometry
            self.n_samples = int(n_samples)
            L = g.frames(self.n_samples)
            sd = self.state_dict()
            spec = wavlm_state_dict_spec(g)
            keep = [sd[n].detach().to(torch.float32).contiguous() for n, _ in spec]       # device tensors are fine
            arr = (ctypes.c_void_p * len(spec))(*[t.data_ptr() for t in keep])
            # compute_bias(L, L) of layer 0 (modules_WavLM.py:444-455): [heads, L, L]
            ctx = torch.arange(L)[:, None]
            mem = torch.arange(L)[None, :]
            bucket = _relative_position_bucket(mem - ctx, g.num_buckets, g.max_distance)
            pb = sd["encoder.layers.0.self_attn.relative_attention_bias.weight"].detach().float().cpu()[bucket].permute(2, 0, 1).contiguous()
            h = ctypes.c_void_p()
            _check(lib, lib.dsg_wavlm_create(dev.index or 0, self.max_batch, self.n_samples, arr, len(spec), _ptr(pb), ctypes.byref(h)))
            self._h, self._stale, self._dev = h, False, dev
        return self._h

    @property
    def launches(self):
        return int(self.lib.dsg_wavlm_launch_count(self._h)) if self._h is not None else 0

    def _forward(self, source, n_poses):
        if source.dim() != 2:
            raise ValueError("source must be [batch, samples]")
        h = self._engine(source.shape[1])
        B = source.shape[0]
        L = int(self.lib.dsg_wavlm_frames(h))
        src = source.detach().to(torch.float32).contiguous()
        out = torch.empty(B, n_poses if n_poses > 0 else L, self.geometry.embed_dim, device=self._dev, dtype=torch.float32)
        _check(self.lib, self.lib.dsg_wavlm_forward(h, B, _ptr(src), int(n_poses), _ptr(out), _stream(self._dev)))
        self._keep = src
        return out

    def extract_features(self, source, padding_mask=None, mask=False, ret_conv=False, output_layer=None, ret_layer_results=False):
        """WavLM.extract_features (WavLM.py:323-375): returns (features [B, L, 1024], padding_mask)."""
        if padding_mask is not None or mask or ret_conv or output_layer is not None or ret_layer_results:
            raise NotImplementedError("only extract_features(source) is implemented (what wav2wavlm calls, sample.py:46)")
        return self._forward(source, 0), None

    def wav2wavlm(self, wav_input_16khz, n_poses=88):
        """Fused sample.wav2wavlm (sample.py:44-48): extract_features + linear interpolation to n_poses frames."""
        return self._forward(wav_input_16khz, n_poses)
