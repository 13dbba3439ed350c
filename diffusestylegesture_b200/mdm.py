"""``MDM`` — host mirror of the reference denoiser module (reference main/model/mdm.py:10-358; the "+"
variant BEAT-TWH-main/model/mdm.py:10-267).  It is a real ``torch.nn.Module`` whose ``state_dict`` keys and
shapes are the reference's, so ``load_model_wo_clip(model, torch.load(ckpt))`` / ``model.to(dev).eval()``
keep working — but it holds parameters only.  ``forward`` and the sampling loop execute in libdsg
(hand-written sm_100a CUDA); there is no PyTorch compute path and no CPU fallback.

``precision="bf16"`` (default) is NOT the reference's fp32 arithmetic: bf16 GEMM operands with fp32 accumulation, and in the
persistent clip kernel a bf16 residual stream, fp16 ``linear2`` operands and the tanh form of GELU in packed fp16 instead of
``F.gelu``'s erf form (mdm.py:79-86).  Stated tolerances: tests/test_gpu_tc.py, tests/test_gpu_r2.py; INTEGRATION.md.
``precision="fp32"`` follows the reference's arithmetic to ~3e-6 (validation path)."""
import torch
from torch import nn

from .config import ModelGeometry, state_dict_spec, VARIANT_ZEGGS_ATTN3, VARIANT_BEAT_ATTN4, VARIANT_BEAT_ATTN5
from .engine import Engine
from .synthetic import synthetic_state_dict


class _Node(nn.Module):
    pass


class MDM(nn.Module):
    def __init__(self, modeltype='', njoints=1141, nfeats=1, latent_dim=256, ff_size=1024, num_layers=8, num_heads=4,
                 dropout=0.1, ablation=None, activation="gelu", legacy=False, data_rep='rot6d', dataset='amass',
                 clip_dim=512, arch='trans_enc', emb_trans_dec=False, audio_feat='', n_seed=1, cond_mode='',
                 device='cpu', style_dim=-1, source_audio_dim=-1, audio_feat_dim_latent=-1,
                 n_poses=None, precision='bf16', max_batch=1, **kargs):
        super().__init__()
        if arch != 'trans_enc':
            raise NotImplementedError(f"arch={arch!r}: the engine implements 'trans_enc' only")
        if activation != 'gelu' or nfeats != 1:
            raise NotImplementedError("activation must be 'gelu' and nfeats 1")
        if audio_feat != 'wavlm':
            raise NotImplementedError(f"audio_feat={audio_feat!r}: only 'wavlm' conditioning is implemented")
        if 'cross_local_attention3' in cond_mode and 'style1' in cond_mode:
            g = ModelGeometry(variant=VARIANT_ZEGGS_ATTN3, njoints=njoints, n_poses=n_poses or 88, n_seed=n_seed,
                              latent_dim=latent_dim, ff_size=ff_size, num_layers=num_layers, num_heads=num_heads,
                              local_window=11, audio_dim=source_audio_dim if source_audio_dim > 0 else 1024,
                              audio_latent=audio_feat_dim_latent if audio_feat_dim_latent > 0 else 64,
                              style_in=style_dim if style_dim > 0 else 6, style_latent=64)
        elif 'cross_local_attention4' in cond_mode and 'style1' in cond_mode:
            g = ModelGeometry(variant=VARIANT_BEAT_ATTN4, njoints=njoints, n_poses=n_poses or 150, n_seed=n_seed,
                              latent_dim=latent_dim, ff_size=ff_size, num_layers=num_layers, num_heads=num_heads,
                              local_window=15, audio_dim=source_audio_dim, audio_latent=audio_feat_dim_latent,
                              style_in=style_dim, style_latent=latent_dim)
        elif 'cross_local_attention5' in cond_mode and 'style1' in cond_mode:
            g = ModelGeometry(variant=VARIANT_BEAT_ATTN5, njoints=njoints, n_poses=n_poses or 150, n_seed=n_seed,
                              latent_dim=latent_dim, ff_size=ff_size, num_layers=num_layers, num_heads=num_heads,
                              local_window=15, audio_dim=source_audio_dim, audio_latent=audio_feat_dim_latent,
                              style_in=style_dim, style_latent=latent_dim)
        else:
            raise NotImplementedError(f"cond_mode={cond_mode!r}: cross_local_attention3/4/5 + style1 are implemented")
        self.geometry = g
        self.njoints, self.nfeats, self.latent_dim, self.n_seed = njoints, nfeats, latent_dim, n_seed
        self.cond_mode, self.audio_feat, self.arch = cond_mode, audio_feat, arch
        self.precision, self.max_batch = precision, max_batch
        init = synthetic_state_dict(g, seed=0)
        buffers = ("sequence_pos_encoder.pe", "embed_timestep.sequence_pos_encoder.pe", "rel_pos.inv_freq")
        for name, _ in state_dict_spec(g):
            self._register(name, init[name], parameter=True)
        for name in buffers:
            self._register(name, init[name].clone(), parameter=False)
        self._engine = None
        self._engine_stale = True

    def _register(self, dotted, tensor, parameter):
        node = self
        parts = dotted.split('.')
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        if parameter:
            node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))
        else:
            node.register_buffer(parts[-1], tensor)

    # ---- weights changed -> engine must be rebuilt ------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        self._engine_stale = True
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._engine_stale = True
        return super()._apply(fn, *a, **kw)

    def parameters_wo_clip(self):
        return [p for name, p in self.named_parameters() if not name.startswith('clip_model.')]

    def get_engine(self, batch):
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("MDM is on the CPU: the B200 engine has no CPU path — call model.to('cuda:N') first")
        if self._engine is None or self._engine_stale or batch > self._engine.max_batch:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self.geometry, self.state_dict(), device=dev.index or 0,
                                  max_batch=max(int(batch), int(self.max_batch)), precision=self.precision)
            self._engine_stale = False
        return self._engine

    @staticmethod
    def check_mask_local(y):
        m = y.get('mask_local', None)
        if m is not None and not bool(torch.as_tensor(m).all()):
            raise NotImplementedError("mask_local with False entries (the reference always passes all-True, "
                                      "sample.py:230)")

    def forward(self, x, timesteps, y=None, uncond_info=False):
        """x [B, njoints, nfeats, T] (x_t), timesteps [B] int, y = {style, seed, audio, mask_local} -> predicted x_0
        (mdm.py:166-358).  One libdsg call."""
        if uncond_info:
            raise NotImplementedError("uncond_info=True (classifier-free guidance is dead code in the reference)")
        if y is None:
            raise ValueError("y is required")
        self.check_mask_local(y)
        eng = self.get_engine(x.shape[0])
        eng.set_conditioning(y['style'], y['seed'], y['audio'], y.get('seed_last', None))
        xin = x.detach().to(device=eng.device, dtype=torch.float32).contiguous()
        return eng.denoise(xin, timesteps)
