"""ctypes binding of libdsg.so (include/dsg.h).  PyTorch is used for device memory and streams only.

The library is built in-tree (``__graft_entry__.build()`` or ``python -m diffusestylegesture_b200.build``).
If it is missing, or no sm_100 device is present, every entry point raises: there is no CPU fallback.
"""
import ctypes
import os

import numpy as np
import torch

from .config import ModelGeometry, state_dict_spec

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdsg.so")

PRECISION = {"fp32": 0, "bf16": 1}
SAMPLER = {"ddpm": 0, "ddim": 1, "plms": 2}
LOOP_CONST_NOISE = 1

EXPORTS = ["dsg_engine_create", "dsg_engine_destroy", "dsg_set_schedule", "dsg_set_conditioning", "dsg_denoise",
           "dsg_posterior_step", "dsg_sample_loop", "dsg_sample_loop_ex", "dsg_set_conditioning_ex", "dsg_stitch_segment", "dsg_kernel_launch_count",
           "dsg_debug_read", "dsg_profile", "dsg_profile_read", "dsg_profile_tag_name", "dsg_selftest_gemm", "dsg_wavlm_create", "dsg_wavlm_forward", "dsg_wavlm_frames",
           "dsg_wavlm_launch_count", "dsg_wavlm_destroy", "dsg_last_error", "dsg_version"]


class _Desc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "variant", "njoints", "n_poses", "n_seed", "latent_dim", "ff_size", "num_layers", "num_heads", "local_heads",
        "local_window", "audio_dim", "audio_latent", "style_in", "style_latent", "num_timesteps", "max_batch",
        "precision", "device")]


class _LoopOpts(ctypes.Structure):
    _fields_ = [("flags", ctypes.c_int32), ("plms_order", ctypes.c_int32), ("n_dump", ctypes.c_int32),
                ("dump_iters", ctypes.c_void_p), ("dump_out", ctypes.c_void_p)]


_lib = None


def load_library(path=None):
    """dlopen libdsg.so and declare the prototypes of include/dsg.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the engine has no CPU fallback)")
    lib = ctypes.CDLL(p)
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
    lib.dsg_engine_create.argtypes = [ctypes.POINTER(_Desc), ctypes.POINTER(vp), i32, vp, ctypes.POINTER(vp)]
    lib.dsg_engine_create.restype = ctypes.c_int
    lib.dsg_engine_destroy.argtypes = [vp]
    lib.dsg_engine_destroy.restype = None
    lib.dsg_set_schedule.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.dsg_set_conditioning.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.dsg_denoise.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.dsg_posterior_step.argtypes = [vp, i32, vp, vp, i32, u64, vp, i32, i32, vp]
    lib.dsg_sample_loop.argtypes = [vp, i32, vp, i32, u64, vp, i32, i32, vp, vp]
    lib.dsg_sample_loop_ex.argtypes = [vp, i32, vp, i32, u64, vp, i32, i32, vp, ctypes.POINTER(_LoopOpts), vp]
    lib.dsg_set_conditioning_ex.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.dsg_stitch_segment.argtypes = [vp, i32, vp, vp, i32, vp]
    for f in ("dsg_set_schedule", "dsg_set_conditioning", "dsg_denoise", "dsg_posterior_step", "dsg_sample_loop",
              "dsg_sample_loop_ex", "dsg_set_conditioning_ex", "dsg_stitch_segment"):
        getattr(lib, f).restype = ctypes.c_int
    lib.dsg_kernel_launch_count.argtypes = [vp]
    lib.dsg_kernel_launch_count.restype = i64
    lib.dsg_debug_read.argtypes = [vp, ctypes.c_char_p, i32, vp, i64]
    lib.dsg_debug_read.restype = i64
    lib.dsg_profile.argtypes = [vp, i32]
    lib.dsg_profile.restype = ctypes.c_int
    lib.dsg_profile_read.argtypes = [vp, i32, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double)]
    lib.dsg_profile_read.restype = ctypes.c_int
    lib.dsg_profile_tag_name.argtypes = [i32]
    lib.dsg_profile_tag_name.restype = ctypes.c_char_p
    lib.dsg_selftest_gemm.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.dsg_selftest_gemm.restype = ctypes.c_int
    lib.dsg_last_error.restype = ctypes.c_char_p
    lib.dsg_version.restype = ctypes.c_char_p
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        msg = lib.dsg_last_error().decode()
        if rc == -3:
            raise NotImplementedError(msg)
        raise RuntimeError(f"libdsg error {rc}: {msg}")


def _ptr(t):
    """Raw address of a torch tensor / numpy array (host or device); keeps no reference."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous(), "libdsg needs contiguous buffers"
        return ctypes.c_void_p(t.data_ptr())
    assert t.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(t.ctypes.data)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def selftest_gemm(A, W, bias=None, bn=128, device=0):
    """C = bf16(A) @ bf16(W).T + bias on the tcgen05 GEMM (fp32 accumulate)."""
    lib = load_library()
    A = np.ascontiguousarray(A, dtype=np.float32)
    W = np.ascontiguousarray(W, dtype=np.float32)
    M, K = A.shape
    N = W.shape[0]
    C = np.empty((M, N), dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    _check(lib, lib.dsg_selftest_gemm(device, bn, M, N, K, _ptr(A), _ptr(W), _ptr(b), _ptr(C)))
    return C


class Engine:
    """One libdsg engine bound to one GPU.  Weights: a reference-keyed ``state_dict`` (fp32)."""

    def __init__(self, geometry: ModelGeometry, state_dict, device=0, max_batch=1, precision="fp32",
                 num_timesteps=1000):
        self.lib = load_library()
        self.g = geometry
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.max_batch = int(max_batch)
        self.precision = precision
        spec = state_dict_spec(geometry)
        keep = []
        arr = (ctypes.c_void_p * len(spec))()
        for i, (name, shape) in enumerate(spec):
            if name not in state_dict:
                raise KeyError(f"state_dict is missing '{name}'")
            t = state_dict[name].detach().to(torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected {shape}, got {tuple(t.shape)}")
            keep.append(t)
            arr[i] = t.data_ptr()
        pe = state_dict["sequence_pos_encoder.pe"].detach().to(torch.float32)[:num_timesteps, 0, :].contiguous()
        if pe.shape[0] < num_timesteps:
            raise ValueError("positional-encoding table shorter than num_timesteps")
        d = _Desc(geometry.variant, geometry.njoints, geometry.n_poses, geometry.n_seed, geometry.latent_dim,
                  geometry.ff_size, geometry.num_layers, geometry.num_heads, geometry.local_heads,
                  geometry.local_window, geometry.audio_dim, geometry.audio_latent, geometry.style_in,
                  geometry.style_latent, num_timesteps, self.max_batch, PRECISION[precision], self.device.index)
        h = ctypes.c_void_p()
        _check(self.lib, self.lib.dsg_engine_create(ctypes.byref(d), arr, len(spec), _ptr(pe), ctypes.byref(h)))
        self.h = h
        self._sched_key = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.dsg_engine_destroy(self.h)
            self.h = None

    __del__ = close

    # -- schedule -----------------------------------------------------------------------------------
    def set_schedule(self, sampler, coef, qsample, timestep_map):
        coef = np.ascontiguousarray(coef, dtype=np.float32)
        qsample = np.ascontiguousarray(qsample, dtype=np.float32)
        tmap = np.ascontiguousarray(timestep_map, dtype=np.int32)
        key = (sampler, coef.tobytes(), tmap.tobytes())
        if key == self._sched_key:
            return
        _check(self.lib, self.lib.dsg_set_schedule(self.h, SAMPLER[sampler], len(tmap), _ptr(coef), _ptr(qsample), _ptr(tmap)))
        self._sched_key = key

    # -- conditioning / denoiser ---------------------------------------------------------------------
    def set_conditioning(self, style, seed, audio, seed_last=None):
        g = self.g
        B = style.shape[0]
        style = self._f32(style).reshape(B, g.style_in)
        seed = self._f32(seed).reshape(B, g.njoints, 1, g.n_seed)
        audio = self._f32(audio).reshape(B, g.audio_frames, g.audio_dim)
        if seed_last is not None:              # the "++" variant (BEAT-TWH-main/model/mdm.py:229)
            seed_last = self._f32(seed_last).expand(B, g.njoints, 1, g.n_seed).contiguous()
        _check(self.lib, self.lib.dsg_set_conditioning_ex(self.h, B, _ptr(style), _ptr(seed), _ptr(audio), _ptr(seed_last),
                                                          _stream(self.device)))
        self._keep = (style, seed, audio, seed_last)      # keep staged host sources alive until the stream consumed them

    def denoise(self, x, timesteps, out=None):
        B = x.shape[0]
        x = self._f32(x)
        if out is None:
            out = torch.empty_like(x)
        t = np.ascontiguousarray(timesteps.detach().cpu().numpy() if isinstance(timesteps, torch.Tensor) else timesteps,
                                 dtype=np.int32)
        _check(self.lib, self.lib.dsg_denoise(self.h, B, _ptr(x), _ptr(t), _ptr(out), _stream(self.device)))
        return out

    def posterior_step(self, x, x0, index, seed, clip_ids=None, segment=0, draw=1):
        B = x.shape[0]
        ids = None if clip_ids is None else np.ascontiguousarray(clip_ids, dtype=np.int64)
        _check(self.lib, self.lib.dsg_posterior_step(self.h, B, _ptr(x), _ptr(self._f32(x0)), int(index), int(seed),
                                                     _ptr(ids), int(segment), int(draw), _stream(self.device)))
        return x

    def sample_loop(self, x, noise_given, seed, clip_ids=None, segment=0, skip_timesteps=0, init_image=None,
                    const_noise=False, dump_steps=None, plms_order=0):
        """dsg_sample_loop_ex.  Returns x (final sample, in place), or the list of dumped samples when ``dump_steps`` is
        given (gaussian_diffusion.py:647-669: the loop then returns ``dump``, not the final sample)."""
        B = x.shape[0]
        ids = None if clip_ids is None else np.ascontiguousarray(clip_ids, dtype=np.int64)
        init = None if init_image is None else self._f32(init_image)
        opts = _LoopOpts(LOOP_CONST_NOISE if const_noise else 0, int(plms_order), 0, None, None)
        iters = dump = None
        if dump_steps is not None:
            iters = np.ascontiguousarray(sorted(set(int(i) for i in dump_steps if int(i) >= 0)), dtype=np.int32)
            dump = torch.empty((len(iters),) + tuple(x.shape), dtype=torch.float32, device=x.device)
            opts.n_dump, opts.dump_iters, opts.dump_out = len(iters), iters.ctypes.data, dump.data_ptr()
        _check(self.lib, self.lib.dsg_sample_loop_ex(self.h, B, _ptr(x), int(bool(noise_given)), int(seed), _ptr(ids),
                                                     int(segment), int(skip_timesteps), _ptr(init), ctypes.byref(opts),
                                                     _stream(self.device)))
        if dump_steps is not None:
            return [dump[i] for i in range(len(iters))]
        return x

    def stitch_segment(self, prev_tail, sample, smoothing=True):
        B = sample.shape[0]
        _check(self.lib, self.lib.dsg_stitch_segment(self.h, B, _ptr(self._f32(prev_tail)), _ptr(sample), int(bool(smoothing)),
                                                     _stream(self.device)))
        return sample

    # -- introspection --------------------------------------------------------------------------------
    @property
    def launches(self):
        return int(self.lib.dsg_kernel_launch_count(self.h))

    def profile(self, enable):
        _check(self.lib, self.lib.dsg_profile(self.h, int(bool(enable))))

    def profile_read(self):
        """{kernel class: (launches, total device ms)} accumulated since profile(True)."""
        out, tag = {}, 0
        while True:
            name = self.lib.dsg_profile_tag_name(tag)
            if name is None:
                break
            n, ms = ctypes.c_int64(), ctypes.c_double()
            _check(self.lib, self.lib.dsg_profile_read(self.h, tag, ctypes.byref(n), ctypes.byref(ms)))
            if n.value:
                out[name.decode()] = (int(n.value), float(ms.value))
            tag += 1
        return out

    def debug_enable(self):
        r = self.lib.dsg_debug_read(self.h, b"enable", 1, None, 0)
        if r < 0:
            _check(self.lib, int(r))

    def clip_profile(self):
        """Cycle counters of the persistent clip kernel's CTA 0 (set DSG_CLIP_PROF=1 before the loop call)."""
        names = ["total", "mma_wait_weights", "mma_wait_other", "producer_wait_empty", "w_stage_x", "w_in_wait", "w_in_epilogue",
                 "w_local_attention", "w_qkv_wait", "w_extract_attention", "w_ln_wait", "w_layernorm", "w_gelu_wait", "w_gelu",
                 "w_head_wait", "w_head_posterior", "w_noise_wait", "w_att_extract", "w_att_sync1", "w_att_mma", "w_att_merge",
                 "mma_ffn_total", "mma_ffn_wait_weights", "mma_ffn_wait_workers", "mma_qkv_total", "mma_qkv_wait_weights",
                 "mma_qkv_wait_workers"]
        out = torch.empty(32, dtype=torch.float32)
        r = self.lib.dsg_debug_read(self.h, b"clipprof", 1, _ptr(out), 32)
        if r < 0:
            _check(self.lib, int(r))
        return {n: float(out[i]) for i, n in enumerate(names)}

    def debug_read(self, name, batch):
        g = self.g
        shape = {"h_in": (batch, g.n_poses, g.latent_dim), "tok": (batch, g.latent_dim)}.get(
            name, (batch, g.seq_len, g.latent_dim))
        out = torch.empty(shape, dtype=torch.float32)
        r = self.lib.dsg_debug_read(self.h, name.encode(), batch, _ptr(out), out.numel())
        if r < 0:
            _check(self.lib, int(r))
        return out

    @staticmethod
    def _f32(t):
        if isinstance(t, torch.Tensor):
            return t.detach().to(torch.float32).contiguous()
        return np.ascontiguousarray(t, dtype=np.float32)
