"""ORACLE — test infrastructure only.  NOT part of the product path.

CPU restatement (torch fp32/fp64 + numpy) of the DiffuseStyleGesture sampling hot path:
schedule tables, the MDM denoiser forward, the DDPM/DDIM posterior step, the sampling loop,
the segment driver of ``sample.py`` and the numeric part of ``pose2bvh``.  Every function
cites the reference file:line it follows (paths relative to /root/reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product (``diffusestylegesture_b200``) must
never import it and has no CPU fallback.

Pinning status: the reference ships no tests or golden vectors for this path ("parity
unpinned" by the reference itself).  This oracle is therefore pinned
against OUTPUTS OF THE REFERENCE ITSELF, run in the authoring container by
``oracle/gen_golden.py`` (which imports /root/reference unmodified) and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against those vectors.

Noise: torch's CPU (mt19937) and CUDA (Philox) generators already disagree in the reference,
so parity uses one counter-based stream defined here and implemented identically in the CUDA
engine: Philox4x32-10, key = 64-bit seed, counter = (element/4, draw, clip, segment),
Box-Muller on the 4 outputs (see ``philox_normal``).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Philox4x32-10 counter-based normal stream (shared definition with csrc/dsg_rng.cuh)
# --------------------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  Inputs: uint32 arrays/scalars."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def _u01(r):
    """uint32 -> float32 in (0,1): top 24 bits, centred (exact in fp32)."""
    return ((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24) + np.float32(2.0 ** -25))


def philox_normal(seed, clip, segment, draw, n):
    """n standard normals (float32) for one clip tensor, element e -> counter (e//4, draw, clip, segment),
    lane e%4.  Lanes (0,1) = Box-Muller(cos,sin) of outputs (0,1); lanes (2,3) of outputs (2,3)."""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint32)
    r = philox4x32_10(q, np.uint32(draw), np.uint32(clip & 0xFFFFFFFF), np.uint32(segment),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    two_pi = np.float32(6.283185307179586)
    out = np.empty((nq, 4), dtype=np.float32)
    for p in range(2):
        u1, u2 = _u01(r[2 * p]), _u01(r[2 * p + 1])
        rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
        ang = (two_pi * u2).astype(np.float32)
        out[:, 2 * p] = rad * np.cos(ang)
        out[:, 2 * p + 1] = rad * np.sin(ang)
    return out.reshape(-1)[:n]


def noise_tensor(seed, clip_ids, segment, draw, shape_per_clip):
    """[B, *shape_per_clip] float32 torch tensor of stream normals."""
    n = int(np.prod(shape_per_clip))
    arr = np.stack([philox_normal(seed, int(c), segment, draw, n) for c in clip_ids])
    return torch.from_numpy(arr.reshape((len(clip_ids),) + tuple(shape_per_clip)))


# --------------------------------------------------------------------------------------
# Schedule (main/diffusion/gaussian_diffusion.py:21-65, 161-198; main/diffusion/respace.py:8-87)
# --------------------------------------------------------------------------------------
def cosine_betas(num_steps, max_beta=0.999):
    """get_named_beta_schedule('cosine') -> betas_for_alpha_bar (gaussian_diffusion.py:39-65)."""
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - ab((i + 1) / num_steps) / ab(i / num_steps), max_beta)
                     for i in range(num_steps)], dtype=np.float64)


def space_timesteps(num_timesteps, section_counts):
    """respace.py:8-61."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[4:])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired:
                    return set(range(0, num_timesteps, i))
            raise ValueError("cannot create exactly %d steps with an integer stride" % num_timesteps)
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, cnt in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < cnt:
            raise ValueError("cannot divide section of %d steps into %d" % (size, cnt))
        stride = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        cur = 0.0
        for _ in range(cnt):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


class Schedule:
    """float64 tables of GaussianDiffusion.__init__ (gaussian_diffusion.py:161-198) after
    SpacedDiffusion's beta re-derivation (respace.py:73-87)."""

    def __init__(self, num_steps=1000, respacing=None):
        base = cosine_betas(num_steps)
        use = space_timesteps(num_steps, respacing if respacing else [num_steps])
        ac = np.cumprod(1.0 - base)
        last, betas, tmap = 1.0, [], []
        for i, a in enumerate(ac):
            if i in use:
                betas.append(1 - a / last)
                last = a
                tmap.append(i)
        betas = np.array(betas, dtype=np.float64)
        self.timestep_map = tmap
        self.num_timesteps = len(betas)
        self.betas = betas
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)


# --------------------------------------------------------------------------------------
# MDM denoiser forward (main/model/mdm.py:166-233,357; BEAT-TWH-main/model/mdm.py:187-224)
# --------------------------------------------------------------------------------------
def _rope(x, n_heads):
    """apply_rotary_pos_emb on the [B, n, D] hidden state viewed as n_heads slices
    (rotary.py:6-25; mdm.py:207-212, 221-229): z*cos + cat(-z[half:], z[:half])*sin, pos = row."""
    B, n, D = x.shape
    hd = D // n_heads
    inv_freq = 1.0 / (10000 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    t = torch.arange(n, dtype=torch.float32)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    freqs = torch.cat((freqs, freqs), dim=-1).to(x.dtype)          # [n, hd]
    z = x.view(B, n, n_heads, hd)
    z1, z2 = z[..., : hd // 2], z[..., hd // 2:]
    rot = torch.cat((-z2, z1), dim=-1)
    out = z * freqs.cos()[None, :, None, :] + rot * freqs.sin()[None, :, None, :]
    return out.reshape(B, n, D)


def _local_attention(h, n_heads, window):
    """LocalAttention.forward with q=k=v=h, causal, look_backward=1, mask all-True
    (local_attention.py:91-199).  h: [B, T, D] -> [B, T, D]."""
    B, T, D = h.shape
    hd = D // n_heads
    z = h.view(B, T, n_heads, hd).permute(0, 2, 1, 3)                # [B,H,T,hd]
    sim = torch.einsum("bhie,bhje->bhij", z, z) * (hd ** -0.5)
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    lo = (i // window - 1) * window
    allowed = (j <= i) & (j >= lo)
    sim = sim.masked_fill(~allowed, -torch.finfo(sim.dtype).max)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhje->bhie", attn, z)
    return out.permute(0, 2, 1, 3).reshape(B, T, D)


def _encoder_layer(xs, W, p, n_heads):
    """nn.TransformerEncoderLayer, post-norm, gelu(erf), eval (mdm.py:79-86): xs [B,S,D]."""
    B, S, D = xs.shape
    hd = D // n_heads
    qkv = F.linear(xs, W[p + "self_attn.in_proj_weight"], W[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, n_heads, hd).transpose(1, 2)
    k = k.view(B, S, n_heads, hd).transpose(1, 2)
    v = v.view(B, S, n_heads, hd).transpose(1, 2)
    att = torch.softmax((q * (hd ** -0.5)) @ k.transpose(-1, -2), dim=-1) @ v
    att = att.transpose(1, 2).reshape(B, S, D)
    att = F.linear(att, W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
    xs = F.layer_norm(xs + att, (D,), W[p + "norm1.weight"], W[p + "norm1.bias"], 1e-5)
    ff = F.linear(F.gelu(F.linear(xs, W[p + "linear1.weight"], W[p + "linear1.bias"])),
                  W[p + "linear2.weight"], W[p + "linear2.bias"])
    return F.layer_norm(xs + ff, (D,), W[p + "norm2.weight"], W[p + "norm2.bias"], 1e-5)


def mdm_forward(W, g, x, t, y, taps=None):
    """x [B,J,1,T], t [B] (ORIGINAL timestep ids, after _WrappedModel mapping), y dict -> [B,J,1,T].
    ``taps`` (dict) collects named intermediates for per-op golden checks."""
    B, J, _, T = x.shape
    D = g.latent_dim
    dt = x.dtype
    W = {k: v.to(dt) if v.is_floating_point() else v for k, v in W.items()}
    pe = W["sequence_pos_encoder.pe"][:, 0, :]                     # [5000, D]
    # TimestepEmbedder (mdm.py:447-448)
    emb_t = F.linear(F.silu(F.linear(pe[t], W["embed_timestep.time_embed.0.weight"],
                                     W["embed_timestep.time_embed.0.bias"])),
                     W["embed_timestep.time_embed.2.weight"], W["embed_timestep.time_embed.2.bias"])
    style = F.linear(y["style"].to(dt), W["embed_style.weight"], W["embed_style.bias"])
    seed = y["seed"].to(dt).squeeze(2)                              # [B,J,n_seed]
    audio = F.linear(y["audio"].to(dt), W["WavEncoder.audio_feature_map.weight"],
                     W["WavEncoder.audio_feature_map.bias"])       # [B,Ta,A]
    if g.variant == 3:      # mdm.py:180-183, 190
        emb_1 = torch.cat((style, F.linear(seed.reshape(B, -1), W["embed_text.weight"], W["embed_text.bias"])), 1)
        enc = audio                                                 # [B,T,A]
    else:                   # BEAT-TWH-main/model/mdm.py:187-190, 198
        emb_1 = style
        et = F.linear(seed.permute(0, 2, 1), W["embed_text.weight"], W["embed_text.bias"])   # [B,n_seed,A]
        enc = torch.cat((et, audio), dim=1)                         # [B,T,A]
        if g.variant == 5:  # "++": BEAT-TWH-main/model/mdm.py:226-230 — audio covers T - 2 n_seed frames
            last = y["seed_last"].to(dt).squeeze(2)
            enc = torch.cat((enc, F.linear(last.permute(0, 2, 1), W["embed_text_last.weight"], W["embed_text_last.bias"])), dim=1)
    tok = emb_1 + emb_t                                             # [B,D]
    # InputProcess (mdm.py:461-467) + input_process2 on cat[tok | x_ | enc] (mdm.py:202-206)
    x_ = F.linear(x.squeeze(2).permute(0, 2, 1), W["input_process.poseEmbedding.weight"],
                  W["input_process.poseEmbedding.bias"])           # [B,T,D]
    cat = torch.cat((tok[:, None, :].expand(B, T, D), x_, enc), dim=-1)
    h = F.linear(cat, W["input_process2.weight"], W["input_process2.bias"])   # [B,T,D]
    if taps is not None:
        taps["emb_t"], taps["tok"], taps["h_in"] = emb_t, tok, h
    h = _rope(h, g.local_heads)
    if taps is not None:
        taps["h_rope"] = h
    h = _local_attention(h, g.local_heads, g.local_window)
    if taps is not None:
        taps["h_local"] = h
    xs = _rope(torch.cat((tok[:, None, :], h), dim=1), g.local_heads)          # [B,S,D]
    if taps is not None:
        taps["xs0"] = xs
    for l in range(g.num_layers):
        xs = _encoder_layer(xs, W, f"seqTransEncoder.layers.{l}.", g.num_heads)
        if taps is not None:
            taps[f"xs{l + 1}"] = xs
    out = F.linear(xs[:, 1:], W["output_process.poseFinal.weight"], W["output_process.poseFinal.bias"])
    return out.permute(0, 2, 1).unsqueeze(2).contiguous()           # [B,J,1,T]


# --------------------------------------------------------------------------------------
# Sampling loop (gaussian_diffusion.py:506-558, 608-740, 742-792, 889-1003)
# --------------------------------------------------------------------------------------
def p_sample_loop(W, g, sched, y, batch, *, seed=123456, clip_ids=None, segment=0, sampler="ddpm",
                  skip_timesteps=0, init_image=None, noise=None, dtype=torch.float32,
                  record_every=0, model_fn=None, const_noise=False, dump_steps=None, order=2):
    """Returns (final sample [B,J,1,T], list of (loop_index, x) snapshots).

    Draw numbering of the shared noise stream: draw 0 = x_T (``th.randn(*shape)``,
    gaussian_diffusion.py:704); draw 1+k = ``randn_like`` of loop iteration k (:542), k = 0 for the
    first (noisiest) step.  The reference also draws (and discards) noise at t == 0.
    ``const_noise`` (:544-545): every clip takes clip 0's step noise.  ``dump_steps`` (:647-669): the second return
    value becomes the list of samples after the listed loop iterations.  ``sampler="plms"``: plms_sample, :1005-1103.
    """
    clip_ids = list(range(batch)) if clip_ids is None else list(clip_ids)
    shp = (g.njoints, 1, g.n_poses)
    f = lambda a, i: torch.tensor(float(np.float32(a[i])), dtype=dtype)      # .float() of a float64 table entry
    img = noise.to(dtype) if noise is not None else noise_tensor(seed, clip_ids, segment, 0, shp).to(dtype)
    n = sched.num_timesteps
    indices = list(range(n - skip_timesteps))[::-1]
    if skip_timesteps and init_image is None:
        init_image = torch.zeros_like(img)
    if init_image is not None:     # q_sample (gaussian_diffusion.py:236-254, 706-713)
        i0 = indices[0]
        img = f(sched.sqrt_alphas_cumprod, i0) * init_image.to(dtype) + f(sched.sqrt_one_minus_alphas_cumprod, i0) * img
    snaps = []
    model_fn = model_fn or (lambda x, t: mdm_forward(W, g, x, t, y))
    for k, i in enumerate(indices):
        t_orig = torch.full((batch,), sched.timestep_map[i], dtype=torch.long)
        x0 = model_fn(img, t_orig)
        z = noise_tensor(seed, [clip_ids[0]] * batch if const_noise else clip_ids, segment, 1 + k, shp).to(dtype)
        nz = 0.0 if i == 0 else 1.0
        if sampler == "ddpm":      # p_sample + q_posterior_mean_variance, FIXED_SMALL (:264-271, :349-362, :557)
            mean = f(sched.posterior_mean_coef1, i) * x0 + f(sched.posterior_mean_coef2, i) * img
            img = mean + nz * torch.exp(0.5 * f(sched.posterior_log_variance_clipped, i)) * z
        elif sampler == "ddim":    # ddim_sample, eta = 0 (:742-792); eps from x0 (:417-421)
            eps = (f(sched.sqrt_recip_alphas_cumprod, i) * img - x0) / f(sched.sqrt_recipm1_alphas_cumprod, i)
            abar, abar_prev = f(sched.alphas_cumprod, i), f(sched.alphas_cumprod_prev, i)
            sigma = 0.0 * torch.sqrt((1 - abar_prev) / (1 - abar)) * torch.sqrt(1 - abar / abar_prev)
            img = x0 * torch.sqrt(abar_prev) + torch.sqrt(1 - abar_prev - sigma ** 2) * eps + nz * sigma * z
        elif sampler == "plms":    # plms_sample (:1005-1103): eps history + Adams-Bashforth, pseudo improved Euler first
            eps_of = lambda xt, x0_, j: (f(sched.sqrt_recip_alphas_cumprod, j) * xt - x0_) / f(sched.sqrt_recipm1_alphas_cumprod, j)
            sq_prev, sq_1m = torch.sqrt(f(sched.alphas_cumprod_prev, i)), torch.sqrt(1 - f(sched.alphas_cumprod_prev, i))
            eps = eps_of(img, x0, i)
            if k == 0:
                assert order > 1, "order 1 fails inside the reference on the first step (old_out is None, :1069)"
                old_eps = [eps]
                mean_pred = x0 * sq_prev + sq_1m * eps
                x0_2 = model_fn(mean_pred, torch.full((batch,), sched.timestep_map[i - 1], dtype=torch.long))
                eps_prime = (eps + eps_of(mean_pred, x0_2, i - 1)) / 2
            else:
                old_eps.append(eps)
                cur = min(order, len(old_eps))
                if cur == 1:
                    eps_prime = old_eps[-1]
                elif cur == 2:
                    eps_prime = (3 * old_eps[-1] - old_eps[-2]) / 2
                elif cur == 3:
                    eps_prime = (23 * old_eps[-1] - 16 * old_eps[-2] + 5 * old_eps[-3]) / 12
                else:
                    eps_prime = (55 * old_eps[-1] - 59 * old_eps[-2] + 37 * old_eps[-3] - 9 * old_eps[-4]) / 24
            pred_prime = f(sched.sqrt_recip_alphas_cumprod, i) * img - f(sched.sqrt_recipm1_alphas_cumprod, i) * eps_prime
            mean_pred = pred_prime * sq_prev + sq_1m * eps_prime
            if len(old_eps) >= order:
                old_eps.pop(0)
            img = mean_pred * nz + x0 * (1 - nz)
        else:
            raise ValueError(sampler)
        if dump_steps is not None and k in dump_steps:
            snaps.append(img.clone())
        elif dump_steps is None and record_every and (k % record_every == record_every - 1 or k == len(indices) - 1):
            snaps.append((k, img.clone()))
    return img, snaps


def inference_clip_beat(W, g, sched, textaudio, style, seed_gesture, *, seed=123456, clip_id=0, skip_timesteps=0,
                        dtype=torch.float32, seed_last=None, division=3):
    """BEAT-TWH `inference` (BEAT-TWH-main/mydiffusion_beat_twh/sample.py:44-201) for one clip, "+" / "++" variants.
    textaudio [n, audio_dim] features; seed_gesture [n_seed, J] = the velocity/acceleration seed of :118-136 (already
    normalised and concatenated); returns the normalised motion [real_n, J // division] (:176-193, before de-normalising)."""
    real_n = textaudio.shape[0]
    stride = g.n_poses - g.n_seed
    nsub = 1 if real_n < stride else math.ceil(real_n / stride)                      # :56-62
    n_frames = nsub * stride
    ta = torch.cat((textaudio.to(dtype), torch.zeros(n_frames - real_n, textaudio.shape[1], dtype=dtype)), 0)     # :71-72
    audio = ta.reshape(nsub, stride, -1)                                              # :73 (segment-major here)
    outs = []
    seed_pose = seed_gesture.to(dtype).T[None, :, None, :]                            # [1,J,1,n_seed]  (:135-136)
    for i in range(nsub):
        a = audio[i]
        if g.variant == 5:
            a = a[:-g.n_seed]                                                         # :106, :145
        y = {"audio": a[None], "style": style[None].to(dtype), "seed": seed_pose}
        if g.variant == 5:
            y["seed_last"] = seed_last
        sample, _ = p_sample_loop(W, g, sched, y, 1, seed=seed, clip_ids=[clip_id], segment=i,
                                  skip_timesteps=skip_timesteps, dtype=dtype)
        if outs and g.n_seed:                                                         # :163-180 (no root shift in this variant)
            tail = outs[-1][..., -g.n_seed:]
            outs[-1] = outs[-1][..., :-g.n_seed]
            sample = stitch_segment(tail, sample, smoothing=False)
        outs.append(sample)
        seed_pose = sample[..., -g.n_seed:]                                           # :147
    Jd = g.njoints // division
    # :189-199: all but the last segment were trimmed by n_seed; the last keeps its full length; then drop the first n_seed
    seq = torch.cat([o[:, :Jd] for o in outs], dim=-1)[0, :, 0, :].T
    return seq[g.n_seed:][:real_n]


# --------------------------------------------------------------------------------------
# Segment driver (main/mydiffusion_zeggs/sample.py:210-326), features instead of raw wav
# --------------------------------------------------------------------------------------
def stitch_segment(prev_tail, sample, smoothing=True):
    """sample.py:266-288 for ONE clip (batch dim 1): root shift + the n=1 'blend' quirk.
    prev_tail: [1,J,1,n_seed] (last n_seed frames of the previous segment), sample [1,J,1,T] (mutated copy)."""
    sample = sample.clone()
    if smoothing:
        delta = (sample[:, 0:3, :, 0] - prev_tail[:, 0:3, :, 0]).unsqueeze(-1)
        sample[:, 0:3] = sample[:, 0:3] - delta
    n = prev_tail.shape[0]          # == 1: ``len(last_poses)`` is the batch dimension (quirk, sample.py:284-288)
    for j in range(n):
        sample[..., j] = prev_tail[..., j] * (n - j) / (n + 1) + sample[..., j] * (j + 1) / (n + 1)
    return sample


def inference_clip(W, g, sched, features, style, *, seed=123456, clip_id=0, sampler="ddpm",
                   skip_timesteps=0, dtype=torch.float32, model_fn_factory=None):
    """One clip, all segments.  features: list of per-segment conditioning [Ta, audio_dim];
    style [style_in].  Returns normalised motion [n_frames - n_seed, J] (sample.py:291-296)."""
    outs = []
    seed_pose = torch.zeros(1, g.njoints, 1, g.n_seed, dtype=dtype)
    for s, feat in enumerate(features):
        y = {"audio": feat[None].to(dtype), "style": style[None].to(dtype), "seed": seed_pose}
        mf = model_fn_factory(y) if model_fn_factory else None
        sample, _ = p_sample_loop(W, g, sched, y, 1, seed=seed, clip_ids=[clip_id], segment=s,
                                  sampler=sampler, skip_timesteps=skip_timesteps, dtype=dtype, model_fn=mf)
        if outs and g.n_seed:
            tail = outs[-1][..., -g.n_seed:]
            outs[-1] = outs[-1][..., :-g.n_seed]
            sample = stitch_segment(tail, sample)
        outs.append(sample)
        seed_pose = sample[..., -g.n_seed:]
    outs[-1] = outs[-1][..., :-g.n_seed]
    seq = torch.cat(outs, dim=-1)[0, :, 0, :].T                     # [n_frames, J]
    return seq[g.n_seed:]


def denormalise(seq, mean, std):
    """sample.py:320-326: x * clip(std, 0.01) + mean (float64 numpy)."""
    return np.multiply(np.asarray(seq, dtype=np.float32), np.clip(std, 0.01, None)) + mean


# --------------------------------------------------------------------------------------
# pose2bvh numeric tail (main/process/process_zeggs_bvh.py:219-275 and the ZEGGS anim helpers)
# --------------------------------------------------------------------------------------
ZEGGS_PARENTS = np.array([-1, 0, 1, 2, 3, 4, 5, 6, 7, 4, 9, 10, 11, 12, 13, 14, 15, 12, 17, 18, 19, 12, 21, 22, 23,
                          12, 25, 26, 27, 12, 29, 30, 31, 12, 11, 4, 35, 36, 37, 38, 39, 40, 41, 38, 43, 44, 45, 38,
                          47, 48, 49, 38, 51, 52, 53, 38, 55, 56, 57, 38, 37, 0, 61, 62, 63, 64, 63, 62, 0, 68, 69,
                          70, 71, 70, 69], dtype=np.int32)


def _quat_mul(x, y):
    x0, x1, x2, x3 = x[..., 0:1], x[..., 1:2], x[..., 2:3], x[..., 3:4]
    y0, y1, y2, y3 = y[..., 0:1], y[..., 1:2], y[..., 2:3], y[..., 3:4]
    return np.concatenate([y0 * x0 - y1 * x1 - y2 * x2 - y3 * x3, y0 * x1 + y1 * x0 - y2 * x3 + y3 * x2,
                           y0 * x2 + y1 * x3 + y2 * x0 - y3 * x1, y0 * x3 - y1 * x2 + y2 * x1 + y3 * x0], axis=-1)


def _quat_mul_vec(q, v):
    t = 2.0 * np.cross(q[..., 1:], v)
    return v + q[..., 0][..., None] * t + np.cross(q[..., 1:], t)


def _quat_from_xform(ts, eps=1e-10):
    """anim/quat.py:166-206."""
    t = ts[..., 0, 0] + ts[..., 1, 1] + ts[..., 2, 2]
    qs = np.zeros(ts.shape[:-2] + (4,), dtype=ts.dtype)
    s = 0.5 / np.sqrt(np.maximum(t + 1, eps))
    cand = np.stack([0.25 / s, s * (ts[..., 2, 1] - ts[..., 1, 2]), s * (ts[..., 0, 2] - ts[..., 2, 0]),
                     s * (ts[..., 1, 0] - ts[..., 0, 1])], -1)
    qs = np.where((t > 0)[..., None], cand, qs)
    c0 = (ts[..., 0, 0] > ts[..., 1, 1]) & (ts[..., 0, 0] > ts[..., 2, 2])
    s0 = 2.0 * np.sqrt(np.maximum(1.0 + ts[..., 0, 0] - ts[..., 1, 1] - ts[..., 2, 2], eps))
    cand = np.stack([(ts[..., 2, 1] - ts[..., 1, 2]) / s0, s0 * 0.25, (ts[..., 0, 1] + ts[..., 1, 0]) / s0,
                     (ts[..., 0, 2] + ts[..., 2, 0]) / s0], -1)
    qs = np.where(((t <= 0) & c0)[..., None], cand, qs)
    c1 = (~c0) & (ts[..., 1, 1] > ts[..., 2, 2])
    s1 = 2.0 * np.sqrt(np.maximum(1.0 + ts[..., 1, 1] - ts[..., 0, 0] - ts[..., 2, 2], eps))
    cand = np.stack([(ts[..., 0, 2] - ts[..., 2, 0]) / s1, (ts[..., 0, 1] + ts[..., 1, 0]) / s1, s1 * 0.25,
                     (ts[..., 1, 2] + ts[..., 2, 1]) / s1], -1)
    qs = np.where(((t <= 0) & c1)[..., None], cand, qs)
    c2 = (~c0) & (~c1)
    s2 = 2.0 * np.sqrt(np.maximum(1.0 + ts[..., 2, 2] - ts[..., 0, 0] - ts[..., 1, 1], eps))
    cand = np.stack([(ts[..., 1, 0] - ts[..., 0, 1]) / s2, (ts[..., 0, 2] + ts[..., 2, 0]) / s2,
                     (ts[..., 1, 2] + ts[..., 2, 1]) / s2, s2 * 0.25], -1)
    return np.where(((t <= 0) & c2)[..., None], cand, qs)


def _quat_to_euler_zyx(x):
    """anim/quat.py:111-118."""
    x0, x1, x2, x3 = x[..., 0:1], x[..., 1:2], x[..., 2:3], x[..., 3:4]
    return np.concatenate([np.arctan2(2.0 * (x0 * x3 + x1 * x2), 1.0 - 2.0 * (x2 * x2 + x3 * x3)),
                           np.arcsin(np.clip(2.0 * (x0 * x2 - x3 * x1), -1.0, 1.0)),
                           np.arctan2(2.0 * (x0 * x1 + x2 * x3), 1.0 - 2.0 * (x1 * x1 + x2 * x2))], axis=-1)


def pose2bvh_values(poses, length, smoothing=True):
    """process_zeggs_bvh.py:219-275 + utils_zeggs.py:47-87 up to the arrays handed to bvh.save:
    returns (positions [3*length,75,3], euler_degrees [3*length,75,3])."""
    from scipy.signal import savgol_filter
    nj = 75
    poses = np.asarray(poses, dtype=np.float64)
    if smoothing:
        out = np.zeros_like(poses)
        for i in range(poses.shape[1]):
            out[:, i] = savgol_filter(poses[:, i], 15, 2)
    else:
        out = poses
    root_pos, root_rot = out[:, 0:3], out[:, 3:7]
    lpos = out[:, 13: 13 + nj * 3].reshape(length, nj, 3)
    ltxy = torch.as_tensor(out[:, 13 + nj * 3: 13 + nj * 9].reshape(length, nj, 2, 3), dtype=torch.float32)
    # txform.xform_orthogonalize_from_xy (anim/txform.py:23-34), float32 torch as in the reference
    xaxis = ltxy[..., 0:1, :]
    zaxis = torch.cross(xaxis, ltxy[..., 1:2, :], dim=-1)
    yaxis = torch.cross(zaxis, xaxis, dim=-1)
    eps = 1e-10
    m = torch.cat([xaxis / (torch.norm(xaxis, 2, dim=-1)[..., None] + eps),
                   yaxis / (torch.norm(yaxis, 2, dim=-1)[..., None] + eps),
                   zaxis / (torch.norm(zaxis, 2, dim=-1)[..., None] + eps)], dim=-2).transpose(-1, -2)
    lrot = _quat_from_xform(m.numpy())
    root_pos, root_rot = root_pos.repeat(3, axis=0), root_rot.repeat(3, axis=0)
    lpos, lrot = lpos.repeat(3, axis=0).copy(), lrot.repeat(3, axis=0).copy()
    lpos[:, 0] = _quat_mul_vec(root_rot, lpos[:, 0]) + root_pos
    lrot[:, 0] = _quat_mul(root_rot, lrot[:, 0])
    return lpos, np.degrees(_quat_to_euler_zyx(lrot))
