"""ZEGGS sampling CLI and segment driver — host mirror of reference main/mydiffusion_zeggs/sample.py
(argparse :400-407, YAML merge :408-414, create_model_and_diffusion :51-56, inference :210-338, main :341-384).

    python -m diffusestylegesture_b200.sample --config <yml> --gpu 0 --model_path model.pt \
           --audiowavlm_path 015_Happy_4_x_1_0.wav --max_len 320

Differences from the reference, all additive:
  * every diffusion step runs inside libdsg (``diffusion.p_sample_loop`` -> one C call per segment);
  * ``inference_batch`` runs B independent clips in lock-step (the reference is batch 1 only); per clip it
    reproduces the reference semantics exactly, including the n == 1 "blend" quirk (sample.py:284-288);
  * the WavLM-Large conditioning forward (``wavlm_init`` / ``wav2wavlm``, sample.py:28-48) runs in libdsg too
    (dsg_wavlm_forward) and takes all segments of a clip as one batch; conditioning may also be given as
    precomputed WavLM-shaped features;
  * new optional YAML keys: ``precision`` (bf16|fp32), ``sampler`` (ddpm|ddim), ``timestep_respacing``, ``max_batch``;
  * ``--batch manifest.csv`` (``main_batch``): many clips per run — wav load, style parsing, batched WavLM + sampling
    of clips with equal segment counts, BVH files written by a thread pool.
"""
import argparse
import math
import os
from datetime import datetime
from pprint import pprint

import numpy as np
import torch
import torch.nn.functional as F
import yaml

from .mdm import MDM
from .model_util import create_gaussian_diffusion, load_model_wo_clip
from .process_zeggs_bvh import pose2bvh

style2onehot = {
    'Happy': [1, 0, 0, 0, 0, 0],
    'Sad': [0, 1, 0, 0, 0, 0],
    'Neutral': [0, 0, 1, 0, 0, 0],
    'Old': [0, 0, 0, 1, 0, 0],
    'Angry': [0, 0, 0, 0, 1, 0],
    'Relaxed': [0, 0, 0, 0, 0, 1],
}

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CONFIG = os.path.join(_HERE, 'configs', 'DiffuseStyleGesture.yml')
DEFAULT_STATS = os.path.join(_HERE, 'configs', 'zeggs_mean_std.npz')


class Config(dict):
    """Attribute-style dict (stands in for easydict.EasyDict, sample.py:414).  Missing attributes raise
    ``AttributeError`` (as easydict does), so ``getattr(cfg, k, default)`` / ``hasattr`` / ``copy`` / ``pickle`` work."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    __setattr__ = dict.__setitem__


def create_model_and_diffusion(args):
    """sample.py:51-56 — same hard-coded ZEGGS geometry; only ``args.audio_feat`` is read by the reference."""
    model = MDM(modeltype='', njoints=1141, nfeats=1, translation=True, pose_rep='rot6d', glob=True, glob_rot=True,
                cond_mode='cross_local_attention3_style1', clip_version='ViT-B/32', action_emb='tensor',
                audio_feat=args.audio_feat, arch='trans_enc', latent_dim=256, n_seed=8,
                n_poses=_get(args, 'n_poses', 88),
                precision=_get(args, 'precision', 'bf16'), max_batch=max(1, int(_get(args, 'max_batch', 1) or 1)))
    diffusion = create_gaussian_diffusion(_get(args, 'timestep_respacing', ''))
    return model, diffusion


def _get(args, key, default):
    if isinstance(args, dict):
        return args.get(key, default)
    return getattr(args, key, default)


def wavlm_init(device=None, wavlm_model_path='./WavLM/WavLM-Large.pt', max_batch=16):
    """sample.py:28-39: load ``WavLM-Large.pt`` ({'cfg', 'model'}) into the engine-backed WavLM mirror."""
    from .wavlm import WavLM, WavLMConfig
    checkpoint = torch.load(wavlm_model_path, map_location=torch.device('cpu'))
    model = WavLM(WavLMConfig(checkpoint['cfg']), max_batch=max_batch)
    model.load_state_dict(checkpoint['model'])
    model = model.to(device if device is not None else torch.device('cuda:0'))
    model.eval()
    return model


def wav2wavlm(model, wav_input_16khz, device, n_poses=88):
    """sample.py:44-48 (ZEGGS: no waveform layer-norm).  ``model`` is the engine-backed
    ``diffusestylegesture_b200.wavlm.WavLM`` (fused forward + interpolation in libdsg) or any module exposing
    ``extract_features``."""
    if hasattr(model, "wav2wavlm"):          # diffusestylegesture_b200.wavlm.WavLM: fused in libdsg
        return model.wav2wavlm(wav_input_16khz.to(device), n_poses)
    rep = model.extract_features(wav_input_16khz.to(device))[0]
    return F.interpolate(rep.transpose(1, 2), size=n_poses, align_corners=True, mode='linear').transpose(1, 2)


def segment_plan(n_frames, n_poses, n_seed):
    """sample.py:216-222: stride = n_poses - n_seed, floor(n_frames / stride) segments (at least one)."""
    stride = n_poses - n_seed
    if n_frames < stride:
        return 1, n_frames
    nseg = math.floor(n_frames / stride)
    return nseg, nseg * stride


@torch.no_grad()
def inference_batch(model, diffusion, features, styles, *, seed=123456, clip_ids=None, smoothing=True,
                    skip_timesteps=0, sampler='ddpm', device=None, out_device='cpu', out=None, seed_pose0=None,
                    seed_last=None, keep_last_tail=False):
    """B clips x S segments through the engine.

    features: sequence (len = segments) of [B, audio_frames, audio_dim] tensors (host or device);  styles: [B, style_in].
    Returns the normalised motion [B, n_frames - n_seed, J] float32 (sample.py:291-296) on ``out_device`` — or in ``out``
    (e.g. a pinned host tensor of that shape: one asynchronous copy instead of a pageable one).
    Host feature tensors in pinned memory are copied one segment ahead on a side stream, under the previous segment's loop.
    ``seed_pose0`` [B, J, 1, n_seed]: seed of the first segment (default zeros, sample.py:244; the BEAT-TWH driver passes
    the velocity/acceleration seed of a recorded gesture).  ``seed_last``: y['seed_last'] of the "++" variant.
    ``keep_last_tail``: the BEAT-TWH driver keeps the last segment's final n_seed frames (its sample.py:189-199).
    """
    g = model.geometry
    nseg = len(features) if not callable(features) else None
    if nseg is None:
        raise ValueError("pass a list of per-segment feature tensors")
    B = styles.shape[0]
    eng = model.get_engine(B)
    dev = eng.device
    if callable(diffusion) and not hasattr(diffusion, 'p_sample_loop'):
        sample_fn = diffusion                    # any callable with the p_sample_loop signature (sample.py:253)
    else:
        sample_fn = {'ddpm': 'p_sample_loop', 'ddim': 'ddim_sample_loop', 'plms': 'plms_sample_loop'}[sampler]
        sample_fn = getattr(diffusion, sample_fn)
    styles = torch.as_tensor(styles, dtype=torch.float32)
    shape_ = (B, g.njoints, 1, g.n_poses)
    if seed_pose0 is None:
        seed_pose = torch.zeros(B, g.njoints, 1, g.n_seed, device=dev)          # sample.py:244
    else:
        seed_pose = torch.as_tensor(seed_pose0, dtype=torch.float32).to(dev).expand(B, g.njoints, 1, g.n_seed).contiguous()
    pieces, prev = [], None
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev) if any((not f.is_cuda) and f.is_pinned() for f in features) else None

    def fetch(i):                                                               # host (pinned) -> device, asynchronously
        f = features[i]
        if side is None or f.is_cuda or not f.is_pinned():
            return f
        with torch.cuda.stream(side):                                           # not ordered after `main`: it must overlap
            return f.to(dev, non_blocking=True)

    cur = fetch(0)
    for i in range(nseg):
        if side is not None:
            main.wait_stream(side)
        y = {'style': styles, 'seed': seed_pose, 'audio': cur,
             'mask_local': torch.ones(1, g.n_poses, dtype=torch.bool),
             'noise_seed': seed, 'segment': i, 'clip_ids': clip_ids}
        if seed_last is not None:
            y['seed_last'] = seed_last
        sample = sample_fn(model, shape_, clip_denoised=False, model_kwargs={'y': y}, skip_timesteps=skip_timesteps,
                           init_image=None, progress=False, dump_steps=None, noise=None, const_noise=False)
        if isinstance(cur, torch.Tensor) and cur.is_cuda:
            cur.record_stream(main)
        cur = fetch(i + 1) if i + 1 < nseg else None                            # overlaps with segment i's loop
        if prev is not None and g.n_seed != 0:                                    # sample.py:266-288
            tail = prev[..., -g.n_seed:].contiguous()
            pieces.append(prev[..., :-g.n_seed])
            eng.stitch_segment(tail, sample, smoothing=smoothing)
        elif prev is not None:
            pieces.append(prev)
        prev = sample
        seed_pose = sample[..., -g.n_seed:].contiguous()                          # sample.py:249
    pieces.append(prev[..., :-g.n_seed] if (g.n_seed != 0 and not keep_last_tail) else prev)        # sample.py:292
    seq = torch.cat(pieces, dim=-1)[:, :, 0, :].transpose(1, 2)                   # [B, n_frames, J]
    seq = seq[:, g.n_seed:] if g.n_seed != 0 else seq                             # sample.py:296
    if out is not None:
        if tuple(out.shape) != tuple(seq.shape) or out.dtype != torch.float32:
            raise ValueError(f"out must be float32 {tuple(seq.shape)}")
        out.copy_(seq, non_blocking=True)
        if not out.is_cuda:
            main.synchronize()
        return out
    return seq.contiguous().to(out_device)


def denormalise(sampled_seq, stats_path=DEFAULT_STATS):
    """sample.py:320-326: x * clip(std, 0.01) + mean  (float64 numpy)."""
    st = np.load(stats_path)
    std = np.clip(np.array(st['std']).squeeze(), a_min=0.01, a_max=None)
    return np.multiply(np.asarray(sampled_seq), std) + np.array(st['mean']).squeeze()


def inference(args, wavlm_model, audio, sample_fn, model, n_frames=0, smoothing=False, SG_filter=False, minibatch=False,
              skip_timesteps=0, n_seed=8, style=None, seed=123456, features=None, save_dir='sample_dir',
              stats_path=DEFAULT_STATS):
    """Reference signature (sample.py:210) for ONE clip.  ``audio`` is the 16 kHz waveform (numpy); alternatively
    pass ``features`` = list of per-segment [1, n_poses, 1024] tensors and ``audio=None``.  Writes the BVH."""
    if not minibatch:
        raise NotImplementedError("minibatch=False references an undefined variable in the reference (sample.py:302)")
    torch.manual_seed(seed)
    g = model.geometry
    n_poses = _get(args, 'n_poses', g.n_poses)
    if features is None:
        avail = audio.shape[0] * 20 // 16000
        n_frames = avail if n_frames == 0 else min(n_frames, avail)
        if n_frames < n_poses - n_seed:
            raise ValueError(f"{n_frames} frames of audio: shorter than one segment stride ({n_poses - n_seed} frames)")
        nseg, n_frames = segment_plan(n_frames, n_poses, n_seed)
        audio = audio[:int(n_frames * 16000 / 20)]
        stride = n_poses - n_seed
        dev = next(model.parameters()).device
        chunks = torch.from_numpy(audio).to(torch.float32).reshape(nseg, int(stride * 16000 / 20))
        pad = int(n_seed * 16000 / 20)
        # sample.py:238-251: segment i hears [last n_seed frames of segment i-1 (zeros for i = 0) | its own stride frames];
        # the segments are independent for WavLM, so they go through it as ONE batch
        wavs = torch.stack([torch.cat((torch.zeros(pad) if i == 0 else chunks[i - 1, -pad:], chunks[i])) for i in range(nseg)])
        feats = wav2wavlm(wavlm_model, wavs, dev, n_poses)
        features = [feats[i:i + 1] for i in range(nseg)]
    else:
        nseg = len(features)
        n_frames = nseg * (n_poses - n_seed)
    # the reference calls sample_fn(model, shape, ...) as a plain callable (sample.py:253): bound sampler methods of the
    # diffusion object, functools.partial of them and user wrappers all work
    seq = inference_batch(model, sample_fn, features, torch.as_tensor([style], dtype=torch.float32), seed=seed,
                          smoothing=smoothing, skip_timesteps=skip_timesteps)
    out_poses = denormalise(seq[0].numpy(), stats_path)
    print(out_poses.shape)
    prefix = str(datetime.now().strftime('%Y%m%d_%H%M%S'))
    if smoothing: prefix += '_smoothing'
    if SG_filter: prefix += '_SG'
    if minibatch: prefix += '_minibatch'
    prefix += '_%s' % (n_frames)
    prefix += '_' + str(style)
    prefix += '_' + str(seed)
    os.makedirs(save_dir, exist_ok=True)
    path = os.path.join(save_dir, prefix + '.bvh')
    pose2bvh(out_poses, path, length=n_frames - n_seed, smoothing=SG_filter)
    return path, out_poses


def load_wav_16k(path):
    """Mono 16 kHz float32 waveform (stands in for librosa.load(path, sr=16000), sample.py:346): PCM 8 / 16 / 24 / 32 bit and
    IEEE-float wav files, channels averaged, integer samples scaled by 2^-(bits-1) as libsndfile does.  Files at 16 kHz (the
    ZEGGS data) come out exactly as librosa returns them; other rates are resampled with a polyphase FIR
    (scipy.signal.resample_poly) — librosa's default resampler (soxr / kaiser_best, version dependent) differs from it by
    ~1e-3 of full scale."""
    from scipy.io import wavfile
    try:
        sr, d = wavfile.read(path)
    except ValueError as ex:
        raise NotImplementedError(f"{path}: unsupported wav encoding ({ex})") from ex
    if d.dtype == np.uint8:
        x = (d.astype(np.float32) - 128.0) / 128.0
    elif d.dtype == np.int16:
        x = d.astype(np.float32) / 32768.0
    elif d.dtype == np.int32:                     # 32-bit PCM, and 24-bit PCM (scipy left-justifies it in int32)
        x = (d.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif d.dtype in (np.float32, np.float64):
        x = d.astype(np.float32)
    else:
        raise NotImplementedError(f"{path}: sample type {d.dtype}")
    if x.ndim > 1:
        x = x.mean(axis=1)
    if sr != 16000:
        from scipy.signal import resample_poly
        gdiv = math.gcd(int(sr), 16000)
        x = resample_poly(x, 16000 // gdiv, int(sr) // gdiv).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32), 16000


def main(args, save_dir, model_path, audio_path=None, mfcc_path=None, audiowavlm_path=None, max_len=0, wavlm_model=None,
         features=None):
    """sample.py:341-384."""
    os.makedirs(save_dir, exist_ok=True)
    print("Creating model and diffusion...")
    model, diffusion = create_model_and_diffusion(args)
    print(f"Loading checkpoints from [{model_path}]...")
    state_dict = torch.load(model_path, map_location='cpu')
    load_model_wo_clip(model, state_dict)
    model.to(torch.device('cuda:' + str(args.gpu)))
    model.eval()
    sample_fn = diffusion.p_sample_loop if _get(args, 'sampler', 'ddpm') == 'ddpm' else diffusion.ddim_sample_loop
    style = style2onehot[style_from_filename(audiowavlm_path)]
    print(style)
    audio = None
    if features is None:
        if wavlm_model is None:                                                  # sample.py:383
            wavlm_model = wavlm_init(torch.device('cuda:' + str(args.gpu)),
                                     _get(args, 'wavlm_path', './WavLM/WavLM-Large.pt'))
        audio, _ = load_wav_16k(audiowavlm_path)
    return inference(args, wavlm_model, audio, sample_fn, model, n_frames=max_len, smoothing=True, SG_filter=True,
                     minibatch=True, skip_timesteps=0, style=style, seed=123456, features=features, save_dir=save_dir)


def style_from_filename(path):
    """sample.py:378: the style is token 1 of the '_'-separated file name (e.g. 015_Happy_4_x_1_0.wav)."""
    parts = os.path.basename(path).split('_')
    if len(parts) < 2:
        raise ValueError(f"{path}: no style given and the file name has no '_<Style>_' token (sample.py:378)")
    return parts[1]


def auto_max_batch(n_clips, device):
    """Clips per engine launch when the config does not pin it: the clip kernel runs one persistent CTA (pair) per clip,
    so up to two waves of the SM count keep every SM busy; fewer clips than that run as one batch."""
    sms = torch.cuda.get_device_properties(device).multi_processor_count if torch.cuda.is_available() else 148
    return max(1, min(int(n_clips), 2 * sms))


def read_manifest(path):
    """Batch manifest (CSV, '#' comments, optional header `wav,style,clip_id`): one clip per row.
    `style` = a name of `style2onehot`, an index 0..5, or empty -> token 1 of the file name (sample.py:378);
    `clip_id` (optional) keys the clip's noise stream (default: the row number)."""
    import csv
    rows = []
    with open(path, newline='') as fh:
        for rec in csv.reader(fh):
            rec = [c.strip() for c in rec]
            if not rec or not rec[0] or rec[0].startswith('#') or rec[0].lower() in ('wav', 'wav_path', 'path'):
                continue
            wav = rec[0]
            tok = rec[1] if len(rec) > 1 and rec[1] else style_from_filename(wav)
            if tok in style2onehot:
                style = style2onehot[tok]
            elif tok.isdigit() and int(tok) < 6:
                style = [1 if i == int(tok) else 0 for i in range(6)]
            else:
                raise ValueError(f"{path}: unknown style '{tok}' for {wav}")
            cid = int(rec[2]) if len(rec) > 2 and rec[2] else len(rows)
            rows.append({'wav': wav, 'style': style, 'style_name': tok, 'clip_id': cid})
    if not rows:
        raise ValueError(f"{path}: empty manifest")
    return rows


def segment_windows(audio, n_frames, n_poses, n_seed):
    """sample.py:224-251 for one clip: [nseg, n_poses * 800] waveform windows = the last n_seed frames of the previous
    stride (zeros for the first) followed by the segment's own stride."""
    nseg, n_frames = segment_plan(n_frames, n_poses, n_seed)
    stride = n_poses - n_seed
    chunks = torch.from_numpy(np.ascontiguousarray(audio[:int(n_frames * 16000 / 20)])).to(torch.float32).reshape(nseg, int(stride * 16000 / 20))
    pad = int(n_seed * 16000 / 20)
    return torch.stack([torch.cat((torch.zeros(pad) if i == 0 else chunks[i - 1, -pad:], chunks[i])) for i in range(nseg)]), n_frames


def plan_batches(lengths, max_batch):
    """Group clips with the same number of segments (they run in lock-step), at most max_batch per group.
    lengths: segments per clip.  Returns a list of index lists, longest clips first."""
    by = {}
    for i, n in enumerate(lengths):
        by.setdefault(n, []).append(i)
    out = []
    for n in sorted(by, reverse=True):
        idx = by[n]
        out += [idx[i:i + max_batch] for i in range(0, len(idx), max_batch)]
    return out


def wav_frames_16k(path):
    """Motion frames (20 fps) a wav file yields after `load_wav_16k`, from the header alone (every rank of a sharded run needs
    every clip's length to size the gather; only the owner decodes the samples)."""
    import wave
    with wave.open(path, 'rb') as w:
        sr, n = w.getframerate(), w.getnframes()
    if sr != 16000:
        gdiv = math.gcd(sr, 16000)
        n = -(-n * (16000 // gdiv) // (sr // gdiv))          # resample_poly: ceil(n * up / down)
    return n * 20 // 16000


def main_batch(args, save_dir, model_path, manifest, max_len=0, wavlm_model=None, model=None, diffusion=None,
               seed=123456, stats_path=DEFAULT_STATS, writers=8):
    """Many clips at once (additive to the reference CLI): every clip of the manifest goes through WavLM and the sampler in
    batches of clips with equal segment counts; BVH files are written by a thread pool while the GPU runs the next batch.

    Under ``torchrun`` (WORLD_SIZE > 1: one process per GPU) the manifest is sharded contiguously over the ranks
    (``distributed.shard_bounds``; noise is keyed by clip id, so results do not depend on the sharding), every rank samples
    its clips on ``cuda:LOCAL_RANK``, and ONE gather brings the finished motions to rank 0, which writes every BVH:
        torchrun --nproc-per-node 8 -m diffusestylegesture_b200.sample --batch manifest.csv --model_path ...
    Returns the list of BVH paths (rank 0) or None (other ranks)."""
    from concurrent.futures import ThreadPoolExecutor
    from .distributed import init_from_env, shard_bounds, gather_motions
    rows = read_manifest(manifest) if isinstance(manifest, str) else manifest
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank, local = 0, None
    if world > 1:
        rank, world, local = init_from_env(os.environ.get('DSG_DIST_BACKEND') or None)
    os.makedirs(save_dir, exist_ok=True)
    dev = torch.device('cuda:' + str(_get(args, 'gpu', '0') if local is None or os.environ.get('DSG_DIST_SAME_GPU') else local))
    lo, hi = shard_bounds(len(rows), rank, world)
    mine = list(range(lo, hi))
    max_batch = int(_get(args, 'max_batch', 0) or 0) or auto_max_batch(len(mine), dev)
    if model is None:
        cfg = dict(args) if isinstance(args, dict) else dict(vars(args))
        cfg['max_batch'] = max_batch
        model, diffusion = create_model_and_diffusion(Config(cfg))
        load_model_wo_clip(model, torch.load(model_path, map_location='cpu'))
        model.to(dev).eval()
    if wavlm_model is None:
        wavlm_model = wavlm_init(dev, _get(args, 'wavlm_path', './WavLM/WavLM-Large.pt'))
    g = model.geometry
    sampler = _get(args, 'sampler', 'ddpm')
    stride = g.n_poses - g.n_seed

    def frames_of(r, audio=None):
        avail = (audio.shape[0] * 20 // 16000) if audio is not None else \
            (r['audio'].shape[0] * 20 // 16000 if 'audio' in r else wav_frames_16k(r['wav']))
        n = min(max_len, avail) if max_len else avail
        if n < stride:                # the reference would run one ragged segment and fail inside the local attention
            raise ValueError(f"{r['wav']}: {n} frames of audio, shorter than one segment stride ({stride} frames = "
                             f"{stride / 20:.1f} s); pad the clip or drop it from the manifest")
        return segment_plan(n, g.n_poses, g.n_seed)[1]

    n_all = [frames_of(r) for r in rows]                  # every rank validates the whole manifest (and sizes the gather)
    wins = {}
    for i in mine:
        audio = rows[i]['audio'] if 'audio' in rows[i] else load_wav_16k(rows[i]['wav'])[0]
        wins[i] = segment_windows(audio, frames_of(rows[i], audio), g.n_poses, g.n_seed)
    paths = [None] * len(rows)

    def write(i, seq, n_frames):
        r = rows[i]
        name = os.path.splitext(os.path.basename(r['wav']))[0]
        path = os.path.join(save_dir, f"{name}_{n_frames}_{r['style_name']}_{seed}_{r['clip_id']}.bvh")
        pose2bvh(denormalise(seq, stats_path), path, length=n_frames - g.n_seed, smoothing=True)
        paths[i] = path

    nmax = max(n_all) - g.n_seed
    shard = torch.zeros(len(mine), nmax, g.njoints, dtype=torch.float32, device=dev) if world > 1 else None
    with ThreadPoolExecutor(max_workers=writers) as pool:
        jobs = []
        for grp in plan_batches([wins[i][0].shape[0] for i in mine], max_batch):
            idx = [mine[k] for k in grp]
            nseg = wins[idx[0]][0].shape[0]
            feats = [wav2wavlm(wavlm_model, torch.stack([wins[i][0][s] for i in idx]), dev, g.n_poses) for s in range(nseg)]
            styles = torch.tensor([rows[i]['style'] for i in idx], dtype=torch.float32)
            seqs = inference_batch(model, diffusion, feats, styles, seed=seed, clip_ids=[rows[i]['clip_id'] for i in idx],
                                   smoothing=True, sampler=sampler, out_device='cpu' if world == 1 else dev)
            if world == 1:
                seqs = seqs.numpy()
                jobs += [pool.submit(write, i, seqs[k], wins[i][1]) for k, i in enumerate(idx)]
            else:
                for k, i in enumerate(idx):
                    shard[i - lo, :seqs.shape[1]] = seqs[k]
        if world > 1:
            if torch.distributed.get_backend() == 'gloo':                # CPU test configuration: gloo moves host tensors
                shard = shard.cpu()
            everything = gather_motions(shard, len(rows))                  # the single collective of the path
            if rank == 0:
                everything = everything.cpu().numpy()
                jobs += [pool.submit(write, i, everything[i, :n_all[i] - g.n_seed], n_all[i]) for i in range(len(rows))]
        for j in jobs:
            j.result()
    if world > 1:
        torch.distributed.barrier()
    return paths if rank == 0 else None


def parse_cli(argv=None):
    parser = argparse.ArgumentParser(description='DiffuseStyleGesture')
    parser.add_argument('--config', default=DEFAULT_CONFIG)
    parser.add_argument('--gpu', type=str, default='0')
    parser.add_argument('--no_cuda', type=list, default=['0'])     # parsed and never read, exactly as in the reference (sample.py:403)
    parser.add_argument('--model_path', type=str, default='./model000450000.pt')
    parser.add_argument('--audiowavlm_path', type=str, default='')
    parser.add_argument('--max_len', type=int, default=0)
    parser.add_argument('--batch', type=str, default='', help='CSV manifest (wav,style[,clip_id]) -> one BVH per row (additive)')
    parser.add_argument('--save_dir', type=str, default='sample_dir')
    parser.add_argument('--wavlm_path', type=str, default='./WavLM/WavLM-Large.pt')
    parser.add_argument('--max_batch', type=int, default=None,
                        help='clips per engine launch for --batch (default: config value; 0 = auto)')
    args = parser.parse_args(argv)
    with open(args.config) as f:
        config = yaml.safe_load(f)
    for k, v in vars(args).items():
        if k == 'max_batch' and v is None:       # flag not given: keep the YAML value
            continue
        config[k] = v
    return Config(config)


if __name__ == '__main__':
    config = parse_cli()
    pprint(dict(config))
    if config.batch:
        if int(os.environ.get('WORLD_SIZE', '1')) == 1:
            torch.cuda.set_device(int(config.gpu))
        for p in main_batch(config, config.save_dir, config.model_path, config.batch, max_len=config.max_len) or []:
            print(p)
    else:
        torch.cuda.set_device(int(config.gpu))
        main(config, config.save_dir, config.model_path, audio_path=None, mfcc_path=None,
             audiowavlm_path=config.audiowavlm_path, max_len=config.max_len)
