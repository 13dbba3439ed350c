"""``sample_beat_twh.py`` — host mirror of the DiffuseStyleGesture+ / ++ driver (reference
BEAT-TWH-main/mydiffusion_beat_twh/sample.py): same ``create_model_and_diffusion(args)`` / ``inference(...)`` / ``main(...)``
entry points, CLI flags (:275-289) and preset resolution (:296-321); the sampling loop runs in libdsg (sm_100a CUDA).

    python -m diffusestylegesture_b200.sample_beat_twh --config configs/DiffuseStyleGesture_beat.yml --dataset BEAT \
        --model_path ./BEAT_mymodel4_512_v0/model001260000.pt --tst_path <dir> --tst_prefix 2_scott_0_1_1

Differences, all additive: ``inference_batch_beat`` runs B clips per engine call (BASELINE config 4: batch 32 sharded over
4 GPUs through ``distributed.shard_bounds``); the recorded seed gesture the reference reads from the dataset tree
(:118-127) can be passed as an array; the BVH tail needs the reference's pickled pymo pipelines (``process_beat_twh_bvh``).
"""
import argparse
import copy
import math
import os

import numpy as np
import torch
import yaml

from .mdm import MDM
from .model_util import create_gaussian_diffusion, load_model_wo_clip
from .sample import Config, _get, inference_batch

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CONFIG = os.path.join(_HERE, 'configs', 'DiffuseStyleGesture_beat.yml')
STATS = {'BEAT': os.path.join(_HERE, 'configs', 'beat_mean_std_v0.npz'),
         'TWH': os.path.join(_HERE, 'configs', 'twh_mean_std_v0.npz')}

speaker_id_dict = {2: 0, 10: 1}           # sample.py:24-27
id_speaker_dict = {0: 2, 1: 10}           # sample.py:29-32
COND_MODE = {'DiffuseStyleGesture': 'cross_local_attention3_style1_sample',          # sample.py:298-304
             'DiffuseStyleGesture+': 'cross_local_attention4_style1_sample',
             'DiffuseStyleGesture++': 'cross_local_attention5_style1_sample'}


def resolve_presets(config):
    """sample.py:296-321: cond_mode from ``name``; dataset / version presets overwrite the YAML geometry."""
    assert config.name in COND_MODE
    config.cond_mode = COND_MODE[config.name]
    if config.dataset == 'BEAT':
        config.style_dim = 2
        config.audio_feature_dim = 1434
        if 'v0' in config.version:
            config.motion_dim, config.njoints = 684, 2052
        elif 'v2' in config.version:
            config.motion_dim, config.njoints = 1141, 1141
    elif config.dataset == 'TWH':
        if 'v0' in config.version:
            config.motion_dim, config.njoints = 744, 2232
            config.latent_dim, config.audio_feat_dim_latent = 512, 128
            config.style_dim = 17
            config.audio_feature_dim = 1435          # with laugh
    else:
        raise NotImplementedError
    return config


def create_model_and_diffusion(args):
    """sample.py:35-41."""
    if args.cond_mode == COND_MODE['DiffuseStyleGesture']:
        raise NotImplementedError("BEAT-TWH with cross_local_attention3 (the embed_text of njoints * n_seed inputs) is not "
                                  "wired here; the ZEGGS driver (sample.py) covers attention3")
    model = MDM(modeltype='', njoints=args.njoints, nfeats=1, cond_mode=args.cond_mode, audio_feat=args.audio_feat,
                arch='trans_enc', latent_dim=args.latent_dim, n_seed=args.n_seed, cond_mask_prob=_get(args, 'cond_mask_prob', 0.1),
                style_dim=args.style_dim, source_audio_dim=args.audio_feature_dim,
                audio_feat_dim_latent=args.audio_feat_dim_latent, n_poses=args.n_poses,
                precision=_get(args, 'precision', 'bf16'), max_batch=max(1, int(_get(args, 'max_batch', 1) or 1)))
    return model, create_gaussian_diffusion(_get(args, 'timestep_respacing', ''))


def load_stats(dataset, version='v0', path=None):
    """gesture_{BEAT,TWH}_{mean,std}_v0.npy of the reference (sample.py:76-84), shipped as one npz per dataset."""
    st = np.load(path or STATS[dataset])
    return np.array(st['mean']), np.array(st['std'])


def seed_from_gesture(seed_gesture, mean, std):
    """sample.py:129-136: normalise n_seed + 2 recorded frames, append velocity and acceleration -> [1, 3 * motion_dim, 1, n_seed]."""
    g = (np.asarray(seed_gesture) - mean) / std
    vel = g[1:] - g[:-1]
    acc = vel[1:] - vel[:-1]
    s = np.concatenate((g[2:], vel[1:], acc), axis=1)                            # (n_seed, njoints)
    return torch.from_numpy(s).float().transpose(0, 1).unsqueeze(0).unsqueeze(2)


def plan_subdivision(n_frames, n_poses, n_seed):
    """sample.py:54-62: ceil (not floor, unlike the ZEGGS driver), padded with zero features."""
    stride = n_poses - n_seed
    if n_frames < stride:
        return 1, stride
    nsub = math.ceil(n_frames / stride)
    return nsub, nsub * stride


@torch.no_grad()
def inference_batch_beat(args, textaudio, sample_fn, model, styles, seed_gestures, *, n_frames=0, skip_timesteps=0, seed=123456,
                         dataset='BEAT', clip_ids=None, seed_last_gesture=None, stats=None, out_device='cpu'):
    """B clips through sample.py:44-193.  textaudio [B, n, audio_feature_dim] (equal lengths), styles [B, style_dim],
    seed_gestures [B, n_seed + 2, motion_dim] raw recorded frames (or [n_seed + 2, motion_dim] for all).
    Returns de-normalised poses [B, real_n_frames, motion_dim] (float64 numpy, as handed to pose2bvh in :190-200)."""
    g = model.geometry
    textaudio = torch.as_tensor(textaudio, dtype=torch.float32)
    if textaudio.dim() == 2:
        textaudio = textaudio[None]
    B = textaudio.shape[0]
    if n_frames == 0:
        n_frames = textaudio.shape[1]
    else:
        textaudio = textaudio[:, :n_frames]
    real_n_frames = copy.deepcopy(n_frames)
    nsub, n_frames = plan_subdivision(n_frames, g.n_poses, g.n_seed)
    stride = g.n_poses - g.n_seed
    mean, std = stats if stats is not None else load_stats(dataset, _get(args, 'version', 'v0'))
    pad = torch.zeros(B, n_frames - real_n_frames, textaudio.shape[2], dtype=torch.float32, device=textaudio.device)
    audio = torch.cat((textaudio, pad), 1).reshape(B, nsub, stride, -1)                      # :71-73
    name = _get(args, 'name', 'DiffuseStyleGesture+')
    if name == 'DiffuseStyleGesture++':
        feats = [audio[:, i, :-g.n_seed].contiguous() for i in range(nsub)]                   # :106, :145
    else:
        feats = [audio[:, i].contiguous() for i in range(nsub)]                               # :104, :143
    sg = np.asarray(seed_gestures)
    if sg.ndim == 2:
        sg = np.broadcast_to(sg, (B,) + sg.shape)
    seed0 = torch.cat([seed_from_gesture(sg[b][:g.n_seed + 2], mean, std) for b in range(B)], 0)       # :118-136
    seed_last = None
    if name == 'DiffuseStyleGesture++':                                                        # :88-96
        sl = sg if seed_last_gesture is None else np.broadcast_to(np.asarray(seed_last_gesture), sg.shape)
        seed_last = torch.cat([seed_from_gesture(sl[b][:g.n_seed + 2], mean, std) for b in range(B)], 0)
    seq = inference_batch(model, sample_fn, feats, torch.as_tensor(np.asarray(styles), dtype=torch.float32).reshape(B, -1),
                          seed=seed, clip_ids=clip_ids, smoothing=False, skip_timesteps=skip_timesteps, seed_pose0=seed0,
                          seed_last=seed_last, keep_last_tail=True, out_device=out_device)     # [B, nsub * stride, J]
    division = 3 if 'v0' in _get(args, 'version', 'v0') else 1                                # :173-179
    seq = seq[:, :, :g.njoints // division]
    if out_device != 'cpu':
        return seq[:, :real_n_frames]
    out_poses = np.multiply(seq.numpy(), std) + mean                                           # :190
    return out_poses[:, :real_n_frames]


def inference(args, save_dir, prefix, textaudio, sample_fn, model, n_frames=0, smoothing=False, skip_timesteps=0, style=None,
              seed=123456, dataset='BEAT', seed_gesture=None, pipeline=None, stats=None):
    """Reference signature (sample.py:44) for one clip.  ``seed_gesture``: the n_seed + 2 recorded frames the reference loads
    from ``../../<dataset>_dataset/processed/gesture_<dataset>/<file>.npy`` (:118-127); if None that file is read.
    Writes ``<save_dir>/<prefix>_generated.bvh`` when the pymo pipeline is loadable, and returns the poses either way."""
    torch.manual_seed(seed)
    style = np.asarray(style)
    if dataset == 'BEAT':
        speaker = id_speaker_dict[int(np.argwhere(style == 1)[0][0])]
        assert speaker in speaker_id_dict.keys()
    elif dataset == 'TWH':
        speaker = int(np.where(style == np.max(style))[0][0])
    else:
        raise NotImplementedError
    if seed_gesture is None:
        if dataset == 'BEAT':
            fn = {2: "2_scott_0_1_1.npy", 10: "10_kieks_0_95_95.npy"}[speaker]
            seed_gesture = np.load("../../BEAT_dataset/processed/gesture_BEAT/" + fn)
        else:
            seed_gesture = np.load("../../TWH_dataset/processed/gesture_TWH/val_2023_v0_014_main-agent.npy")
    poses = inference_batch_beat(args, textaudio, sample_fn, model, style[None], np.asarray(seed_gesture)[None],
                                 n_frames=n_frames, skip_timesteps=skip_timesteps, seed=seed, dataset=dataset, stats=stats)[0]
    print(poses.shape, poses.shape[0])
    from . import process_beat_twh_bvh as PB
    try:
        if dataset == 'BEAT':
            if "v0" in _get(args, 'version', 'v0'):
                PB.pose2bvh_bugfix(save_dir, prefix, poses,
                                   pipeline=pipeline or '../process/resource/data_pipe_30fps_speaker' + str(speaker) + '.sav')
            else:
                raise NotImplementedError("BEAT v2 (ZEGGS-style features) goes through process_zeggs_bvh.pose2bvh")
        else:
            PB.pose2bvh_twh(poses, save_dir, prefix, pipeline_path=pipeline or "../process/resource/pipeline_rotmat_62.sav")
    except PB.PipelineUnavailable as ex:
        print(f"BVH not written: {ex}")
    return poses


def main(args, save_dir, model_path, tst_path=None, max_len=0, skip_timesteps=0, tst_prefix=None, dataset='BEAT', wav_path=None,
         txt_path=None, wavlm_path=None, word2vector_path=None):
    """sample.py:204-272 (the pre-extracted feature branch; the wav + tsv branch needs librosa / parselmouth feature
    extractors that are outside the hot path and not available offline)."""
    os.makedirs(save_dir, exist_ok=True)
    print("Creating model and diffusion...")
    model, diffusion = create_model_and_diffusion(args)
    print(f"Loading checkpoints from [{model_path}]...")
    load_model_wo_clip(model, torch.load(model_path, map_location='cpu'))
    model.to(torch.device('cuda:' + str(args.gpu)))
    model.eval()
    sample_fn = diffusion.p_sample_loop
    if tst_path is None:
        raise NotImplementedError("wav_path / txt_path: the audio + text feature extractors (process_TWH_bvh.load_audio, "
                                  "load_tsv) are not part of this engine; pass --tst_path with pre-extracted features")
    if dataset == 'TWH':
        from .process_beat_twh_bvh import load_metadata
        _, metadict_byfname, _ = load_metadata(os.path.join(tst_path, "metadata.csv"), "main-agent")
    out = []
    for filename in tst_prefix:
        print(f"Processing: {filename}")
        speaker = np.zeros([args.style_dim])
        if dataset == 'BEAT':
            speaker[speaker_id_dict[int(filename.split('_')[0])]] = 1
        else:
            speaker[metadict_byfname[filename][1]] = 1
        audio = np.load(os.path.join(tst_path, 'audio_' + dataset, filename + '.npy'))
        text = np.load(os.path.join(tst_path, 'text_' + dataset, filename + '.npy'))
        textaudio = torch.FloatTensor(np.concatenate((audio, text), axis=-1))
        out.append(inference(args, save_dir, filename, textaudio, sample_fn, model, n_frames=max_len, smoothing=True,
                             skip_timesteps=skip_timesteps, style=speaker, seed=123456, dataset=dataset))
    return out


def parse_cli(argv=None):
    parser = argparse.ArgumentParser(description='DiffuseStyleGesture')                      # sample.py:275-289
    parser.add_argument('--config', default=DEFAULT_CONFIG)
    parser.add_argument('--gpu', type=str, default='0')
    parser.add_argument('--tst_prefix', nargs='+')
    parser.add_argument('--no_cuda', type=list, default=['0'])
    parser.add_argument('--model_path', type=str, default='./model000450000.pt')
    parser.add_argument('--tst_path', type=str, default=None)
    parser.add_argument('--wav_path', type=str, default=None)
    parser.add_argument('--txt_path', type=str, default=None)
    parser.add_argument('--save_dir', type=str, default='sample_dir')
    parser.add_argument('--max_len', type=int, default=0)
    parser.add_argument('--skip_timesteps', type=int, default=0)
    parser.add_argument('--dataset', type=str, default='BEAT')
    parser.add_argument('--wavlm_path', type=str, default='./WavLM/WavLM-Large.pt')
    parser.add_argument('--word2vector_path', type=str, default='./crawl-300d-2M.vec')
    args = parser.parse_args(argv)
    with open(args.config) as f:
        config = yaml.safe_load(f)
    for k, v in vars(args).items():
        config[k] = v
    return resolve_presets(Config(config))


if __name__ == '__main__':
    config = parse_cli()
    torch.cuda.set_device(int(config.gpu))
    main(config, config.save_dir, config.model_path, tst_path=config.tst_path, max_len=config.max_len,
         skip_timesteps=config.skip_timesteps, tst_prefix=config.tst_prefix, dataset=config.dataset, wav_path=config.wav_path,
         txt_path=config.txt_path, wavlm_path=config.wavlm_path, word2vector_path=config.word2vector_path)
