"""GPU (B200): the WavLM-Large conditioning forward (dsg_wavlm_*: tcgen05 conv/linear GEMMs, mma.sync flash attention with
the gated relative-position bias) against the reference golden vectors (tests/golden/wavlm_large.npz, generated from the
reference WavLM class) and the oracle.  bf16 operands, fp32 residual stream: stated tolerance on the interpolated
[B, 88, 1024] features (values are LayerNorm outputs, |x| <= ~4.5): max |err| < 0.08, rms < 0.015 (measured 0.026 / 0.0059)."""
import os

import numpy as np
import pytest
import torch

from diffusestylegesture_b200.wavlm import WavLM
from diffusestylegesture_b200.wavlm_config import WAVLM_LARGE, synthetic_wavlm_state_dict, synthetic_wav

pytestmark = pytest.mark.gpu


def test_wav2wavlm_vs_reference_golden(gold_dir):
    gold = np.load(os.path.join(gold_dir, "wavlm_large.npz"))
    m = WavLM(max_batch=2)
    m.load_state_dict(synthetic_wavlm_state_dict(WAVLM_LARGE, seed=0))
    m.to('cuda:0').eval()
    wav = synthetic_wav(2, 70400)
    out = m.wav2wavlm(wav, 88).cpu()                      # host waveform in, device features out
    assert out.shape == (2, 88, 1024)
    d = (out.double() - torch.from_numpy(gold["out"]).double())
    mx, rms = float(d.abs().max()), float(d.pow(2).mean().sqrt())
    print(f"wavlm features vs reference: max {mx:.3g} rms {rms:.3g} (|ref|max {np.abs(gold['out']).max():.3g})")
    assert np.isfinite(mx) and mx < 0.08 and rms < 0.015
    feats, pm = m.extract_features(wav.cuda())
    assert feats.shape == (2, 219, 1024) and pm is None
    # sub-batching: 3 clips through max_batch = 2 equals the per-clip results
    wav3 = torch.cat([wav, wav[:1]])
    out3 = m.wav2wavlm(wav3, 88).cpu()
    assert float((out3[:2] - out).abs().max()) == 0.0 and float((out3[2] - out[0]).abs().max()) == 0.0
    assert m.launches > 0
    with pytest.raises(NotImplementedError):
        m.extract_features(wav, mask=True)


def test_wav_to_bvh_through_both_engines(tmp_path):
    """Raw 16 kHz waveform -> WavLM (libdsg) -> 2 segments x 20-step DDPM (libdsg) -> BVH, the reference `inference` call
    (sample.py:210-338) with every device op in the engine; batched segment conditioning equals per-segment calls."""
    from diffusestylegesture_b200 import sample as S
    from diffusestylegesture_b200.config import ZEGGS
    from diffusestylegesture_b200.mdm import MDM
    from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
    from diffusestylegesture_b200.synthetic import synthetic_state_dict
    wm = WavLM(max_batch=4)
    wm.load_state_dict(synthetic_wavlm_state_dict(WAVLM_LARGE, seed=0))
    wm.to('cuda:0').eval()
    model = MDM(njoints=ZEGGS.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=8, precision='bf16', max_batch=1)
    load_model_wo_clip(model, synthetic_state_dict(ZEGGS, seed=0))
    model.to('cuda:0').eval()
    d = create_gaussian_diffusion([20])
    audio = synthetic_wav(1, 160 * 800 + 1234)[0].numpy()             # a little more than 2 strides of 80 frames
    path, poses = S.inference(S.Config(n_poses=88, audio_feat="wavlm"), wm, audio, d.p_sample_loop, model, n_frames=0, smoothing=True,
                              SG_filter=True, minibatch=True, style=[1, 0, 0, 0, 0, 0], seed=7, save_dir=str(tmp_path))
    assert poses.shape == (160 - 8, 1141) and np.isfinite(poses).all()
    lines = open(path).read().splitlines()
    assert lines[0] == "HIERARCHY" and any(l.startswith("Frames: 456") for l in lines)      # 152 frames at 20 fps -> 60 fps
    # batched conditioning == per-segment conditioning (reference loop order, sample.py:238-251)
    a = torch.from_numpy(audio[:160 * 800]).reshape(2, 64000)
    w0 = torch.cat((torch.zeros(6400), a[0]))[None]
    w1 = torch.cat((a[0, -6400:], a[1]))[None]
    f01 = S.wav2wavlm(wm, torch.cat((w0, w1)), 'cuda:0', 88)
    assert torch.equal(f01[0], S.wav2wavlm(wm, w0, 'cuda:0', 88)[0]) and torch.equal(f01[1], S.wav2wavlm(wm, w1, 'cuda:0', 88)[0])


def test_batch_manifest_matches_single_clip_runs(tmp_path):
    """`sample.main_batch` (clips of different lengths, batched per segment count) gives every clip the result of running it
    alone with the same clip id: batching changes nothing (counter-based noise; row-independent kernels)."""
    from diffusestylegesture_b200 import sample as S
    from diffusestylegesture_b200.config import ZEGGS
    from diffusestylegesture_b200.mdm import MDM
    from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
    from diffusestylegesture_b200.synthetic import synthetic_state_dict
    wm = WavLM(max_batch=4)
    wm.load_state_dict(synthetic_wavlm_state_dict(WAVLM_LARGE, seed=0))
    wm.to('cuda:0').eval()
    model = MDM(njoints=ZEGGS.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=8, precision='bf16', max_batch=2)
    load_model_wo_clip(model, synthetic_state_dict(ZEGGS, seed=0))
    model.to('cuda:0').eval()
    d = create_gaussian_diffusion([20])
    wavs = synthetic_wav(3, 170 * 800).numpy()
    rows = [{'wav': 'a_Happy_0.wav', 'style': [1, 0, 0, 0, 0, 0], 'style_name': 'Happy', 'clip_id': 0, 'audio': wavs[0]},
            {'wav': 'b_Old_0.wav', 'style': [0, 0, 0, 1, 0, 0], 'style_name': 'Old', 'clip_id': 5, 'audio': wavs[1][:100 * 800]},
            {'wav': 'c_Sad_0.wav', 'style': [0, 1, 0, 0, 0, 0], 'style_name': 'Sad', 'clip_id': 9, 'audio': wavs[2]}]
    paths = S.main_batch(S.Config(n_poses=88, audio_feat="wavlm", gpu="0", max_batch=2), str(tmp_path), None, rows,
                         wavlm_model=wm, model=model, diffusion=d, seed=7)
    assert len(paths) == 3 and all(os.path.getsize(p) > 1e5 for p in paths)
    frames = [int([l for l in open(p).read().splitlines() if l.startswith("Frames:")][0].split()[1]) for p in paths]
    assert frames == [456, 216, 456]                                  # (160 - 8) * 3, (80 - 8) * 3
    # clip 2 alone, same clip id
    w, n_frames = S.segment_windows(wavs[2], 170, 88, 8)
    feats = [S.wav2wavlm(wm, w[s:s + 1], 'cuda:0', 88) for s in range(2)]
    seq = S.inference_batch(model, d, feats, torch.tensor([[0, 1, 0, 0, 0, 0]], dtype=torch.float32), seed=7, clip_ids=[9], smoothing=True)
    single = str(tmp_path / "single.bvh")
    S.pose2bvh(S.denormalise(seq[0].numpy()), single, length=n_frames - 8, smoothing=True)
    assert open(single).read() == open(paths[2]).read()
