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
