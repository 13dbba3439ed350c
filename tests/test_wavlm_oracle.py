"""CPU: the WavLM-Large oracle (oracle/wavlm_oracle.py) against the golden vectors produced by the reference WavLM class
(oracle/gen_golden_wavlm.py), and the host mirror's parameter inventory / position-bias table."""
import os

import numpy as np
import pytest
import torch

from diffusestylegesture_b200.wavlm_config import WAVLM_LARGE, synthetic_wavlm_state_dict, synthetic_wav, wavlm_state_dict_spec
from oracle import wavlm_oracle as WO


@pytest.fixture(scope="module")
def weights():
    return synthetic_wavlm_state_dict(WAVLM_LARGE, seed=0)


def test_geometry_and_inventory(weights):
    g = WAVLM_LARGE
    assert g.frames(70400) == 219 and g.frames(16000) == 49
    spec = wavlm_state_dict_spec(g)
    assert len(spec) == 21 + 8 + 24 * 19 + 2
    assert sum(int(np.prod(s)) for _, s in spec) + 1024 == 315453120          # + mask_emb = WavLM-Large parameter count
    for n, s in spec:
        assert tuple(weights[n].shape) == tuple(s)


def test_oracle_matches_reference_golden(gold_dir, weights):
    gold = np.load(os.path.join(gold_dir, "wavlm_large.npz"))
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    taps = {}
    with torch.no_grad():
        out = WO.wav2wavlm(weights, WAVLM_LARGE, synthetic_wav(2, 70400), 88, taps)
    assert float(np.abs(out.numpy() - gold["out"]).max()) < 2e-4
    sub = lambda t: t[:, ::4, ::8].numpy()
    assert float(np.abs(sub(taps["conv0"][:, :2000]) - gold["conv0"]).max()) < 1e-4
    assert float(np.abs(taps["conv6"].numpy()[:, :, ::4] - gold["conv6"]).max()) < 1e-4
    for k in ("x_pos", "layer0", "layer11", "layer23"):
        assert float(np.abs(sub(taps[k]) - gold[k]).max()) < 5e-4, k


def test_host_position_bias_equals_oracle(weights):
    from diffusestylegesture_b200.wavlm import _relative_position_bucket
    L = 219
    rel = torch.arange(L)[None, :] - torch.arange(L)[:, None]
    assert torch.equal(_relative_position_bucket(rel, 320, 800), WO.relative_position_bucket(rel, 320, 800))
    b = WO.relative_position_bucket(torch.tensor([[-900, -80, -79, 0, 79, 80, 900]]))
    assert b.tolist() == [[159, 80, 79, 0, 239, 240, 319]]


def test_host_mirror_state_dict_and_no_cpu_path(weights):
    from diffusestylegesture_b200.wavlm import WavLM, WavLMConfig
    m = WavLM(WavLMConfig())
    assert set(m.state_dict()) == set(weights)
    m.load_state_dict(weights)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.extract_features(torch.zeros(1, 70400))
    with pytest.raises(NotImplementedError):
        WavLM(dict(encoder_layers=12))
