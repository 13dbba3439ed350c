"""CPU: the oracle against the golden vectors generated from the UNMODIFIED reference
(oracle/gen_golden.py; tests/golden/GOLDEN_REPORT.txt records the generation-time agreement)."""
import os

import numpy as np
import pytest
import torch

from diffusestylegesture_b200.config import ZEGGS
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
from oracle import dsg_oracle as O

SEED = 123456


@pytest.fixture(scope="module")
def sd():
    return synthetic_state_dict(ZEGGS, seed=0)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_noise_stream_moments():
    z = O.philox_normal(SEED, 3, 1, 17, 1141 * 88)
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    assert np.isfinite(z).all()
    z2 = O.philox_normal(SEED, 4, 1, 17, 1141 * 88)
    assert abs(float(np.corrcoef(z, z2)[0, 1])) < 0.02


def test_schedule_known_answers(gold_dir):
    gold = np.load(os.path.join(gold_dir, "schedule.npz"))
    for tag, resp in (("ddpm1000", None), ("ddpm50", [50]), ("ddim100", "ddim100")):
        s = O.Schedule(1000, resp)
        for k in ("betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                  "posterior_log_variance_clipped", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
            assert np.array_equal(getattr(s, k), gold[f"{tag}/{k}"]), (tag, k)
        assert list(gold[f"{tag}/timestep_map"]) == s.timestep_map
    s = O.Schedule(1000)
    # SURVEY.md section 3.3 known answers
    assert s.posterior_mean_coef1[0] == 1.0 and s.posterior_mean_coef2[0] == 0.0
    np.testing.assert_allclose(s.posterior_mean_coef1[500], 0.0043678669, rtol=1e-8)
    np.testing.assert_allclose(s.posterior_mean_coef2[999], 0.0316226999, rtol=1e-8)
    np.testing.assert_allclose(np.exp(0.5 * s.posterior_log_variance_clipped[999]), 0.999498662, rtol=1e-8)
    np.testing.assert_allclose(s.betas[0], 4.12842248e-05, rtol=1e-8)
    assert s.betas[999] == 0.999
    assert sorted(O.space_timesteps(1000, "ddim100")) == list(range(0, 1000, 10))
    k50 = sorted(O.space_timesteps(1000, [50]))
    assert k50[:4] == [0, 20, 41, 61] and k50[-3:] == [958, 979, 999] and len(k50) == 50


def test_mdm_forward_matches_reference(gold_dir, sd):
    gold = np.load(os.path.join(gold_dir, "mdm_forward_zeggs.npz"))
    g = ZEGGS
    y = synthetic_conditioning(g, 2, segment=0)
    y["seed"] = torch.from_numpy(gold["seed_pose"])
    x = O.noise_tensor(SEED, [0, 1], 0, 0, (g.njoints, 1, g.n_poses))
    with torch.no_grad():
        out = O.mdm_forward(sd, g, x, torch.from_numpy(gold["t"]), y)
    assert float((out - torch.from_numpy(gold["out"])).abs().max()) < 2e-5      # fp32 tolerance, |out| ~ 2.4


@pytest.mark.parametrize("tag,resp,sampler,skip", [("ddpm50", [50], "ddpm", 0), ("ddim100", "ddim100", "ddim", 0),
                                                   ("ddpm1000_skip950", None, "ddpm", 950)])
def test_sampling_loops_match_reference(gold_dir, sd, tag, resp, sampler, skip):
    gold = np.load(os.path.join(gold_dir, "loops_zeggs.npz"))[tag]
    g = ZEGGS
    y = synthetic_conditioning(g, 2, segment=0)
    with torch.no_grad():
        out, _ = O.p_sample_loop(sd, g, O.Schedule(1000, resp), y, 2, seed=SEED, segment=0, sampler=sampler,
                                 skip_timesteps=skip)
    assert float((out - torch.from_numpy(gold)).abs().max()) < 5e-5


def test_inference_and_bvh_match_reference(gold_dir, sd):
    gold = np.load(os.path.join(gold_dir, "inference_zeggs_50.npz"))
    st = np.load(os.path.join(gold_dir, "zeggs_mean_std.npz"))
    g = ZEGGS
    feats = [synthetic_conditioning(g, 1, segment=s)["audio"][0] for s in range(2)]
    with torch.no_grad():
        seq = O.inference_clip(sd, g, O.Schedule(1000, [50]), feats, torch.tensor(gold["style"], dtype=torch.float32),
                               seed=SEED, clip_id=0)
    poses = O.denormalise(seq.numpy(), st["mean"], st["std"])
    assert np.abs(poses - gold["poses"]).max() < 5e-4
    pos, eul = O.pose2bvh_values(gold["poses"].astype(np.float64), int(gold["length"]), smoothing=True)
    assert np.abs(pos - gold["positions"]).max() < 1e-4          # cm
    assert np.abs(eul - gold["rotations"]).max() < 2e-2          # degrees (poses stored as fp32)


def test_bvh_tail_full_clip_golden(gold_dir):
    gold = np.load(os.path.join(gold_dir, "inference_zeggs_1000.npz"))
    assert gold["poses"].shape == (312, 1141) and int(gold["length"]) == 312
    pos, eul = O.pose2bvh_values(gold["poses"].astype(np.float64), 312, smoothing=True)
    assert pos.shape == (936, 75, 3)
    assert np.abs(pos - gold["positions"]).max() < 1e-4
    assert np.abs(eul - gold["rotations"]).max() < 2e-2


@pytest.mark.parametrize("tag", ["beat", "twh"])
def test_plus_variant_matches_reference(gold_dir, tag):
    """DiffuseStyleGesture+ (BEAT-TWH-main/model/mdm.py:187-224, cross_local_attention4) — oracle vs reference golden."""
    from diffusestylegesture_b200.config import BEAT_PLUS, TWH_PLUS
    g = BEAT_PLUS if tag == "beat" else TWH_PLUS
    gold = np.load(os.path.join(gold_dir, "beat_twh_plus.npz"))
    sdg = synthetic_state_dict(g, seed=0)
    y = synthetic_conditioning(g, 2, segment=0)
    seed_pose = 0.5 * O.noise_tensor(SEED, [0, 1], 7, 99, (g.njoints, 1, g.n_seed))
    assert np.array_equal(seed_pose[:, ::4].numpy(), gold[f"{tag}/seed_pose_sub"])
    y["seed"] = seed_pose
    x = O.noise_tensor(SEED, [0, 1], 0, 0, (g.njoints, 1, g.n_poses))
    with torch.no_grad():
        out = O.mdm_forward(sdg, g, x, torch.from_numpy(gold[f"{tag}/t"]), y)
    assert float((out[:, ::4] - torch.from_numpy(gold[f"{tag}/out_sub"])).abs().max()) < 2e-5
