"""GPU (B200): the tcgen05 tensor-core path (DSG_PRECISION_BF16) — bf16 operands, fp32 accumulate, fp32
residual stream / LayerNorm / posterior.

Stated tolerances (normalised motion is O(1), |x| <= ~2.5; the oracle / reference golden is fp32):
  tcgen05 GEMM building block vs fp64 on bf16-rounded operands : 2e-3 relative to the row scale (fp32 accumulate order)
  one denoiser call (8 layers)                                  : max |err| < 0.03, rms < 0.006   (measured 0.0092 / 0.0020)
  sampling loops (50 / 100 / 1000 steps)                        : max |err| < 0.05, rms < 0.008   (measured 0.014 / 0.0030)
  final BVH of the 320-frame 1000-step clip                     : positions < 0.4 cm, Euler angles < 1.0 degree
                                                                  (measured 0.147 cm, 0.31 deg max, 0.014 deg mean)
The measured values are printed on every run.
"""
import os

import numpy as np
import pytest
import torch

from diffusestylegesture_b200.config import ZEGGS
from diffusestylegesture_b200.engine import Engine, selftest_gemm
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
from diffusestylegesture_b200 import sample as S
from diffusestylegesture_b200 import process_zeggs_bvh as PB
from oracle import dsg_oracle as O

pytestmark = pytest.mark.gpu
SEED = 123456
G = ZEGGS


def _bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float64).numpy()


@pytest.mark.parametrize("bn,M,N,K", [(128, 300, 200, 256), (128, 128, 128, 64), (256, 300, 512, 1152), (256, 1000, 256, 1024),
                                      (128, 89, 1141, 256)])
def test_tcgen05_gemm_building_block(bn, M, N, K):
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C = selftest_gemm(A, W, bias, bn=bn)
    want = _bf16_round(A) @ _bf16_round(W).T + bias.astype(np.float64)
    err = np.abs(C - want).max()
    print(f"tcgen05 gemm bn={bn} {M}x{N}x{K}: max err {err:.3g}")
    assert err < 2e-3, err


@pytest.fixture(scope="module")
def sd():
    return synthetic_state_dict(G, seed=0)


def _model(sd, max_batch=4):
    m = MDM(njoints=G.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=G.n_seed,
            precision="bf16", max_batch=max_batch)
    load_model_wo_clip(m, sd)
    return m.to('cuda:0').eval()


def _err(a, b):
    d = (torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu())
    return float(d.abs().max()), float(d.pow(2).mean().sqrt())


def test_denoiser_bf16_vs_reference_golden(gold_dir, sd):
    gold = np.load(os.path.join(gold_dir, "mdm_forward_zeggs.npz"))
    eng = Engine(G, sd, device=0, max_batch=2, precision="bf16")
    y = synthetic_conditioning(G, 2, segment=0)
    y["seed"] = torch.from_numpy(gold["seed_pose"])
    x = O.noise_tensor(SEED, [0, 1], 0, 0, (G.njoints, 1, G.n_poses))
    eng.debug_enable()
    eng.set_conditioning(y["style"], y["seed"], y["audio"])
    out = eng.denoise(x, gold["t"])
    for tap, tol in (("tok", 1e-4), ("h_in", 0.01), ("xs0", 0.01), ("xs1", 0.03), ("xs8", 0.05)):
        got = eng.debug_read(tap, 2)
        mx, rms = _err(got, gold["tap_" + tap])
        print(f"bf16 tap {tap}: max {mx:.3g} rms {rms:.3g}")
        assert mx < tol, (tap, mx)
    mx, rms = _err(out, gold["out"])
    print(f"bf16 denoiser output: max {mx:.3g} rms {rms:.3g}")
    assert mx < 0.03 and rms < 0.006
    eng.close()


@pytest.mark.parametrize("tag,resp,sampler,skip", [("ddpm50", [50], "ddpm", 0), ("ddim100", "ddim100", "ddim", 0),
                                                   ("ddpm1000_skip950", '', "ddpm", 950)])
def test_sampling_loops_bf16_vs_reference_golden(gold_dir, sd, tag, resp, sampler, skip):
    gold = np.load(os.path.join(gold_dir, "loops_zeggs.npz"))[tag]
    model = _model(sd)
    d = create_gaussian_diffusion(resp)
    y = synthetic_conditioning(G, 2, segment=0)
    y.update(noise_seed=SEED, segment=0)
    fn = d.p_sample_loop if sampler == "ddpm" else d.ddim_sample_loop
    out = fn(model, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, skip_timesteps=skip)
    mx, rms = _err(out, gold)
    print(f"bf16 loop {tag}: max {mx:.3g} rms {rms:.3g}")
    assert mx < 0.05 and rms < 0.008, tag
    # graph replay is deterministic and reusable across calls
    out2 = fn(model, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, skip_timesteps=skip)
    assert torch.equal(out, out2)


def test_bf16_matches_fp32_engine_and_sharding(sd):
    m16 = _model(sd, max_batch=4)
    d = create_gaussian_diffusion([20])
    y4 = synthetic_conditioning(G, 4, segment=0)
    shp = (4, G.njoints, 1, G.n_poses)
    full = d.p_sample_loop(m16, shp, clip_denoised=False,
                           model_kwargs={'y': dict(y4, noise_seed=SEED, segment=0, clip_ids=[0, 1, 2, 3])})
    ys = {k: (v[2:4] if isinstance(v, torch.Tensor) and v.shape[0] == 4 else v) for k, v in y4.items()}
    part = d.p_sample_loop(m16, (2,) + shp[1:], clip_denoised=False,
                           model_kwargs={'y': dict(ys, noise_seed=SEED, segment=0, clip_ids=[2, 3])})
    mx, _ = _err(part, full[2:4])
    assert mx < 1e-5, mx          # same clips, different batch composition: identical arithmetic per clip


@pytest.mark.slow
def test_full_clip_1000_steps_bf16_bvh_vs_reference_golden(gold_dir, sd):
    """BASELINE.json configs[1]: 320-frame clip, 4 segments x 1000 DDPM steps, bf16 — final BVH joint values."""
    gold = np.load(os.path.join(gold_dir, "inference_zeggs_1000.npz"))
    st = np.load(os.path.join(gold_dir, "zeggs_mean_std.npz"))
    model = _model(sd, max_batch=1)
    d = create_gaussian_diffusion()
    feats = [synthetic_conditioning(G, 1, segment=s)["audio"] for s in range(4)]
    seq = S.inference_batch(model, d, feats, torch.tensor([list(gold["style"])], dtype=torch.float32), seed=SEED)
    poses = O.denormalise(seq[0].numpy(), st["mean"], st["std"])
    pos, eul = PB.pose2bvh_arrays(poses, 312, smoothing=True)
    d_pos = np.abs(pos - gold["positions"]).max()
    d_eul = np.abs((eul - gold["rotations"] + 180.0) % 360.0 - 180.0)
    print(f"bf16 full clip: poses max err {np.abs(poses - gold['poses']).max():.3g}; BVH positions {d_pos:.3g} cm; "
          f"Euler max {d_eul.max():.3g} deg, mean {d_eul.mean():.3g} deg")
    assert d_pos < 0.4 and d_eul.max() < 1.0


def test_clip_kernel_matches_multikernel_path(sd):
    """The persistent per-clip kernel (default) against the multi-kernel graph path (DSG_TC_MODE=kernels): same bf16
    GEMM operands; the clip kernel additionally keeps the residual stream in bf16 on chip.  Per-layer taps of the last
    step and the final sample are compared; 148+ clips also exercises CTAs that run more than one clip."""
    d = create_gaussian_diffusion([6])
    B = 3
    y = synthetic_conditioning(G, B, segment=0)
    shp = (B, G.njoints, 1, G.n_poses)
    res = {}
    for mode in ("kernels", "clip"):
        os.environ["DSG_TC_MODE"] = mode
        eng_model = _model(sd, max_batch=B)
        eng = eng_model.get_engine(B)
        eng.debug_enable()
        out = d.p_sample_loop(eng_model, shp, clip_denoised=False, model_kwargs={'y': dict(y, noise_seed=SEED, segment=0)})
        res[mode] = (out.cpu(), {k: eng.debug_read(k, B) for k in ("xs0", "xs1", "xs4", "xs8")})
    os.environ.pop("DSG_TC_MODE")
    for k in ("xs0", "xs1", "xs4", "xs8"):
        mx, rms = _err(res["clip"][1][k], res["kernels"][1][k])
        print(f"clip vs kernels tap {k}: max {mx:.3g} rms {rms:.3g}")
        assert mx < 0.08 and rms < 0.01, k
    mx, rms = _err(res["clip"][0], res["kernels"][0])
    print(f"clip vs kernels final sample: max {mx:.3g} rms {rms:.3g}")
    assert mx < 0.05 and rms < 0.008
    want, _ = O.p_sample_loop(sd, G, O.Schedule(1000, [6]), y, B, seed=SEED, segment=0)
    mx, rms = _err(res["clip"][0], want)
    print(f"clip vs oracle final sample: max {mx:.3g} rms {rms:.3g}")
    assert mx < 0.05 and rms < 0.008


def test_clip_kernel_many_clips_per_cta(sd, monkeypatch):
    """More clips than SMs: CTAs loop over several clips; results must equal a run where every clip has its own CTA.
    (Both runs in the one-CTA-per-clip mode: the CTA-pair mode small batches select sums linear2 in another order.)"""
    monkeypatch.setenv("DSG_CLIP_PAIR", "0")
    d = create_gaussian_diffusion([3])
    B = 150
    y = synthetic_conditioning(G, B, segment=0)
    m = _model(sd, max_batch=B)
    full = d.p_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False,
                           model_kwargs={'y': dict(y, noise_seed=SEED, segment=0, clip_ids=list(range(B)))})
    ys = {k: (v[147:150] if isinstance(v, torch.Tensor) and v.shape[0] == B else v) for k, v in y.items()}
    part = d.p_sample_loop(m, (3, G.njoints, 1, G.n_poses), clip_denoised=False,
                           model_kwargs={'y': dict(ys, noise_seed=SEED, segment=0, clip_ids=[147, 148, 149])})
    assert _err(part, full[147:150])[0] < 1e-5


def test_clip_pair_mode_matches_single_cta_mode(sd, monkeypatch):
    """Batches of at most SMs / 2 clips run a CLUSTER of two CTAs per clip (dsg_clip_kernel.cuh, CL = 2: heads, FFN chunks and
    pose-head tiles split over the pair, exchanges through distributed shared memory).  Same arithmetic except for the order of
    the linear2 sum and its bf16 hand-over: against the one-CTA mode and the fp32 oracle within the loop tolerance, per-layer
    taps included; bitwise reproducible; independent of the batch composition; the largest pair batch (74 clips) included."""
    d = create_gaussian_diffusion([6])
    B = 3
    y = synthetic_conditioning(G, B, segment=0)
    shp = (B, G.njoints, 1, G.n_poses)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DSG_CLIP_PAIR", mode)
        m = _model(sd, max_batch=B)
        eng = m.get_engine(B)
        eng.debug_enable()
        out = d.p_sample_loop(m, shp, clip_denoised=False, model_kwargs={'y': dict(y, noise_seed=SEED, segment=0)})
        res[mode] = (out.cpu(), {k: eng.debug_read(k, B) for k in ("xs0", "xs1", "xs4", "xs8")})
    for k in ("xs0", "xs1", "xs4", "xs8"):
        mx, rms = _err(res["1"][1][k], res["0"][1][k])
        print(f"pair vs single tap {k}: max {mx:.3g} rms {rms:.3g}")
        assert mx < 0.08 and rms < 0.01, k
    want, _ = O.p_sample_loop(sd, G, O.Schedule(1000, [6]), y, B, seed=SEED, segment=0)
    for mode in ("0", "1"):
        mx, rms = _err(res[mode][0], want)
        print(f"{'pair' if mode == '1' else 'single'} vs oracle, 6 steps: max {mx:.3g} rms {rms:.3g}")
        assert mx < 0.05 and rms < 0.008
    # 40 steps, 74 clips (148 CTAs): reproducible, finite, equal to the same clips run as a batch of 2, close to the one-CTA mode
    d = create_gaussian_diffusion([40])
    B = 74
    y = synthetic_conditioning(G, B, segment=1)
    m = _model(sd, max_batch=B)
    kw = lambda yy, ids: {'y': dict(yy, noise_seed=SEED, segment=1, clip_ids=ids)}
    monkeypatch.setenv("DSG_CLIP_PAIR", "1")
    a = d.p_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs=kw(y, list(range(B)))).clone()
    b = d.p_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs=kw(y, list(range(B)))).clone()
    assert bool(torch.isfinite(a).all()) and torch.equal(a, b)
    ys = {k: (v[72:74] if isinstance(v, torch.Tensor) and v.shape[0] == B else v) for k, v in y.items()}
    part = d.p_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs=kw(ys, [72, 73]))
    assert torch.equal(part, a[72:74])
    monkeypatch.setenv("DSG_CLIP_PAIR", "0")
    single = d.p_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs=kw(y, list(range(B))))
    mx, rms = _err(a, single)
    print(f"pair vs single, 74 clips x 40 steps: max {mx:.3g} rms {rms:.3g}")
    assert mx < 0.05 and rms < 0.008


def test_clip_kernel_is_bitwise_reproducible(sd):
    """Race detector: the persistent kernel synchronises 16 warps through ~40 mbarriers, named barriers and proxy fences
    with no host involvement; any missing edge shows up as run-to-run differences.  Same inputs -> identical bits, for
    one clip per CTA, several clips per CTA, and a single clip (different relative timing of the roles)."""
    d = create_gaussian_diffusion([40])
    for B in (148, 150, 1):                 # (B = 1 runs the CTA-pair mode)
        y = synthetic_conditioning(G, B, segment=1)
        m = _model(sd, max_batch=B)
        outs = []
        for _ in range(3):
            outs.append(d.p_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False,
                                        model_kwargs={'y': dict(y, noise_seed=SEED, segment=1, clip_ids=list(range(B)))}).clone())
        assert bool(torch.isfinite(outs[0]).all())
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), f"B={B}: results differ between identical runs"
