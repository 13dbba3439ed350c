"""GPU (B200): round-2 rows — the "+" / "++" denoisers on the tcgen05 path (D = 384 / 512), the remaining sampler options
(const_noise, dump_steps, PLMS), BASELINE config 3 (B = 64, DDIM-100, six styles) and the BEAT-TWH `inference` driver.

Tolerances (fp32 reference golden, normalised motion |x| <= ~3):
  fp32 engine                       : loops 2e-3 (as in test_gpu_parity.py)
  bf16 engine, one denoiser call    : max 0.03 / rms 0.006 (D = 256), max 0.05 / rms 0.008 (D = 384 / 512: K and J are 1.5-2x larger)
  bf16 engine, loops                : max 0.05 / rms 0.008 (ZEGGS), max 0.08 / rms 0.012 ("+" geometries)
"""
import os

import numpy as np
import pytest
import torch

from diffusestylegesture_b200.config import ZEGGS, BEAT_PLUS, TWH_PLUS, BEAT_PLUSPLUS
from diffusestylegesture_b200.engine import Engine
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
from oracle import dsg_oracle as O

pytestmark = pytest.mark.gpu
SEED = 123456
G = ZEGGS


def _err(a, b):
    d = (torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu())
    return float(d.abs().max()), float(d.pow(2).mean().sqrt())


def _zeggs_model(precision, max_batch=2):
    m = MDM(njoints=G.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=G.n_seed,
            precision=precision, max_batch=max_batch)
    load_model_wo_clip(m, synthetic_state_dict(G, seed=0))
    return m.to('cuda:0').eval()


def _plus_model(g, precision, max_batch=2):
    mode = {4: 'cross_local_attention4_style1_sample', 5: 'cross_local_attention5_style1_sample'}[g.variant]
    m = MDM(njoints=g.njoints, cond_mode=mode, audio_feat='wavlm', n_seed=g.n_seed, latent_dim=g.latent_dim,
            style_dim=g.style_in, source_audio_dim=g.audio_dim, audio_feat_dim_latent=g.audio_latent, precision=precision,
            max_batch=max_batch)
    load_model_wo_clip(m, synthetic_state_dict(g, seed=0))
    return m.to('cuda:0').eval()


# ------------------------------------------------------------------------------------------------ a18: "+" on tcgen05
@pytest.mark.parametrize("tag", ["beat", "twh"])
def test_plus_variant_bf16_tcgen05_vs_reference_golden(gold_dir, tag):
    """DiffuseStyleGesture+ (D = 384 / 512, T = 150, S = 151, window 15, J = 2052 / 2232) on the tensor-core engine: every
    Linear a tcgen05 GEMM (LayerNorm GEMMs as one 384- / 512-column tile), attention on mma.sync — against the BEAT-TWH
    reference golden (BEAT-TWH-main/model/mdm.py:187-224; every 4th channel is stored)."""
    g = BEAT_PLUS if tag == "beat" else TWH_PLUS
    gold = np.load(os.path.join(gold_dir, "beat_twh_plus.npz"))
    sdg = synthetic_state_dict(g, seed=0)
    e = Engine(g, sdg, device=0, max_batch=2, precision="bf16")
    y = synthetic_conditioning(g, 2, segment=0)
    y["seed"] = 0.5 * O.noise_tensor(SEED, [0, 1], 7, 99, (g.njoints, 1, g.n_seed))
    x = O.noise_tensor(SEED, [0, 1], 0, 0, (g.njoints, 1, g.n_poses))
    e.set_conditioning(y["style"], y["seed"], y["audio"])
    l0 = e.launches
    out = e.denoise(x, gold[f"{tag}/t"])
    assert e.launches > l0
    mx, rms = _err(out[:, ::4], gold[f"{tag}/out_sub"])
    print(f"{tag}+ bf16 denoiser: max {mx:.3g} rms {rms:.3g}")
    assert mx < 0.05 and rms < 0.008
    e.close()
    m = _plus_model(g, "bf16")
    d = create_gaussian_diffusion([20])
    yy = dict(y, noise_seed=SEED, segment=0)
    loop = d.p_sample_loop(m, (2, g.njoints, 1, g.n_poses), clip_denoised=False, model_kwargs={'y': yy})
    mx, rms = _err(loop[:, ::4], gold[f"{tag}/loop20_sub"])
    print(f"{tag}+ bf16 ddpm20 loop: max {mx:.3g} rms {rms:.3g}")
    assert mx < 0.08 and rms < 0.012
    loop2 = d.p_sample_loop(m, (2, g.njoints, 1, g.n_poses), clip_denoised=False, model_kwargs={'y': yy})
    assert torch.equal(loop, loop2)                  # graph replay: deterministic


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-4), ("bf16", 0.05)])
def test_plusplus_attn5_vs_reference_golden(gold_dir, precision, tol):
    """cross_local_attention5 ("++", BEAT-TWH-main/model/mdm.py:226-264): y['seed_last'] embedded per frame behind the audio."""
    g = BEAT_PLUSPLUS
    gold = np.load(os.path.join(gold_dir, "r2_beat.npz"))
    y = synthetic_conditioning(g, 2, segment=0)
    y["seed"] = 0.5 * O.noise_tensor(SEED, [0, 1], 7, 99, (g.njoints, 1, g.n_seed))
    y["seed_last"] = 0.5 * O.noise_tensor(SEED, [0, 1], 7, 98, (g.njoints, 1, g.n_seed))
    assert np.allclose(y["seed_last"][:, ::4].numpy(), gold["pp/seed_last_sub"])
    x = O.noise_tensor(SEED, [0, 1], 0, 0, (g.njoints, 1, g.n_poses))
    m = _plus_model(g, precision)
    out = m(x, torch.from_numpy(gold["pp/t"]), y=y)
    mx, rms = _err(out[:, ::4], gold["pp/out_sub"])
    print(f"beat++ {precision} denoiser: max {mx:.3g} rms {rms:.3g}")
    assert mx < tol
    d = create_gaussian_diffusion([20])
    loop = d.p_sample_loop(m, (2, g.njoints, 1, g.n_poses), clip_denoised=False, model_kwargs={'y': dict(y, noise_seed=SEED, segment=0)})
    mx, rms = _err(loop[:, ::4], gold["pp/loop20_sub"])
    print(f"beat++ {precision} ddpm20 loop: max {mx:.3g} rms {rms:.3g}")
    assert mx < (2e-3 if precision == "fp32" else 0.08)
    with pytest.raises(RuntimeError, match="seed_last"):
        m(x, torch.from_numpy(gold["pp/t"]), y={k: v for k, v in y.items() if k != "seed_last"})


# ------------------------------------------------------------------------------------------------ f4: sampler options
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 0.05)])
def test_const_noise_and_dump_steps_vs_reference_golden(gold_dir, precision, tol):
    """p_sample_loop(const_noise=True, dump_steps=[0, 10, 49]) (gaussian_diffusion.py:544-545, 647-669): returns the list of
    dumped samples; every clip is noised with clip 0's per-step noise.  bf16 runs the persistent clip kernel in three launches."""
    gold = np.load(os.path.join(gold_dir, "r2_zeggs.npz"))
    m = _zeggs_model(precision)
    d = create_gaussian_diffusion([50])
    y = synthetic_conditioning(G, 2, segment=0)
    y.update(noise_seed=SEED, segment=0)
    steps = [int(i) for i in gold["opt_dump_steps"]]
    dump = d.p_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, dump_steps=steps,
                           const_noise=True)
    assert isinstance(dump, list) and len(dump) == len(steps)
    for i, (got, want) in enumerate(zip(dump, gold["opt_dump"])):
        mx, rms = _err(got[:, ::2], want)
        print(f"{precision} const_noise dump[{steps[i]}]: max {mx:.3g} rms {rms:.3g}")
        assert mx < tol, (steps[i], mx)
    # const_noise really changes the result, dump_steps does not: the last dump of a run without const_noise is the plain loop
    plain = d.p_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y})
    cut = d.p_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, dump_steps=[7, 49])
    assert torch.equal(cut[-1], plain)
    assert _err(dump[-1][1], plain[1])[0] > 0.05 and _err(dump[-1][0], plain[0])[0] < 1e-6      # clip 0 keeps its own noise
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, dump_steps=[1])


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-3), ("bf16", 0.06)])
@pytest.mark.parametrize("order", [2, 3])
def test_plms_sample_loop_vs_reference_golden(gold_dir, precision, tol, order):
    """plms_sample_loop (gaussian_diffusion.py:1005-1200): pseudo improved Euler first step (two model calls), then
    Adams-Bashforth of the given order on the eps re-derived from the predicted x_start."""
    gold = np.load(os.path.join(gold_dir, "r2_zeggs.npz"))
    m = _zeggs_model(precision)
    d = create_gaussian_diffusion([50])
    y = synthetic_conditioning(G, 2, segment=0)
    y.update(noise_seed=SEED, segment=0)
    out = d.plms_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, order=order)
    mx, rms = _err(out[:, ::2], gold[f"plms{order}"])
    print(f"{precision} plms order {order}: max {mx:.3g} rms {rms:.3g}")
    assert mx < tol
    with pytest.raises(TypeError):
        d.plms_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, order=1)
    with pytest.raises(ValueError):
        d.plms_sample_loop(m, (2, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y}, order=5)


# ------------------------------------------------------------------------------------------------ BASELINE config 3
def test_config3_b64_ddim100_six_styles_vs_reference_golden(gold_dir):
    """BASELINE.json configs[2]: batch = 64 ZEGGS clips, style one-hot i mod 6, DDIM-100 (eta 0), clip kernel (bf16) —
    all 64 clips against the reference's own ddim_sample_loop (every 8th channel stored in fp16, per-clip sums in fp64)."""
    gold = np.load(os.path.join(gold_dir, "r2_zeggs.npz"))
    B = 64
    m = _zeggs_model("bf16", max_batch=B)
    d = create_gaussian_diffusion("ddim100")
    y = synthetic_conditioning(G, B, segment=0)
    assert set(int(i) for i in y["style"].argmax(1)) == set(range(6))
    y.update(noise_seed=SEED, segment=0, clip_ids=list(range(B)))
    out = d.ddim_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False, model_kwargs={'y': y})
    want = torch.from_numpy(gold["c3_sub"].astype(np.float32))
    err = (out[:, ::8].cpu() - want).abs()
    per_clip = err.amax(dim=(1, 2, 3))
    rms = float(err.pow(2).mean().sqrt())
    print(f"config 3: worst clip max err {float(per_clip.max()):.3g} (clip {int(per_clip.argmax())}), rms {rms:.3g}")
    assert float(per_clip.max()) < 0.05 and rms < 0.008
    # size-independent property over the FULL tensors: per-clip mean of the sample (fp64 sums from the reference)
    mean_err = (out.double().sum(dim=(1, 2, 3)).cpu().numpy() - gold["c3_sum"]) / (G.njoints * G.n_poses)
    assert np.abs(mean_err).max() < 2e-3, np.abs(mean_err).max()


# ------------------------------------------------------------------------------------------------ a18: BEAT-TWH inference driver
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_beat_plus_inference_driver_vs_reference_golden(gold_dir, precision):
    """BEAT-TWH-main/mydiffusion_beat_twh/sample.py:44-201 through the host mirror: 900 frames of features -> ceil(900/120) = 8
    segments x 50 DDPM steps, velocity/acceleration seed, 1/2-1/2 first-frame blend, first J/3 channels, de-normalise —
    against the reference's own `inference` (poses as handed to pose2bvh_bugfix)."""
    from diffusestylegesture_b200 import sample_beat_twh as SB
    gold = np.load(os.path.join(gold_dir, "r2_beat.npz"))
    g = BEAT_PLUS
    m = _plus_model(g, precision, max_batch=1)
    d = create_gaussian_diffusion([50])
    gen = torch.Generator().manual_seed(int(gold["inf/textaudio_seed"][0]))
    textaudio = torch.randn(900, g.audio_dim, generator=gen)
    args = SB.Config(dict(n_poses=g.n_poses, n_seed=g.n_seed, version="v0", name="DiffuseStyleGesture+"))
    poses = SB.inference_batch_beat(args, textaudio, d.p_sample_loop, m, gold["inf/style"][None], gold["inf/seed_raw"][None],
                                    seed=SEED, dataset='BEAT')[0]
    assert poses.shape == gold["inf/poses"].shape == (900, g.njoints // 3)
    mean, std = SB.load_stats('BEAT')
    err_n = np.abs((poses - gold["inf/poses"]) / std)          # error in normalised units (std spans 1e-6 .. 0.53)
    mx, rms = float(err_n.max()), float(np.sqrt((err_n ** 2).mean()))
    print(f"beat+ inference {precision}: normalised max {mx:.3g} rms {rms:.3g}; de-normalised max {np.abs(poses - gold['inf/poses']).max():.3g}")
    # bf16: 8 chained segments (each seeded by the previous one's last 30 frames) x 50 steps; measured max 0.119 / rms 0.0042
    assert mx < (5e-3 if precision == "fp32" else 0.25) and rms < (1e-3 if precision == "fp32" else 0.01)


def test_beat_plus_batch_equals_single_clips():
    """Config 4 shape (batch of long-form clips): a batch of 3 clips with clip ids 5, 6, 7 == the three clips run alone."""
    from diffusestylegesture_b200 import sample_beat_twh as SB
    g = BEAT_PLUS
    m = _plus_model(g, "bf16", max_batch=3)
    d = create_gaussian_diffusion([4])
    gen = torch.Generator().manual_seed(5)
    ta = torch.randn(3, 300, g.audio_dim, generator=gen)                      # 300 frames -> 3 segments (ceil(300/120)), zero-padded
    styles = np.array([[1, 0], [0, 1], [1, 0]], dtype=np.float32)
    rng = np.random.default_rng(3)
    mean, std = SB.load_stats('BEAT')
    seeds = mean + std * np.cumsum(0.05 * rng.standard_normal((3, g.n_seed + 2, g.njoints // 3)), axis=1)
    args = SB.Config(dict(n_poses=g.n_poses, n_seed=g.n_seed, version="v0", name="DiffuseStyleGesture+"))
    full = SB.inference_batch_beat(args, ta, d.p_sample_loop, m, styles, seeds, seed=SEED, clip_ids=[5, 6, 7])
    assert full.shape == (3, 300, g.njoints // 3)
    for b in range(3):
        one = SB.inference_batch_beat(args, ta[b], d.p_sample_loop, m, styles[b:b + 1], seeds[b:b + 1], seed=SEED, clip_ids=[5 + b])
        assert np.abs(one[0] - full[b]).max() < 1e-5


# ------------------------------------------------------------------------------------------------ (e): sharded product path
_SHARD_WORKER = r'''
import os, sys, torch
sys.path.insert(0, %(root)r)
from diffusestylegesture_b200 import sample as S
from diffusestylegesture_b200.config import ZEGGS
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.synthetic import synthetic_state_dict
from diffusestylegesture_b200.wavlm import WavLM
from diffusestylegesture_b200.wavlm_config import WAVLM_LARGE, synthetic_wavlm_state_dict, synthetic_wav
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0")) if not os.environ.get("DSG_DIST_SAME_GPU") else 0
dev = "cuda:%%d" %% local
torch.cuda.set_device(local)
wm = WavLM(max_batch=4); wm.load_state_dict(synthetic_wavlm_state_dict(WAVLM_LARGE, seed=0)); wm.to(dev).eval()
model = MDM(njoints=ZEGGS.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=8, precision='bf16', max_batch=4)
load_model_wo_clip(model, synthetic_state_dict(ZEGGS, seed=0)); model.to(dev).eval()
d = create_gaussian_diffusion([12])
wavs = synthetic_wav(5, 170 * 800).numpy()
names = ["Happy", "Old", "Sad", "Angry", "Neutral"]
rows = [{'wav': '%%s_%%s_0.wav' %% (chr(97 + i), n), 'style': S.style2onehot[n], 'style_name': n, 'clip_id': 3 * i,
         'audio': wavs[i][:(100 if i == 1 else 170) * 800]} for i, n in enumerate(names)]
paths = S.main_batch(S.Config(n_poses=88, audio_feat="wavlm", gpu=str(local), max_batch=0), %(out)r, None, rows, wavlm_model=wm,
                     model=model, diffusion=d, seed=11)
if int(os.environ.get("RANK", "0")) == 0:
    assert len(paths) == 5
    print("PATHS " + " ".join(paths))
else:
    assert paths is None
'''


def _run_sharded(tmp_path, tag, nproc, env_extra):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / tag
    script = tmp_path / (tag + ".py")
    script.write_text(_SHARD_WORKER % {"root": root, "out": str(out)})
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, str(script)] if nproc == 1 else \
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
         "--master-port", "29633", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("PATHS ")][0]
    return line.split(" ")[1:]


def test_manifest_sharded_over_ranks_writes_identical_bvh(tmp_path):
    """`sample.main_batch` under torchrun: the manifest is sharded over the ranks, ONE gather brings the motions to rank 0,
    which writes every BVH — byte-identical to the single-process run (noise keyed by clip id).  With >= 2 GPUs the two
    ranks use NCCL on cuda:0 / cuda:1; on a one-GPU box both ranks share cuda:0 and the gather runs over gloo."""
    single = _run_sharded(tmp_path, "single", 1, {})
    if torch.cuda.device_count() >= 2:
        sharded = _run_sharded(tmp_path, "nccl2", 2, {})
    else:
        sharded = _run_sharded(tmp_path, "gloo2", 2, {"DSG_DIST_BACKEND": "gloo", "DSG_DIST_SAME_GPU": "1"})
    assert len(single) == len(sharded) == 5
    for a, b in zip(single, sharded):
        assert os.path.basename(a) == os.path.basename(b)
        assert open(a, "rb").read() == open(b, "rb").read(), os.path.basename(a)


# ------------------------------------------------------------------------------------------------ bf16 BVH parity, widened
@pytest.mark.slow
def test_six_clips_six_styles_1000_steps_bf16_bvh_vs_reference_golden(gold_dir):
    """BASELINE configs[1] over six more clips (ids 1..6, one per style) as ONE batch of the clip kernel: 320 frames = 4 segments
    x 1000 DDPM steps, bf16 — final BVH joint values against the reference's own `sample.inference` per clip.
    Same stated bound as the single-clip test of round 1: positions < 0.4 cm, Euler angles < 1.0 degree."""
    from diffusestylegesture_b200 import sample as S
    from diffusestylegesture_b200 import process_zeggs_bvh as PBZ
    gold = np.load(os.path.join(gold_dir, "r2_clips6.npz"))
    st = np.load(os.path.join(gold_dir, "zeggs_mean_std.npz"))
    ids = list(range(1, 7))
    m = _zeggs_model("bf16", max_batch=6)
    d = create_gaussian_diffusion()
    feats = [torch.cat([synthetic_conditioning(G, 1, segment=s, clip_offset=c)["audio"] for c in ids]) for s in range(4)]
    styles = torch.tensor(np.stack([gold[f"c{c}/style"] for c in ids]), dtype=torch.float32)
    assert sorted(int(i) for i in styles.argmax(1)) == list(range(6))
    seq = S.inference_batch(m, d, feats, styles, seed=SEED, clip_ids=ids)
    worst = (0.0, 0.0, 0.0)
    for k, c in enumerate(ids):
        err_n = np.abs(seq[k].numpy()[:, ::4] - gold[f"c{c}/norm_sub"].astype(np.float32)).max()
        poses = O.denormalise(seq[k].numpy(), st["mean"], st["std"])
        pos, eul = PBZ.pose2bvh_arrays(poses, 312, smoothing=True)
        d_pos = np.abs(pos[::3] - gold[f"c{c}/positions"]).max()
        d_eul = np.abs((eul[::3] - gold[f"c{c}/rotations"] + 180.0) % 360.0 - 180.0).max()
        print(f"clip {c} (style {int(styles[k].argmax())}): normalised max err {err_n:.3g}; BVH positions {d_pos:.3g} cm, Euler {d_eul:.3g} deg")
        worst = (max(worst[0], err_n), max(worst[1], d_pos), max(worst[2], d_eul))
    assert worst[1] < 0.4 and worst[2] < 1.0, worst


# ------------------------------------------------------------------------------------------------ large-activation regime
@pytest.mark.parametrize("gain", [2.0, 4.0])
def test_bf16_clip_kernel_large_activation_regime(gain):
    """The bf16 engine deviates from the reference arithmetic (tanh-form GELU in packed fp16, fp16 linear2 operands, bf16 residual
    stream).  All golden vectors use default-init-scale weights, where FFN pre-activations are O(1).  Here every
    Linear weight is scaled by `gain` (pre-activations up to ~9 at gain 4).  Eight post-norm layers with such weights amplify ANY
    rounding: merely rounding the weights to bf16 (everything else fp32, on the oracle) already moves a 2-step loop by 0.3 rms of
    a 6.8-wide output at gain 4.  So the bound is relative: the clip kernel may be at most 3x as far from the fp32 oracle as the
    oracle with bf16-rounded weights is."""
    g = G
    sd = synthetic_state_dict(g, seed=0, gain=gain)
    sdb = {k: (v.to(torch.bfloat16).float() if v.is_floating_point() and v.dim() == 2 else v) for k, v in sd.items()}
    y = synthetic_conditioning(g, 2, segment=0)
    d = create_gaussian_diffusion([2])
    with torch.no_grad():
        want, _ = O.p_sample_loop(sd, g, O.Schedule(1000, [2]), y, 2, seed=SEED, segment=0)
        wb, _ = O.p_sample_loop(sdb, g, O.Schedule(1000, [2]), y, 2, seed=SEED, segment=0)
    m = MDM(njoints=g.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=g.n_seed, precision="bf16", max_batch=2)
    load_model_wo_clip(m, sd)
    m.to('cuda:0').eval()
    # a 2-step loop runs the persistent clip kernel (the production path with the fp16 tanh GELU), not the multi-kernel denoiser
    got = d.p_sample_loop(m, (2, g.njoints, 1, g.n_poses), clip_denoised=False, model_kwargs={'y': dict(y, noise_seed=SEED, segment=0)}).cpu()
    assert bool(torch.isfinite(got).all())
    mx, rms = _err(got, want)
    mx_b, rms_b = _err(wb, want)
    print(f"gain {gain}: |out|max {float(want.abs().max()):.3g}; clip kernel vs fp32 oracle max {mx:.3g} rms {rms:.3g}; "
          f"bf16-weights-only oracle max {mx_b:.3g} rms {rms_b:.3g}")
    assert rms < 3.0 * rms_b + 1e-3 and mx < 3.0 * mx_b + 1e-2
