"""Round-2 golden vectors from the UNMODIFIED reference (/root/reference, read-only), and the pin of the oracle's new
restatements (const_noise, dump_steps, PLMS, attn5, the BEAT-TWH `inference` driver) to it.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_r2.py zeggs     # config 3 (B = 64, DDIM-100, six styles), options, PLMS
    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_r2.py beat      # "++" (attn5) denoiser, BEAT "+" inference (8 segments)

(two processes: both reference trees use the top-level module names `model` / `diffusion`).  Shims as in gen_golden.py;
for `inference` of BEAT-TWH-main/mydiffusion_beat_twh/sample.py the modules process_BEAT_bvh / process_TWH_bvh (pymo,
textgrid, h5py ... — the BVH tail, not the path under test) are replaced by stubs that capture `out_poses`, and the
dataset file the function reads its seed gesture from is a synthetic array in a temporary tree.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.dont_write_bytecode = True

from diffusestylegesture_b200.config import ZEGGS, BEAT_PLUS, BEAT_PLUSPLUS  # noqa: E402
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning  # noqa: E402
from oracle import dsg_oracle as O  # noqa: E402
from oracle.gen_golden import StreamNoise, import_reference_zeggs, make_ref_diffusion, maxdiff  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
SEED = 123456
CSUB = 8          # config 3 keeps every 8th joint channel (fp16) + full-tensor sums per clip


def ref_model_zeggs(ref_sample, g, sd):
    ref_sample.mydevice = torch.device('cpu')
    args = types.SimpleNamespace(audio_feat='wavlm')
    model, _ = ref_sample.create_model_and_diffusion(args)
    model.load_state_dict(sd, strict=False)
    return model.eval()


def zeggs():
    g = ZEGGS
    torch.set_num_threads(os.cpu_count())
    ref_sample, gd, SpacedDiffusion, space_timesteps = import_reference_zeggs()
    sd = synthetic_state_dict(g, seed=0)
    model = ref_model_zeggs(ref_sample, g, sd)
    report, out = [], {}
    shape1 = (g.njoints, 1, g.n_poses)

    # ---- config 3: batch 64, six styles (clip i -> style i mod 6), DDIM-100, eta 0
    B = 64
    y = synthetic_conditioning(g, B, segment=0)
    diff = make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, "ddim100")
    with StreamNoise() as sn, torch.no_grad():
        sn.reset(list(range(B)), 0)
        r = diff.ddim_sample_loop(model, (B,) + shape1, clip_denoised=False, model_kwargs={'y': y}, skip_timesteps=0,
                                  init_image=None, progress=False, dump_steps=None, noise=None, const_noise=False)
    o4, _ = O.p_sample_loop(sd, g, O.Schedule(1000, "ddim100"), {k: (v[:4] if k != "mask_local" else v) for k, v in y.items()}, 4,
                            seed=SEED, clip_ids=[0, 1, 2, 3], segment=0, sampler="ddim")
    d = maxdiff(r[:4], o4)
    assert d < 1e-4, d
    report.append(f"config 3 (B=64, ddim100, styles i mod 6): |ref-oracle|max on clips 0..3 = {d:.3g}, |out|max={float(r.abs().max()):.3g}")
    out["c3_sub"] = r[:, ::CSUB].numpy().astype(np.float16)
    out["c3_sum"] = r.double().sum(dim=(1, 2, 3)).numpy()
    out["c3_abssum"] = r.double().abs().sum(dim=(1, 2, 3)).numpy()

    # ---- const_noise + dump_steps through p_sample_loop (gaussian_diffusion.py:544-545, 647-669)
    B = 2
    y = synthetic_conditioning(g, B, segment=0)
    diff = make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, [50])
    dump_steps = [0, 10, 49]

    class ConstNoise(StreamNoise):              # randn_like(x) then noise[[0]].repeat: draw per clip, the reference keeps clip 0's
        pass
    with ConstNoise() as sn, torch.no_grad():
        sn.reset([0, 1], 0)
        dump = diff.p_sample_loop(model, (B,) + shape1, clip_denoised=False, model_kwargs={'y': y}, skip_timesteps=0,
                                  init_image=None, progress=False, dump_steps=dump_steps, noise=None, const_noise=True)
    assert isinstance(dump, list) and len(dump) == 3
    _, od = O.p_sample_loop(sd, g, O.Schedule(1000, [50]), y, B, seed=SEED, segment=0, const_noise=True, dump_steps=dump_steps)
    d = max(maxdiff(a, b) for a, b in zip(dump, od))
    assert d < 1e-4, d
    report.append(f"ddpm50 const_noise=True dump_steps={dump_steps} (B=2): |ref-oracle|max = {d:.3g}")
    out["opt_dump_steps"] = np.array(dump_steps)
    out["opt_dump"] = torch.stack(dump).numpy()[:, :, ::2]

    # ---- PLMS (gaussian_diffusion.py:1005-1200), orders 2 and 3
    for order in (2, 3):
        with StreamNoise() as sn, torch.no_grad():
            sn.reset([0, 1], 0)
            r = diff.plms_sample_loop(model, (B,) + shape1, clip_denoised=False, model_kwargs={'y': y}, skip_timesteps=0,
                                      init_image=None, progress=False, order=order)
        o, _ = O.p_sample_loop(sd, g, O.Schedule(1000, [50]), y, B, seed=SEED, segment=0, sampler="plms", order=order)
        d = maxdiff(r, o)
        assert d < 1e-3, d
        report.append(f"plms50 order {order} (B=2): |ref-oracle|max = {d:.3g}, |out|max={float(r.abs().max()):.3g}")
        out[f"plms{order}"] = r.numpy()[:, ::2]
    np.savez_compressed(os.path.join(GOLD, "r2_zeggs.npz"), **out)
    with open(os.path.join(GOLD, "GOLDEN_REPORT_R2_ZEGGS.txt"), "w") as fh:
        fh.write("Generated by oracle/gen_golden_r2.py zeggs against /root/reference/main\n" + "\n".join(report) + "\n")
    print("\n".join(report))


def clips6():
    """Six more 320-frame x 1000-step clips through the reference's own `sample.inference` (clip ids 1..6, style (id - 1) mod 6,
    features / noise keyed by the clip id): widens the bf16 BVH parity test beyond the one Neutral clip of round 1."""
    g = ZEGGS
    torch.set_num_threads(os.cpu_count())
    ref_sample, gd, SpacedDiffusion, space_timesteps = import_reference_zeggs()
    sd = synthetic_state_dict(g, seed=0)
    model = ref_model_zeggs(ref_sample, g, sd)
    diff = make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, None)
    st = np.load(os.path.join(GOLD, "zeggs_mean_std.npz"))
    A = types.SimpleNamespace(n_poses=88, audio_feat='wavlm')
    out, report = {}, []
    sn = StreamNoise()
    for cid in range(1, 7):
        style = [0] * 6
        style[(cid - 1) % 6] = 1
        feats = [synthetic_conditioning(g, 1, segment=sg, clip_offset=cid)["audio"] for sg in range(4)]
        calls = {"i": 0}

        def fake_wav2wavlm(model_, wav, device=None):
            f = feats[calls["i"]]
            sn.reset([cid], calls["i"])
            calls["i"] += 1
            return f
        ref_sample.wav2wavlm = fake_wav2wavlm
        ref_sample.mydevice = torch.device("cpu")
        ref_sample.batch_size = 1
        ref_sample.save_dir = tempfile.mkdtemp()
        cap = {}
        real = ref_sample.pose2bvh

        def spy(poses, path, length, smoothing=False):
            cap["poses"], cap["length"] = np.array(poses), length
        ref_sample.pose2bvh = spy
        import time
        t0 = time.time()
        with sn, torch.no_grad():
            ref_sample.inference(A, None, np.zeros(320 * 800, dtype=np.float32), diff.p_sample_loop, model, n_frames=320,
                                 smoothing=True, SG_filter=True, minibatch=True, skip_timesteps=0, style=style, seed=SEED)
        ref_sample.pose2bvh = real
        pos, eul = O.pose2bvh_values(cap["poses"], cap["length"], smoothing=True)       # tail pinned bit-exact in GOLDEN_REPORT.txt
        norm = (cap["poses"] - np.array(st["mean"]).squeeze()) / np.clip(np.array(st["std"]).squeeze(), 0.01, None)
        out[f"c{cid}/style"] = np.array(style)
        out[f"c{cid}/norm_sub"] = norm[:, ::4].astype(np.float16)
        out[f"c{cid}/positions"] = pos[::3].astype(np.float32)
        out[f"c{cid}/rotations"] = eul[::3].astype(np.float32)
        report.append(f"clip {cid} style {style}: reference inference {time.time() - t0:.1f} s, |norm|max {np.abs(norm).max():.3g}")
        print(report[-1], flush=True)
    np.savez_compressed(os.path.join(GOLD, "r2_clips6.npz"), **out)
    with open(os.path.join(GOLD, "GOLDEN_REPORT_R2_CLIPS6.txt"), "w") as fh:
        fh.write("Generated by oracle/gen_golden_r2.py clips6 against /root/reference/main (sample.inference, 4 segments x 1000 steps)\n"
                 + "\n".join(report) + "\n")


def beat():
    refb = os.path.join(REF, "BEAT-TWH-main")
    torch.set_num_threads(os.cpu_count())
    report, out = [], {}
    # ---- stubs for the BVH tail / feature extractors that sample.py imports at module level
    for name in ("librosa", "easydict", "process_BEAT_bvh", "process_TWH_bvh"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["easydict"].EasyDict = type("EasyDict", (dict,), {"__getattr__": dict.__getitem__, "__setattr__": dict.__setitem__})
    captured = {}

    def capture(save_dir, prefix, poses, pipeline=None):
        captured["poses"] = np.array(poses)
    pb = sys.modules["process_BEAT_bvh"]
    pb.wav2wavlm = pb.pose2bvh = None
    pb.pose2bvh_bugfix = capture
    pt = sys.modules["process_TWH_bvh"]
    pt.pose2bvh = pt.wavlm_init = pt.load_metadata = None

    # ---- a temporary tree with the relative paths `inference` reads (sample.py:76-82, 118-127): the real mean/std files of
    # the reference, a synthetic seed-gesture file
    tmp = tempfile.mkdtemp(prefix="dsg_beat_")
    cwd = os.path.join(tmp, "BEAT-TWH-main", "mydiffusion_beat_twh")
    os.makedirs(cwd)
    os.makedirs(os.path.join(tmp, "BEAT-TWH-main", "process"))
    os.makedirs(os.path.join(tmp, "BEAT_dataset", "processed", "gesture_BEAT"))
    g = BEAT_PLUS
    Jd = g.njoints // 3
    mean = np.load(os.path.join(refb, "process", "gesture_BEAT_mean_v0.npy"))
    std = np.load(os.path.join(refb, "process", "gesture_BEAT_std_v0.npy"))
    assert mean.shape[-1] == Jd and std.shape[-1] == Jd, (mean.shape, std.shape)
    np.save(os.path.join(tmp, "BEAT-TWH-main", "process", "gesture_BEAT_mean_v0.npy"), mean)
    np.save(os.path.join(tmp, "BEAT-TWH-main", "process", "gesture_BEAT_std_v0.npy"), std)
    rng = np.random.default_rng(7)
    walk = np.cumsum(0.05 * rng.standard_normal((g.n_seed + 2, Jd)), axis=0)
    seed_raw = (mean + std * walk).astype(np.float64)                  # a smooth synthetic "recorded" gesture
    np.save(os.path.join(tmp, "BEAT_dataset", "processed", "gesture_BEAT", "2_scott_0_1_1.npy"), seed_raw)
    os.chdir(cwd)
    for p in [os.path.join(refb, "mydiffusion_beat_twh"), refb, os.path.join(refb, "process"), os.path.join(refb, "model")]:
        sys.path.append(p)
    import sample as ref_sample                     # BEAT-TWH-main/mydiffusion_beat_twh/sample.py (module body is import-only)
    from model.mdm import MDM
    from diffusion import gaussian_diffusion as gd
    from diffusion.respace import SpacedDiffusion, space_timesteps

    def make(gg, cond_mode):
        m = MDM(modeltype='', njoints=gg.njoints, nfeats=1, cond_mode=cond_mode, audio_feat='wavlm', arch='trans_enc',
                latent_dim=gg.latent_dim, n_seed=gg.n_seed, cond_mask_prob=0.1, device='cpu', style_dim=gg.style_in,
                source_audio_dim=gg.audio_dim, audio_feat_dim_latent=gg.audio_latent)
        sd_ = synthetic_state_dict(gg, seed=0)
        ref_sd = m.state_dict()
        assert set(ref_sd) == set(sd_), set(ref_sd) ^ set(sd_)
        m.load_state_dict(sd_)
        return m.eval(), sd_

    # ---- "++" (cross_local_attention5): forward + 20-step loop, B = 2
    g5 = BEAT_PLUSPLUS
    model5, sd5 = make(g5, 'cross_local_attention5_style1_sample')
    B = 2
    y = synthetic_conditioning(g5, B, segment=0)
    y["seed"] = 0.5 * O.noise_tensor(SEED, [0, 1], 7, 99, (g5.njoints, 1, g5.n_seed))
    y["seed_last"] = 0.5 * O.noise_tensor(SEED, [0, 1], 7, 98, (g5.njoints, 1, g5.n_seed))
    x = O.noise_tensor(SEED, [0, 1], 0, 0, (g5.njoints, 1, g5.n_poses))
    t = torch.tensor([12, 850])
    with torch.no_grad():
        ref = model5(x, t, y=y)
        ora = O.mdm_forward(sd5, g5, x, t, y)
    d = maxdiff(ref, ora)
    assert d < 5e-5, d
    report.append(f"beat++ (attn5, D={g5.latent_dim}) forward B=2: |ref-oracle|max={d:.3g}, |out|max={float(ref.abs().max()):.3g}")
    out["pp/t"] = t.numpy()
    out["pp/seed_sub"] = y["seed"][:, ::4].numpy()
    out["pp/seed_last_sub"] = y["seed_last"][:, ::4].numpy()
    out["pp/out_sub"] = ref[:, ::4].numpy()
    diff20 = make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, [20])
    with StreamNoise() as sn, torch.no_grad():
        sn.reset([0, 1], 0)
        r = diff20.p_sample_loop(model5, (B, g5.njoints, 1, g5.n_poses), clip_denoised=False, model_kwargs={'y': y},
                                 skip_timesteps=0, init_image=None, progress=False, dump_steps=None, noise=None, const_noise=False)
    o, _ = O.p_sample_loop(sd5, g5, O.Schedule(1000, [20]), y, B, seed=SEED, segment=0)
    d = maxdiff(r, o)
    assert d < 5e-4, d
    report.append(f"beat++ ddpm20 loop B=2: |ref-oracle|max={d:.3g}")
    out["pp/loop20_sub"] = r[:, ::4].numpy()

    # ---- BEAT "+" `inference` (sample.py:44-201): 900 frames of features -> ceil(900 / 120) = 8 segments, 50-step DDPM
    model4, sd4 = make(g, 'cross_local_attention4_style1_sample')
    n_frames = 900
    gen = torch.Generator().manual_seed(2024)
    textaudio = torch.randn(n_frames, g.audio_dim, generator=gen)
    style = np.array([1.0, 0.0])                                    # speaker 2 (id_speaker_dict, sample.py:29-32)
    args = ref_sample.EasyDict(dict(n_poses=g.n_poses, n_seed=g.n_seed, audio_feature_dim=g.audio_dim, njoints=g.njoints,
                                    version="v0", name="DiffuseStyleGesture+"))
    ref_sample.mydevice = torch.device("cpu")
    ref_sample.batch_size = 1
    diff50 = make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, [50])

    class SegNoise(StreamNoise):                                    # one segment counter per p_sample_loop call
        pass
    sn = SegNoise()
    seg = {"i": -1}
    loop = diff50.p_sample_loop

    def sample_fn(*a, **kw):
        seg["i"] += 1
        sn.reset([0], seg["i"])
        return loop(*a, **kw)
    with sn, torch.no_grad():
        ref_sample.inference(args, tmp, "golden", textaudio, sample_fn, model4, n_frames=0, smoothing=True, skip_timesteps=0,
                             style=style, seed=SEED, dataset='BEAT')
    poses = captured["poses"]                                       # de-normalised [900, 684]
    assert poses.shape == (n_frames, Jd), poses.shape
    # oracle restatement of the same driver
    sg = (seed_raw - mean) / std
    vel = sg[1:] - sg[:-1]
    acc = vel[1:] - vel[:-1]
    seed_gesture = torch.from_numpy(np.concatenate((sg[2:], vel[1:], acc), axis=1)).float()          # [n_seed, J]
    seq = O.inference_clip_beat(sd4, g, O.Schedule(1000, [50]), textaudio, torch.tensor(style, dtype=torch.float32),
                                seed_gesture, seed=SEED, clip_id=0)
    ora = np.multiply(seq.numpy(), std) + mean
    d = float(np.abs(ora - poses).max())
    scale = float(np.abs(poses).max())
    assert d < 2e-3 * max(1.0, scale), (d, scale)
    report.append(f"beat+ inference (900 frames = 8 segments x ddpm50): |ref-oracle|max on de-normalised poses = {d:.3g} "
                  f"(|poses|max {scale:.3g})")
    out["inf/textaudio_seed"] = np.array([2024])
    out["inf/seed_raw"] = seed_raw
    out["inf/style"] = style
    out["inf/poses"] = poses.astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "r2_beat.npz"), **out)
    with open(os.path.join(GOLD, "GOLDEN_REPORT_R2_BEAT.txt"), "w") as fh:
        fh.write("Generated by oracle/gen_golden_r2.py beat against /root/reference/BEAT-TWH-main\n" + "\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    {"zeggs": zeggs, "beat": beat, "clips6": clips6}[sys.argv[1]]()
