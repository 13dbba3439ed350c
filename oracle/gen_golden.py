"""Generate the committed golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, read-only) on CPU in the authoring container, and pin oracle/dsg_oracle.py to it.

Run:  PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [--fast]

The reference cannot travel to the GPU box, so what this script writes (small .npz files) is what
the tests use there.  Shims (none of them touch the reference tree):
  * ``numpy.float = float`` (removed alias used by an import chain of gaussian_diffusion.py:19);
  * stub modules for librosa / easydict / omegaconf (import-only dependencies of sample.py);
  * ``torch.randn`` / ``torch.randn_like`` read the shared Philox stream while the reference samples.
Weights are the deterministic synthetic ``state_dict`` (no checkpoint ships with the reference).
"""
import argparse
import os
import sys
import tempfile
import time
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.dont_write_bytecode = True

from diffusestylegesture_b200.config import ZEGGS, BEAT_PLUS, state_dict_spec  # noqa: E402
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning  # noqa: E402
from oracle import dsg_oracle as O  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
SEED = 123456


class StreamNoise:
    """Replaces torch.randn / randn_like with the shared counter-based stream."""

    def __init__(self):
        self.clip_ids, self.segment, self.draw = [0], 0, 0
        self._randn, self._randn_like = torch.randn, torch.randn_like

    def reset(self, clip_ids, segment):
        self.clip_ids, self.segment, self.draw = list(clip_ids), segment, 0

    def _next(self, shape):
        assert shape[0] == len(self.clip_ids), (shape, self.clip_ids)
        t = O.noise_tensor(SEED, self.clip_ids, self.segment, self.draw, tuple(shape[1:]))
        self.draw += 1
        return t

    def __enter__(self):
        torch.randn = lambda *shape, **kw: self._next(shape[0] if isinstance(shape[0], (tuple, list)) else shape)
        torch.randn_like = lambda x, **kw: self._next(tuple(x.shape)).to(x.dtype)
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def import_reference_zeggs():
    np.float = float  # noqa
    for name in ("librosa", "easydict", "omegaconf"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["easydict"].EasyDict = type("EasyDict", (dict,), {"__getattr__": dict.__getitem__,
                                                                  "__setattr__": dict.__setitem__})
    sys.modules["omegaconf"].DictConfig = dict
    os.chdir(os.path.join(REF, "main", "mydiffusion_zeggs"))
    for p in ['.', '..', '../process', '../model', '../../ubisoft-laforge-ZeroEGGS-main',
              '../../ubisoft-laforge-ZeroEGGS-main/ZEGGS']:
        sys.path.append(os.path.abspath(p))
    import sample as ref_sample
    from diffusion import gaussian_diffusion as gd
    from diffusion.respace import SpacedDiffusion, space_timesteps
    return ref_sample, gd, SpacedDiffusion, space_timesteps


def make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, respacing):
    """create_gaussian_diffusion (main/utils/model_util.py:59-100) with a respacing argument."""
    betas = gd.get_named_beta_schedule('cosine', 1000, 1.)
    return SpacedDiffusion(use_timesteps=space_timesteps(1000, respacing if respacing else [1000]), betas=betas,
                           model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                           loss_type=gd.LossType.MSE, rescale_timesteps=False)


def maxdiff(a, b):
    return float((torch.as_tensor(a, dtype=torch.float64) - torch.as_tensor(b, dtype=torch.float64)).abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fast", action="store_true", help="skip the 1000-step 4-segment clip")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    report = []
    g = ZEGGS
    ref_sample, gd, SpacedDiffusion, space_timesteps = import_reference_zeggs()

    # ---- 0. normalisation fixtures (data, not source): mean/std of the ZEGGS feature vector
    m = np.load(os.path.join(REF, "ubisoft-laforge-ZeroEGGS-main/data/processed_v1/processed/mean.npz"))['mean'].squeeze()
    s = np.load(os.path.join(REF, "ubisoft-laforge-ZeroEGGS-main/data/processed_v1/processed/std.npz"))['std'].squeeze()
    np.savez_compressed(os.path.join(GOLD, "zeggs_mean_std.npz"), mean=m, std=s)

    # ---- 1. schedule tables (known answers from the reference objects)
    sched_gold = {}
    for tag, resp in (("ddpm1000", None), ("ddpm50", [50]), ("ddim100", "ddim100")):
        d = make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, resp)
        o = O.Schedule(1000, resp)
        assert d.timestep_map == o.timestep_map
        for k in ("betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                  "posterior_log_variance_clipped", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
            assert np.array_equal(getattr(d, k), getattr(o, k)), (tag, k)
            sched_gold[f"{tag}/{k}"] = getattr(d, k)
        sched_gold[f"{tag}/timestep_map"] = np.array(d.timestep_map)
    np.savez_compressed(os.path.join(GOLD, "schedule.npz"), **sched_gold)
    report.append("schedule: oracle tables bit-equal to reference (ddpm1000, ddpm50, ddim100)")

    # ---- 2. reference model with synthetic weights
    class A:  # args for create_model_and_diffusion (sample.py:51-56 reads only audio_feat)
        audio_feat = "wavlm"
        n_poses = 88
    model, diffusion = ref_sample.create_model_and_diffusion(A)
    sd = synthetic_state_dict(g, seed=0)
    ref_sd = model.state_dict()
    spec = dict(state_dict_spec(g))
    assert set(ref_sd.keys()) == set(sd.keys()), set(ref_sd.keys()) ^ set(sd.keys())
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
        if k.endswith(".pe") or k == "rel_pos.inv_freq":
            assert torch.equal(v, sd[k]), k
    assert sum(int(np.prod(sh)) for sh in spec.values()) == 9001781
    from utils.model_util import load_model_wo_clip
    load_model_wo_clip(model, sd)
    model.eval()
    report.append("state_dict: %d keys, shapes equal to reference MDM, 9,001,781 parameters" % len(sd))

    # ---- 3. one denoiser call, B=2, per-op taps from the oracle + reference output
    B = 2
    y = synthetic_conditioning(g, B, segment=0)
    y["seed"] = 0.5 * O.noise_tensor(SEED, [0, 1], 7, 99, (g.njoints, 1, g.n_seed))   # non-zero seed pose
    x = O.noise_tensor(SEED, [0, 1], 0, 0, (g.njoints, 1, g.n_poses))
    t = torch.tensor([12, 850])
    with torch.no_grad():
        ref_out = model(x, t, y=y)
        taps = {}
        ora_out = O.mdm_forward(sd, g, x, t, y, taps=taps)
        ora64 = O.mdm_forward(sd, g, x.double(), t, {k: (v.double() if v.is_floating_point() else v) for k, v in y.items()})
    d = maxdiff(ref_out, ora_out)
    report.append("mdm forward B=2: |ref-oracle32|max=%.3g  |ref-oracle64|max=%.3g  |out|max=%.3g"
                  % (d, maxdiff(ref_out, ora64), float(ref_out.abs().max())))
    assert d < 2e-5, d
    np.savez_compressed(os.path.join(GOLD, "mdm_forward_zeggs.npz"), t=t.numpy(), seed_pose=y["seed"].numpy(),
                        out=ref_out.numpy(), **{"tap_" + k: v.numpy() for k, v in taps.items()
                                                if k in ("tok", "h_in", "h_local", "xs0", "xs1", "xs8")})

    # ---- 4. sampling loops through the reference sampler (stream noise)
    sn = StreamNoise()

    def ref_loop(diff, fn_name, B, y, seg, skip=0):
        sn.reset(list(range(B)), seg)
        with sn, torch.no_grad():
            fn = getattr(diff, fn_name)
            kw = dict(clip_denoised=False, model_kwargs={'y': y}, skip_timesteps=skip, init_image=None,
                      progress=False, dump_steps=None, noise=None)
            if fn_name == "p_sample_loop":
                kw["const_noise"] = False
            return fn(model, (B, g.njoints, 1, g.n_poses), **kw)

    traj = {}
    y2 = synthetic_conditioning(g, 2, segment=0)
    for tag, resp, fn, sampler, skip in (("ddpm50", [50], "p_sample_loop", "ddpm", 0),
                                         ("ddim100", "ddim100", "ddim_sample_loop", "ddim", 0),
                                         ("ddpm1000_skip950", None, "p_sample_loop", "ddpm", 950)):
        dref = make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, resp)
        t0 = time.time()
        r = ref_loop(dref, fn, 2, y2, 0, skip)
        tr = time.time() - t0
        o, _ = O.p_sample_loop(sd, g, O.Schedule(1000, resp), y2, 2, seed=SEED, segment=0, sampler=sampler,
                               skip_timesteps=skip)
        dd = maxdiff(r, o)
        report.append("%s B=2: |ref-oracle|max=%.3g (|x|max=%.3g, ref %.1fs)" % (tag, dd, float(r.abs().max()), tr))
        assert dd < 5e-4, (tag, dd)
        traj[tag] = r.numpy()
    np.savez_compressed(os.path.join(GOLD, "loops_zeggs.npz"), **traj)

    # ---- 5. full sample.inference() through the reference driver: 2 segments x 50 steps, then BVH
    def run_inference(diff, n_frames, tag):
        nseg = n_frames // 80
        feats = [synthetic_conditioning(g, 1, segment=sg)["audio"] for sg in range(nseg)]
        calls = {"i": 0}

        def fake_wav2wavlm(model_, wav, device=None):
            f = feats[calls["i"]]
            sn.reset([0], calls["i"])      # noise stream: new segment
            calls["i"] += 1
            return f
        ref_sample.wav2wavlm = fake_wav2wavlm
        ref_sample.mydevice = torch.device("cpu")
        ref_sample.batch_size = 1
        tmp = tempfile.mkdtemp()
        ref_sample.save_dir = tmp
        captured = {}
        real_pose2bvh = ref_sample.pose2bvh

        def spy_pose2bvh(poses, path, length, smoothing=False):
            captured["poses"] = np.array(poses)
            captured["length"] = length
            import utils_zeggs
            real_save = utils_zeggs.bvh.save

            def spy_save(fname, data, **kw):
                captured["positions"] = np.array(data["positions"])
                captured["rotations"] = np.array(data["rotations"])
                return real_save(fname, data, **kw)
            import process_zeggs_bvh as pz
            pz.write_bvh.__globals__["bvh"].save = spy_save
            try:
                return real_pose2bvh(poses, path, length=length, smoothing=smoothing)
            finally:
                pz.write_bvh.__globals__["bvh"].save = real_save
        ref_sample.pose2bvh = spy_pose2bvh
        audio = np.zeros(n_frames * 800, dtype=np.float32)
        style = [0, 0, 1, 0, 0, 0]
        with sn, torch.no_grad():
            ref_sample.inference(A, None, audio, diff.p_sample_loop, model, n_frames=n_frames, smoothing=True,
                                 SG_filter=True, minibatch=True, skip_timesteps=0, style=style, seed=SEED)
        ref_sample.pose2bvh = real_pose2bvh
        bvh_files = [f for f in os.listdir(tmp) if f.endswith(".bvh")]
        assert len(bvh_files) == 1
        with open(os.path.join(tmp, bvh_files[0])) as fh:
            txt = fh.read()
        captured["bvh_text_head"] = txt[:txt.index("MOTION")]
        captured["feats"] = feats
        captured["style"] = style
        return captured

    def check_inference(cap, resp, tag, steps_desc):
        sched = O.Schedule(1000, resp)
        seq = O.inference_clip(sd, g, sched, [f[0] for f in cap["feats"]], torch.tensor(cap["style"], dtype=torch.float32),
                               seed=SEED, clip_id=0)
        poses = O.denormalise(seq.numpy(), m, s)
        dpos = float(np.abs(poses - cap["poses"]).max())
        pos, eul = O.pose2bvh_values(poses, cap["length"], smoothing=True)
        dp, de = float(np.abs(pos - cap["positions"]).max()), float(np.abs(eul - cap["rotations"]).max())
        # also the pure tail: oracle pose2bvh on the REFERENCE's poses
        pos2, eul2 = O.pose2bvh_values(cap["poses"], cap["length"], smoothing=True)
        dp2, de2 = float(np.abs(pos2 - cap["positions"]).max()), float(np.abs(eul2 - cap["rotations"]).max())
        report.append("%s (%s): |poses ref-oracle|max=%.3g; BVH end-to-end pos %.3g cm, euler %.3g deg; "
                      "BVH tail on ref poses pos %.3g, euler %.3g" % (tag, steps_desc, dpos, dp, de, dp2, de2))
        assert dp2 < 1e-9 and de2 < 1e-4, (dp2, de2)
        return seq

    cap = run_inference(make_ref_diffusion(gd, SpacedDiffusion, space_timesteps, [50]), 160, "inference50")
    check_inference(cap, [50], "inference 160 frames", "2 segments x 50 steps")
    np.savez_compressed(os.path.join(GOLD, "inference_zeggs_50.npz"), poses=cap["poses"].astype(np.float32),
                        positions=cap["positions"].astype(np.float32), rotations=cap["rotations"].astype(np.float32),
                        length=cap["length"], style=np.array(cap["style"]))
    with open(os.path.join(GOLD, "bvh_header_zeggs.txt"), "w") as fh:
        fh.write(cap["bvh_text_head"])

    if not args.fast:
        t0 = time.time()
        cap = run_inference(diffusion, 320, "inference1000")
        tr = time.time() - t0
        report.append("reference sample.inference 320 frames x 1000 steps on %d cores: %.1f s -> %.2f motion frames/s"
                      % (os.cpu_count(), tr, 320 / tr))
        check_inference(cap, None, "inference 320 frames", "4 segments x 1000 steps")
        np.savez_compressed(os.path.join(GOLD, "inference_zeggs_1000.npz"), poses=cap["poses"].astype(np.float32),
                            positions=cap["positions"].astype(np.float32),
                            rotations=cap["rotations"].astype(np.float32), length=cap["length"],
                            style=np.array(cap["style"]))

    with open(os.path.join(GOLD, "GOLDEN_REPORT.txt"), "w") as fh:
        fh.write("Generated by oracle/gen_golden.py against /root/reference (torch %s, %d cores)\n"
                 % (torch.__version__, os.cpu_count()))
        fh.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
