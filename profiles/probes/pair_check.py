"""CTA-pair mode of the clip kernel (DSG_CLIP_PAIR): results against the single-CTA mode and the oracle, reproducibility, step time.
    python profiles/probes/pair_check.py [--steps 200]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch

from diffusestylegesture_b200.config import ZEGGS as G
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--batches", default="1,8,64")
ap.add_argument("--no-parity", action="store_true")
ap.add_argument("--no-taps", action="store_true")
a = ap.parse_args()
sd = synthetic_state_dict(G, seed=0)


def model(B):
    m = MDM(njoints=G.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=G.n_seed, precision="bf16", max_batch=B)
    load_model_wo_clip(m, sd)
    return m.to('cuda:0').eval()


def run(m, d, B, y, seg=0):
    return d.p_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False,
                           model_kwargs={'y': dict(y, noise_seed=123456, segment=seg, clip_ids=list(range(B)))}).clone()


if not a.no_parity:
    from oracle import dsg_oracle as O
    for B, n in ((1, 2), (2, 6), (3, 40)):
        d = create_gaussian_diffusion([n])
        y = synthetic_conditioning(G, B, segment=0)
        m = model(B)
        os.environ["DSG_CLIP_PAIR"] = "0"
        single = run(m, d, B, y)
        os.environ["DSG_CLIP_PAIR"] = "1"
        pair = [run(m, d, B, y) for _ in range(3)]
        torch.cuda.synchronize()
        e = (pair[0] - single).abs()
        print(f"B={B} steps={n}: pair vs single max {float(e.max()):.3g} rms {float(e.pow(2).mean().sqrt()):.3g}; finite {bool(torch.isfinite(pair[0]).all())}; "
              f"reproducible {torch.equal(pair[0], pair[1]) and torch.equal(pair[0], pair[2])}", flush=True)
        if n <= 6:
            want, _ = O.p_sample_loop(sd, G, O.Schedule(1000, [n]), y, B, seed=123456, segment=0)
            for name, got in (("single", single), ("pair", pair[0])):
                e = (got.cpu() - want).abs()
                print(f"   {name} vs oracle: max {float(e.max()):.3g} rms {float(e.pow(2).mean().sqrt()):.3g}", flush=True)

# per-layer taps of ONE step from the same x_T (xs0 = after the local attention, xs<l+1> = after layer l): where the two modes part
B = 2
d = create_gaussian_diffusion([2])
y = synthetic_conditioning(G, B, segment=0)
taps = {}
for mode in (() if a.no_taps else ("0", "1", "1b")):
    os.environ["DSG_CLIP_PAIR"] = mode[0]
    m = model(B)
    eng = m.get_engine(B)
    eng.debug_enable()
    out = d.p_sample_loop(m, (B, G.njoints, 1, G.n_poses), clip_denoised=False, skip_timesteps=1,
                          model_kwargs={'y': dict(y, noise_seed=123456, segment=0, clip_ids=list(range(B)))}).clone()
    taps[mode] = {k: eng.debug_read(k, B).clone().reshape(B, 89, 256) for k in ["xs%d" % i for i in range(9)]}
    taps[mode]["out"] = out.cpu().reshape(B, G.njoints, G.n_poses).permute(0, 2, 1)
rms = lambda t: float(t.pow(2).mean().sqrt())
for k in ([] if a.no_taps else taps["0"]):
    for other in ("1",):
        e = (taps[other][k] - taps["0"][k])
        line = f"tap {k} pair{other[1:]} vs single: max {float(e.abs().max()):.3g} rms {rms(e):.3g}"
        if k != "out":
            line += " | clip0 rows 0-31 %.3g 32-63 %.3g 64-88 %.3g | cols %s" % (rms(e[0, :32]), rms(e[0, 32:64]), rms(e[0, 64:]),
                                                                                    " ".join("%.3g" % rms(e[0, :, c:c + 64]) for c in range(0, 256, 64)))
        else:
            line += " | clip0 channels <640 %.3g >=640 %.3g" % (rms(e[0, :, :640]), rms(e[0, :, 640:]))
        print(line, flush=True)
    e2 = (taps["1"][k] - taps["1b"][k]).abs()
    print(f"      pair vs pair: max {float(e2.max()):.3g}; per clip {[float(e2[b].max()) for b in range(B)]}", flush=True)

# in-kernel cycle counters (CTA 0) of both modes, one clip
os.environ["DSG_CLIP_PROF"] = "1"
d50 = create_gaussian_diffusion([50])
y = synthetic_conditioning(G, 1, segment=0)
prof = {}
for mode in ("0", "1"):
    os.environ["DSG_CLIP_PAIR"] = mode
    m = model(1)
    run(m, d50, 1, y)
    torch.cuda.synchronize()
    prof[mode] = m.get_engine(1).clip_profile()
os.environ.pop("DSG_CLIP_PROF")
print("phase (us per step at 1.9 GHz)      single    pair")
for k in prof["0"]:
    print(f"  {k:32s} {prof['0'][k] / 50 / 1900:8.1f} {prof['1'][k] / 50 / 1900:8.1f}", flush=True)

d = create_gaussian_diffusion([a.steps])
for B in [int(b) for b in a.batches.split(",")]:
    y = synthetic_conditioning(G, B, segment=0)
    m = model(B)
    for mode in ("0", "1"):
        os.environ["DSG_CLIP_PAIR"] = mode
        run(m, d, B, y)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(m, d, B, y)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"B={B} pair={mode}: {dt / a.steps * 1e6:.1f} us per step  ({B * 80 / (dt / a.steps * 1000):.0f} frames/s at 1000 steps)", flush=True)
