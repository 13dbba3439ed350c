import os, sys
sys.path.insert(0, os.getcwd())
import torch
from diffusestylegesture_b200.config import ZEGGS as G
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
sd = synthetic_state_dict(G, seed=0)
def model(B):
    m = MDM(njoints=G.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=G.n_seed, precision='bf16', max_batch=B)
    load_model_wo_clip(m, sd); m.to('cuda:0').eval(); return m
for nsteps in (2, 6):
    d = create_gaussian_diffusion([nsteps]); B = 3
    y = synthetic_conditioning(G, B, segment=0); shp = (B, G.njoints, 1, G.n_poses)
    res = {}
    for mode in ("kernels", "clip"):
        os.environ["DSG_TC_MODE"] = mode
        m = model(B); eng = m.get_engine(B); eng.debug_enable()
        out = d.p_sample_loop(m, shp, clip_denoised=False, model_kwargs={'y': dict(y, noise_seed=5, segment=0)})
        res[mode] = (out.cpu(), {f"xs{k}": eng.debug_read(f"xs{k}", B) for k in range(9)})
    for k in range(9):
        a, b = res["clip"][1][f"xs{k}"], res["kernels"][1][f"xs{k}"]
        dd = (a - b).abs()
        print(nsteps, f"xs{k}: max {float(dd.max()):.4g} rms {float(dd.pow(2).mean().sqrt()):.4g}  worst row {int(dd.amax(dim=(0,2)).argmax())} col {int(dd.amax(dim=(0,1)).argmax())}")
    dd = (res["clip"][0] - res["kernels"][0]).abs()
    print(nsteps, "final: max %.4g rms %.4g" % (float(dd.max()), float(dd.pow(2).mean().sqrt())), "worst joint", int(dd.amax(dim=(0,2,3)).argmax()), "frame", int(dd.amax(dim=(0,1,2)).argmax()))
