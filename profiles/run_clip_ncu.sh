#!/bin/bash
# ncu --set full of the persistent clip kernel (one launch = B clips x 12 DDPM steps).  Run under gpurun.
#   default: B=148 (one CTA per clip) -> gpurun_out/clip_kernel_r02.ncu-rep
#   B=64 OUT=clip_kernel_pair_r02     -> the CTA-pair kernel (128 CTAs in clusters of two)
set -e
mkdir -p gpurun_out
cat > /tmp/clip_once.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from diffusestylegesture_b200.config import ZEGGS
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
B, steps = int(os.environ.get('B', '148')), 12
g = ZEGGS
m = MDM(njoints=g.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=g.n_seed, precision="bf16", max_batch=B)
load_model_wo_clip(m, synthetic_state_dict(g, seed=0)); m.to('cuda:0').eval()
d = create_gaussian_diffusion([steps])
y = synthetic_conditioning(g, B, segment=0); y.update(noise_seed=1, segment=0)
for _ in range(2):
    d.p_sample_loop(m, (B, g.njoints, 1, g.n_poses), clip_denoised=False, model_kwargs={'y': y})
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:clip_kernel -s 1 -c 1 -o gpurun_out/${OUT:-clip_kernel_r02} -f python /tmp/clip_once.py > gpurun_out/ncu_clip.log 2>&1
tail -3 gpurun_out/ncu_clip.log
ls -la gpurun_out/${OUT:-clip_kernel_r02}.ncu-rep
