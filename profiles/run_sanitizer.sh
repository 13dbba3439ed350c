#!/bin/bash
# compute-sanitizer racecheck + synccheck + memcheck of the persistent clip kernel (3 DDPM steps, 2 clips, bf16).  Run under gpurun.
#   PAIR=0 (default) : the one-CTA-per-clip kernel (DSG_CLIP_PAIR=0)          -> gpurun_out/r02_sanitizer.log
#   PAIR=1 TOOLS="synccheck memcheck" : the CTA-pair kernel (cluster of two)  -> gpurun_out/r02_sanitizer_pair.log
# The kernel synchronises 16 warps through 45 mbarriers, two named barriers and async-proxy fences; racecheck tracks
# shared-memory hazards between generic-proxy accesses (st.shared / ld.shared of the epilogues).
set -e
mkdir -p gpurun_out
cat > /tmp/clip_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from diffusestylegesture_b200.config import ZEGGS
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning
B, steps = 2, 3
g = ZEGGS
m = MDM(njoints=g.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=g.n_seed, precision="bf16", max_batch=B)
load_model_wo_clip(m, synthetic_state_dict(g, seed=0)); m.to('cuda:0').eval()
d = create_gaussian_diffusion([steps])
y = synthetic_conditioning(g, B, segment=0); y.update(noise_seed=1, segment=0)
out = d.p_sample_loop(m, (B, g.njoints, 1, g.n_poses), clip_denoised=False, model_kwargs={'y': y})
torch.cuda.synchronize()
print("finite", bool(torch.isfinite(out).all()), "absmax", float(out.abs().max()))
PY
PAIR=${PAIR:-0}
TOOLS=${TOOLS:-"synccheck racecheck memcheck"}
LOG=gpurun_out/r02_sanitizer$([ "$PAIR" = 1 ] && echo _pair).log
export DSG_CLIP_PAIR=$PAIR
: > $LOG
for tool in $TOOLS; do
  echo "=== compute-sanitizer --tool $tool (DSG_CLIP_PAIR=$PAIR) ===" >> $LOG
  timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=clip_kernel --print-limit 20 python /tmp/clip_small.py >> $LOG 2>&1 || echo "(exit $?)" >> $LOG
done
grep -E "===|ERROR SUMMARY|RACECHECK SUMMARY|finite|exit" $LOG
