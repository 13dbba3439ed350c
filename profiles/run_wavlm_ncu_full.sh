#!/bin/bash
# ncu --set full of the WavLM layer kernels (persistent GEMM with the GELU and the residual epilogue, gated-bias attention).
set -e
mkdir -p gpurun_out
cat > /tmp/wl_once.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from diffusestylegesture_b200.wavlm import WavLM
from diffusestylegesture_b200.wavlm_config import WAVLM_LARGE, synthetic_wavlm_state_dict, synthetic_wav
m = WavLM(max_batch=32); m.load_state_dict(synthetic_wavlm_state_dict(WAVLM_LARGE, 0)); m.to('cuda:0')
w = synthetic_wav(32, 70400).cuda()
m.wav2wavlm(w, 88); torch.cuda.synchronize()
m.wav2wavlm(w, 88); torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_persistent|flash_attn' -s 96 -c 5 -o gpurun_out/wavlm_r02 -f python /tmp/wl_once.py > gpurun_out/ncu_wavlm.log 2>&1
tail -2 gpurun_out/ncu_wavlm.log
