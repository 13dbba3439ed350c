#!/bin/bash
# ncu launch list of one WavLM-Large forward (16 segments) — per-kernel durations, cold-cache/serialised (shares only).
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
torch.cuda.cudart().cudaProfilerStart()
m.wav2wavlm(w, 88); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
PY
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/wavlm_launches.csv python /tmp/wl_once.py
