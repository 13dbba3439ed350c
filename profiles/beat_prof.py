"""Per-kernel-class CUDA-event timing of the multi-kernel tcgen05 path on the BEAT "+" geometry (dsg_profile).
Usage on the GPU box:  python profiles/beat_prof.py [batch] [steps] [beat+|twh+]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from diffusestylegesture_b200.config import BEAT_PLUS, TWH_PLUS  # noqa: E402
from diffusestylegesture_b200.mdm import MDM  # noqa: E402
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip  # noqa: E402
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = TWH_PLUS if (len(sys.argv) > 3 and sys.argv[3] == "twh+") else BEAT_PLUS
m = MDM(njoints=g.njoints, cond_mode='cross_local_attention4_style1_sample', audio_feat='wavlm', n_seed=g.n_seed, latent_dim=g.latent_dim,
        style_dim=g.style_in, source_audio_dim=g.audio_dim, audio_feat_dim_latent=g.audio_latent, precision="bf16", max_batch=B)
load_model_wo_clip(m, synthetic_state_dict(g, seed=0))
m.to('cuda:0').eval()
d = create_gaussian_diffusion([steps])
y = synthetic_conditioning(g, B, segment=0)
y.update(noise_seed=1, segment=0)
shape = (B, g.njoints, 1, g.n_poses)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    d.p_sample_loop(m, shape, clip_denoised=False, model_kwargs={'y': y})
    b.record()
    torch.cuda.synchronize()
print(f"B={B} steps={steps} D={g.latent_dim}: graph replay {a.elapsed_time(b) * 1e3 / steps:.1f} us per DDPM step")
eng = m.get_engine(B)
eng.profile(True)
d.p_sample_loop(m, shape, clip_denoised=False, model_kwargs={'y': y})
prof = eng.profile_read()
eng.profile(False)
tot = sum(ms for _, ms in prof.values())
for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:22s} {n:6d} launches  {ms / n * 1e3:8.1f} us each  {ms / steps * 1e3:8.1f} us per step  {100 * ms / tot:5.1f} %")
print(f"  sum of kernel times: {tot / steps * 1e3:.1f} us per step")
