"""Where does the persistent clip kernel spend its cycles?  (DSG_CLIP_PROF=1: CTA 0 accumulates clock64() deltas per role.)
Usage on the GPU box:  python profiles/clip_prof.py [batch] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DSG_CLIP_PROF"] = "1"
import torch  # noqa: E402
from diffusestylegesture_b200.config import ZEGGS  # noqa: E402
from diffusestylegesture_b200.mdm import MDM  # noqa: E402
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip  # noqa: E402
from diffusestylegesture_b200.synthetic import synthetic_state_dict, synthetic_conditioning  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
g = ZEGGS
m = MDM(njoints=g.njoints, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=g.n_seed, precision="bf16", max_batch=B)
load_model_wo_clip(m, synthetic_state_dict(g, seed=0))
m.to('cuda:0').eval()
d = create_gaussian_diffusion([steps])
y = synthetic_conditioning(g, B, segment=0)
y.update(noise_seed=1, segment=0)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    d.p_sample_loop(m, (B, g.njoints, 1, g.n_poses), clip_denoised=False, model_kwargs={'y': y})
    b.record()
    torch.cuda.synchronize()
prof = m.get_engine(B).clip_profile()
mhz = 1965.0
print(f"B={B} steps={steps}: segment call {a.elapsed_time(b):.2f} ms -> {a.elapsed_time(b) * 1e3 / steps:.1f} us/step (event-timed)")
tot = prof["total"]
for k, v in prof.items():
    print(f"  {k:24s} {v / steps / mhz:9.1f} us/step  {100 * v / tot:5.1f} % of the MMA warp's wall time")
