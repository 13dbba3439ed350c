"""CPU: host-side mirror of the reference interface, the C ABI surface, and the N>1 sharding logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from diffusestylegesture_b200 import engine as E
from diffusestylegesture_b200.config import ZEGGS, BEAT_PLUS, TWH_PLUS, state_dict_spec
from diffusestylegesture_b200.distributed import shard_bounds
from diffusestylegesture_b200.gaussian_diffusion import ModelVarType
from diffusestylegesture_b200.mdm import MDM
from diffusestylegesture_b200.model_util import create_gaussian_diffusion, load_model_wo_clip
from diffusestylegesture_b200.respace import space_timesteps
from diffusestylegesture_b200.synthetic import synthetic_state_dict
from diffusestylegesture_b200 import process_zeggs_bvh as PB
from diffusestylegesture_b200 import sample as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schedule_tables_bit_equal_to_reference(gold_dir):
    gold = np.load(os.path.join(gold_dir, "schedule.npz"))
    for tag, resp in (("ddpm1000", ''), ("ddpm50", [50]), ("ddim100", "ddim100")):
        d = create_gaussian_diffusion(resp)
        for k in ("betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                  "posterior_log_variance_clipped", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
            assert np.array_equal(getattr(d, k), gold[f"{tag}/{k}"]), (tag, k)
        assert list(gold[f"{tag}/timestep_map"]) == d.timestep_map
    d = create_gaussian_diffusion()
    assert d.num_timesteps == 1000 and d.model_var_type == ModelVarType.FIXED_SMALL
    coef, qs, tmap = d.engine_tables("ddpm")
    assert coef.shape == (1000, 4) and coef.dtype == np.float32 and coef[0, 0] == 1.0 and coef[0, 1] == 0.0
    np.testing.assert_allclose(coef[999, 2], 0.999498662, rtol=1e-6)
    coef, _, tmap = create_gaussian_diffusion("ddim100").engine_tables("ddim")
    assert coef.shape == (100, 4) and list(tmap[:3]) == [0, 10, 20]
    assert coef[0, 2] == 1.0 and coef[0, 3] == 0.0            # abar_prev[0] = 1: the last DDIM step returns x0


def test_space_timesteps_errors_and_sections():
    assert space_timesteps(300, [10, 15, 20]) == space_timesteps(300, "10,15,20")
    assert len(space_timesteps(300, [10, 15, 20])) == 45
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")
    with pytest.raises(ValueError):
        space_timesteps(10, [20])


def test_mdm_state_dict_surface():
    m = MDM(njoints=1141, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=8)
    sd = m.state_dict()
    assert len(sd) == 115 and sum(p.numel() for p in m.parameters()) == 9001781
    assert sd["seqTransEncoder.layers.7.self_attn.in_proj_weight"].shape == (768, 256)
    assert sd["embed_text.weight"].shape == (192, 9128)
    load_model_wo_clip(m, synthetic_state_dict(ZEGGS, seed=3))
    assert torch.equal(m.state_dict()["embed_style.weight"], synthetic_state_dict(ZEGGS, seed=3)["embed_style.weight"])
    bad = synthetic_state_dict(ZEGGS, seed=3)
    bad["not_a_key"] = torch.zeros(1)
    with pytest.raises(AssertionError):
        load_model_wo_clip(m, bad)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.get_engine(1)                                   # model on the CPU: must fail loudly
    for g in (BEAT_PLUS, TWH_PLUS):
        m2 = MDM(njoints=g.njoints, cond_mode='cross_local_attention4_style1_sample', audio_feat='wavlm',
                 n_seed=g.n_seed, latent_dim=g.latent_dim, style_dim=g.style_in, source_audio_dim=g.audio_dim,
                 audio_feat_dim_latent=g.audio_latent)
        assert set(m2.state_dict()) >= {n for n, _ in state_dict_spec(g)}
    from diffusestylegesture_b200.config import BEAT_PLUSPLUS
    g = BEAT_PLUSPLUS                                     # "++": two more tensors (embed_text_last), audio over T - 2 n_seed frames
    m5 = MDM(njoints=g.njoints, cond_mode='cross_local_attention5_style1_sample', audio_feat='wavlm', n_seed=g.n_seed,
             latent_dim=g.latent_dim, style_dim=g.style_in, source_audio_dim=g.audio_dim, audio_feat_dim_latent=g.audio_latent)
    assert m5.state_dict()["embed_text_last.weight"].shape == (g.audio_latent, g.njoints)
    assert [n for n, _ in state_dict_spec(g)][-2:] == ["embed_text_last.weight", "embed_text_last.bias"]
    assert g.audio_frames == g.n_poses - 2 * g.n_seed and BEAT_PLUS.audio_frames == 120
    with pytest.raises(NotImplementedError):
        MDM(njoints=1141, cond_mode='cross_local_attention2_style1', audio_feat='wavlm', n_seed=8)
    with pytest.raises(NotImplementedError):
        MDM(njoints=1141, cond_mode='cross_local_attention3_style1', audio_feat='mfcc', n_seed=8)


def test_unsupported_sampler_options_raise():
    d = create_gaussian_diffusion()
    m = MDM(njoints=1141, cond_mode='cross_local_attention3_style1', audio_feat='wavlm', n_seed=8)
    for kw in ({"clip_denoised": True}, {"clip_denoised": False, "randomize_class": True},
               {"clip_denoised": False, "denoised_fn": lambda x: x}, {"clip_denoised": False, "cond_fn": lambda *a: 0}):
        with pytest.raises(NotImplementedError):
            d.p_sample_loop(m, (1, 1141, 1, 88), model_kwargs={'y': {}}, **kw)
    with pytest.raises(NotImplementedError):
        d.ddim_sample_loop(m, (1, 1141, 1, 88), clip_denoised=False, eta=0.5, model_kwargs={'y': {}})
    # the reference's ddim_sample_loop raises NotImplementedError for these two (gaussian_diffusion.py:913-916); so do we
    for kw in ({"dump_steps": [1]}, {"const_noise": True}):
        with pytest.raises(NotImplementedError):
            d.ddim_sample_loop(m, (1, 1141, 1, 88), clip_denoised=False, model_kwargs={'y': {}}, **kw)
    with pytest.raises(ValueError):                       # plms_sample: 'order is invalid' (:1023-1024)
        d.plms_sample_loop(m, (1, 1141, 1, 88), clip_denoised=False, model_kwargs={'y': {}}, order=7)
    with pytest.raises(RuntimeError, match="no CPU path"):  # const_noise / dump_steps are engine options now: no CPU path
        d.p_sample_loop(m, (1, 1141, 1, 88), clip_denoised=False, const_noise=True, dump_steps=[0],
                        model_kwargs={'y': {'style': 0, 'seed': 0, 'audio': 0}})


def test_bvh_tail_matches_reference_values(gold_dir, tmp_path):
    gold = np.load(os.path.join(gold_dir, "inference_zeggs_50.npz"))
    poses = gold["poses"].astype(np.float64)
    pos, eul = PB.pose2bvh_arrays(poses, int(gold["length"]), smoothing=True)
    assert np.abs(pos - gold["positions"]).max() < 1e-4
    assert np.abs(eul - gold["rotations"]).max() < 2e-2
    out = tmp_path / "clip.bvh"
    PB.pose2bvh(poses, str(out), length=int(gold["length"]), smoothing=True)
    txt = out.read_text()
    head = txt[:txt.index("MOTION")]
    want = open(os.path.join(gold_dir, "bvh_header_zeggs.txt")).read()
    strip = lambda s: re.sub(r"OFFSET [-0-9. e]+", "OFFSET", s)      # offsets come from fp32-rounded golden poses
    assert strip(head) == strip(want)
    off = lambda s: np.array([[float(v) for v in m.split()] for m in re.findall(r"OFFSET ([-0-9. e]+)\n", s)])
    assert np.abs(off(head) - off(want)).max() < 1e-3
    lines = txt[txt.index("MOTION"):].splitlines()
    assert lines[1] == "Frames: %d" % (3 * int(gold["length"])) and lines[2] == "Frame Time: 0.016667"
    assert len(lines[3].split()) == 3 + 75 * 3


def test_segment_plan_and_style_table():
    assert S.segment_plan(320, 88, 8) == (4, 320)
    assert S.segment_plan(333, 88, 8) == (4, 320)
    assert S.segment_plan(50, 88, 8) == (1, 50)
    assert S.style2onehot['Neutral'] == [0, 0, 1, 0, 0, 0] and len(S.style2onehot) == 6
    cfg = S.parse_cli(["--gpu", "1", "--max_len", "320"])
    assert cfg.n_poses == 88 and cfg.audio_feat == "wavlm" and cfg.gpu == "1" and cfg.max_len == 320


def test_c_abi_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "dsg.h")).read()
    declared = set(re.findall(r"\b(dsg_[a-z_]+)\s*\(", header))
    assert declared == set(E.EXPORTS), declared ^ set(E.EXPORTS)
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, nm), name
    lib2 = E.load_library()
    assert b"sm_100a" in lib2.dsg_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_engine_fails_loudly_without_gpu(built_lib):
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        E.Engine(ZEGGS, synthetic_state_dict(ZEGGS, seed=0), device=0, max_batch=1, precision="fp32")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diffusestylegesture_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                for line in src.splitlines():           # the reference tree is cited in comments only, never opened
                    code = line.split("#")[0].split("//")[0]
                    assert not re.search(r"(open|path\.(insert|append)|chdir)\(.*/root/reference", code), (f, line)


def test_shard_bounds_cover_everything():
    for total in (1, 7, 8, 64, 513):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, torch
sys.path.insert(0, %r)
import torch.distributed as dist
from diffusestylegesture_b200.distributed import init_from_env, shard_bounds, gather_motions, barrier_max_ms
rank, world, _ = init_from_env("gloo")
total = 5
lo, hi = shard_bounds(total, rank, world)
local = torch.stack([torch.full((3, 4), float(c)) for c in range(lo, hi)]) if hi > lo else torch.zeros(0, 3, 4)
out = gather_motions(local, total)
ms = barrier_max_ms(10.0 * (rank + 1))
assert ms == 10.0 * world, ms
if rank == 0:
    assert out.shape == (5, 3, 4) and [float(out[c, 0, 0]) for c in range(5)] == [0., 1., 2., 3., 4.]
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
'''


def test_world_size_2_gather_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK" in r.stdout


def test_manifest_windows_and_batch_plan(tmp_path):
    """Batch front-end: manifest parsing, per-clip waveform windows (sample.py:224-251), batch plan."""
    from diffusestylegesture_b200 import sample as S
    man = tmp_path / "m.csv"
    man.write_text("wav,style,clip_id\n# comment\n/a/015_Happy_4_x_1_0.wav\n/a/b.wav,Old\n/a/c.wav,4,17\n")
    rows = S.read_manifest(str(man))
    assert [r["style"] for r in rows] == [[1, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0]]
    assert [r["clip_id"] for r in rows] == [0, 1, 17]
    man.write_text("/a/x_Bogus_1.wav\n")
    with pytest.raises(ValueError):
        S.read_manifest(str(man))
    audio = np.arange(170 * 800, dtype=np.float32)
    w, n_frames = S.segment_windows(audio, 170, 88, 8)
    assert w.shape == (2, 70400) and n_frames == 160
    assert float(w[0, :6400].abs().max()) == 0.0 and float(w[0, 6400]) == 0.0 and float(w[0, -1]) == 80 * 800 - 1
    assert float(w[1, 0]) == 80 * 800 - 6400 and float(w[1, -1]) == 160 * 800 - 1       # 8 seed frames of the previous stride
    assert S.plan_batches([4, 2, 4, 1, 2, 4], 2) == [[0, 2], [5], [1, 4], [3]]
    assert S.segment_plan(50, 88, 8) == (1, 50)


def test_bvh_writer_is_vectorised_and_stable(tmp_path):
    """One C-level format call per file and one savgol call per clip: same text as the per-row / per-column reference loops."""
    from scipy.signal import savgol_filter
    from diffusestylegesture_b200 import process_zeggs_bvh as PB
    from diffusestylegesture_b200.sample import denormalise
    rng = np.random.default_rng(3)
    poses = denormalise(rng.standard_normal((40, 1141)).astype(np.float32) * 0.1)
    pos, eul = PB.pose2bvh_arrays(poses, 40, smoothing=True)
    loop = np.stack([savgol_filter(poses[:, c], 15, 2) for c in range(poses.shape[1])], 1)
    pos2, eul2 = PB.pose2bvh_arrays(loop, 40, smoothing=False)
    assert np.abs(pos - pos2).max() < 1e-9 and np.abs(eul - eul2).max() < 1e-7
    path = str(tmp_path / "a.bvh")
    PB.pose2bvh(poses, path, 40, smoothing=True)
    lines = open(path).read().splitlines()
    k = lines.index("MOTION")
    assert lines[k + 1] == "Frames: 120" and len(lines) == k + 3 + 120
    row = lines[k + 3].split(" ")
    assert len(row) == 3 + 75 * 3 + 1 and row[-1] == ""                           # "%f " per value, trailing space
    assert abs(float(row[0]) - pos[0, 0, 0]) < 1e-6


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line on stdout with the contract keys,
    the same metric / unit / direction as our arm, a cpu_baseline describing the run, zero copy bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-sample-steps", "2", "--ref-batch", "2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("motion frames/sec") and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    from oracle import build_ref                      # kind follows what is staged: the reference itself, else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if build_ref.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
