"""Recipe: stage the UNMODIFIED reference implementation of the hot path under oracle/_ref/ (git-ignored, travels to
the GPU box with the snapshot like a built .so) so that ``bench.py --impl reference`` and the CPU-baseline leg time the
reference ITSELF on the box's host cores (``cpu_baseline.kind == "reference"``), not the oracle port.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under diffusestylegesture_b200/ imports oracle/ (tests/test_host_logic.py
checks it).  No reference source is committed: this script copies the files where they lie under /root/reference into
oracle/_ref/ at build time (``__graft_entry__.build()`` calls it when /root/reference exists) and never edits them.

Staged (SURVEY.md section 8(a) rows a1-a14, a18 — the sampler and the denoiser, nothing else):
    main/model/mdm.py, main/model/local_attention/*.py, main/diffusion/{gaussian_diffusion,respace,nn,losses}.py,
    main/utils/model_util.py and the import-only chain of gaussian_diffusion.py:19
    (data_loaders/humanml/{scripts/motion_process,common/skeleton,common/quaternion,utils/paramUtil}.py);
    BEAT-TWH-main/{model/mdm.py, model/local_attention/*.py, diffusion/*.py} for the "+" variant.
Import shim (applied by ``load_reference`` below, outside the staged files): ``numpy.float = float`` (quaternion.py:13 uses
the alias numpy removed).
"""
import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")

ZEGGS_FILES = [
    "main/model/mdm.py",
    "main/model/local_attention/__init__.py", "main/model/local_attention/local_attention.py",
    "main/model/local_attention/rotary.py", "main/model/local_attention/transformer.py",
    "main/diffusion/gaussian_diffusion.py", "main/diffusion/respace.py", "main/diffusion/nn.py", "main/diffusion/losses.py",
    "main/utils/model_util.py",
    "main/data_loaders/humanml/scripts/motion_process.py", "main/data_loaders/humanml/common/skeleton.py",
    "main/data_loaders/humanml/common/quaternion.py", "main/data_loaders/humanml/utils/paramUtil.py",
]
BEAT_FILES = [
    "BEAT-TWH-main/model/mdm.py",
    "BEAT-TWH-main/model/local_attention/__init__.py", "BEAT-TWH-main/model/local_attention/local_attention.py",
    "BEAT-TWH-main/model/local_attention/rotary.py", "BEAT-TWH-main/model/local_attention/transformer.py",
    "BEAT-TWH-main/diffusion/gaussian_diffusion.py", "BEAT-TWH-main/diffusion/respace.py", "BEAT-TWH-main/diffusion/nn.py",
    "BEAT-TWH-main/diffusion/losses.py",
]


def build(verbose=True):
    """Copy the in-scope reference files into oracle/_ref/ (idempotent).  Returns True when the copy exists afterwards."""
    if not os.path.isdir(REF):
        return os.path.isdir(DST)
    n = 0
    for rel in ZEGGS_FILES + BEAT_FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified files of /root/reference staged by oracle/build_ref.py (git-ignored; measurement infrastructure).\n")
    if verbose:
        print(f"staged {n} reference files under {DST}")
    return True


def available():
    return os.path.exists(os.path.join(DST, "main", "model", "mdm.py"))


def load_reference(flavour="zeggs"):
    """Import the staged reference.  Returns (MDM class, gaussian_diffusion module, SpacedDiffusion, space_timesteps).
    flavour: "zeggs" (main/) or "beat" (BEAT-TWH-main/)."""
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float  # noqa: the alias the reference's quaternion.py:13 still uses
    root = os.path.join(DST, "main" if flavour == "zeggs" else "BEAT-TWH-main")
    if not os.path.isdir(root):
        raise FileNotFoundError(f"{root}: run `python oracle/build_ref.py` where /root/reference exists")
    sys.dont_write_bytecode = True
    # the two flavours use the same top-level module names: drop the other one's modules before importing
    for name in list(sys.modules):
        if name.split('.')[0] in ("model", "diffusion", "local_attention", "data_loaders", "utils") and \
                getattr(sys.modules[name], "__file__", None) and DST in (sys.modules[name].__file__ or ""):
            del sys.modules[name]
    paths = [root, os.path.join(root, "model")]
    sys.path[:] = [p for p in sys.path if not p.startswith(DST)]
    for p in reversed(paths):
        sys.path.insert(0, p)
    importlib.invalidate_caches()
    mdm = importlib.import_module("model.mdm")
    gd = importlib.import_module("diffusion.gaussian_diffusion")
    rs = importlib.import_module("diffusion.respace")
    return mdm.MDM, gd, rs.SpacedDiffusion, rs.space_timesteps


def make_reference_diffusion(gd, SpacedDiffusion, space_timesteps, respacing=None):
    """create_gaussian_diffusion (main/utils/model_util.py:59-100) with a respacing argument."""
    betas = gd.get_named_beta_schedule('cosine', 1000, 1.)
    return SpacedDiffusion(use_timesteps=space_timesteps(1000, respacing if respacing else [1000]), betas=betas,
                           model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                           loss_type=gd.LossType.MSE, rescale_timesteps=False)


def make_reference_model(MDM, g, state_dict):
    """The reference's own constructor call (main/mydiffusion_zeggs/sample.py:51-56; BEAT-TWH sample.py:35-41)."""
    if g.variant == 3:
        m = MDM(modeltype='', njoints=g.njoints, nfeats=1, translation=True, pose_rep='rot6d', glob=True, glob_rot=True,
                cond_mode='cross_local_attention3_style1', clip_version='ViT-B/32', action_emb='tensor', audio_feat='wavlm',
                arch='trans_enc', latent_dim=g.latent_dim, n_seed=g.n_seed)
    else:
        m = MDM(modeltype='', njoints=g.njoints, nfeats=1, cond_mode='cross_local_attention4_style1_sample', audio_feat='wavlm',
                arch='trans_enc', latent_dim=g.latent_dim, n_seed=g.n_seed, cond_mask_prob=0.1, device='cpu',
                style_dim=g.style_in, source_audio_dim=g.audio_dim, audio_feat_dim_latent=g.audio_latent)
    missing, unexpected = m.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    return m.eval()


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref", "ready" if ok else "unavailable (no /root/reference here and no staged copy)")
