"""Golden vectors for the numeric part of the BEAT / TWH BVH tails, from the UNMODIFIED reference functions
(BEAT-TWH-main/process/process_BEAT_bvh.py:108-131 `pose2bvh_bugfix`, process_TWH_bvh.py:201-226 `pose2bvh`).

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_bvh_tail.py

The two functions end in `pipeline.inverse_transform` + `BVHWriter.write` on pickled pymo objects, which cannot be loaded in this
container (pymo imports transforms3d, which is not installed).  Shims, none touching the tree: stub modules for the import-only
dependencies (transforms3d, textgrid, h5py, librosa, ...), `joblib.load` returns a recorder whose `inverse_transform` captures
the array the reference hands to the pipeline, `BVHWriter.write` is a no-op.  What is recorded is therefore computed by the
reference's own statements (Savitzky-Golay smoothing, rotation matrix -> Euler conversion)."""
import os
import sys
import tempfile
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/BEAT-TWH-main/process"
sys.dont_write_bytecode = True
GOLD = os.path.join(REPO, "tests", "golden")


class Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})


def main():
    for name in ("transforms3d", "textgrid", "h5py", "librosa", "librosa.display", "parselmouth", "pydub", "soundfile", "pyarrow",
                 "lmdb", "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors", "matplotlib.patheffects",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "IPython", "IPython.display", "seaborn", "essentia", "essentia.standard",
                 "python_speech_features"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = Anything(name)
    sys.path[:0] = [REF, os.path.join(REF, "..")]
    os.chdir(REF)
    captured = {}

    class Recorder:
        def inverse_transform(self, xs):
            captured["x"] = np.array(xs[0])
            return [None]
    import joblib
    joblib.load = lambda path: Recorder()
    rng = np.random.default_rng(11)
    from scipy.spatial.transform import Rotation as R
    out = {}
    n = 64
    # ---- BEAT: [n, 9 k] rotation matrices (k = 7 joints), slightly noisy so the smoothing matters
    import process_BEAT_bvh as PB
    PB.BVHWriter = lambda: types.SimpleNamespace(write=lambda *a, **k: None)
    k = 7
    eul = np.cumsum(rng.normal(0, 3.0, (n, k, 3)), axis=0)
    mats = R.from_euler('XYZ', eul.reshape(-1, 3), degrees=True).as_matrix().reshape(n, k * 9)
    poses = mats + rng.normal(0, 0.01, mats.shape)
    with tempfile.TemporaryDirectory() as td:
        PB.pose2bvh_bugfix(td, "g", poses, pipeline="unused.sav")
    out["beat_poses"] = poses
    out["beat_euler"] = captured["x"]
    # ---- TWH 'rotmat' mode: [n, 12 k] (position 3 | rotation matrix 9)
    import process_TWH_bvh as PT
    PT.BVHWriter = lambda: types.SimpleNamespace(write=lambda *a, **k: None)
    k = 5
    eul = np.cumsum(rng.normal(0, 3.0, (n, k, 3)), axis=0)
    mats = R.from_euler('ZXY', eul.reshape(-1, 3), degrees=True).as_matrix().reshape(n, k, 9)
    pos = np.cumsum(rng.normal(0, 0.5, (n, k, 3)), axis=0)
    g = np.concatenate((pos, mats), axis=2).reshape(n, k * 12) + rng.normal(0, 0.01, (n, k * 12))
    with tempfile.TemporaryDirectory() as td:
        PT.pose2bvh(g, td, "g", pipeline_path="pipeline_rotmat_62.sav")
    out["twh_gesture"] = g
    out["twh_pos_euler"] = captured["x"]
    # ---- load_tsv (process_TWH_bvh.py:168-200) on a small transcript
    with tempfile.TemporaryDirectory() as td:
        tsv = os.path.join(td, "t.tsv")
        open(tsv, "w").write("0.10\t0.50\thello\n0.50\t1.20\tbig world\n1.30\t1.60\t#laugh#\n2.00\t2.40\tunknownword,\n")
        w2v = {w: rng.normal(size=300) for w in ("hello", "big", "world", "laugh")}
        feats = PT.load_tsv(tsv, w2v, 90)
    out["tsv_words"] = np.array(sorted(w2v))
    out["tsv_vecs"] = np.stack([w2v[w] for w in sorted(w2v)])
    out["tsv_feats"] = feats
    np.savez_compressed(os.path.join(GOLD, "bvh_tail_beat_twh.npz"), **out)
    print({k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
