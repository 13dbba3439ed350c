"""BVH / feature tails of the BEAT-TWH driver (reference BEAT-TWH-main/process/process_BEAT_bvh.py:108-131,
process_TWH_bvh.py:139-262) — host numpy, outside the GPU hot path.

The numeric part of ``pose2bvh_bugfix`` / ``pose2bvh`` (Savitzky-Golay smoothing per channel, rotation matrices -> Euler
angles) is restated here, vectorised.  The last step of both functions, ``pipeline.inverse_transform`` + ``BVHWriter``, runs
pickled *reference* objects (sklearn pipelines of ``pymo`` classes, ``resource/*.sav``): they are loaded with joblib when the
reference's ``pymo`` package (and its ``transforms3d`` dependency) is importable and are never re-implemented here;
otherwise ``PipelineUnavailable`` is raised after the numeric part (callers keep the pose arrays).
"""
import io
import os
import string

import numpy as np
from scipy.signal import savgol_filter
from scipy.spatial.transform import Rotation as R

BEAT_EULER_ORDER = 'XYZ'          # process_BEAT_bvh.py:48


class PipelineUnavailable(RuntimeError):
    pass


def smooth_channels(poses, window=15, order=2):
    """`for i in range(poses.shape[1]): savgol_filter(poses[:, i], 15, 2)` in one call (the filter is per channel)."""
    return savgol_filter(np.asarray(poses, dtype=np.float64), window, order, axis=0)


def beat_euler(poses, order=BEAT_EULER_ORDER):
    """process_BEAT_bvh.py:113-126: [n, 9 k] smoothed rotation matrices -> [n, 3 k] Euler angles (degrees)."""
    out = smooth_channels(poses)
    n = out.shape[0]
    mats = out.reshape(n, -1, 3, 3)
    eul = R.from_matrix(mats.reshape(-1, 3, 3)).as_euler(order, degrees=True)
    return eul.reshape(n, -1)


def twh_pos_euler(predicted_gesture):
    """process_TWH_bvh.py:205-220 ('rotmat' mode): [n, 12 k] (position 3 | rotation matrix 9) -> [n, 6 k] (position | ZXY Euler)."""
    out = smooth_channels(predicted_gesture)
    n = out.shape[0]
    blocks = out.reshape(n, -1, 12)
    eul = R.from_matrix(blocks[:, :, 3:].reshape(-1, 3, 3)).as_euler('ZXY', degrees=True).reshape(n, -1, 3)
    return np.concatenate((blocks[:, :, :3], eul), axis=2).reshape(n, -1)


def _load_pipeline(pipeline):
    if hasattr(pipeline, 'inverse_transform'):
        return pipeline
    try:
        import joblib as jl
        return jl.load(pipeline)
    except (ImportError, ModuleNotFoundError, FileNotFoundError, AttributeError) as ex:
        raise PipelineUnavailable(f"cannot load the pymo pipeline {pipeline!r} ({type(ex).__name__}: {ex}); put the reference's "
                                  "BEAT-TWH-main/process on PYTHONPATH (pymo needs transforms3d)") from ex


def _writer(writer):
    if writer is not None:
        return writer
    try:
        from pymo.writers import BVHWriter
        return BVHWriter()
    except (ImportError, ModuleNotFoundError) as ex:
        raise PipelineUnavailable(f"pymo.writers.BVHWriter is not importable ({ex})") from ex


def pose2bvh_bugfix(save_path, filename_prefix, poses, pipeline='./resource/data_pipe_30fps.sav', writer=None):
    """process_BEAT_bvh.py:108-131."""
    out_euler = beat_euler(poses)
    pipe = _load_pipeline(pipeline)
    bvh_data = pipe.inverse_transform([out_euler])
    out_bvh_path = os.path.join(save_path, filename_prefix + '_generated.bvh')
    w = _writer(writer)
    with open(out_bvh_path, 'w') as f:
        w.write(bvh_data[0], f)
    return out_bvh_path


def pose2bvh_twh(predicted_gesture, output_dir, name, pipeline_path="./pipeline_expmap_25.sav", writer=None):
    """process_TWH_bvh.py:201-226."""
    mode = os.path.basename(pipeline_path).split("_")[1] if isinstance(pipeline_path, str) else 'rotmat'
    data = twh_pos_euler(predicted_gesture) if mode == 'rotmat' else np.asarray(predicted_gesture)
    pipe = _load_pipeline(pipeline_path)
    bvh_data = pipe.inverse_transform([data])[0]
    path = os.path.join(output_dir, f"{name}.bvh")
    w = _writer(writer)
    with open(path, 'w') as f:
        w.write(bvh_data, f, framerate=30)
    return path


def load_metadata(metadata, participant):
    """process_TWH_bvh.py:229-262 -> (num_speakers, by file name, by index)."""
    assert participant in ("main-agent", "interloctr"), "`participant` must be either 'main-agent' or 'interloctr'"
    by_name, by_index, speaker_ids = {}, {}, []
    with open(metadata, "r") as f:
        for i, line in enumerate(f.readlines()[1:]):          # the first line is the csv header
            fname, main_id, main_finger, iloc_id, iloc_finger = line.strip().split(",")
            if participant == "main-agent":
                has_finger, speaker_id = main_finger == "finger_incl", int(main_id) - 1
            else:
                has_finger, speaker_id = iloc_finger == "finger_incl", int(iloc_id) - 1
            speaker_ids.append(speaker_id)
            by_index[i] = has_finger, speaker_id
            by_name[fname + f"_{participant}"] = has_finger, speaker_id
    num_speakers = int(np.unique(np.array(speaker_ids)).shape[0]) if speaker_ids else 0
    return num_speakers, by_name, by_index


def load_wordvectors(fname):
    """process_TWH_bvh.py:157-165 (fastText .vec text format)."""
    data = {}
    with io.open(fname, 'r', encoding='utf-8', newline='\n', errors='ignore') as fin:
        fin.readline()
        for line in fin:
            tokens = line.rstrip().split(' ')
            data[tokens[0]] = np.array([float(v) for v in tokens[1:]])
    return data


def load_tsv(tsvpath, word2vector, clip_len):
    """process_TWH_bvh.py:139-200: word vectors aligned to 30 fps frames + [has_laughter, is_silence] columns."""
    sentence = []
    with open(tsvpath, "r") as f:
        for line in f.readlines():
            parts = line.strip().split("\t")
            if len(parts) == 3:
                sentence.append([float(parts[0]), float(parts[1]), parts[2]])
    feats = np.zeros([clip_len, 300 + 2])
    feats[:, -1] = 1
    for start, end, raw_word in sentence:
        has_laughter = "#" in raw_word
        s, e = int(start * 30), int(end * 30)
        feats[s:e, -1] = 0
        word = raw_word.translate(str.maketrans('', '', string.punctuation)).strip().replace("  ", " ")
        if len(word) > 0 and word[0] == " ":
            word = word[1:]
        if " " in word:
            ww = word.split(" ")
            dur = (e - s) / len(ww)
            for j, w in enumerate(ww):
                vec = word2vector.get(w)
                if vec is not None:
                    feats[s + int(dur * j):s + int(dur * (j + 1)), :300] = vec
        else:
            vec = word2vector.get(word)
            if vec is not None:
                feats[s:e, :300] = vec
        feats[s:e, -2] = has_laughter
    return feats
