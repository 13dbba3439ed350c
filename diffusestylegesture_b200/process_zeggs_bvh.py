"""ZEGGS pose vector -> BVH (host tail of the path; numpy/scipy on the CPU, as in the reference).

Host mirror of ``pose2bvh`` (reference main/process/process_zeggs_bvh.py:219-275), ``write_bvh``
(ubisoft-laforge-ZeroEGGS-main/ZEGGS/utils_zeggs.py:47-87), the quaternion helpers it needs
(ZEGGS/anim/quat.py:17-38, 111-118, 166-206; anim/txform.py:23-34) and the BVH text writer
(ZEGGS/anim/bvh.py:137-234).  SURVEY.md section 7.11: this is ~0.0x % of the time and stays on the host.
"""
import numpy as np
from scipy.signal import savgol_filter

NJOINTS = 75
ORDER = 'zyx'

bone_names = [
    "Hips", "Spine", "Spine1", "Spine2", "Spine3", "Neck", "Neck1", "Head", "HeadEnd", "RightShoulder", "RightArm",
    "RightForeArm", "RightHand", "RightHandThumb1", "RightHandThumb2", "RightHandThumb3", "RightHandThumb4",
    "RightHandIndex1", "RightHandIndex2", "RightHandIndex3", "RightHandIndex4", "RightHandMiddle1",
    "RightHandMiddle2", "RightHandMiddle3", "RightHandMiddle4", "RightHandRing1", "RightHandRing2", "RightHandRing3",
    "RightHandRing4", "RightHandPinky1", "RightHandPinky2", "RightHandPinky3", "RightHandPinky4", "RightForeArmEnd",
    "RightArmEnd", "LeftShoulder", "LeftArm", "LeftForeArm", "LeftHand", "LeftHandThumb1", "LeftHandThumb2",
    "LeftHandThumb3", "LeftHandThumb4", "LeftHandIndex1", "LeftHandIndex2", "LeftHandIndex3", "LeftHandIndex4",
    "LeftHandMiddle1", "LeftHandMiddle2", "LeftHandMiddle3", "LeftHandMiddle4", "LeftHandRing1", "LeftHandRing2",
    "LeftHandRing3", "LeftHandRing4", "LeftHandPinky1", "LeftHandPinky2", "LeftHandPinky3", "LeftHandPinky4",
    "LeftForeArmEnd", "LeftArmEnd", "RightUpLeg", "RightLeg", "RightFoot", "RightToeBase", "RightToeBaseEnd",
    "RightLegEnd", "RightUpLegEnd", "LeftUpLeg", "LeftLeg", "LeftFoot", "LeftToeBase", "LeftToeBaseEnd", "LeftLegEnd",
    "LeftUpLegEnd"]

parents = np.array([-1, 0, 1, 2, 3, 4, 5, 6, 7, 4, 9, 10, 11, 12, 13, 14, 15, 12, 17, 18, 19, 12, 21, 22, 23, 12, 25,
                    26, 27, 12, 29, 30, 31, 12, 11, 4, 35, 36, 37, 38, 39, 40, 41, 38, 43, 44, 45, 38, 47, 48, 49, 38,
                    51, 52, 53, 38, 55, 56, 57, 38, 37, 0, 61, 62, 63, 64, 63, 62, 0, 68, 69, 70, 71, 70, 69],
                   dtype=np.int32)


# ---- quaternion helpers (w, x, y, z) -----------------------------------------------------------------
def quat_mul(a, b):
    aw, ax, ay, az = (a[..., i:i + 1] for i in range(4))
    bw, bx, by, bz = (b[..., i:i + 1] for i in range(4))
    return np.concatenate([bw * aw - bx * ax - by * ay - bz * az,
                           bw * ax + bx * aw - by * az + bz * ay,
                           bw * ay + bx * az + by * aw - bz * ax,
                           bw * az - bx * ay + by * ax + bz * aw], axis=-1)


def _cross(a, b):
    out = np.empty(np.broadcast(a, b).shape)
    out[..., 0] = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    out[..., 1] = a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2]
    out[..., 2] = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    return out


def quat_mul_vec(q, v):
    t = 2.0 * _cross(q[..., 1:], v)
    return v + q[..., 0][..., np.newaxis] * t + _cross(q[..., 1:], t)


def quat_from_xform(m, eps=1e-10):
    """Rotation matrices [..., 3, 3] -> quaternions, branch on the largest diagonal term."""
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    tr = m00 + m11 + m22
    q = np.zeros(m.shape[:-2] + (4,), dtype=m.dtype)

    def put(mask, w, x, y, z):
        nonlocal q
        q = np.where(mask[..., None], np.stack([w, x, y, z], axis=-1), q)

    s = 0.5 / np.sqrt(np.maximum(tr + 1, eps))
    put(tr > 0, 0.25 / s, s * (m[..., 2, 1] - m[..., 1, 2]), s * (m[..., 0, 2] - m[..., 2, 0]), s * (m[..., 1, 0] - m[..., 0, 1]))
    neg = tr <= 0
    x_big = (m00 > m11) & (m00 > m22)
    y_big = (~x_big) & (m11 > m22)
    z_big = (~x_big) & (~y_big)
    sx = 2.0 * np.sqrt(np.maximum(1.0 + m00 - m11 - m22, eps))
    put(neg & x_big, (m[..., 2, 1] - m[..., 1, 2]) / sx, sx * 0.25, (m[..., 0, 1] + m[..., 1, 0]) / sx, (m[..., 0, 2] + m[..., 2, 0]) / sx)
    sy = 2.0 * np.sqrt(np.maximum(1.0 + m11 - m00 - m22, eps))
    put(neg & y_big, (m[..., 0, 2] - m[..., 2, 0]) / sy, (m[..., 0, 1] + m[..., 1, 0]) / sy, sy * 0.25, (m[..., 1, 2] + m[..., 2, 1]) / sy)
    sz = 2.0 * np.sqrt(np.maximum(1.0 + m22 - m00 - m11, eps))
    put(neg & z_big, (m[..., 1, 0] - m[..., 0, 1]) / sz, (m[..., 0, 2] + m[..., 2, 0]) / sz, (m[..., 1, 2] + m[..., 2, 1]) / sz, sz * 0.25)
    return q


def quat_to_euler(q, order='zyx'):
    if order != 'zyx':
        raise NotImplementedError('Cannot convert to ordering %s' % order)
    w, x, y, z = (q[..., i:i + 1] for i in range(4))
    return np.concatenate([np.arctan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z)),
                           np.arcsin(np.clip(2.0 * (w * y - z * x), -1.0, 1.0)),
                           np.arctan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))], axis=-1)


def xform_orthogonalize_from_xy(xy, eps=1e-10):
    """Two-axis (x, y) encoding [..., 2, 3] (float32) -> orthonormal rotation matrices [..., 3, 3]."""
    xy = np.asarray(xy, dtype=np.float32)
    xa = xy[..., 0, :]
    za = np.cross(xa, xy[..., 1, :]).astype(np.float32)
    ya = np.cross(za, xa).astype(np.float32)

    def unit(v):
        return v / (np.sqrt(np.sum(v * v, axis=-1, dtype=np.float32))[..., None] + np.float32(eps))
    rows = np.stack([unit(xa), unit(ya), unit(za)], axis=-2)
    return np.swapaxes(rows, -1, -2)


def pose2bvh_arrays(poses, length, smoothing=False):
    """Numeric part of pose2bvh + write_bvh: returns (positions [3*length, 75, 3], euler degrees [3*length, 75, 3])."""
    poses = np.asarray(poses)
    if smoothing:
        # the reference filters column by column (process_zeggs_bvh.py:224-227); one call along axis 0 is the same filter
        # (differences ~3e-15, far below the %f of the writer) and 5x faster.  NOTE(reference): smoothing rotation matrices is not optimal
        poses = savgol_filter(np.asarray(poses, dtype=np.float64), 15, 2, axis=0)
    nj = NJOINTS
    root_pos, root_rot = poses[:, 0:3], poses[:, 3:7]
    lpos = poses[:, 13: 13 + nj * 3].reshape([length, nj, 3])
    ltxy = poses[:, 13 + nj * 3: 13 + nj * 9].reshape([length, nj, 2, 3])
    lrot = quat_from_xform(xform_orthogonalize_from_xy(ltxy))
    # 20 fps -> 60 fps by frame repetition
    root_pos, root_rot = root_pos.repeat(3, axis=0), root_rot.repeat(3, axis=0)
    lpos, lrot = lpos.repeat(3, axis=0).copy(), lrot.repeat(3, axis=0).copy()
    lpos[:, 0] = quat_mul_vec(root_rot, lpos[:, 0]) + root_pos
    lrot[:, 0] = quat_mul(root_rot, lrot[:, 0])
    return lpos, np.degrees(quat_to_euler(lrot, ORDER))


_CH = {'x': 'Xrotation', 'y': 'Yrotation', 'z': 'Zrotation'}


def save_bvh(filename, positions, rotations, offsets, frametime=1.0 / 60.0, names=bone_names, order=ORDER,
             parent_idx=parents):
    """BVH text writer: root has 6 channels, every other joint 3; leaf joints get a zero End Site."""
    children = {i: [j for j in range(len(parent_idx)) if parent_idx[j] == i] for i in range(len(parent_idx))}
    seq, lines = [0], []
    rot_names = "%s %s %s" % (_CH[order[0]], _CH[order[1]], _CH[order[2]])

    def joint(i, tabs):
        seq.append(i)
        lines.append("%sJOINT %s\n%s{\n" % (tabs, names[i], tabs))
        inner = tabs + '\t'
        lines.append("%sOFFSET %f %f %f\n" % ((inner,) + tuple(offsets[i])))
        lines.append("%sCHANNELS 3 %s\n" % (inner, rot_names))
        for c in children[i]:
            joint(c, inner)
        if not children[i]:
            lines.append("%sEnd Site\n%s{\n%s\tOFFSET %f %f %f\n%s}\n" % (inner, inner, inner, 0.0, 0.0, 0.0, inner))
        lines.append("%s}\n" % tabs)

    lines.append("HIERARCHY\nROOT %s\n{\n" % names[0])
    lines.append("\tOFFSET %f %f %f\n" % tuple(offsets[0]))
    lines.append("\tCHANNELS 6 Xposition Yposition Zposition %s \n" % rot_names)
    for c in children[0]:
        joint(c, '\t')
    lines.append("}\nMOTION\nFrames: %i\nFrame Time: %f\n" % (len(rotations), frametime))
    # motion block: "%f %f %f " per (root position, joint rotations in hierarchy order), one C-level format call per file
    rot_seq = rotations[:, seq, :]
    table = np.concatenate([positions[:, 0, :], rot_seq.reshape(rot_seq.shape[0], -1)], axis=1)
    row_fmt = "%f " * table.shape[1] + "\n"
    lines.append((row_fmt * table.shape[0]) % tuple(table.ravel().tolist()))
    with open(filename, 'w') as f:
        f.write("".join(lines))


def pose2bvh(poses, outpath, length, smoothing=False, smooth_foot=False):
    """Reference signature (process_zeggs_bvh.py:219)."""
    if smooth_foot:
        raise NotImplementedError("smooth_foot drops into pdb in the reference (process_zeggs_bvh.py:253-254)")
    positions, eulers = pose2bvh_arrays(poses, length, smoothing=smoothing)
    save_bvh(outpath, positions, eulers, offsets=positions[0], frametime=1 / 60)
