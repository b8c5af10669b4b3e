"""TEST INFRASTRUCTURE — numpy restatement of the reference's skeleton feature pipeline steps (the step before the hot path):
JointToBone (pyskl/datasets/pipelines/pose_related.py:340-372), ToMotion (:375-397), GenSkeFeat / MergeSkeFeat (:400-442),
FormatGCNInput (:468-514).  Pinned against the unmodified reference classes by tests/test_pipeline.py (live, where /root/reference
exists).  The product package never imports this file."""
import numpy as np

PAIRS = {
    "nturgb+d": [(0, 1), (1, 20), (2, 20), (3, 2), (4, 20), (5, 4), (6, 5), (7, 6), (8, 20), (9, 8), (10, 9), (11, 10), (12, 0), (13, 12),
                 (14, 13), (15, 14), (16, 0), (17, 16), (18, 17), (19, 18), (21, 22), (20, 20), (22, 7), (23, 24), (24, 11)],
    "openpose": [(0, 0), (1, 0), (2, 1), (3, 2), (4, 3), (5, 1), (6, 5), (7, 6), (8, 2), (9, 8), (10, 9), (11, 5), (12, 11), (13, 12),
                 (14, 0), (15, 0), (16, 14), (17, 15)],
    "coco": [(0, 0), (1, 0), (2, 0), (3, 1), (4, 2), (5, 0), (6, 0), (7, 5), (8, 6), (9, 7), (10, 8), (11, 0), (12, 0), (13, 11), (14, 12),
             (15, 13), (16, 14)],
}


def joint_to_bone(kp, dataset):
    bone = np.zeros(kp.shape, dtype=np.float32)
    two_d = kp.shape[-1] == 3 and dataset in ("openpose", "coco")
    for a, b in PAIRS[dataset]:
        bone[..., a, :] = kp[..., a, :] - kp[..., b, :]
        if two_d:
            bone[..., a, 2] = (kp[..., a, 2] + kp[..., b, 2]) / 2
    return bone


def to_motion(x, dataset):
    mo = np.zeros_like(x)
    T = x.shape[1]
    mo[:, :T - 1] = x[:, 1:] - x[:, :-1]
    if x.shape[-1] == 3 and dataset in ("openpose", "coco"):
        mo[:, :T - 1, :, 2] = (x[:, :T - 1, :, 2] + x[:, 1:, :, 2]) / 2
    return mo


def gen_ske_feat(kp, score, dataset, feats):
    if score is not None:
        kp = np.concatenate([kp, score[..., None]], -1)
    d = {"j": kp}
    if "b" in feats or "bm" in feats:
        d["b"] = joint_to_bone(kp, dataset)
    if "jm" in feats:
        d["jm"] = to_motion(d["j"], dataset)
    if "bm" in feats:
        d["bm"] = to_motion(d["b"], dataset)
    return np.concatenate([d[f] for f in feats], -1)


def format_gcn_input(kp, num_person, mode, num_clips):
    M = kp.shape[0]
    if M < num_person:
        kp = np.concatenate([kp, np.zeros((num_person - M,) + kp.shape[1:], dtype=kp.dtype)], 0)
        if mode == "loop":
            for i in range(1, num_person):
                kp[i] = kp[0]
    elif M > num_person:
        kp = kp[:num_person]
    M, T, V, C = kp.shape
    return np.ascontiguousarray(kp.reshape(M, num_clips, T // num_clips, V, C).transpose(1, 0, 2, 3, 4))
