"""Input stage of the GCN configs on the device: the skeleton feature generators the reference runs in
numpy inside the data-loader workers, as batched torch ops on whatever device the keypoints live on, so the four streams of a
config-3 ensemble (joint, bone, joint-motion, bone-motion) share one skeleton upload.

  JointToBone   pyskl/datasets/pipelines/pose_related.py:340-372
  ToMotion      pyskl/datasets/pipelines/pose_related.py:375-397
  GenSkeFeat    pyskl/datasets/pipelines/pose_related.py:417-442  (MergeSkeFeat :400-414)
  FormatGCNInput pyskl/datasets/pipelines/pose_related.py:468-514

All functions take [..., M, T, V, C] (any leading batch dimensions) and return new tensors; C in {2, 3}.  For the 2-D layouts
('openpose', 'coco') the third channel is a keypoint score and is averaged instead of differenced, as in the reference.
"""
import torch

BONE_PAIRS = {
    "nturgb+d": ((0, 1), (1, 20), (2, 20), (3, 2), (4, 20), (5, 4), (6, 5), (7, 6), (8, 20), (9, 8), (10, 9), (11, 10), (12, 0), (13, 12),
                 (14, 13), (15, 14), (16, 0), (17, 16), (18, 17), (19, 18), (21, 22), (20, 20), (22, 7), (23, 24), (24, 11)),
    "openpose": ((0, 0), (1, 0), (2, 1), (3, 2), (4, 3), (5, 1), (6, 5), (7, 6), (8, 2), (9, 8), (10, 9), (11, 5), (12, 11), (13, 12),
                 (14, 0), (15, 0), (16, 14), (17, 15)),
    "coco": ((0, 0), (1, 0), (2, 0), (3, 1), (4, 2), (5, 0), (6, 0), (7, 5), (8, 6), (9, 7), (10, 8), (11, 0), (12, 0), (13, 11), (14, 12),
             (15, 13), (16, 14)),
}


def _check(x, dataset):
    if dataset not in BONE_PAIRS:
        raise ValueError(f"The dataset type {dataset} is not supported")
    if x.dim() < 4 or x.shape[-1] not in (2, 3):
        raise ValueError(f"expected [..., M, T, V, C] with C in (2, 3), got {tuple(x.shape)}")


def joint_to_bone(keypoint, dataset="nturgb+d"):
    """bone[v1] = joint[v1] - joint[v2] for every (v1, v2) pair; 2-D layouts: score channel = mean of the two scores."""
    _check(keypoint, dataset)
    pairs = BONE_PAIRS[dataset]
    V = keypoint.shape[-2]
    src = torch.arange(V, device=keypoint.device)                 # joints without a pair keep a zero bone (x - x)
    has = torch.zeros(V, dtype=torch.bool, device=keypoint.device)
    for v1, v2 in pairs:
        src[v1] = v2
        has[v1] = True
    x = keypoint.float()
    other = x.index_select(-2, src)
    bone = (x - other) * has.view(V, 1).to(x.dtype)
    if keypoint.shape[-1] == 3 and dataset in ("openpose", "coco"):
        score = (x[..., 2] + other[..., 2]) / 2 * has.to(x.dtype)
        bone = torch.cat([bone[..., :2], score.unsqueeze(-1)], -1)
    return bone


def to_motion(data, dataset="nturgb+d"):
    """motion[t] = data[t+1] - data[t], last frame zero; 2-D layouts: score channel = mean of consecutive scores."""
    _check(data, dataset)
    x = data.float()
    motion = torch.zeros_like(x)
    motion[..., :-1, :, :] = x[..., 1:, :, :] - x[..., :-1, :, :]
    if data.shape[-1] == 3 and dataset in ("openpose", "coco"):
        motion[..., :-1, :, 2] = (x[..., :-1, :, 2] + x[..., 1:, :, 2]) / 2
    return motion


def gen_ske_feat(keypoint, keypoint_score=None, dataset="nturgb+d", feats=("j",), axis=-1):
    """GenSkeFeat: feats from {'j', 'b', 'jm', 'bm'} concatenated along `axis` in the order given."""
    if keypoint_score is not None:
        if dataset == "nturgb+d" or keypoint.shape[-1] != 2:
            raise ValueError("Only 2D keypoints have keypoint_score.")
        keypoint = torch.cat([keypoint, keypoint_score.unsqueeze(-1).to(keypoint.dtype)], -1)
    out = {"j": keypoint.float()}
    if "b" in feats or "bm" in feats:
        out["b"] = joint_to_bone(keypoint, dataset)
    if "jm" in feats:
        out["jm"] = to_motion(out["j"], dataset)
    if "bm" in feats:
        out["bm"] = to_motion(out["b"], dataset)
    for f in feats:
        if f not in out:
            raise ValueError(f"unknown feature {f}")
    return torch.cat([out[f] for f in feats], axis)


def format_gcn_input(keypoint, num_person=2, mode="zero", num_clips=1, keypoint_score=None):
    """FormatGCNInput for one sample [M, T, V, C] -> [num_clips, num_person, T / num_clips, V, C] (contiguous)."""
    if mode not in ("zero", "loop"):
        raise ValueError(mode)
    if keypoint_score is not None:
        keypoint = torch.cat([keypoint, keypoint_score.unsqueeze(-1).to(keypoint.dtype)], -1)
    M = keypoint.shape[0]
    if M < num_person:
        pad = torch.zeros((num_person - M,) + tuple(keypoint.shape[1:]), dtype=keypoint.dtype, device=keypoint.device)
        keypoint = torch.cat([keypoint, pad], 0)
        if mode == "loop":
            keypoint = keypoint.clone()
            keypoint[1:] = keypoint[0]
    elif M > num_person:
        keypoint = keypoint[:num_person]
    M, T, V, C = keypoint.shape
    if T % num_clips != 0:
        raise ValueError("T must be a multiple of num_clips")
    return keypoint.reshape(M, num_clips, T // num_clips, V, C).permute(1, 0, 2, 3, 4).contiguous()
