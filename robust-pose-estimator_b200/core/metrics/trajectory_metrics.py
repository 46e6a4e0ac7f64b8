"""Trajectory accuracy metrics with the reference's interface (/root/reference/core/metrics/trajectory_metrics.py:7-112):
ATE-RMSE after a closed-form (Horn / Umeyama without scale) alignment, relative pose error, path length.  Host-side numpy on
(n,4,4) pose matrices -- accuracy reporting around the pose path, not part of it."""
import numpy as np
import torch


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def _align(model, data):
    """Rigid transform (4x4) that maps the 3xn point set ``model`` onto ``data`` in the least-squares sense."""
    mu_m, mu_d = model.mean(1, keepdims=True), data.mean(1, keepdims=True)
    cov = (data - mu_d) @ (model - mu_m).T                      # sum of outer products, transposed like the reference's W^T
    U, _, Vh = np.linalg.svd(cov)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0:
        S[2, 2] = -1.0
    rot = U @ S @ Vh
    T = np.eye(4)
    T[:3, :3] = rot
    T[:3, 3] = (mu_d - rot @ mu_m).ravel()
    return T


def absolute_trajectory_error(gt_poses, predicted_poses, prealign=True, ret_align_T=False, ignore_failed_pos=False):
    """-> (ATE-RMSE, per-pose translation errors[, alignment, valid]).  ``ignore_failed_pos``: drop poses equal to their
    predecessor (the tracker repeats the last pose when a pair fails)."""
    gt, pred = _np(gt_poses).astype(np.float64), _np(predicted_poses).astype(np.float64)
    assert len(gt) == len(pred)
    valid = np.ones(len(pred), dtype=bool)
    if ignore_failed_pos:
        valid[1:] = np.array([(pred[i] - pred[i + 1]).sum() != 0 for i in range(len(pred) - 1)], dtype=bool)
    T = None
    if prealign:
        T = _align(pred[valid, :3, 3].T, gt[valid, :3, 3].T)
        pred = T[None] @ pred
    err2 = ((gt[valid, :3, 3] - pred[valid, :3, 3]) ** 2).sum(1)
    ate = float(np.sqrt(err2.mean()))
    if ret_align_T:
        return ate, np.sqrt(err2), T, valid
    return ate, np.sqrt(err2)


def relative_pose_error(gt_poses, predicted_poses, delta=1, ignore_failed_pos=False):
    """-> (translation errors, rotation errors [rad]) of the relative motions over ``delta`` frames."""
    gt, pred = _np(gt_poses).astype(np.float64), _np(predicted_poses).astype(np.float64)
    assert len(gt) == len(pred)
    te, re = [], []
    for i in range(len(gt) - delta):
        if ignore_failed_pos and (pred[i] - pred[i + 1]).sum() == 0:
            continue
        gt_rel = np.linalg.inv(gt[i]) @ gt[i + delta]
        pr_rel = np.linalg.inv(pred[i]) @ pred[i + delta]
        e = np.linalg.inv(gt_rel) @ pr_rel
        te.append(np.sqrt((e[:3, 3] ** 2).sum()))
        re.append(np.arccos(np.clip(0.5 * (np.trace(e[:3, :3]) - 1.0), -1.0, 1.0)))
    return np.asarray(te), np.asarray(re)


def total_trajectory_length(gt_list):
    def loc(g):
        if hasattr(g, "matrix"):                                 # SE3
            return _np(g.matrix()).reshape(4, 4)[:3, 3]
        return _np(g)[:3, 3]
    locs = np.stack([loc(g) for g in gt_list])
    return float(np.sqrt(((locs[1:] - locs[:-1]) ** 2).sum(1)).sum())
