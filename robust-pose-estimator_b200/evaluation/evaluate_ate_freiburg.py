"""``eval(gt, pred)`` with the reference's interface (/root/reference/evaluation/evaluate_ate_freiburg.py:6-33): ATE-RMSE / RPE
of a ``trajectory.freiburg`` file (or {timestamp: SE3} dict) against a ground-truth trajectory with the same time stamps."""
import numpy as np

from ..core.metrics.trajectory_metrics import absolute_trajectory_error, relative_pose_error, total_trajectory_length
from ..core.utils.trajectory import read_freiburg


def _as_dict(t):
    if isinstance(t, dict):
        return t
    poses, stamps = read_freiburg(t, ret_stamps=True)
    return {int(k): poses[i] for i, k in enumerate(stamps)}


def eval(gt_list, pred_list, delta=1, offset=0, ret_align_T=False, ignore_failed_pos=False):
    gt, pred = _as_dict(gt_list), _as_dict(pred_list)
    gt_max = max(gt.keys())
    pred_poses, gt_poses = [], []
    for k in sorted(pred.keys()):
        if (k + offset > 0) and (k + offset < gt_max):                  # exact synchronisation is assumed (same stamps)
            pred_poses.append(pred[k].matrix().reshape(4, 4).double().cpu().numpy())
            gt_poses.append(gt[k + offset].matrix().reshape(4, 4).double().cpu().numpy())
    pred_poses, gt_poses = np.stack(pred_poses), np.stack(gt_poses)
    ate, trans_err, T, valid = absolute_trajectory_error(gt_poses, pred_poses, ret_align_T=True, ignore_failed_pos=ignore_failed_pos)
    rpe_t, rpe_r = relative_pose_error(gt_poses, pred_poses, delta=delta, ignore_failed_pos=ignore_failed_pos)
    if ret_align_T:
        return ate, np.mean(rpe_t), np.mean(rpe_r), trans_err, rpe_t, rpe_r, T, gt_poses, valid
    return ate, np.mean(rpe_t), np.mean(rpe_r), trans_err, rpe_t, rpe_r


def get_traj_length(gt_list, pred_list=None, offset=0):
    gt = _as_dict(gt_list)
    if pred_list is None:
        return total_trajectory_length(list(gt.values()))
    pred = _as_dict(pred_list)
    gt_max = max(gt.keys())
    return total_trajectory_length([gt[k + offset] for k in sorted(pred.keys()) if (k + offset > 0) and (k + offset < gt_max)])
