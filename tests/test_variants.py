"""BASELINE.json configs 4 and 5, against golden outputs of the unmodified reference (oracle/make_golden.py --variants):

  * infer_f2f_nw (configuration/infer_f2f_nw.yaml: conf_weighing False, weights = 1): 5-frame trajectory and the
    ``trajectory.freiburg`` text written by the reference's own writer (core/utils/trajectory.py:17-23)
    -> tests/golden/e2e_nw_384x352.npz (committed);
  * only3d_1a7ix98y.pth with the 2-D term switched off (loss_weight[1] = 0; new-repo mode
    ``residuals: '3d'``) at 1280x1024 -> oracle/_ref/golden_only3d_1280x1024.npz (git-ignored, ships with gpurun).
"""
import os

import numpy as np
import pytest
import torch

from oracle import se3_np
from oracle.detrand import unpack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAINED = os.path.join(ROOT, "oracle", "_ref", "trained")
NW = os.path.join(ROOT, "tests", "golden", "e2e_nw_384x352.npz")
ONLY3D = os.path.join(ROOT, "oracle", "_ref", "golden_only3d_1280x1024.npz")
SLAM_NW = {"frame2frame": True, "dist_thr": 0.05, "depth_clipping": [1, 250], "debug": False, "conf_weighing": False,
           "average_pts": False, "lbgfs_iters": 20}


def _pose_err(a, b):
    d = se3_np.mul(se3_np.inv(a.astype(np.float64)), b.astype(np.float64))
    return np.linalg.norm(se3_np.log(d)[3:]), np.linalg.norm(a[:3] - b[:3]) / max(np.linalg.norm(b[:3]), 1e-12)


def _frames(g, dev=None):
    W, H = [int(v) for v in g["size"]]
    L = torch.from_numpy(g["imgs_l"].astype(np.float32))
    R = torch.from_numpy(g["imgs_r"].astype(np.float32))
    M = torch.from_numpy(np.stack([unpack(m, (1, H, W)) for m in g["masks_in"]]))
    if dev is not None:
        L, R, M = L.to(dev), R.to(dev), M.to(dev)
    return L, R, M, (W, H)


# ---------------------------------------------------------------------------------------------------------------
# CPU: host logic (wire format) and the oracle port on the nw configuration
# ---------------------------------------------------------------------------------------------------------------
def test_freiburg_writer_matches_reference_bytes(tmp_path):
    """save_trajectory reproduces the reference writer byte for byte; read_freiburg round-trips it (mm <-> m)."""
    import rpe_b200  # noqa: F401
    from rpe_b200.core.utils.trajectory import read_freiburg, save_trajectory
    from rpe_b200.lie import SE3
    g = np.load(NW)
    traj = [{"camera-pose": SE3(torch.from_numpy(p)[None]), "timestamp": i} for i, p in enumerate(g["traj"])]
    save_trajectory(traj, str(tmp_path))
    ours = open(tmp_path / "trajectory.freiburg", "rb").read()
    assert ours == g["freiburg"].tobytes()
    back = read_freiburg(str(tmp_path / "trajectory.freiburg")).vec().numpy()
    np.testing.assert_allclose(back, g["traj"], rtol=1e-6, atol=1e-6)


@pytest.mark.skipif(not os.path.isfile(os.path.join(TRAINED, "poseNet_2xf8up4b.pth")), reason="reference checkpoint not present")
def test_oracle_tracker_nw_matches_reference():
    from oracle import pipeline_ref
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = np.load(NW)
    L, R, M, _ = _frames(g)
    sd = torch.load(os.path.join(TRAINED, "poseNet_2xf8up4b.pth"), map_location="cpu", weights_only=False)["state_dict"]
    trk = pipeline_ref.RefTracker(sd, g["K"], float(g["bf"]), conf_weighing=False)
    for k in range(3):                                                  # 2 pairs keep the CPU suite short
        pose = trk.step(L[k:k + 1], R[k:k + 1], M[k:k + 1].clone())
        rot, trans = _pose_err(pose, g["traj"][k])
        assert rot < 1e-6 and (k == 0 or trans < 1e-5)
        if k > 0:
            assert trk.last["n_evals"] == int(g["n_evals"][k - 1])


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_gpu_nw_trajectory_and_freiburg(tmp_path, precision):
    """Per-frame tracker and the batched engine on the nw configuration; trajectory.freiburg within the pose gate."""
    ckpt = os.path.join(TRAINED, "poseNet_2xf8up4b.pth")
    if not os.path.isfile(ckpt):
        pytest.skip("reference checkpoint not shipped")
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    from rpe_b200.core.utils.trajectory import read_freiburg, save_trajectory
    from rpe_b200.lie import SE3
    g = np.load(NW)
    dev = torch.device("cuda:0")
    L, R, M, size = _frames(g, dev)
    cfg = dict(SLAM_NW, precision=precision)
    est = PoseEstimator(cfg, torch.tensor(g["K"]), float(g["bf"]), ckpt, size).to(dev)
    traj = []
    for k in range(L.shape[0]):
        pose, _, _, _ = est(L[k:k + 1], R[k:k + 1], M[k:k + 1].clone())
        traj.append({"camera-pose": SE3(pose.data.clone()), "timestamp": k})
    for k in range(1, L.shape[0]):
        rot, trans = _pose_err(traj[k]["camera-pose"].vec().cpu().numpy().reshape(7), g["traj"][k])
        print(f"nw/{precision}: frame {k} rot {rot:.2e} rad, rel. trans {trans:.2e}")
        assert rot < 1e-4 and trans < 1e-4
    save_trajectory(traj, str(tmp_path))
    ours = np.loadtxt(tmp_path / "trajectory.freiburg")
    ref = np.loadtxt(g["freiburg"].tobytes().decode().splitlines())
    assert ours.shape == ref.shape and np.array_equal(ours[:, 0], ref[:, 0])
    assert np.abs(ours[:, 1:4] - ref[:, 1:4]).max() < 1e-4 * np.abs(ref[:, 1:4]).max()      # metres
    assert np.abs(ours[:, 4:] - ref[:, 4:]).max() < 1e-4
    # batched engine (one chunk of 3 + a ragged chunk of 1) == per-frame tracker
    est2 = PoseEstimator(cfg, torch.tensor(g["K"]), float(g["bf"]), ckpt, size).to(dev)
    out, failed = est2.infer_sequence(L, R, M.clone(), chunk=3)
    assert not bool(failed.any())
    for k in range(1, L.shape[0]):
        rot, trans = _pose_err(out[k].numpy(), g["traj"][k])
        assert rot < 1e-4 and trans < 1e-4
    assert read_freiburg(str(tmp_path / "trajectory.freiburg")).vec().shape == (L.shape[0], 7)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_gpu_only3d_1280x1024(precision):
    """Config 4: only3d checkpoint, 3-D residual only, full-resolution 1280x1024 stereo."""
    ckpt = os.path.join(TRAINED, "only3d_1a7ix98y.pth")
    if not (os.path.isfile(ckpt) and os.path.isfile(ONLY3D)):
        pytest.skip("only3d checkpoint / 1280x1024 golden not shipped (oracle/make_golden.py --variants)")
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    g = np.load(ONLY3D)
    dev = torch.device("cuda:0")
    L, R, M, size = _frames(g, dev)
    assert size == (1280, 1024)
    cfg = dict(SLAM_NW, conf_weighing=True, precision=precision, residuals="3d")
    est = PoseEstimator(cfg, torch.tensor(g["K"]), float(g["bf"]), ckpt, size).to(dev)
    assert float(est.model.loss_weight[1]) == 0.0 and float(est.model.loss_weight[0]) > 0.0
    for k in range(2):
        pose, _, flow, _ = est(L[k:k + 1], R[k:k + 1], M[k:k + 1].clone())
    p = pose.vec().cpu().numpy().reshape(7)
    rot, trans = _pose_err(p, g["traj"][1])
    epe_t = np.sqrt(((flow[0, :, ::4, ::4].cpu().numpy() - g["s_time_flow_ds4"]) ** 2).sum(0))
    epe_s = np.sqrt(((est.frame.flow[0, :, ::4, ::4].cpu().numpy() - g["s_stereo_flow2_ds4"]) ** 2).sum(0))
    print(f"only3d/{precision}: rot {rot:.2e} rad, rel. trans {trans:.2e}, abs trans {np.linalg.norm(p[:3] - g['traj'][1][:3]):.2e} mm, "
          f"flow EPE {epe_t.mean():.2e} / {epe_s.mean():.2e}")
    assert epe_t.mean() < 1e-2 and epe_s.mean() < 1e-2
    # Tolerances: flow 1e-2 px EPE and rotation 1e-4 rad as in the north star.  Relative translation: 3e-4 for THIS config only.
    # The frame-to-frame translation of this pair is 0.27 mm, so 1e-4 relative is 27 nm, and the reference does not reproduce
    # ITSELF to that: multiplying the outputs of its own fp32 convolutions by (1 + 6e-8 N(0,1)) -- one ulp, i.e. another BLAS or
    # summation order -- moves its pose by 0.7e-4 .. 1.9e-4 relative translation (three seeds; 2.7e-4 at 1e-5 noise;
    # tools/sensitivity_study.py --only3d, profiles/r2_sensitivity_study_only3d_1280x1024.txt).  Measured here: cuDNN fp32 trunk
    # 1.06e-4, fp16x3 trunk 2.19e-4 (5.8e-5 mm absolute).  Configs 1-3 and 5 keep the 1e-4 gate on their well-conditioned pairs.
    assert rot < 1e-4 and trans < 3e-4 and np.linalg.norm(p[:3] - g["traj"][1][:3]) < 1e-4
