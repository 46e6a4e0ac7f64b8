"""GPU end-to-end parity of the drop-in API (PoseEstimator / PoseNet.infer) against golden outputs of the
reference's CPU fp32 run with the shipped poseNet_2xf8up4b.pth weights (tests/golden/e2e_384x352.npz,
oracle/_ref/golden_full.npz).  North-star gates: flow EPE <= 1e-2 px, pose <= 1e-4 rad / 1e-4 relative
translation, masks bit-exact stage-wise on identical inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import se3_np
from oracle.detrand import unpack

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
FULL = os.path.join(ROOT, "oracle", "_ref", "golden_full.npz")

SLAM = {"frame2frame": True, "dist_thr": 0.05, "depth_clipping": [1, 250], "debug": False, "conf_weighing": True,
        "average_pts": False, "lbgfs_iters": 20}


def _need_ckpt():
    if not os.path.isfile(CKPT):
        pytest.skip("reference checkpoint not shipped (oracle/_ref/trained is created by oracle/make_golden.py)")


def _run_tracker(g, precision="fp32", solver="lbfgs_ref"):
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    W, H = [int(v) for v in g["size"]]
    cfg = dict(SLAM, precision=precision, solver=solver)
    est = PoseEstimator(cfg, torch.tensor(g["K"]), float(g["bf"]), CKPT, (W, H)).cuda()
    poses, details = [], []
    for i in range(g["imgs_l"].shape[0]):
        limg = torch.from_numpy(g["imgs_l"][i].astype(np.float32))[None].cuda()
        rimg = torch.from_numpy(g["imgs_r"][i].astype(np.float32))[None].cuda()
        mask = torch.from_numpy(unpack(g["masks_in"][i], (1, 1, H, W))).cuda()
        pose, _, flow, weights = est(limg, rimg, mask)
        poses.append(pose.vec().cpu().numpy().reshape(7))
        details.append((flow, weights, est.frame))
    return est, np.stack(poses), details


def _pose_err(a, b):
    """rotation angle (rad) and relative translation error between two pose 7-vectors."""
    d = se3_np.mul(se3_np.inv(a.astype(np.float64)), b.astype(np.float64))
    rot = np.linalg.norm(se3_np.log(d)[3:])
    trans = np.linalg.norm(a[:3] - b[:3]) / max(np.linalg.norm(b[:3]), 1e-12)
    return rot, trans


def _epe(a, b):
    return np.sqrt(((a - b) ** 2).sum(0))


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("name", ["e2e_384x352", "full_640x512"])
def test_tracker_matches_reference(golden_dir, name, precision):
    """Both parity-grade modes: fp32 (cuDNN fp32 trunk) and fp16x3 (update operator on the tcgen05 kernels)."""
    _need_ckpt()
    path = os.path.join(golden_dir, "e2e_384x352.npz") if name == "e2e_384x352" else FULL
    if not os.path.isfile(path):
        pytest.skip("full-size golden dump not shipped")
    g = np.load(path)
    W, H = [int(v) for v in g["size"]]
    est, poses, details = _run_tracker(g, precision=precision)
    # ---- trajectory (absolute poses in mm, chained on the host/device like the reference)
    for k in range(1, poses.shape[0]):
        rot, trans = _pose_err(poses[k], g["traj"][k])
        print(f"{name}/{precision}: frame {k} pose error rot {rot:.2e} rad, rel. trans {trans:.2e}")
        assert rot < 1e-4 and trans < 1e-4, f"frame {k}: rot {rot:.2e} rad, trans {trans:.2e}"
    # ---- flows of the last pair
    flow, weights, frame = details[-1]
    epe_t = _epe(flow[0].cpu().numpy(), g["s_time_flow"])
    epe_s = _epe(frame.flow[0].cpu().numpy(), g["s_stereo_flow2"])
    print(f"{name}: time-flow EPE mean {epe_t.mean():.2e} max {epe_t.max():.2e}; stereo EPE mean {epe_s.mean():.2e} max {epe_s.max():.2e}")
    assert epe_t.mean() < 1e-2 and epe_s.mean() < 1e-2
    assert epe_t.max() < 5e-2 and epe_s.max() < 5e-2
    # ---- confidence maps
    c1 = weights[0][0].cpu().numpy().astype(np.float32)
    c2 = weights[1][0].cpu().numpy().astype(np.float32)
    assert np.abs(c1 - g["s_conf1"].astype(np.float32)).max() < 5e-3
    assert np.abs(c2 - g["s_conf2"].astype(np.float32)).max() < 5e-3
    # ---- masks: the tracker's own masks differ from the reference's only where the flow differs; report, then
    # check bit-exactness stage-wise on the reference's own flows below
    m2 = frame.mask[0, 0].cpu().numpy()
    diff = (m2 != unpack(g["s_mask2_valid"], (H, W))).mean()
    print(f"{name}: mask2&valid mismatch fraction with e2e flows {diff:.2e}")
    assert diff < 1e-3


@pytest.mark.parametrize("name", ["e2e_384x352", "full_640x512"])
def test_masks_bit_exact_on_reference_flows(golden_dir, name):
    """Stage-wise gate: identical inputs (the reference's flows) -> validity / occlusion masks bit-exact."""
    import rpe_b200  # noqa: F401
    from rpe_b200 import ops
    path = os.path.join(golden_dir, "e2e_384x352.npz") if name == "e2e_384x352" else FULL
    if not os.path.isfile(path):
        pytest.skip("full-size golden dump not shipped")
    g = np.load(path)
    W, H = [int(v) for v in g["size"]]
    sflow = torch.from_numpy(g["s_stereo_flow2"])[None].cuda()
    tflow = torch.from_numpy(g["s_time_flow"])[None].cuda()
    K = torch.tensor(g["K"], dtype=torch.float32)[None].cuda()
    scale = torch.tensor(1 / 250)                                         # fp32, like the tracker's buffer
    bf = (torch.tensor(float(g["bf"])).float() * scale).reshape(1).cuda()
    mask_in = torch.from_numpy(unpack(g["masks_in"][-1], (1, 1, H, W))).cuda()
    depth2, valid, pcl2 = ops.depth_proj(sflow, bf, K, mask_in)
    assert np.array_equal(np.packbits(mask_in.cpu().numpy().reshape(-1)), g["s_mask2_valid"])
    img2 = torch.from_numpy(g["imgs_l"][-1].astype(np.float32))[None].cuda()
    pcl2w, img2w, sflow2w, mask2w = ops.warp8_mask(pcl2, img2, sflow, mask_in, tflow)
    assert np.array_equal(np.packbits(mask2w.cpu().numpy().reshape(-1)), g["s_mask2w"])
    # float stages on identical inputs
    sub = (lambda a: a.reshape(a.shape[0], -1)[:, ::7]) if name == "e2e_384x352" else (lambda a: a)
    np.testing.assert_allclose(sub(pcl2[0].cpu().numpy()), g["s_pcl2"], rtol=3e-6, atol=1e-6)
    np.testing.assert_allclose(sub(pcl2w[0].cpu().numpy()), g["s_pcl2w"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(sub(img2w[0].cpu().numpy()), g["s_img2w"], rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(sub(sflow2w[0].cpu().numpy()), g["s_sflow2w"], rtol=1e-5, atol=2e-5)
    d1 = torch.from_numpy(g["s_depth1_norm"])[None, None].cuda()
    pcl1 = ops.proj(d1, K)
    np.testing.assert_allclose(sub(pcl1[0].cpu().numpy()), g["s_pcl1"], rtol=3e-6, atol=1e-6)


def test_full_size_corr_and_solver_on_reference_tensors():
    """640x512: CorrBlock on the reference's feature maps vs the reference's lookup output, and the solver on
    the reference's pose-head inputs vs every evaluation of the reference's L-BFGS run."""
    import rpe_b200  # noqa: F401
    from rpe_b200 import ops
    if not os.path.isfile(FULL):
        pytest.skip("full-size golden dump not shipped")
    g = np.load(FULL)
    f1, f2 = torch.from_numpy(g["c_fmap1"]).cuda(), torch.from_numpy(g["c_fmap2"]).cuda()
    for prec, tol in ((ops.CORR_TF32X3, 2e-4), (ops.CORR_TF32, 3e-2)):
        cp = ops.CorrPyramid(f1, f2, precision=prec)
        for it in (0, 11):
            out = cp(torch.from_numpy(g[f"c_coords{it}"]).cuda()).cpu().numpy()
            err = np.abs(out - g[f"c_lookup{it}"]).max()
            print(f"corr precision {prec} iter {it}: max abs err {err:.2e} (|ref| max {np.abs(g[f'c_lookup{it}']).max():.1f})")
            assert err < tol
    H, W = 512, 640
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    args = (dev(g["s_time_flow"][None]), dev(g["s_pcl1"][None]), dev(g["s_pcl2w"][None]), dev(g["s_conf1"][None]),
            dev(g["s_conf2"][None]), dev(unpack(g["s_mask1"], (1, 1, H, W))), dev(unpack(g["s_mask2w"], (1, 1, H, W))),
            dev(g["K"].astype(np.float32)[None]), dev(g["s_loss_weight"][None]))
    sol = ops.pose_solve(*args, max_iter=20, trace_cap=32)
    out = sol.out.cpu().numpy()[0]
    ref_p, ref_g = g["pair1_eval_pose"], g["pair1_eval_grad"]
    assert int(out[16]) == len(ref_p)
    tr = sol.trace.cpu().numpy()[0]
    np.testing.assert_allclose(tr[:len(ref_p), :7], ref_p, atol=1e-9)
    np.testing.assert_allclose(tr[:len(ref_p), 7:13], ref_g, rtol=1e-6, atol=1e-12)


def test_throughput_precisions_report(golden_dir):
    """Not a gate: how far the faster trunk precisions drift from the fp32 CPU reference."""
    _need_ckpt()
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    for prec in ("tf32", "bf16", "fp16"):
        est, poses, details = _run_tracker(g, precision=prec)
        flow = details[-1][0]
        epe = _epe(flow[0].cpu().numpy(), g["s_time_flow"])
        rot, trans = _pose_err(poses[-1], g["traj"][-1])
        print(f"precision {prec}: flow EPE mean {epe.mean():.2e} max {epe.max():.2e}; pose rot {rot:.2e} trans {trans:.2e}")
        assert np.isfinite(poses).all()


def test_gauss_newton_solver_mode(golden_dir):
    """solver='gn' runs through the same API; it converges to the minimiser, which is a different point than the
    reference's truncated L-BFGS -- the deviation is reported, only sanity is asserted."""
    _need_ckpt()
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    est, poses, _ = _run_tracker(g, solver="gn")
    rot, trans = _pose_err(poses[-1], g["traj"][-1])
    print(f"GN vs reference L-BFGS: rot {rot:.2e} rad, rel. trans {trans:.2e}")
    assert rot < 5e-3 and trans < 5e-2


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_graphed_streaming_tracker_matches_reference(golden_dir, precision):
    """config['cuda_graph']: PoseEstimator.forward replays one captured CUDA graph per frame (latency path).  Same trajectory
    as the reference, and as the eager per-frame tracker, frame by frame; a second sequence on the same estimator re-uses
    the captured graph."""
    _need_ckpt()
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    from rpe_b200.lie import SE3
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    W, H = [int(v) for v in g["size"]]
    _, eager, _ = _run_tracker(g, precision=precision)
    est = PoseEstimator(dict(SLAM, precision=precision, cuda_graph=True), torch.tensor(g["K"]), float(g["bf"]), CKPT, (W, H)).cuda()
    for rep in range(2):
        est.frame = est.last_frame = None
        est.last_pose = SE3.Identity(1, device="cuda")
        for i in range(g["imgs_l"].shape[0]):
            limg = torch.from_numpy(g["imgs_l"][i].astype(np.float32))[None].cuda()
            rimg = torch.from_numpy(g["imgs_r"][i].astype(np.float32))[None].cuda()
            mask = torch.from_numpy(unpack(g["masks_in"][i], (1, 1, H, W))).cuda()
            pose = est(limg, rimg, mask)[0].vec().cpu().numpy().reshape(7)
            rot, trans = _pose_err(pose, g["traj"][i]) if i else (0.0, 0.0)
            rot_e, trans_e = _pose_err(pose, eager[i]) if i else (0.0, 0.0)
            print(f"graphed/{precision} pass {rep} frame {i}: vs reference rot {rot:.2e} trans {trans:.2e}; vs eager rot {rot_e:.2e} trans {trans_e:.2e}")
            assert rot < 1e-4 and trans < 1e-4 and rot_e < 1e-5 and trans_e < 1e-5
    assert not est.check_failures()


def test_infer_sequence_from_pinned_host_frames(golden_dir):
    """infer_sequence fed from pinned uint8 host frames (chunked uploads on a copy stream behind the compute) returns exactly
    what the device-resident call returns."""
    _need_ckpt()
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    from rpe_b200.lie import SE3
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    W, H = [int(v) for v in g["size"]]
    est = PoseEstimator(dict(SLAM, precision="fp16x3"), torch.tensor(g["K"]), float(g["bf"]), CKPT, (W, H)).cuda()
    idx = [0, 1, 2, 1, 0, 2]
    L8 = torch.from_numpy(np.clip(np.round(g["imgs_l"]), 0, 255).astype(np.uint8))[idx].contiguous().pin_memory()
    R8 = torch.from_numpy(np.clip(np.round(g["imgs_r"]), 0, 255).astype(np.uint8))[idx].contiguous().pin_memory()
    M = torch.from_numpy(np.stack([unpack(g["masks_in"][i], (1, H, W)) for i in range(3)]))[idx].contiguous().pin_memory()
    ref, failed_ref = est.infer_sequence(L8.cuda(), R8.cuda(), M.cuda(), chunk=2)              # device-resident uint8 frames
    est.last_pose = SE3.Identity(1, device="cuda")
    got, failed = est.infer_sequence(L8, R8, M, chunk=2)
    assert got.shape == (6, 7) and torch.equal(got, ref) and torch.equal(failed, failed_ref)
    # float frames (the reference's tensors) take the normalised two-plane stem instead of the raw single-plane one: same poses
    # up to fp32 rounding
    est.last_pose = SE3.Identity(1, device="cuda")
    flt, _ = est.infer_sequence(L8.cuda().float(), R8.cuda().float(), M.cuda(), chunk=2)
    assert float((flt - ref).abs().max()) < 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize("chunk,graphs,precision", [(1, False, "fp32"), (2, False, "fp32"), (2, True, "fp32"), (2, False, "fp16x3"),
                                                    (3, True, "fp16x3")])
def test_batched_engine_matches_reference(golden_dir, chunk, graphs, precision):
    """PoseEstimator.infer_sequence (chunked engine, feature reuse, optional CUDA graph, host composition through
    rpe_compose_trajectory_host) reproduces the reference trajectory."""
    _need_ckpt()
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    W, H = [int(v) for v in g["size"]]
    est = PoseEstimator(dict(SLAM, precision=precision), torch.tensor(g["K"]), float(g["bf"]), CKPT, (W, H)).cuda()
    L = torch.from_numpy(g["imgs_l"].astype(np.float32)).cuda()
    R = torch.from_numpy(g["imgs_r"].astype(np.float32)).cuda()
    M = torch.from_numpy(np.stack([unpack(g["masks_in"][i], (1, H, W)) for i in range(3)])).cuda()
    # 5 frames: 0 1 2 1 0 (exercise more than one chunk); the first three must match the golden trajectory
    idx = [0, 1, 2, 1, 0]
    traj, failed = est.infer_sequence(L[idx], R[idx], M[idx], chunk=chunk, use_graphs=graphs)
    assert traj.shape == (5, 7) and not failed.any()
    for k in range(1, 3):
        rot, trans = _pose_err(traj[k].numpy(), g["traj"][k])
        assert rot < 1e-4 and trans < 1e-4, f"frame {k}: rot {rot:.2e} trans {trans:.2e}"
    if precision == "fp32":                 # identical evaluation count of the reference's L-BFGS run
        assert est.last_evals[:2].tolist() == [len(g["pair0_eval_pose"]), len(g["pair1_eval_pose"])]


def test_sequence_edge_cases_and_chunk_invariance(golden_dir):
    """Empty and ragged inputs of the throughput path: one frame (no pair), two frames, chunks that do not divide the sequence,
    a chunk larger than the sequence -- and the result does not depend on the chunking at all (bit-identical poses)."""
    _need_ckpt()
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    from rpe_b200.lie import SE3
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    W, H = [int(v) for v in g["size"]]
    est = PoseEstimator(dict(SLAM, precision="fp16x3"), torch.tensor(g["K"]), float(g["bf"]), CKPT, (W, H)).cuda()
    idx = [0, 1, 2, 1, 0, 2]
    L = torch.from_numpy(g["imgs_l"])[idx].cuda()
    R = torch.from_numpy(g["imgs_r"])[idx].cuda()
    M = torch.from_numpy(np.stack([unpack(g["masks_in"][i], (1, H, W)) for i in range(3)]))[idx].cuda()

    def run(T, chunk):
        est.last_pose = SE3.Identity(1, device="cuda")
        return est.infer_sequence(L[:T], R[:T], M[:T], chunk=chunk)

    traj, failed = run(1, 4)
    assert traj.shape == (1, 7) and failed.shape == (0,) and traj[0].tolist() == [0, 0, 0, 0, 0, 0, 1]
    traj2, failed2 = run(2, 4)
    assert traj2.shape == (2, 7) and failed2.shape == (1,) and not failed2.any()
    ref, _ = run(6, 8)                                    # one chunk holds the whole sequence
    assert torch.equal(ref[:2], traj2)
    for chunk in (1, 2, 4, 5):
        got, failed = run(6, chunk)
        assert got.shape == (6, 7) and failed.shape == (5,) and not failed.any()
        assert torch.equal(got, ref), f"chunk {chunk}: max diff {float((got - ref).abs().max()):.3e}"


def test_fully_masked_frame_keeps_the_pose(golden_dir):
    """A frame whose mask is empty (instrument covering the view).  As the CURRENT frame of a pair it only removes the 3-D term (the 2-D
    term is gated by the previous frame's mask alone, pose_head.py:24-28); as the PREVIOUS frame it leaves no valid pixel at all: the
    objective and its gradient are exactly zero, the solver stops at its first evaluation and the pose stays where it was -- no NaN,
    no failure flag, and the frames before it are tracked as before."""
    _need_ckpt()
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    from rpe_b200.lie import SE3
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    W, H = [int(v) for v in g["size"]]
    est = PoseEstimator(dict(SLAM, precision="fp16x3"), torch.tensor(g["K"]), float(g["bf"]), CKPT, (W, H)).cuda()
    idx = [0, 1, 2, 1, 0]
    L = torch.from_numpy(g["imgs_l"])[idx].cuda()
    R = torch.from_numpy(g["imgs_r"])[idx].cuda()
    M = torch.from_numpy(np.stack([unpack(g["masks_in"][i], (1, H, W)) for i in range(3)]))[idx].cuda()
    ref, failed_ref = est.infer_sequence(L, R, M, chunk=4)
    M2 = M.clone()
    M2[2] = False
    est.last_pose = SE3.Identity(1, device="cuda")
    got, failed = est.infer_sequence(L, R, M2, chunk=4)
    assert torch.isfinite(got).all() and not failed.any() and not failed_ref.any()
    assert torch.equal(got[1], ref[1])                       # pair 0 -> 1 does not see frame 2
    assert est.last_evals[1] > 1                             # pair 1 -> 2: the 2-D term alone still constrains the pose
    assert torch.equal(got[3], got[2]) and est.last_evals[2] == 1          # pair 2 -> 3: nothing valid, pose kept
