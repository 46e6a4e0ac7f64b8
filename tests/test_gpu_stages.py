"""GPU parity: the CUDA stages (through the C ABI) against the numpy oracle on identical inputs and
against the golden outputs of the reference itself.  Bit-exact for masks / validity / indices; stated
tolerances for floats."""
import os

import numpy as np
import pytest
import torch

from oracle import geom_np, pose_np
from oracle.detrand import det_uniform, unpack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    import rpe_b200  # noqa: F401
    from rpe_b200 import ops as _ops
    return _ops


def dev(a, device="cuda"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


# ---------------------------------------------------------------------------------------------------
# stage 2: depth + back-projection
# ---------------------------------------------------------------------------------------------------
def test_depth_proj_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "stages_small.npz"))
    mask = dev(np.ones((1, 1, 40, 56), bool))
    depth, valid, pcl = ops.depth_proj(dev(g["geom_sflow"]), dev(np.array([8.8], np.float32)), dev(g["geom_K"]), mask)
    assert np.array_equal(depth.cpu().numpy(), g["geom_depth"][:, None])               # IEEE division: exact
    assert np.array_equal(np.packbits(valid.cpu().numpy().reshape(-1)), g["geom_valid"])  # bit-exact
    assert np.array_equal(mask.cpu().numpy(), valid.cpu().numpy())                      # mask &= valid in place
    np.testing.assert_allclose(pcl.cpu().numpy(), g["geom_pcl"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("shape", [(512, 640), (352, 384), (1024, 1280)])
def test_depth_proj_vs_oracle_full_size(ops, shape):
    H, W = shape
    n = 2
    sflow = det_uniform((n, 2, H, W), 5, -40.0, 2.0)
    sflow[0, 0, 0, :4] = [0.0, -8.8, np.nan, np.inf]
    K = np.array([[[0.625 * W, 0, W / 2], [0, 0.625 * W, H / 2], [0, 0, 1]]] * n, np.float32)
    K[1, 0, 1] = 0.3                                                                  # general (skewed) K
    bf = np.array([8.8, 5.5], np.float32)
    m_in = det_uniform((n, 1, H, W), 6, 0, 1) > 0.1
    mask = dev(m_in.copy())
    depth, valid, pcl = ops.depth_proj(dev(sflow), dev(bf), dev(K), mask)
    for b in range(n):
        d, v = geom_np.depth_from_stereo_flow(sflow[b], bf[b])
        assert np.array_equal(depth[b, 0].cpu().numpy(), d)
        assert np.array_equal(valid[b, 0].cpu().numpy(), v)
        assert np.array_equal(mask[b, 0].cpu().numpy(), m_in[b, 0] & v)
        np.testing.assert_allclose(pcl[b].cpu().numpy(), geom_np.proj(d, K[b]), rtol=3e-6, atol=1e-6)


def test_proj_rescale_round_trip(ops):
    H, W = 64, 96
    depth = det_uniform((1, 1, H, W), 9, 0.01, 1.0)
    K = np.array([[[60.0, 0, 48], [0, 60.0, 32], [0, 0, 1]]], np.float32)
    scale = np.float32(1.0 / 250.0)
    got = ops.proj(dev(depth), dev(K), rescale=float(scale)).cpu().numpy()
    d1 = ((depth[0, 0] / scale).astype(np.float32) * scale).astype(np.float32)        # pose_estimator.py:115,121
    np.testing.assert_allclose(got[0], geom_np.proj(d1, K[0]), rtol=3e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------
# stage 5: warps
# ---------------------------------------------------------------------------------------------------
def test_warp_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "stages_small.npz"))
    x8 = det_uniform((1, 8, 40, 56), 21, -2.0, 2.0)
    m = det_uniform((1, 1, 40, 56), 23, 0.0, 1.0) > 0.2
    flow = dev(g["warp_flow"])
    a, b, c, mw = ops.warp8_mask(dev(x8[:, 0:3]), dev(x8[:, 3:6]), dev(x8[:, 6:8]), dev(m), flow)
    got = torch.cat((a, b, c), 1).cpu().numpy()
    np.testing.assert_allclose(got, g["warp_bilinear"], atol=2e-6)
    assert np.array_equal(np.packbits(mw.cpu().numpy().reshape(-1)), g["warp_mask_out"])  # bit-exact


@pytest.mark.parametrize("shape", [(512, 640), (352, 384)])
def test_warp_vs_oracle_full_size(ops, shape):
    H, W = shape
    n = 2
    x8 = det_uniform((n, 8, H, W), 31, -3.0, 3.0)
    flow = det_uniform((n, 2, H, W), 32, -30.0, 30.0)
    flow[0, :, 0, :8] = np.array([[0.5, 1.5, 2.5, 3.5, -0.5, 640.0, -700.0, 0.0]] * 2)  # ties + out of range
    flow[1, 0] = 0.5                                                                   # a whole image of ties
    m = det_uniform((n, 1, H, W), 33, 0.0, 1.0) > 0.3
    a, b, c, mw = ops.warp8_mask(dev(x8[:, 0:3]), dev(x8[:, 3:6]), dev(x8[:, 6:8]), dev(m), dev(flow))
    got = torch.cat((a, b, c), 1).cpu().numpy()
    for k in range(n):
        np.testing.assert_allclose(got[k], geom_np.remap_from_flow(x8[k], flow[k]), atol=3e-6)
        assert np.array_equal(mw[k, 0].cpu().numpy(), geom_np.remap_mask_nearest(m[k, 0], flow[k]))


def test_warp_identity_and_idempotence(ops):
    H, W = 128, 160
    x = det_uniform((1, 3, H, W), 41)
    z = torch.zeros((1, 2, H, W), device="cuda")
    m = det_uniform((1, 1, H, W), 42, 0, 1) > 0.5
    a, _, _, mw = ops.warp8_mask(dev(x), None, None, dev(m), z)
    # zero flow is the identity up to the fp32 normalise/un-normalise round trip (<= 1e-4 px)
    np.testing.assert_allclose(a.cpu().numpy(), x, atol=2e-3)
    assert np.array_equal(mw.cpu().numpy(), m)


def test_downsample8_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "stages_small.npz"))
    x8 = det_uniform((1, 8, 40, 56), 21, -2.0, 2.0)
    got = ops.downsample8_cat([dev(x8[:, 0:2]), dev(x8[:, 2:5]), dev(x8[:, 5:8])]).cpu().numpy()
    np.testing.assert_allclose(got, g["down8"], atol=1e-6)


# ---------------------------------------------------------------------------------------------------
# stages 3 + 4: fused residual reduction and solver
# ---------------------------------------------------------------------------------------------------
def _real(golden_dir):
    g = np.load(os.path.join(golden_dir, "posehead_real.npz"))
    h, w = g["shape"]
    args = (g["flow"], g["pcl1"], g["pcl2w"], g["w1"], g["w2"], unpack(g["m1"], (h, w)), unpack(g["m2w"], (h, w)),
            g["K"], g["lw"])
    return g, args


def _to_dev(args, reps=1):
    flow, p1, p2, w1, w2, m1, m2, K, lw = args
    return (dev(np.stack([flow] * reps)), dev(np.stack([p1] * reps)), dev(np.stack([p2] * reps)),
            dev(np.stack([w1[None]] * reps)), dev(np.stack([w2[None]] * reps)), dev(np.stack([m1[None]] * reps)),
            dev(np.stack([m2[None]] * reps)), dev(np.stack([K.astype(np.float32)] * reps)),
            dev(np.stack([lw.astype(np.float32)] * reps)))


def test_pose_eval_matches_reference_autograd(ops, golden_dir):
    g, args = _real(golden_dir)
    d_args = _to_dev(args)
    for row in g["objective_probe"]:
        sol = ops.pose_solve(*d_args, mode=ops.SOLVER_EVAL_ONLY, init_pose=dev(row[None, :7].copy()), with_hessian=True)
        out = sol.out.cpu().numpy()[0]
        assert abs(out[13] - row[7]) <= 1e-11 * abs(row[7])           # f
        assert abs(out[14] - row[8]) <= 1e-11 * abs(row[8])           # L2d
        assert abs(out[15] - row[9]) <= 1e-11 * abs(row[9])           # L3d
        np.testing.assert_allclose(out[19:25], row[10:], rtol=1e-9, atol=1e-14)
        # GN Hessian against the oracle
        dd = pose_np._prep(*args)
        Hm = pose_np.evaluate(dd, row[:7], want_hessian=True)[3]
        np.testing.assert_allclose(sol.hessian().cpu().numpy()[0], Hm, rtol=1e-9, atol=1e-16)


def test_pose_lbfgs_trajectory_matches_reference(ops, golden_dir):
    """Every objective evaluation of the reference's L-BFGS run is reproduced: same poses, same gradients,
    same number of evaluations, same final pose (tolerance 1e-9; north-star gate is 1e-4)."""
    g, args = _real(golden_dir)
    sol = ops.pose_solve(*_to_dev(args), mode=ops.SOLVER_LBFGS_REF, max_iter=20, trace_cap=32)
    out = sol.out.cpu().numpy()[0]
    n_ref = len(g["eval_pose"])
    assert int(out[16]) == n_ref
    tr = sol.trace.cpu().numpy()[0]
    np.testing.assert_allclose(tr[:n_ref, :7], g["eval_pose"], atol=1e-9)
    np.testing.assert_allclose(tr[:n_ref, 7:13], g["eval_grad"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(out[:7], g["sol_vec"], atol=1e-9)
    np.testing.assert_allclose(out[7:13], g["sol_log"], atol=1e-9)
    np.testing.assert_allclose(sol.pose.cpu().numpy()[0], g["sol_vec"].astype(np.float32), atol=1e-7)
    np.testing.assert_allclose(sol.log.cpu().numpy()[0], g["sol_log"].astype(np.float32), atol=1e-7)


def test_pose_unit_recipe(ops, golden_dir):
    """The reference's own unit test (tests/unit_test_pose_head.py:38-50), lbgfs_iters=100, as a batch of 2
    independent pairs."""
    u = np.load(os.path.join(golden_dir, "posehead_unit.npz"))
    R = 128
    valid = unpack(u["valid"], (2, 1, R, R))
    ones = torch.ones((2, 1, R, R), device="cuda")
    K = dev(np.stack([u["K"]] * 2).astype(np.float32))
    lw = dev(np.array([[0.001, 1.0]] * 2, np.float32))
    sol = ops.pose_solve(dev(u["flow"]), dev(u["pcl"]), dev(u["pcl_t"]), ones, ones, dev(valid),
                         torch.ones((2, 1, R, R), dtype=torch.bool, device="cuda"), K, lw, max_iter=100)
    out = sol.out.cpu().numpy()
    assert out[:, 16].astype(int).tolist() == u["n_evals"].tolist()
    np.testing.assert_allclose(out[:, :7], u["sol_vec"], atol=1e-8)
    assert (np.abs(out[:, 7:13] - u["xi_gt"]).sum(1) < 0.05).all()
    # objective at the solution ~ 0 (unit_test_pose_head.py:45-46)
    assert (out[:, 13] < 1e-5).all()


def test_pose_batch_independence_and_grouping(ops, golden_dir):
    """Pairs in one launch are solved independently: a batch of 5 copies + a perturbed pair gives
    the single-pair answers bit for bit, for any number of concurrent CTA groups."""
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib
    g, args = _real(golden_dir)
    single = ops.pose_solve(*_to_dev(args), max_iter=20).out.cpu().numpy()[0]
    d6 = list(_to_dev(args, reps=6))
    d6[0] = d6[0].clone()
    d6[0][5] += 0.25                                                       # pair 5 differs
    lone = ops.pose_solve(*[t[5:6].contiguous() for t in d6], max_iter=20).out.cpu().numpy()[0]
    try:
        for groups in (1, 2, 8):
            assert _lib.lib().rpe_pose_set_groups(groups) == 0
            out = ops.pose_solve(*d6, max_iter=20).out.cpu().numpy()
            for k in range(5):
                assert np.array_equal(out[k, :19], single[:19])
            assert np.array_equal(out[5, :19], lone[:19])
            assert not np.array_equal(out[5, :7], single[:7])
    finally:
        _lib.lib().rpe_pose_set_groups(8)


def test_pose_gauss_newton_matches_oracle(ops, golden_dir):
    g, args = _real(golden_dir)
    X, lg, it = pose_np.gn_solve(*args, max_iter=10)
    sol = ops.pose_solve(*_to_dev(args), mode=ops.SOLVER_GN, max_iter=10)
    out = sol.out.cpu().numpy()[0]
    assert out[18] == 0
    np.testing.assert_allclose(out[:7], X, atol=1e-10)
    np.testing.assert_allclose(out[7:13], lg, atol=1e-10)
    assert np.abs(out[19:25]).max() < 1e-9                                   # gradient vanishes at the GN solution


def test_pose_full_size_vs_oracle(ops):
    """640x512 synthetic problem (the bench shape) with masks, weights, out-of-image flow."""
    H, W = 512, 640
    rng_pose = np.array([0.012, -0.008, 0.02, 0.004, -0.006, 0.003])
    from oracle import se3_np
    T = se3_np.exp(rng_pose)
    K = np.array([[400.0, 0, 320], [0, 400.0, 256], [0, 0, 1]], np.float32)
    depth = det_uniform((H, W), 51, 0.2, 0.9)
    pcl1 = geom_np.proj(depth, K)
    p2 = (se3_np.act(T, pcl1.reshape(3, -1).T.astype(np.float64)).T).reshape(3, H, W)
    q = K.astype(np.float64) @ p2.reshape(3, -1)
    uv = geom_np.img_coords(H, W)[:2].astype(np.float64)
    flow = (q[:2] / q[2] - uv).reshape(2, H, W).astype(np.float32) + det_uniform((2, H, W), 52, -0.3, 0.3)
    pcl2 = (p2 + det_uniform((3, H, W), 53, -0.002, 0.002)).astype(np.float32)
    w1 = det_uniform((H, W), 54, 0.0, 1.0)
    w2 = det_uniform((H, W), 55, 0.0, 1.0)
    m1 = det_uniform((H, W), 56, 0, 1) > 0.1
    m2 = det_uniform((H, W), 57, 0, 1) > 0.1
    lw = np.array([0.9296, 1.0004], np.float32)
    args = (flow, pcl1, pcl2, w1, w2, m1, m2, K, lw)
    tr = []
    X, lg, ne = pose_np.lbfgs_solve(*args, max_iter=20, trace=tr)
    sol = ops.pose_solve(*_to_dev(args), max_iter=20, trace_cap=32)
    out = sol.out.cpu().numpy()[0]
    assert int(out[16]) == ne
    np.testing.assert_allclose(out[:7], X, atol=1e-9)
    trg = sol.trace.cpu().numpy()[0]
    for k, (p, gr, f) in enumerate(tr):
        np.testing.assert_allclose(trg[k, :7], p, atol=1e-9)
        assert abs(trg[k, 13] - f) <= 1e-10 * abs(f)
    assert np.abs(lg - rng_pose).max() < 5e-3                                # and it recovers the motion


def _small_problem(H=48, W=64, seed=70):
    from oracle import se3_np
    T = se3_np.exp(np.array([0.01, -0.006, 0.015, 0.003, -0.004, 0.002]))
    K = np.array([[60.0, 0, W / 2], [0, 60.0, H / 2], [0, 0, 1]], np.float32)
    pcl1 = geom_np.proj(det_uniform((H, W), seed, 0.2, 0.9), K)
    p2 = (se3_np.act(T, pcl1.reshape(3, -1).T.astype(np.float64)).T).reshape(3, H, W)
    q = K.astype(np.float64) @ p2.reshape(3, -1)
    uv = geom_np.img_coords(H, W)[:2].astype(np.float64)
    flow = (q[:2] / q[2] - uv).reshape(2, H, W).astype(np.float32)
    ones = np.ones((H, W), np.float32)
    return [flow, pcl1, p2.astype(np.float32), ones.copy(), ones.copy(), np.ones((H, W), bool), np.ones((H, W), bool), K,
            np.array([0.9296, 1.0004], np.float32)]


def test_pose_degenerate_inputs_match_the_oracle(ops):
    """The domain's nulls: a pair without a single valid pixel (objective and gradient are exactly 0: one evaluation, identity pose,
    not flagged), non-finite flow / points / weights on valid pixels (pose_head.py:26 zeroes non-finite 2-D residuals; everything
    else propagates like the reference's arithmetic does), and a singular Gauss-Newton system."""
    # (a) nothing valid
    args = _small_problem()
    args[5] = np.zeros_like(args[5])
    X, lg, ne = pose_np.lbfgs_solve(*args, max_iter=20)
    sol = ops.pose_solve(*_to_dev(args), max_iter=20)
    out = sol.out.cpu().numpy()[0]
    assert ne == 1 and int(out[16]) == 1
    assert np.array_equal(out[:7], np.array([0, 0, 0, 0, 0, 0, 1.0])) and np.array_equal(X, out[:7])
    assert np.array_equal(sol.log.cpu().numpy()[0], np.zeros(6, np.float32))
    # (b) a non-finite flow on a valid pixel: the reference zeroes the residual (pose_head.py:26) but its autograd still multiplies the
    #     zero gradient by the non-finite difference, so the pose comes out NaN (checked with oracle/pose_torch.py, the autograd
    #     restatement) and the tracker's guard keeps the previous pose (pose_estimator.py:81-85).  Same here, on both sides.
    for seed, (c, y, x, v) in ((71, (0, 3, 5, np.nan)), (72, (1, 7, 9, np.inf))):
        args = _small_problem(seed=seed)
        args[0][c, y, x] = v
        X, _, _ = pose_np.lbfgs_solve(*args, max_iter=20)
        out = ops.pose_solve(*_to_dev(args), max_iter=20).out.cpu().numpy()[0]
        assert np.isnan(X).any() and np.isnan(out[:7]).any(), (X, out[:7])
    # (c) a NaN point poisons the 3-D term the same way
    args = _small_problem(seed=73)
    args[1][2, 11, 13] = np.nan
    X, _, _ = pose_np.lbfgs_solve(*args, max_iter=20)
    out = ops.pose_solve(*_to_dev(args), max_iter=20).out.cpu().numpy()[0]
    assert np.isnan(X).any() and np.isnan(out[:7]).any()
    # (d) Gauss-Newton without a valid pixel: singular normal equations -> flagged, identity kept
    args = _small_problem(seed=73)
    args[5] = np.zeros_like(args[5])
    sol = ops.pose_solve(*_to_dev(args), mode=ops.SOLVER_GN, max_iter=10)
    out = sol.out.cpu().numpy()[0]
    assert np.array_equal(out[:7], np.array([0, 0, 0, 0, 0, 0, 1.0]))


def test_pose_rejects_bad_arguments(ops):
    from rpe_b200._lib import RpeError
    z = torch.zeros((1, 2, 8, 6), device="cuda")                            # H*W = 48 ok, but wrong companions
    with pytest.raises(RpeError):
        ops.pose_solve(z, z, z, None, None, z.bool(), z.bool(), torch.eye(3, device="cuda")[None],
                       torch.ones((1, 2), device="cuda"))


# ---------------------------------------------------------------------------------------------------
# convex up-sampling
# ---------------------------------------------------------------------------------------------------
def test_convex_upsample_vs_torch_formula(ops):
    """raft.py:66-77 restated with torch ops (test-only) vs the CUDA kernel."""
    import torch.nn.functional as F
    B, h, w = 2, 44, 48
    flow = dev(det_uniform((B, 2, h, w), 61, -4, 4))
    mask = dev(det_uniform((B, 576, h, w), 62, -3, 3))
    m = torch.softmax(mask.view(B, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(B, 2, 9, 1, 1, h, w)
    ref = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(B, 2, 8 * h, 8 * w)
    got = ops.convex_upsample8(flow, mask)
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), atol=2e-5)
