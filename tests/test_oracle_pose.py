"""CPU: the numpy pose oracle (oracle/pose_np.py, se3_np.py) against the reference's own outputs:
 - tests/golden/posehead_real.npz: objective, autograd gradient, every L-BFGS evaluation and the final
   pose of DPoseSE3Head.solve on decimated real network data;
 - tests/golden/posehead_unit.npz: the reference's unit-test recipe (tests/unit_test_pose_head.py);
 - the lietorch stand-in the reference's unit tests pass with (oracle/lietorch)."""
import os

import numpy as np
import pytest

from oracle import pose_np, se3_np
from oracle.detrand import unpack


@pytest.fixture(scope="module")
def real(golden_dir):
    g = np.load(os.path.join(golden_dir, "posehead_real.npz"))
    h, w = g["shape"]
    args = (g["flow"], g["pcl1"], g["pcl2w"], g["w1"], g["w2"], unpack(g["m1"], (h, w)), unpack(g["m2w"], (h, w)),
            g["K"], g["lw"])
    return g, args


def test_objective_and_gradient_match_reference_autograd(real):
    g, args = real
    d = pose_np._prep(*args)
    for row in g["objective_probe"]:
        f, (l2, l3), grad = pose_np.evaluate(d, row[:7])
        assert abs(f - row[7]) <= 1e-12 * abs(row[7])
        assert abs(l2 - row[8]) <= 1e-12 * abs(row[8])
        assert abs(l3 - row[9]) <= 1e-12 * abs(row[9])
        np.testing.assert_allclose(grad, row[10:], rtol=1e-10, atol=1e-14)


def test_lbfgs_trajectory_matches_reference(real):
    g, args = real
    tr = []
    X, lg, n_evals = pose_np.lbfgs_solve(*args, max_iter=20, trace=tr)
    assert n_evals == len(g["eval_pose"])
    for k, (pose, grad, _) in enumerate(tr):
        np.testing.assert_allclose(pose, g["eval_pose"][k], atol=1e-12)
        np.testing.assert_allclose(grad, g["eval_grad"][k], rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(X, g["sol_vec"], atol=1e-12)
    np.testing.assert_allclose(lg, g["sol_log"], atol=1e-12)


def test_gauss_newton_differs_from_reference_stop_point(real):
    """the reference stops L-BFGS early; the GN minimiser is a different point."""
    g, args = real
    X, lg, it = pose_np.gn_solve(*args)
    assert it <= 10
    d = pose_np._prep(*args)
    f_gn = pose_np.evaluate(d, X)[0]
    f_ref = pose_np.evaluate(d, g["sol_vec"].astype(np.float64))[0]
    assert f_gn <= f_ref + 1e-15
    assert np.abs(pose_np.evaluate(d, X)[2]).max() < 1e-9


def test_unit_recipe_matches_reference(golden_dir):
    u = np.load(os.path.join(golden_dir, "posehead_unit.npz"))
    R = 128
    valid = unpack(u["valid"], (2, R, R))
    ones = np.ones((R, R), np.float32)
    for i in range(2):
        X, lg, ne = pose_np.lbfgs_solve(u["flow"][i], u["pcl"][i], u["pcl_t"][i], ones, ones, valid[i],
                                        np.ones((R, R), bool), u["K"], np.array([0.001, 1.0]), max_iter=100)
        assert ne == u["n_evals"][i]
        np.testing.assert_allclose(X, u["sol_vec"][i], atol=1e-8)
        # the reference's own acceptance bound (unit_test_pose_head.py:48-49)
        assert np.abs(lg - u["xi_gt"][i]).sum() < 0.05


def test_se3_matches_standin():
    import torch
    from oracle.lietorch import SE3
    rng = np.random.default_rng(0)
    for _ in range(20):
        xi = rng.normal(size=6) * 0.3
        X = se3_np.exp(xi)
        Xt = SE3.exp(torch.tensor(xi))
        np.testing.assert_allclose(X, Xt.data.numpy(), atol=1e-14)
        np.testing.assert_allclose(se3_np.log(X), xi, atol=1e-13)
        Y = se3_np.exp(rng.normal(size=6) * 0.2)
        np.testing.assert_allclose(se3_np.mul(X, Y), (Xt * SE3(torch.tensor(Y))).data.numpy(), atol=1e-14)
        np.testing.assert_allclose(se3_np.mul(X, se3_np.inv(X)), se3_np.identity(), atol=1e-14)
        p = rng.normal(size=(5, 3))
        np.testing.assert_allclose(se3_np.act(X, p), (Xt * torch.tensor(p)).numpy(), atol=1e-14)
    np.testing.assert_allclose(se3_np.log(se3_np.exp(np.full(6, 1e-9))), np.full(6, 1e-9), atol=1e-20)


def test_torch_solver_matches_reference(real):
    """oracle/pose_torch.solve (torch.optim.LBFGS over the lietorch stand-in: the solver of bench.py's reference-on-GPU arm)
    stops at the reference's pose after the reference's number of evaluations."""
    import torch
    from oracle import pose_torch
    g, args = real
    t = [torch.from_numpy(np.ascontiguousarray(a)) for a in args]
    flow, pcl1, pcl2w, w1, w2, m1, m2w, K, lw = t
    X, lg, n_evals = pose_torch.solve(flow[None], pcl1[None], pcl2w[None], w1[None, None], w2[None, None], m1[None, None],
                                      m2w[None, None], K[None].float(), lw[None].float(), lbgfs_iters=20)
    assert n_evals == len(g["eval_pose"])
    np.testing.assert_allclose(X[0].numpy(), g["sol_vec"], atol=1e-10)
    np.testing.assert_allclose(lg[0].numpy(), g["sol_log"], atol=1e-10)
