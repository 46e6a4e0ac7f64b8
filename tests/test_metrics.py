"""CPU: trajectory metrics / evaluate_ate_freiburg.eval against the reference's own function (tests/golden/metrics.npz, produced by
oracle/make_golden.py --metrics from /root/reference/evaluation/evaluate_ate_freiburg.py)."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("name,kw", [("plain", {}), ("delta3", {"delta": 3}), ("offset", {"offset": -4}), ("ignore", {"ignore_failed_pos": True})])
def test_evaluate_ate_matches_reference(golden_dir, tmp_path, name, kw):
    import rpe_b200  # noqa: F401
    from rpe_b200.evaluation.evaluate_ate_freiburg import eval as ate_eval, get_traj_length
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    gt, pr = tmp_path / "gt.freiburg", tmp_path / "pred.freiburg"
    gt.write_bytes(g["gt_file"].tobytes()), pr.write_bytes(g["pred_file"].tobytes())
    ate, rpe_t, rpe_r, trans_err, rpe_trans, rpe_rot = ate_eval(str(gt), str(pr), **kw)
    np.testing.assert_allclose([ate, rpe_t, rpe_r], g[name + "_scalars"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(trans_err, g[name + "_trans_error"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(rpe_trans, g[name + "_rpe_trans"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(rpe_rot, g[name + "_rpe_rot"], rtol=1e-6, atol=1e-9)
    assert get_traj_length(str(gt)) > 0
