"""ORACLE / TEST INFRASTRUCTURE ONLY.  numpy float64 restatement of the reference's SE(3) pose head:

  DPoseSE3Head.reprojection_objective   /root/reference/core/pose/pose_head.py:12-33
  DPoseSE3Head.depth_objective          /root/reference/core/pose/pose_head.py:35-51
  DPoseSE3Head.objective                /root/reference/core/pose/pose_head.py:53-58
  DPoseSE3Head.solve                    /root/reference/core/pose/pose_head.py:60-79
  project / transform                   /root/reference/core/geometry/pinhole_transforms.py:28-30,90-99
  torch.optim.LBFGS.step (third party, torch==1.13 pinned in requirements.txt:6; the installed
                          torch 2.11 lbfgs.py:333-537 is the same algorithm)
  torch.nn.utils.clip_grad_norm_(y, 10) /root/reference/core/pose/pose_head.py:76
  DeclarativeFunctionLie.forward        /root/reference/core/optimization/declerative_node_lie.py:224-247

Pinned against outputs of the reference itself (tests/golden/e2e_384x352.npz: per-evaluation poses
and autograd gradients of the reference's L-BFGS run, final poses; tests/golden/posehead_unit.npz:
the reference's unit-test recipe) in tests/test_oracle_pose.py.

All inputs are single-pair (n = 1): flow (2,H,W), pcl1/pcl2 (3,H,W), w1/w2 (H,W),
m1/m2 bool (H,W), K (3,3), lw (2,) = [w3d, w2d].  Arithmetic in float64 from float32 inputs exactly
like solve() (pose_head.py:63-64).
"""
import numpy as np

from . import se3_np as se3


def _prep(flow, pcl1, pcl2, w1, w2, m1, m2, K, lw):
    H, W = flow.shape[-2:]
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    d = dict(H=H, W=W, N=H * W,
             flow=f64(flow).reshape(2, -1), p1=f64(pcl1).reshape(3, -1), p2=f64(pcl2).reshape(3, -1),
             w1=f64(w1).reshape(-1), w2=f64(w2).reshape(-1),
             m1=np.asarray(m1).astype(bool).reshape(-1), m2=np.asarray(m2).astype(bool).reshape(-1),
             K=f64(K).reshape(3, 3), lw=f64(lw).reshape(2))
    v, u = np.meshgrid(np.arange(H, dtype=np.float64) + 0.5, np.arange(W, dtype=np.float64) + 0.5, indexing="ij")
    d["uv"] = np.stack((u.reshape(-1), v.reshape(-1)))            # create_img_coords_t (pinhole_transforms.py:7-19)
    return d


def evaluate(d, pose, want_grad=True, want_hessian=False):
    """f, (L2d, L3d), grad(6) wrt the LEFT perturbation xi=[tau|phi] at `pose`, optional GN Hessian.

    Gradient convention d(Xp)/dxi = [I | -[Xp]x]  (pinhole_transforms.py:39-41; lietorch).
    """
    H, W, N = d["H"], d["W"], d["N"]
    K, lw = d["K"], d["lw"]
    R = se3.rotmat(pose[3:])
    pp = R @ d["p1"] + pose[:3, None]                              # transform()  (3, N)
    # ---- 3D point-to-point term (pose_head.py:43-51)
    r3 = pp - d["p2"]
    e3 = (r3 * r3).sum(0) * d["w2"]
    v3 = d["m1"] & d["m2"]
    e3 = np.where(v3, e3, 0.0)
    L3 = e3.sum() / N                                              # torch.mean over HW
    # ---- 2D reprojection term (pose_head.py:18-29, project: pinhole_transforms.py:90-99)
    q = K @ pp
    den = np.maximum(q[2], 1e-12)                                  # torch.clamp(z, 1e-12, None)
    pi = q[:2] / den
    t = d["uv"] + d["flow"]
    r2 = t - pi
    e2 = (r2 * r2).sum(0) * d["w1"]
    inside = (t[0] > 0) & (t[1] > 0) & (t[0] < W) & (t[1] < H)
    bad = np.isinf(e2) | np.isnan(e2) | ~inside | ~d["m1"]
    e2 = np.where(bad, 0.0, e2)
    L2 = e2.sum() / N / (H * W)
    f = lw[1] * L2 + lw[0] * L3
    out = [f, (L2, L3)]
    if not (want_grad or want_hessian):
        return tuple(out)
    # ---- analytic gradient; zero where the reference zeroes the residual (index assignment)
    s3 = lw[0] / N
    s2 = lw[1] / N / (H * W)
    a3 = np.where(v3, 2.0 * d["w2"], 0.0) * r3                     # dE3/dp'
    clamp_pass = (q[2] >= 1e-12).astype(np.float64)                # clamp gradient: 1 inside [min, inf)
    g_pi = np.where(bad, 0.0, 2.0 * d["w1"]) * (-r2)               # dE2/dpi   (r2 = t - pi)
    g_q = np.stack((g_pi[0] / den, g_pi[1] / den,
                    -(g_pi[0] * q[0] + g_pi[1] * q[1]) / (den * den) * clamp_pass))
    a2 = K.T @ g_q                                                 # dE2/dp'
    a = s3 * a3 + s2 * a2
    g = np.concatenate((a.sum(1), np.cross(pp.T, a.T).sum(0)))     # J^T a,  J = [I | -[p']x]
    out.append(g)
    if want_hessian:
        # Gauss-Newton: sum_i c_i J_r^T J_r with J_r3 = J, J_r2 = -dpi/dp' J
        Hm = np.zeros((6, 6))
        ok3 = v3
        c3 = np.where(ok3, 2.0 * s3 * d["w2"], 0.0)
        c2 = np.where(bad, 0.0, 2.0 * s2 * d["w1"])
        zeros = np.zeros(N)
        ones = np.ones(N)
        X, Y, Z = pp
        J = np.stack([np.stack([ones, zeros, zeros, zeros, Z, -Y]),
                      np.stack([zeros, ones, zeros, -Z, zeros, X]),
                      np.stack([zeros, zeros, ones, Y, -X, zeros])])            # (3, 6, N)
        Hm += np.einsum("n,kan,kbn->ab", c3, J, J)
        dq = np.stack([np.stack([1.0 / den, zeros, -q[0] / (den * den) * clamp_pass]),
                       np.stack([zeros, 1.0 / den, -q[1] / (den * den) * clamp_pass])])   # (2, 3, N) dpi/dq
        dp = np.einsum("ijn,jk->ikn", dq, K)                                    # dpi/dp'  (2, 3, N)
        J2 = np.einsum("ikn,kan->ian", dp, J)                                   # (2, 6, N)
        Hm += np.einsum("n,kan,kbn->ab", c2, J2, J2)
        out.append(Hm)
    return tuple(out)


def clip_grad(g, max_norm=10.0):
    """torch.nn.utils.clip_grad_norm_: g *= clamp(max_norm / (||g||_2 + 1e-6), max=1)."""
    coef = max_norm / (np.linalg.norm(g) + 1e-6)
    return g * min(coef, 1.0)


def lbfgs_solve(flow, pcl1, pcl2, w1, w2, m1, m2, K, lw, max_iter=20, trace=None,
                tolerance_grad=1e-7, tolerance_change=1e-9, history_size=100, lr=1.0):
    """torch.optim.LBFGS(lr=1, max_iter, line_search_fn=None).step(closure) on a LieGroupParameter.

    Returns (pose7 float64, log6 float64, n_evals).  `trace` (list) receives one
    (pose7, raw_grad6, loss) per objective evaluation.
    """
    d = _prep(flow, pcl1, pcl2, w1, w2, m1, m2, K, lw)
    max_eval = max_iter * 5 // 4
    X = se3.identity()

    def closure(X):
        f, _, g = evaluate(d, X)
        if trace is not None:
            trace.append((X.copy(), g.copy(), f))
        return f, clip_grad(g)

    loss, g = closure(X)
    evals = 1
    if np.abs(g).max() <= tolerance_grad:
        return X, se3.log(X), evals
    old_dirs, old_stps, ro = [], [], []
    H_diag = 1.0
    dvec = t = prev_g = prev_loss = None
    n_iter = 0
    while n_iter < max_iter:
        n_iter += 1
        if n_iter == 1:
            dvec = -g
        else:
            y = g - prev_g
            s = dvec * t
            ys = float(y @ s)
            if ys > 1e-10:
                if len(old_dirs) == history_size:
                    old_dirs.pop(0), old_stps.pop(0), ro.pop(0)
                old_dirs.append(y)
                old_stps.append(s)
                ro.append(1.0 / ys)
                H_diag = ys / float(y @ y)
            num_old = len(old_dirs)
            al = [0.0] * num_old
            qv = -g
            for i in range(num_old - 1, -1, -1):
                al[i] = float(old_stps[i] @ qv) * ro[i]
                qv = qv - al[i] * old_dirs[i]
            r = qv * H_diag
            for i in range(num_old):
                be = float(old_dirs[i] @ r) * ro[i]
                r = r + (al[i] - be) * old_stps[i]
            dvec = r
        prev_g = g.copy()
        prev_loss = loss
        t = min(1.0, 1.0 / np.abs(g).sum()) * lr if n_iter == 1 else lr
        gtd = float(g @ dvec)
        if gtd > -tolerance_change:
            break
        X = se3.retract(X, t * dvec)                               # LieGroupParameter.add_(d, alpha=t)
        if n_iter != max_iter:
            loss, g = closure(X)
            evals += 1
        if n_iter == max_iter:
            break
        if evals >= max_eval:
            break
        if np.abs(g).max() <= tolerance_grad:
            break
        if np.abs(dvec * t).max() <= tolerance_change:
            break
        if abs(loss - prev_loss) < tolerance_change:
            break
    return X, se3.log(X), evals


def gn_solve(flow, pcl1, pcl2, w1, w2, m1, m2, K, lw, max_iter=10, damping=0.0, tol=1e-12, trace=None):
    """Gauss-Newton with a 6x6 Cholesky solve and left exp-map update (north-star stage 4; the
    linear solve mirrors ddn's _solve_linear_system, /root/reference/core/ddn/ddn/pytorch/node.py:270-297,
    which the reference only uses in the training backward -- SURVEY.md D1)."""
    d = _prep(flow, pcl1, pcl2, w1, w2, m1, m2, K, lw)
    X = se3.identity()
    n = 0
    for n in range(1, max_iter + 1):
        f, _, g, Hm = evaluate(d, X, want_hessian=True)
        if trace is not None:
            trace.append((X.copy(), g.copy(), f))
        Hm = 0.5 * (Hm + Hm.T) + damping * np.eye(6)
        L = np.linalg.cholesky(Hm)
        step = -np.linalg.solve(L.T, np.linalg.solve(L, g))
        X = se3.retract(X, step)
        if np.abs(step).max() < tol:
            break
    return X, se3.log(X), n
