"""ORACLE / TEST INFRASTRUCTURE ONLY -- torch restatement of the reference's pose head, runnable on any torch device.

Used by oracle/pipeline_ref.RefTracker(solver="torch"): the "reference on the GPU" arm of bench.py, which must behave like
the stock reference on CUDA -- ~100 small ATen kernels per objective evaluation, autograd through the lietorch stand-in,
``torch.optim.LBFGS`` with a ``float(loss)`` host synchronisation per iteration.  Never imported by the product package.

Follows /root/reference/core/pose/pose_head.py:
  reprojection_objective  :12-33     depth_objective  :35-51     objective  :53-58     solve  :60-79
and core/geometry/pinhole_transforms.py: create_img_coords_t :7-19, transform_forward :28-30, project :90-99.
Pinned by tests/test_oracle_pose.py against the same golden L-BFGS runs as oracle/pose_np.py (outputs of the reference).
"""
import torch

from .lietorch import SE3, LieGroupParameter


def img_coords(h, w, device):
    xs = torch.linspace(0, w - 1, w, device=device).repeat(1, h, 1) + 0.5
    ys = torch.linspace(0, h - 1, h, device=device).repeat(1, w, 1).transpose(1, 2) + 0.5
    return torch.vstack([xs.flatten(), ys.flatten(), torch.ones(h * w, device=device)])


def _transform(pts, T):
    return (T * pts.permute(0, 2, 1)).permute(0, 2, 1)


def _project(pts, K, T):
    ipts = torch.bmm(K, _transform(pts, T))
    depth = torch.clamp(ipts[:, -1], 1e-12, None).unsqueeze(1)
    ipts = torch.cat((ipts[:, :2], torch.ones_like(ipts[:, None, 2])), dim=1)
    return ipts / depth


def objective(ic, flow, pcl1, pcl2, w1, w2, m1, m2, K, lw, y):
    n, _, h, w = flow.shape
    # 3-D point-to-point term
    aligned = _transform(pcl1.view(n, 3, -1), y)
    r3 = torch.sum((aligned - pcl2.view(n, 3, -1)) ** 2, dim=1)
    r3 *= w2.view(n, -1)
    r3[~(m1 & m2).view(n, -1)] = 0.0
    loss3d = torch.mean(r3, dim=-1)
    # 2-D reprojection term
    warped = _project(pcl1.view(n, 3, -1), K, y)[:, :2]
    flow_off = ic[None, :2] + flow.view(n, 2, -1)
    r2 = torch.sum((flow_off - warped) ** 2, dim=1)
    r2 *= w1.view(n, -1)
    inside = (flow_off[:, 0] > 0) & (flow_off[:, 1] > 0) & (flow_off[:, 0] < w) & (flow_off[:, 1] < h)
    bad = torch.isinf(r2) | torch.isnan(r2) | ~inside.view(n, -1) | ~m1.view(n, -1)
    r2[bad] = 0.0
    loss2d = torch.mean(r2, dim=1) / (h * w)
    return lw[:, 1] * loss2d + lw[:, 0] * loss3d


def solve(flow, pcl1, pcl2, w1, w2, m1, m2, K, lw, lbgfs_iters=20):
    """-> (pose (n,7) float64 tensor, tangent (n,6), number of objective evaluations).  n must be 1 for parity."""
    xs = [x.detach().clone() for x in (flow, pcl1, pcl2, w1, w2, m1, m2, K, lw)]
    xs = [x.double() if x.dtype == torch.float32 else x for x in xs]
    n, _, h, w = xs[0].shape
    ic = img_coords(h, w, xs[0].device)
    evals = [0]
    with torch.enable_grad():
        y = LieGroupParameter(SE3.Identity(n, 1, device=xs[0].device, requires_grad=True, dtype=torch.float64))
        opt = torch.optim.LBFGS([y], lr=1.0, max_iter=lbgfs_iters, line_search_fn=None)

        def fun():
            opt.zero_grad()
            loss = objective(ic, *xs, y).sum()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(y, 10)
            evals[0] += 1
            return loss

        opt.step(fun)
    return y.group.vec().detach().reshape(n, 7), y.log().detach().reshape(n, 6), evals[0]
