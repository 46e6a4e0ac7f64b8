"""SE(3) pose head with the reference's interface (/root/reference/core/pose/pose_head.py:5-79,
DeclarativeLayerLie: core/optimization/declerative_node_lie.py:224-247, 283-284), backed by the fused
sm_100a residual-reduction + on-device solver (rpe_pose_solve).  Inference forward only: the training
backward of the declarative node is out of scope."""
import torch

from ... import ops
from ...lie import SE3


class PoseParameter:
    """What ``DPoseSE3Head.solve`` returns in place of lietorch's LieGroupParameter: ``.group`` (SE3,
    float64), ``.log()``; plus the device-resident solver record."""

    def __init__(self, solution):
        self.solution = solution
        self.group = SE3(solution.pose64.unsqueeze(1))            # (n,1,7) like SE3.Identity(n, 1)

    def log(self):
        return self.solution.log64.unsqueeze(1)

    def retr(self):
        return self.group


class DPoseSE3Head:
    """solver: "lbfgs_ref" (exact replica of the reference's truncated L-BFGS; parity mode) or "gn"
    (Gauss-Newton with in-kernel 6x6 Cholesky, the north-star variant)."""

    def __init__(self, img_coordinates=None, lbgfs_iters=100, dbg=False, solver="lbfgs_ref", gn_iters=10):
        self.img_coordinates = img_coordinates                    # implied by the pixel index inside the kernel
        self.lbgfs_iters = lbgfs_iters
        self.solver = solver
        self.gn_iters = gn_iters
        self.losses = []

    @staticmethod
    def _prep(xs):
        flow, pcl1, pcl2, w1, w2, m1, m2, K, lw = xs
        f = lambda t: t.detach().float().contiguous()
        b = lambda t: t.detach().bool().contiguous()
        n = flow.shape[0]
        return (f(flow), f(pcl1), f(pcl2), None if w1 is None else f(w1), None if w2 is None else f(w2), b(m1), b(m2),
                f(K), f(lw).expand(n, 2).contiguous())

    def solve(self, *xs):
        args = self._prep(xs)
        if self.solver == "gn":
            sol = ops.pose_solve(*args, mode=ops.SOLVER_GN, max_iter=self.gn_iters)
        else:
            sol = ops.pose_solve(*args, mode=ops.SOLVER_LBFGS_REF, max_iter=self.lbgfs_iters)
        return PoseParameter(sol), None

    def objective(self, *xs, y, backward=False):
        """Objective value per pair at pose y[0] (SE3 / PoseParameter), float64 (pose_head.py:53-58)."""
        pose = y[0]
        pose = pose.group if isinstance(pose, PoseParameter) else pose
        n = xs[0].shape[0]
        init = pose.data.detach().double().reshape(n, 7).contiguous().to(xs[0].device)
        sol = ops.pose_solve(*self._prep(xs), mode=ops.SOLVER_EVAL_ONLY, init_pose=init)
        return sol.loss


class DeclarativeLayerLie(torch.nn.Module):
    """Forward of the reference's declarative layer: (pose embedding (n,1,7) f32, tangent (n,1,6) f32)."""

    def __init__(self, problem):
        super().__init__()
        self.problem = problem

    def forward(self, *inputs):
        with torch.no_grad():
            y, _ = self.problem.solve(*inputs)
        self.last_solution = y.solution
        return y.solution.pose.unsqueeze(1).clone(), y.solution.log.unsqueeze(1).clone()
