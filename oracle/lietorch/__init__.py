"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Pure-torch stand-in for the third-party ``lietorch`` package
(github.com/princeton-vl/lietorch, un-pinned in the reference: /root/reference/README.md:37).
It exists so that the UNMODIFIED reference under /root/reference can be imported and run on
CPU in this container to generate golden vectors (oracle/make_golden.py) and to validate the
numpy restatement in oracle/*.py.

Only the API surface the reference touches on the f2f pose path is provided
(call sites: core/pose/pose_head.py:68, core/geometry/pinhole_transforms.py:29,
core/optimization/declerative_node_lie.py:233-234, core/pose/pose_net.py:96-100,
core/pose/pose_estimator.py:42,81,84,90,91, core/utils/trajectory.py:21).

Published lietorch semantics restated here:
  * embedding  data[..., 7] = [tx ty tz qx qy qz qw]
  * tangent    xi = [tau(3) | phi(3)]  (translation first)
  * Exp: q = [sin(th/2)/th * phi, cos(th/2)], t = V(phi) tau,
         V = I + (1-cos th)/th^2 [phi]x + (th - sin th)/th^3 [phi]x^2  (Taylor below 1e-6)
  * LieGroupParameter: zero 6-vector leaf; retr() = Exp(a) * group  (LEFT perturbation);
    add_(u, alpha) moves the group: group <- Exp(alpha u) * group.
All ops are plain differentiable torch ops, so autograd through ``retr()`` at a = 0 yields the
tangent-space gradient d(Xp)/dxi = [I | -[Xp]x] that lietorch's analytic backward produces
(pinned by the reference's tests/unit_test_pinhole_transforms.py:35-53).
"""
import torch

_EPS = 1e-6


def _cross(a, b):
    a, b = torch.broadcast_tensors(a, b)
    return torch.linalg.cross(a, b, dim=-1)


def _quat_mul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz), dim=-1)


def _quat_rotate(q, p):
    """R(q) p using the uv / uuv form lietorch's Act3 uses."""
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * _cross(qv, p)
    return p + qw * uv + _cross(qv, uv)


def _safe(theta2, small):
    # a denominator that is never 0 on the branch torch.where discards
    return torch.where(small, torch.ones_like(theta2), theta2)


def _so3_exp(phi):
    theta2 = (phi * phi).sum(-1, keepdim=True)
    small = theta2 < _EPS * _EPS
    t2 = _safe(theta2, small)
    theta = torch.sqrt(t2)
    imag = torch.where(small, 0.5 - theta2 / 48.0 + theta2 * theta2 / 3840.0, torch.sin(0.5 * theta) / theta)
    real = torch.where(small, 1.0 - theta2 / 8.0 + theta2 * theta2 / 384.0, torch.cos(0.5 * theta))
    return torch.cat((imag * phi, real), dim=-1)


def _left_jacobian_apply(phi, tau):
    theta2 = (phi * phi).sum(-1, keepdim=True)
    small = theta2 < _EPS * _EPS
    t2 = _safe(theta2, small)
    theta = torch.sqrt(t2)
    c1 = torch.where(small, 0.5 - theta2 / 24.0, (1.0 - torch.cos(theta)) / t2)
    c2 = torch.where(small, 1.0 / 6.0 - theta2 / 120.0, (theta - torch.sin(theta)) / (t2 * theta))
    pxt = _cross(phi, tau)
    return tau + c1 * pxt + c2 * _cross(phi, pxt)


def _so3_log(q):
    qv, qw = q[..., :3], q[..., 3:4]
    n2 = (qv * qv).sum(-1, keepdim=True)
    small = n2 < _EPS * _EPS
    n = torch.sqrt(_safe(n2, small))
    # 2*atan(n/w)/n ; atan (not atan2) as published, valid for w > 0 (all poses on this path)
    big = 2.0 * torch.atan(n / qw) / n
    tiny = 2.0 / qw - (2.0 / 3.0) * n2 / (qw * qw * qw)
    return torch.where(small, tiny, big) * qv


def _left_jacobian_inv_apply(phi, t):
    theta2 = (phi * phi).sum(-1, keepdim=True)
    small = theta2 < _EPS * _EPS
    t2 = _safe(theta2, small)
    theta = torch.sqrt(t2)
    half = 0.5 * theta
    c2 = torch.where(small, torch.full_like(theta2, 1.0 / 12.0),
                     (1.0 - theta * torch.cos(half) / (2.0 * torch.sin(half))) / t2)
    pxt = _cross(phi, t)
    return t - 0.5 * pxt + c2 * _cross(phi, pxt)


def _plain(x):
    return x.as_subclass(torch.Tensor) if type(x) is not torch.Tensor else x


class SE3:
    manifold_dim = 6
    embedded_dim = 7

    def __init__(self, data):
        self.data = data.data if isinstance(data, SE3) else _plain(data)

    # ---- constructors -------------------------------------------------------------------
    @staticmethod
    def _shape(batch_shape):
        if len(batch_shape) == 1 and isinstance(batch_shape[0], (tuple, list, torch.Size)):
            return tuple(batch_shape[0])
        return tuple(batch_shape)

    @classmethod
    def Identity(cls, *batch_shape, **kwargs):
        kwargs.pop('requires_grad', None)
        data = torch.zeros(*cls._shape(batch_shape), 7, **kwargs)
        data[..., 6] = 1.0
        return cls(data)

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, device=G.data.device, dtype=G.data.dtype)

    @classmethod
    def InitFromVec(cls, vec):
        return cls(vec)

    @classmethod
    def Random(cls, *batch_shape, sigma=1.0, **kwargs):
        kwargs.pop('requires_grad', None)
        return cls.exp(sigma * torch.randn(*cls._shape(batch_shape), 6, **kwargs))

    @classmethod
    def exp(cls, xi):
        xi = _plain(xi)
        tau, phi = xi[..., :3], xi[..., 3:]
        return cls(torch.cat((_left_jacobian_apply(phi, tau), _so3_exp(phi)), dim=-1))

    # ---- properties ---------------------------------------------------------------------
    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def tangent_shape(self):
        return self.data.shape[:-1] + (6,)

    @property
    def device(self):
        return self.data.device

    @property
    def dtype(self):
        return self.data.dtype

    # ---- group operations ---------------------------------------------------------------
    def log(self):
        phi = _so3_log(self.data[..., 3:])
        return torch.cat((_left_jacobian_inv_apply(phi, self.data[..., :3]), phi), dim=-1)

    def vec(self):
        return self.data

    def inv(self):
        q = self.data[..., 3:]
        qi = torch.cat((-q[..., :3], q[..., 3:]), dim=-1)
        return SE3(torch.cat((-_quat_rotate(qi, self.data[..., :3]), qi), dim=-1))

    def mul(self, other):
        a, b = torch.broadcast_tensors(self.data, other.data)
        t = a[..., :3] + _quat_rotate(a[..., 3:], b[..., :3])
        return SE3(torch.cat((t, _quat_mul(a[..., 3:], b[..., 3:])), dim=-1))

    def act(self, p):
        d = self.data
        while d.dim() < p.dim():          # SE3.Random(20) * (20, N, 3)  (unit_test_pinhole_transforms.py:26-27)
            d = d.unsqueeze(-2)
        q, t = d[..., 3:], d[..., :3]
        if p.shape[-1] == 3:
            return _quat_rotate(q, p) + t
        xyz = _quat_rotate(q, p[..., :3]) + t * p[..., 3:4]
        return torch.cat((xyz, p[..., 3:4].expand(*xyz.shape[:-1], 1)), dim=-1)

    def retr(self, a):
        return SE3.exp(a).mul(self)

    def matrix(self):
        eye = torch.eye(4, dtype=self.dtype, device=self.device)
        cols = SE3(self.data.unsqueeze(-2)).act(eye.expand(*self.shape, 4, 4))   # rows = T e_i
        return cols.transpose(-1, -2)

    def scale(self, s):
        s = torch.as_tensor(s, dtype=self.dtype, device=self.device)
        return SE3(torch.cat((self.data[..., :3] * s[..., None] if s.dim() else self.data[..., :3] * s,
                              self.data[..., 3:]), dim=-1))

    def __mul__(self, other):
        if isinstance(other, LieGroupParameter):
            other = other.retr()
        return self.mul(other) if isinstance(other, SE3) else self.act(other)

    # ---- tensor-like plumbing -----------------------------------------------------------
    def __getitem__(self, index):
        return SE3(self.data[index])

    def __len__(self):
        return self.data.shape[0]

    def view(self, dims):
        return SE3(self.data.view(*dims, 7))

    def squeeze(self, *dim):
        return SE3(self.data.squeeze(*dim))

    def detach(self):
        return SE3(self.data.detach())

    def clone(self):
        return SE3(self.data.clone())

    def to(self, *args, **kwargs):
        return SE3(self.data.to(*args, **kwargs))

    def cpu(self):
        return SE3(self.data.cpu())

    def cuda(self):
        return SE3(self.data.cuda())

    def float(self, device=None):        # lietorch takes (and ignores) a positional: pose_estimator.py:42
        return SE3(self.data.float())

    def double(self, device=None):
        return SE3(self.data.double())

    def __repr__(self):
        return f"SE3(stand-in): shape={tuple(self.shape)}, dtype={self.dtype}"


class LieGroupParameter(torch.Tensor):
    """Zero tangent leaf + group element; see module docstring."""
    __torch_function__ = torch._C._disabled_torch_function_impl

    def __new__(cls, group, requires_grad=True):
        data = torch.zeros(group.tangent_shape, device=group.device, dtype=group.dtype)
        return torch.Tensor._make_subclass(cls, data, requires_grad)

    def __init__(self, group, requires_grad=True):
        self.group = group

    def retr(self):
        return self.group.retr(self)

    def log(self):
        return self.retr().log()

    def inv(self):
        return self.retr().inv()

    def __mul__(self, other):
        if isinstance(other, LieGroupParameter):
            other = other.retr()
        return self.retr() * other

    def add_(self, update, alpha=1):
        with torch.no_grad():
            self.group = SE3.exp(alpha * _plain(update)).mul(self.group)
        return self

    def __getitem__(self, index):
        return self.retr()[index]
