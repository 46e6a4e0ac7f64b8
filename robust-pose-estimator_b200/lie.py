"""Minimal SE(3) value type replacing the third-party ``lietorch.SE3`` on the f2f pose path
(reference call sites: core/pose/pose_net.py:96-100, core/pose/pose_estimator.py:42,81,84,90,91,
core/utils/trajectory.py:21).  The heavy Lie-group work (exp-map retraction, log, act on 327k
points) happens inside the CUDA solver (csrc/pose.cu); this class only carries the 7-vector and
composes a handful of poses per frame with torch ops on whatever device the data lives on.

data[..., 7] = [tx ty tz qx qy qz qw]; tangent = [tau | phi]; same conventions as lietorch.
"""
import torch

__all__ = ["SE3"]


def _cross(a, b):
    a, b = torch.broadcast_tensors(a, b)
    return torch.linalg.cross(a, b, dim=-1)


def _qrot(q, p):
    qv, qw = q[..., :3], q[..., 3:4]
    uv = 2.0 * _cross(qv, p)
    return p + qw * uv + _cross(qv, uv)


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz), -1)


class SE3:
    manifold_dim, embedded_dim = 6, 7

    def __init__(self, data):
        self.data = data.data if isinstance(data, SE3) else data

    # -- constructors
    @classmethod
    def Identity(cls, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        kw.pop("requires_grad", None)
        d = torch.zeros(*shape, 7, **kw)
        d[..., 6] = 1.0
        return cls(d)

    @classmethod
    def IdentityLike(cls, G):
        return cls.Identity(G.shape, device=G.device, dtype=G.dtype)

    @classmethod
    def InitFromVec(cls, v):
        return cls(v)

    @classmethod
    def exp(cls, xi):
        tau, phi = xi[..., :3], xi[..., 3:]
        th2 = (phi * phi).sum(-1, keepdim=True)
        small = th2 < 1e-12
        t2 = torch.where(small, torch.ones_like(th2), th2)
        th = t2.sqrt()
        imag = torch.where(small, 0.5 - th2 / 48.0, torch.sin(0.5 * th) / th)
        real = torch.where(small, 1.0 - th2 / 8.0, torch.cos(0.5 * th))
        c1 = torch.where(small, 0.5 - th2 / 24.0, (1.0 - torch.cos(th)) / t2)
        c2 = torch.where(small, 1.0 / 6.0 - th2 / 120.0, (th - torch.sin(th)) / (t2 * th))
        pxt = _cross(phi, tau)
        t = tau + c1 * pxt + c2 * _cross(phi, pxt)
        return cls(torch.cat((t, imag * phi, real), -1))

    @classmethod
    def Random(cls, *shape, sigma=1.0, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return cls.exp(sigma * torch.randn(*shape, 6, **kw))

    # -- properties
    shape = property(lambda s: s.data.shape[:-1])
    device = property(lambda s: s.data.device)
    dtype = property(lambda s: s.data.dtype)
    tangent_shape = property(lambda s: s.data.shape[:-1] + (6,))

    # -- group operations
    def vec(self):
        return self.data

    def log(self):
        t, qv, qw = self.data[..., :3], self.data[..., 3:6], self.data[..., 6:7]
        n2 = (qv * qv).sum(-1, keepdim=True)
        small = n2 < 1e-12
        n = torch.where(small, torch.ones_like(n2), n2).sqrt()
        s = torch.where(small, 2.0 / qw - (2.0 / 3.0) * n2 / (qw * qw * qw), 2.0 * torch.atan(n / qw) / n)
        phi = s * qv
        th2 = (phi * phi).sum(-1, keepdim=True)
        sm = th2 < 1e-12
        t2 = torch.where(sm, torch.ones_like(th2), th2)
        th = t2.sqrt()
        c2 = torch.where(sm, torch.full_like(th2, 1.0 / 12.0),
                         (1.0 - th * torch.cos(0.5 * th) / (2.0 * torch.sin(0.5 * th))) / t2)
        pxt = _cross(phi, t)
        return torch.cat((t - 0.5 * pxt + c2 * _cross(phi, pxt), phi), -1)

    def inv(self):
        q = self.data[..., 3:]
        qi = torch.cat((-q[..., :3], q[..., 3:]), -1)
        return SE3(torch.cat((-_qrot(qi, self.data[..., :3]), qi), -1))

    def mul(self, other):
        a, b = torch.broadcast_tensors(self.data, other.data)
        return SE3(torch.cat((a[..., :3] + _qrot(a[..., 3:], b[..., :3]), _qmul(a[..., 3:], b[..., 3:])), -1))

    def act(self, p):
        d = self.data
        while d.dim() < p.dim():
            d = d.unsqueeze(-2)
        if p.shape[-1] == 3:
            return _qrot(d[..., 3:], p) + d[..., :3]
        xyz = _qrot(d[..., 3:], p[..., :3]) + d[..., :3] * p[..., 3:4]
        return torch.cat((xyz, p[..., 3:4].expand(*xyz.shape[:-1], 1)), -1)

    def matrix(self):
        eye = torch.eye(4, dtype=self.dtype, device=self.device).expand(*self.shape, 4, 4)
        return SE3(self.data.unsqueeze(-2)).act(eye).transpose(-1, -2)

    def scale(self, s):
        s = torch.as_tensor(s, dtype=self.dtype, device=self.device)
        return SE3(torch.cat((self.data[..., :3] * s, self.data[..., 3:]), -1))

    def __mul__(self, other):
        return self.mul(other) if isinstance(other, SE3) else self.act(other)

    # -- tensor-like plumbing
    def __getitem__(self, i):
        return SE3(self.data[i])

    def __len__(self):
        return self.data.shape[0]

    def view(self, dims):
        return SE3(self.data.view(*dims, 7))

    def squeeze(self, *d):
        return SE3(self.data.squeeze(*d))

    def detach(self):
        return SE3(self.data.detach())

    def clone(self):
        return SE3(self.data.clone())

    def to(self, *a, **k):
        return SE3(self.data.to(*a, **k))

    def cpu(self):
        return SE3(self.data.cpu())

    def cuda(self):
        return SE3(self.data.cuda())

    def float(self, device=None):      # lietorch's float()/double() take an ignored positional
        return SE3(self.data.float())

    def double(self, device=None):
        return SE3(self.data.double())

    def __repr__(self):
        return f"SE3(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device})"
