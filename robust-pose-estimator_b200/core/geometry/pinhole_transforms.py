"""Pinhole geometry with the reference's function names
(/root/reference/core/geometry/pinhole_transforms.py).  ``reproject`` runs the sm_100a back-projection
kernel; ``project`` / ``transform`` are small torch helpers kept for API compatibility (on the pose path
they are fused into the CUDA residual kernel, csrc/pose.cu)."""
import torch

from ... import ops
from ...lie import SE3


def create_img_coords_t(y, x, b=1, device=torch.device("cpu")):
    """(3, y*x) pixel-centre coordinates [u+.5, v+.5, 1], row-major (pinhole_transforms.py:7-19)."""
    xs = torch.arange(x, device=device, dtype=torch.float32) + 0.5
    ys = torch.arange(y, device=device, dtype=torch.float32) + 0.5
    vv, uu = torch.meshgrid(ys, xs, indexing="ij")
    one = torch.ones(b * y * x, device=device)
    return torch.vstack((uu.repeat(b, 1, 1).flatten(), vv.repeat(b, 1, 1).flatten(), one))


def homogeneous(opts):
    n = opts.shape[0]
    return torch.cat((opts, torch.ones((n, 1, opts.shape[-1]), device=opts.device, dtype=opts.dtype)), dim=1)


def transform(opts, T, double_backward=False):
    """(n,3,N) points, SE3 of shape (n,1) or (n,) -> T * points."""
    return (T * opts.permute(0, 2, 1)).permute(0, 2, 1)


def reproject(depth, intrinsics, img_coords=None):
    """depth (n,1,H,W), intrinsics (n,3,3) or (3,3) -> homogeneous points (n,4,H*W) (pinhole_transforms.py:79-87)."""
    n = depth.shape[0]
    K = intrinsics if intrinsics.dim() == 3 else intrinsics[None].repeat(n, 1, 1)
    pcl = ops.proj(depth.float().contiguous(), K.float().contiguous())
    return homogeneous(pcl.view(n, 3, -1))


def project(opts, intrinsics, T=None, double_backward=False):
    """(n,3,N) points -> (n,3,N) image points [u, v, 1] with depth clamped at 1e-12 (pinhole_transforms.py:90-99)."""
    if T is not None:
        opts = transform(opts, T)
    ipts = torch.bmm(intrinsics, opts)
    depth = torch.clamp(ipts[:, -1], 1e-12, None).unsqueeze(1)
    ipts = torch.cat((ipts[:, :2], torch.ones_like(ipts[:, None, 2])), dim=1)
    return ipts / depth


__all__ = ["create_img_coords_t", "homogeneous", "transform", "reproject", "project", "SE3"]
