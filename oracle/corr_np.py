"""ORACLE / TEST INFRASTRUCTURE ONLY.  numpy restatement of RAFT's all-pairs correlation block:

  CorrBlock.corr       /root/reference/core/RAFT/core/corr.py:52-60   fmap1^T fmap2 / sqrt(C)
  CorrBlock.__init__   /root/reference/core/RAFT/core/corr.py:13-27   3x avg_pool2d(2, 2) pyramid
  CorrBlock.__call__   /root/reference/core/RAFT/core/corr.py:29-50   radius-r bilinear lookup
  bilinear_sampler     /root/reference/core/RAFT/core/utils/utils.py:57-71

Output channel order: level-major, then i*(2r+1)+j where i indexes the X offset.
Pinned by tests/golden/stages_small.npz (reference CorrBlock outputs on exact inputs).
"""
import numpy as np

from .geom_np import grid_sample_bilinear

F32 = np.float32


def corr_volume(f1, f2):
    """f1, f2 (B,C,h,w) f32 -> (B, h*w, h, w) f32."""
    B, C, h, w = f1.shape
    a = f1.reshape(B, C, h * w).astype(F32)
    b = f2.reshape(B, C, h * w).astype(F32)
    v = np.einsum("bcq,bct->bqt", a, b).astype(F32) / np.sqrt(F32(C))
    return v.astype(F32).reshape(B, h * w, h, w)


def avg_pool2(v):
    H, W = v.shape[-2] // 2 * 2, v.shape[-1] // 2 * 2
    v = v[..., :H, :W]
    return ((v[..., 0::2, 0::2] + v[..., 0::2, 1::2] + v[..., 1::2, 0::2] + v[..., 1::2, 1::2]) * F32(0.25)).astype(F32)


def pyramid(f1, f2, num_levels=4):
    lv = [corr_volume(f1, f2)]
    for _ in range(num_levels - 1):
        lv.append(avg_pool2(lv[-1]))
    return lv


def lookup(pyr, coords, radius=4):
    """pyr: list of (B, Q, h_l, w_l); coords (B,2,h,w) (ch0 = x) -> (B, L*(2r+1)^2, h, w) f32."""
    B, _, h, w = coords.shape
    r = radius
    n = 2 * r + 1
    d = np.arange(-r, r + 1, dtype=F32)
    out = np.zeros((B, len(pyr) * n * n, h, w), dtype=F32)
    cx = coords[:, 0].reshape(B, h * w).astype(F32)
    cy = coords[:, 1].reshape(B, h * w).astype(F32)
    for l, vol in enumerate(pyr):
        hl, wl = vol.shape[-2:]
        x = (cx / F32(2 ** l)).astype(F32)[..., None, None] + d[None, None, :, None]     # slow index i -> x offset
        y = (cy / F32(2 ** l)).astype(F32)[..., None, None] + d[None, None, None, :]     # fast index j -> y offset
        x, y = np.broadcast_arrays(x, y)
        gx = (F32(2) * x / F32(wl - 1) - F32(1)).astype(F32)
        gy = (F32(2) * y / F32(hl - 1) - F32(1)).astype(F32)
        for b in range(B):
            for q in range(h * w):
                s = grid_sample_bilinear(vol[b, q][None], gx[b, q], gy[b, q])[0]          # (n, n)
                out[b, l * n * n:(l + 1) * n * n, q // w, q % w] = s.reshape(-1)
    return out
