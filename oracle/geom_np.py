"""ORACLE / TEST INFRASTRUCTURE ONLY.  numpy float32 restatement (same op order as the reference's
CPU fp32 run -- the parity oracle) of the per-pixel geometry and warping stages:

  create_img_coords_t      /root/reference/core/geometry/pinhole_transforms.py:7-19
  PoseNet.proj / reproject /root/reference/core/pose/pose_net.py:121-125, pinhole_transforms.py:79-87
  depth from stereo flow   /root/reference/core/pose/pose_net.py:73-77, 127-135
  remap_from_flow          /root/reference/core/interpol/flow_utils.py:4-14
  remap_from_flow_nearest  /root/reference/core/interpol/flow_utils.py:17-26 (+ pose_net.py:107-108)
  1/8 bilinear downsample  /root/reference/core/pose/pose_net.py:110-113 (F.interpolate, scale 0.125)
  ATen grid_sampler_2d (CPU, vectorised): unnormalize = (g + 1) * ((size-1)/2) for align_corners=True,
      bilinear corner weights from x - floor(x), zeros padding, nearest = round-half-even.

  mask_specularities       /root/reference/dataset/stereo_dataset.py:12-16 (cv2.erode with a box kernel, default border:
      out-of-image pixels do not constrain the minimum)

Pinned by tests/golden/stages_small.npz and e2e_384x352.npz (outputs of the reference itself); mask_specularities by
tests/golden/mask_specularities.npz (output of the reference function, oracle/make_golden.py --mask-spec).
"""
import numpy as np

F32 = np.float32


def img_coords(H, W):
    v, u = np.meshgrid(np.arange(H, dtype=F32) + F32(0.5), np.arange(W, dtype=F32) + F32(0.5), indexing="ij")
    return np.stack((u.reshape(-1), v.reshape(-1), np.ones(H * W, dtype=F32)))


def depth_from_stereo_flow(sflow, bf):
    """sflow (2,H,W) f32, bf scalar f32 -> depth (H,W) f32, valid (H,W) bool.  pose_net.py:73-75."""
    with np.errstate(divide="ignore", invalid="ignore"):
        depth = F32(bf) / -np.asarray(sflow[0], dtype=F32)
    valid = (depth > 0) & (depth <= F32(1.0))
    depth = np.where(valid, depth, F32(1.0)).astype(F32)
    return depth, valid


def proj(depth, K):
    """depth (H,W) f32, K (3,3) f32 -> pcl (3,H,W) f32 = depth * (K^-1 @ [u+.5, v+.5, 1])."""
    H, W = depth.shape
    Kinv = np.linalg.inv(np.asarray(K, dtype=F32)).astype(F32)
    rays = (Kinv @ img_coords(H, W)).astype(F32)
    return (depth.reshape(1, -1) * rays).astype(F32).reshape(3, H, W)


def flow_grid(flow):
    """Normalised sampling grid of remap_from_flow (flow_utils.py:7-10), float32 op by op."""
    H, W = flow.shape[-2:]
    row, col = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    gx = (F32(2) * (flow[0].astype(F32) + col.astype(F32))) / F32(W - 1) - F32(1)
    gy = (F32(2) * (flow[1].astype(F32) + row.astype(F32))) / F32(H - 1) - F32(1)
    return gx.astype(F32), gy.astype(F32)


def unnormalize(g, size):
    return ((g + F32(1)) * F32((size - 1) / 2)).astype(F32)


def grid_sample_bilinear(img, gx, gy):
    """img (C,H,W) f32; gx, gy arbitrary-shape normalised coords -> (C, *gx.shape)."""
    C, H, W = img.shape
    x = unnormalize(gx, W)
    y = unnormalize(gy, H)
    with np.errstate(invalid="ignore"):
        x0 = np.floor(x)
        y0 = np.floor(y)
    wx = (x - x0).astype(F32)
    ex = (F32(1) - wx).astype(F32)
    wy = (y - y0).astype(F32)
    ey = (F32(1) - wy).astype(F32)
    out = np.zeros((C,) + gx.shape, dtype=F32)
    for dy, dx, wgt in ((0, 0, ey * ex), (0, 1, ey * wx), (1, 0, wy * ex), (1, 1, wy * wx)):
        xi = x0 + dx
        yi = y0 + dy
        with np.errstate(invalid="ignore"):
            ok = (xi > -1) & (xi < W) & (yi > -1) & (yi < H)
        xi_c = np.where(ok, xi, 0).astype(np.int64)
        yi_c = np.where(ok, yi, 0).astype(np.int64)
        val = img[:, yi_c, xi_c]
        out += np.where(ok, val * wgt.astype(F32), F32(0)).astype(F32)
    return out


def grid_sample_nearest(img, gx, gy):
    C, H, W = img.shape
    x = np.rint(unnormalize(gx, W))                                # round-half-even (nearbyint)
    y = np.rint(unnormalize(gy, H))
    with np.errstate(invalid="ignore"):
        ok = (x > -1) & (x < W) & (y > -1) & (y < H)
    xi = np.where(ok, x, 0).astype(np.int64)
    yi = np.where(ok, y, 0).astype(np.int64)
    return np.where(ok, img[:, yi, xi], F32(0)).astype(img.dtype)


def remap_from_flow(x, flow):
    gx, gy = flow_grid(flow)
    return grid_sample_bilinear(np.asarray(x, dtype=F32), gx, gy)


def remap_mask_nearest(mask, flow):
    """mask (H,W) bool -> mask2w = valid_mapping & warped  (pose_net.py:107-108)."""
    gx, gy = flow_grid(flow)
    warped = grid_sample_nearest(np.asarray(mask, dtype=F32)[None], gx, gy)[0]
    return (warped > 0) & warped.astype(bool)


def downsample8(x):
    """F.interpolate(x, scale_factor=0.125, mode='bilinear') == mean of pixels (8i+3, 8i+4)^2."""
    x = np.asarray(x, dtype=F32)
    a = x[..., 3::8, :]
    b = x[..., 4::8, :]
    r = (F32(0.5) * a + F32(0.5) * b).astype(F32)
    return (F32(0.5) * r[..., 3::8] + F32(0.5) * r[..., 4::8]).astype(F32)


def mask_specularities(img_hwc, mask=None, spec_thr=0.96, radius=5):
    """img (H,W,3) uint8, mask (H,W) bool or None -> (H,W) uint8 0/1.  stereo_dataset.py:12-16."""
    img_hwc = np.asarray(img_hwc)
    spec = img_hwc.sum(axis=-1) < (3 * 255 * spec_thr)
    m = (np.asarray(mask, dtype=bool) & spec) if mask is not None else spec
    H, W = m.shape
    pad = np.ones((H + 2 * radius, W + 2 * radius), dtype=bool)           # cv2.erode default border: never the minimum
    pad[radius:radius + H, radius:radius + W] = m
    out = np.ones((H, W), dtype=bool)
    for dy in range(2 * radius + 1):
        for dx in range(2 * radius + 1):
            out &= pad[dy:dy + H, dx:dx + W]
    return out.astype(np.uint8)
