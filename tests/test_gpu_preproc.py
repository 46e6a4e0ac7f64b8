"""GPU parity of the input-pipeline row: rpe_mask_specularities against the reference's own dataset function
(golden fixture written from /root/reference/dataset/stereo_dataset.py:12-16) and against the numpy oracle.  Byte work: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import geom_np
from oracle.detrand import det_uniform

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    import rpe_b200  # noqa: F401
    from rpe_b200 import ops as _ops
    return _ops


def _dev_img(img_hwc):
    return torch.from_numpy(np.ascontiguousarray(img_hwc.transpose(2, 0, 1)))[None].cuda()


def test_mask_specularities_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "mask_specularities.npz"))
    for k in range(3):
        H, W = [int(v) for v in g[f"shape{k}"]]
        mask = np.unpackbits(g[f"mask{k}"])[:H * W].reshape(H, W).astype(bool)
        out = ops.mask_specularities(_dev_img(g[f"img{k}"]), torch.from_numpy(mask)[None, None].cuda())
        assert out.dtype == torch.bool and out.shape == (1, 1, H, W)
        assert np.array_equal(np.packbits(out.cpu().numpy().reshape(-1)), g[f"out{k}"])            # bit-exact vs the reference
        out = ops.mask_specularities(_dev_img(g[f"img{k}"]))
        assert np.array_equal(np.packbits(out.cpu().numpy().reshape(-1)), g[f"out_nomask{k}"])


@pytest.mark.parametrize("H,W,radius,thr", [(512, 640, 5, 0.96), (1024, 1280, 5, 0.96), (37, 301, 3, 0.9), (16, 64, 0, 0.5), (5, 7, 8, 0.96)])
def test_mask_specularities_vs_oracle(ops, H, W, radius, thr):
    """Bench-size frames (batch 3) and edge cases (radius 0, frames smaller than the window) against the numpy oracle."""
    n = 3
    imgs = (det_uniform((n, H, W, 3), 71, 0.0, 1.0) ** 0.25 * 255.999).astype(np.uint8)            # skewed towards bright pixels
    masks = det_uniform((n, H, W), 72, 0.0, 1.0) > 0.001
    dimg = torch.from_numpy(np.ascontiguousarray(imgs.transpose(0, 3, 1, 2))).cuda()
    dmask = torch.from_numpy(masks[:, None]).cuda()
    out = ops.mask_specularities(dimg, dmask, spec_thr=thr, radius=radius).cpu().numpy()
    for i in range(n):
        ref = geom_np.mask_specularities(imgs[i], masks[i], spec_thr=thr, radius=radius).astype(bool)
        assert np.array_equal(out[i, 0], ref), f"sample {i}: {(out[i, 0] != ref).sum()} pixels differ"
    assert np.array_equal(dmask.cpu().numpy()[:, 0], masks)                                        # the input mask is not modified
    if H >= 512:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        big = dimg.repeat(22, 1, 1, 1)                                                            # 66 frames, like one engine chunk
        bm = dmask.repeat(22, 1, 1, 1)
        for _ in range(3):
            ops.mask_specularities(big, bm, spec_thr=thr, radius=radius)
        e0.record()
        for _ in range(10):
            ops.mask_specularities(big, bm, spec_thr=thr, radius=radius)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        print(f"mask_specularities 66 x {H}x{W}: {us:.1f} us per launch, {66 * H * W * 5 / us / 1e3:.0f} GB/s algorithmic (5 B/px)")


def test_mask_specularities_rejects_bad_arguments(ops):
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib
    img = torch.zeros((1, 3, 8, 8), dtype=torch.uint8, device="cuda")
    m = torch.ones((1, 1, 8, 8), dtype=torch.bool, device="cuda")
    with pytest.raises(Exception):
        ops.mask_specularities(img, m, radius=9)                                                   # window larger than the kernel's halo
    l = _lib.lib()
    assert l.rpe_mask_specularities(img.data_ptr(), m.data_ptr(), m.data_ptr(), 1, 8, 8, 734, 5, None) == -1      # in place
    with pytest.raises(Exception):
        ops.mask_specularities(img.cpu(), None)                                                    # no CPU path
