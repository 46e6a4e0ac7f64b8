"""ResizeStereo (input pipeline): the numpy oracle (CPU) and the CUDA kernel (GPU) against the reference's own
class (tests/golden/resize_stereo.npz from /root/reference/dataset/transforms.py:20-39)."""
import os

import numpy as np
import pytest

from oracle import resize_np
from oracle.detrand import det_uniform


def _inputs(g, k):
    Hi, Wi = [int(v) for v in g[f"in_shape{k}"]]
    left = np.floor(det_uniform((3, Hi, Wi), 500 + k, 0.0, 255.0)).astype(np.float32)
    right = np.floor(det_uniform((3, Hi, Wi), 520 + k, 0.0, 255.0)).astype(np.float32)
    mask = (det_uniform((1, Hi, Wi), 540 + k, 0.0, 1.0) > 0.3).astype(np.uint8)
    return left, right, mask, tuple(int(v) for v in g[f"size{k}"])


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_oracle_resize_matches_reference(golden_dir, k):
    g = np.load(os.path.join(golden_dir, "resize_stereo.npz"))
    left, right, mask, size = _inputs(g, k)
    np.testing.assert_allclose(resize_np.resize_stereo(left, size), g[f"left{k}"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(resize_np.resize_stereo(right, size), g[f"right{k}"], rtol=0, atol=2e-4)
    assert np.array_equal(resize_np.resize_stereo(mask, size, nearest=True), g[f"mask{k}"])


@pytest.mark.gpu
@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_gpu_resize_stereo_matches_reference(golden_dir, k):
    import torch
    import rpe_b200  # noqa: F401
    from rpe_b200.dataset.transforms import ResizeStereo
    g = np.load(os.path.join(golden_dir, "resize_stereo.npz"))
    left, right, mask, size = _inputs(g, k)
    tr = ResizeStereo(size)
    # uint8 frames (the values are integers): the conversion to float is folded into the kernel
    l, r, m = tr(torch.from_numpy(left.astype(np.uint8)).cuda(), torch.from_numpy(right).cuda(), torch.from_numpy(mask).cuda())
    assert l.dtype == torch.float32 and m.dtype == torch.uint8
    np.testing.assert_allclose(l.cpu().numpy(), g[f"left{k}"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(r.cpu().numpy(), g[f"right{k}"], rtol=0, atol=2e-4)
    assert np.array_equal(m.cpu().numpy(), g[f"mask{k}"])
    # batched (n,3,H,W) input and a bool mask
    lb, _, mb = tr(torch.from_numpy(np.stack((left, right))).cuda(), torch.from_numpy(np.stack((left, right))).cuda(),
                   torch.from_numpy(np.stack((mask, mask)).astype(bool)).cuda())
    np.testing.assert_allclose(lb[1].cpu().numpy(), g[f"right{k}"], rtol=0, atol=2e-4)
    assert mb.dtype == torch.bool and np.array_equal(mb[0].cpu().numpy(), g[f"mask{k}"].astype(bool))
