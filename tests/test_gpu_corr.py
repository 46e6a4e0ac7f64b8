"""GPU parity of stage 1 (tcgen05/TMA correlation GEMM, pyramid pooling, window lookup) against the
numpy oracle / golden outputs of the reference's CorrBlock."""
import os

import numpy as np
import pytest
import torch

from oracle import corr_np
from oracle.detrand import det_uniform

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    import rpe_b200  # noqa: F401
    from rpe_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _tf32_round(x):
    xi = x.contiguous().view(torch.int32)
    return ((xi + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("precision,tol", [("x3", 2e-6), ("x1", 2e-3)])
def test_corr_pyramid_golden(ops, golden_dir, precision, tol):
    g = np.load(os.path.join(golden_dir, "stages_small.npz"))
    f1, f2 = det_uniform((2, 32, 16, 16), 11), det_uniform((2, 32, 16, 16), 12)
    cp = ops.CorrPyramid(dev(f1), dev(f2), precision=ops.CORR_TF32X3 if precision == "x3" else ops.CORR_TF32)
    l0 = cp.level(0).cpu().numpy().reshape(2, 256, 16, 16)
    np.testing.assert_allclose(l0[:1], g["corr_l0"], atol=tol)
    for l in (1, 2, 3):
        got = cp.level(l).cpu().numpy().reshape(2, 256, 16 >> l, 16 >> l)
        np.testing.assert_allclose(got, g[f"corr_l{l}"], atol=tol)
    out = cp(dev(g["corr_coords"])).cpu().numpy()
    np.testing.assert_allclose(out, g["corr_lookup"], atol=4 * tol)


def test_lookup_exact_on_reference_pyramid(ops, golden_dir):
    """Lookup only: overwrite the device pyramid with the oracle's fp32 pyramid (itself pinned to the
    reference) so the comparison isolates the gather/interpolation kernel."""
    g = np.load(os.path.join(golden_dir, "stages_small.npz"))
    f1, f2 = det_uniform((2, 32, 16, 16), 11), det_uniform((2, 32, 16, 16), 12)
    cp = ops.CorrPyramid(dev(f1), dev(f2))
    pyr = corr_np.pyramid(f1, f2)
    for l in range(4):
        cp.level(l).copy_(dev(pyr[l].reshape(cp.level(l).shape)))
    out = cp(dev(g["corr_coords"])).cpu().numpy()
    np.testing.assert_allclose(out, g["corr_lookup"], atol=3e-6)


@pytest.mark.parametrize("h,w,B", [(64, 80, 2), (44, 48, 1), (36, 44, 3), (48, 40, 2), (20, 22, 1)])
def test_corr_gemm_full_size_vs_fp64(ops, h, w, B):
    """Bench-size volume (and sizes that are not multiples of the 128x256 tile / of the 16x16 target block, an odd number of
    query tiles per image = no CTA pairs, a width that rules out vector stores) against an fp64 matmul."""
    C = 256
    f1 = dev(det_uniform((B, C, h, w), 71, -1.5, 1.5))
    f2 = dev(det_uniform((B, C, h, w), 72, -1.5, 1.5))
    Q = h * w
    ref64 = torch.matmul(f1.double().view(B, C, Q).transpose(1, 2), f2.double().view(B, C, Q)) / 16.0
    cp3 = ops.CorrPyramid(f1, f2, precision=ops.CORR_TF32X3)
    got3 = cp3.level(0).view(B, Q, Q).double()
    assert (got3 - ref64).abs().max().item() < 2e-5                         # fp32-class accuracy
    # pooled levels equal the pooled level 0 (avg_pool2d restated with torch ops, test-only)
    import torch.nn.functional as F
    lvl = cp3.level(0)
    for l in (1, 2, 3):
        lvl = F.avg_pool2d(lvl, 2, stride=2)
        assert (cp3.level(l) - lvl).abs().max().item() < 1e-6
    # fp16x3 split (kind::f16 MMAs): hi*hi + lo*hi + hi*lo of the fp16 hi / bf16 lo planes == the exact product minus the lo*lo terms (~2^-22 relative)
    cpb = ops.CorrPyramid(f1, f2, precision=ops.CORR_F16X3)
    gotb = cpb.level(0).view(B, Q, Q).double()
    h1, h2 = f1.to(torch.float16).float(), f2.to(torch.float16).float()              # hi planes fp16, lo planes bf16
    l1, l2 = (f1 - h1).to(torch.bfloat16).float(), (f2 - h2).to(torch.bfloat16).float()
    mm = lambda a, b: torch.matmul(a.double().view(B, C, Q).transpose(1, 2), b.double().view(B, C, Q))
    refb = (mm(h1, h2) + mm(l1, h2) + mm(h1, l2)) / 16.0
    errb, errb64 = (gotb - refb).abs().max().item(), (gotb - ref64).abs().max().item()
    print(f"corr fp16x3 {B}x{h}x{w}: vs split products {errb:.2e}, vs fp64 {errb64:.2e}")
    assert errb < 2e-5 and errb64 < 2e-5
    # the fused epilogue's pooled levels == avg_pool2d of ITS level 0, level by level (floor on odd sizes)
    lvl = cpb.level(0)
    for l in (1, 2, 3):
        assert cpb.level(l).shape == cp3.level(l).shape
        lvl = F.avg_pool2d(lvl, 2, stride=2)
        assert (cpb.level(l) - lvl).abs().max().item() < 1e-6, f"fp16x3 pyramid level {l}"
    cp1 = ops.CorrPyramid(f1, f2, precision=ops.CORR_TF32)
    got1 = cp1.level(0).view(B, Q, Q).double()
    # single pass == exact product of tf32-rounded inputs
    r1, r2 = _tf32_round(f1), _tf32_round(f2)
    ref1 = torch.matmul(r1.double().view(B, C, Q).transpose(1, 2), r2.double().view(B, C, Q)) / 16.0
    assert (got1 - ref1).abs().max().item() < 2e-5
    assert (got1 - ref64).abs().max().item() < 2e-2


def test_lookup_full_size_vs_oracle_sample(ops):
    """64x80 grid: compare a sample of queries against the (slow) numpy oracle, and check linearity."""
    B, C, h, w = 2, 256, 64, 80
    f1 = dev(det_uniform((B, C, h, w), 81))
    f2 = dev(det_uniform((B, C, h, w), 82))
    cp = ops.CorrPyramid(f1, f2, precision=ops.CORR_TF32X3)
    base = np.stack(np.meshgrid(np.arange(w), np.arange(h), indexing="xy"), 0).astype(np.float32)
    coords = base[None].repeat(B, 0) + det_uniform((B, 2, h, w), 83, -20, 20)
    out = cp(dev(coords)).cpu().numpy()
    pyr = [cp.level(l).cpu().numpy().reshape(B, h * w, h >> l, w >> l) for l in range(4)]
    for (b, y, x) in [(0, 0, 0), (0, 13, 77), (1, 63, 79), (1, 31, 40), (0, 5, 5)]:
        q = y * w + x
        sub = [p[b:b + 1, q:q + 1] for p in pyr]
        ref = corr_np.lookup(sub, coords[b:b + 1, :, y:y + 1, x:x + 1])
        np.testing.assert_allclose(out[b, :, y, x], ref[0, :, 0, 0], atol=1e-5)   # shared sub-pixel fraction per window
    # linearity in the volume: lookup(2 * pyramid) == 2 * lookup(pyramid)
    cp.pyramid.mul_(2.0)
    out2 = cp(dev(coords)).cpu().numpy()
    np.testing.assert_allclose(out2, 2.0 * out, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("h,w,B", [(64, 80, 3), (16, 24, 2), (46, 62, 1)])
def test_lookup_planes_match_nchw_lookup(ops, h, w, B):
    """The warp-per-query NHWC bf16 hi/lo lookup (tensor-core path) reproduces the NCHW fp32 lookup to the 16-bit split."""
    import rpe_b200  # noqa: F401
    from rpe_b200 import tc
    f1 = torch.from_numpy(det_uniform((B, 64, h, w), 31)).cuda()
    f2 = torch.from_numpy(det_uniform((B, 64, h, w), 32)).cuda()
    cp = ops.CorrPyramid(f1, f2, precision=ops.CORR_TF32)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    coords = np.stack((xs, ys), 0)[None].astype(np.float32) + det_uniform((B, 2, h, w), 33, -9, 9)
    coords[:, :, 0, 0] = -50.0                                  # fully outside windows read zeros
    coords = torch.from_numpy(np.ascontiguousarray(coords.astype(np.float32))).cuda()
    ref = cp(coords)                                            # (B,324,h,w)
    planes = tc.Planes(B, h, w, 384, f1.device)
    ops.corr_lookup_planes(cp, coords, planes)
    got = planes.float()[..., :324].permute(0, 3, 1, 2)
    assert float(planes.float()[..., 324:].abs().max()) == 0.0
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"lookup planes vs NCHW lookup: max abs err {err:.2e} (|ref| max {scale:.2f})")
    assert err <= scale * 2.0 ** -19
