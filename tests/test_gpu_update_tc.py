"""GPU: the tcgen05 implicit-GEMM convolution (fp16x3) and the tensor-core RAFT update operator against torch fp32
convolutions (test-only reference of a floating-point kernel) and against the reference tracker's golden outputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.detrand import det_uniform, unpack

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")


@pytest.fixture(scope="module")
def ops(cuda):
    import rpe_b200  # noqa: F401
    from rpe_b200 import ops as _ops
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("cin,cout,kh,kw,H,W,act", [
    (64, 64, 1, 1, 8, 16, "none"),          # exactly one tile, one K block
    (128, 256, 3, 3, 64, 80, "relu"),       # flow_head.conv1 shape
    (384, 256, 1, 5, 64, 80, "sigmoid"),    # GRU gates, horizontal
    (384, 128, 5, 1, 64, 80, "tanh"),       # GRU candidate, vertical
    (256, 2, 3, 3, 64, 80, "none"),         # flow_head.conv2 (N tile of 16)
    (256, 126, 3, 3, 44, 48, "relu"),       # odd channel count, image not a multiple of the 8x16 tile
    (324, 256, 1, 1, 36, 44, "relu"),       # convc1: K padded 324 -> 384
    (256, 576, 1, 1, 64, 80, "none"),       # mask head: three N blocks of 192
    (98, 128, 1, 1, 64, 80, "relu"),        # im2col'ed 7x7 flow convolution
])
def test_conv_fp16x3_vs_torch_fp32(ops, cin, cout, kh, kw, H, W, act):
    n = 2
    x = dev(det_uniform((n, cin, H, W), 101, -2.0, 2.0))
    w = dev(det_uniform((cout, cin, kh, kw), 102, -1.0, 1.0)) * (1.0 / np.sqrt(cin * kh * kw))
    b = dev(det_uniform((cout,), 103, -0.5, 0.5))
    with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=(kh // 2, kw // 2))
    ref = {"none": lambda t: t, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}[act](ref).float()
    got = ops.conv2d_f16x3(x, w, b, activation=act)
    err = (got - ref).abs().max().item()
    print(f"conv {cin}->{cout} {kh}x{kw} {act}: max abs err {err:.2e} (|ref| max {ref.abs().max().item():.2f})")
    assert err < 3e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("cin,cout,k,stride,H,W", [
    (64, 96, 3, 2, 64, 80),                 # encoder layer2 entry (3x3 / 2)
    (64, 96, 1, 2, 64, 80),                 # its 1x1 / 2 skip projection
    (96, 128, 3, 2, 32, 40),                # layer3 entry; 96 channels = one full + one half K block
    (96, 96, 3, 1, 40, 48),                 # layer2 body
    (160, 64, 1, 1, 64, 80),                # im2col'ed 7x7 / 2 stem (147 -> 160 channels)
])
def test_conv_strided_and_partial_blocks(ops, cin, cout, k, stride, H, W):
    n = 2
    x = dev(det_uniform((n, cin, H, W), 111, -2.0, 2.0))
    w = dev(det_uniform((cout, cin, k, k), 112, -1.0, 1.0)) * (1.0 / np.sqrt(cin * k * k))
    b = dev(det_uniform((cout,), 113, -0.5, 0.5))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=k // 2).float()
    got = ops.conv2d_f16x3(x, w, b, stride=stride)
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    print(f"conv {cin}->{cout} {k}x{k}/{stride}: max abs err {err:.2e}")
    assert err < 3e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("cin,cout,k,stride,H,W", [(64, 64, 3, 1, 44, 52), (64, 96, 3, 2, 50, 70), (64, 96, 1, 2, 50, 70), (176, 64, 1, 1, 40, 24)])
def test_conv_instnorm_partial_sums(ops, cin, cout, k, stride, H, W):
    """conv.cu kind 6: the fp32-only epilogue also writes per-(tile, lane quarter) sums / sums of squares; reduced by
    rpe_instnorm_stats_from_partials they equal InstanceNorm2d's statistics of the convolution output (extractor.py:23-56).
    Ragged sizes: pixel rows outside the image and the ghost tile of an odd tile count contribute nothing."""
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib, tc
    from rpe_b200.ops import _p, _stream, check
    n = 3
    x = dev(det_uniform((n, cin, H, W), 131, -2.0, 2.0))
    w = dev(det_uniform((cout, cin, k, k), 132, -1.0, 1.0)) * (1.0 / np.sqrt(cin * k * k))
    b = dev(det_uniform((cout,), 133, -0.5, 0.5))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=k // 2)
    OH, OW = ref.shape[-2:]
    planes = tc.Planes(n, H, W, cin, x.device)
    tc.nchw_to_planes(x, planes)
    wts = tc.pack_weight(w.float(), 0, cin, cout)
    out = torch.zeros((n, OH, OW, cout), dtype=torch.float32, device=x.device)
    tiles = ((OW + 7) // 8) * ((OH + 15) // 16)
    part = torch.full((n * tiles * 4, cout, 2), float("nan"), dtype=torch.float32, device=x.device)
    plan = tc.ConvPlan("stats", [(planes, 0, cin, wts)], (n, H, W), k, k, cout, "none", bias=b.float().contiguous(), stride=stride,
                       out_f32=out, stat_partials=part)
    assert plan.tiles_per_image == tiles
    plan.run()
    stats = torch.empty((n, cout, 2), dtype=torch.float32, device=x.device)
    ws = torch.empty(_lib.lib().rpe_instnorm_workspace_bytes(n, cout), dtype=torch.uint8, device=x.device)
    check(_lib.lib().rpe_instnorm_stats_from_partials(_p(part), _p(stats), n, 4 * tiles, cout, cout, OH * OW, 1e-5, _p(ws), ws.numel(),
                                                      _stream()), "rpe_instnorm_stats_from_partials")
    torch.cuda.synchronize()
    assert (out.permute(0, 3, 1, 2).double() - ref).abs().max().item() < 3e-5 * max(1.0, ref.abs().max().item())
    assert torch.isfinite(part).all()                                   # every slot of every real tile was written
    o64 = out.double().reshape(n, OH * OW, cout)
    mean, var = o64.mean(1), o64.var(1, unbiased=False)
    err_m = (stats[..., 0].double() - mean).abs().max().item()
    err_r = (stats[..., 1].double() * torch.sqrt(var + 1e-5) - 1.0).abs().max().item()
    print(f"fused instnorm stats {cin}->{cout} {k}x{k}/{stride} {H}x{W}: mean err {err_m:.2e}, rstd rel err {err_r:.2e}")
    assert err_m < 1e-6 and err_r < 1e-5


@pytest.mark.parametrize("H,W", [(50, 70), (64, 80), (33, 129)])
def test_stem_im2col_vs_unfold(ops, H, W):
    """rpe_im2col7s2_split: K axis k = ky*24 + kx*3 + c of the 7x7 / 2 windows of 2*(img/255)-1 (raft.py:82-83, extractor.py:124),
    zero padding, odd sizes and output rows that do not fill a 32-pixel CTA."""
    import rpe_b200  # noqa: F401
    from rpe_b200.core.RAFT.core.encoder_tc import stem_planes
    n = 2
    img = dev(det_uniform((n, 3, H, W), 141, 0.0, 255.0))
    pl = stem_planes(img)
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    x = 2.0 * (img / 255.0) - 1.0
    cols = F.unfold(x, 7, padding=3, stride=2).reshape(n, 3, 7, 7, OH, OW)          # (n, c, ky, kx, oy, ox)
    ref = torch.zeros((n, OH, OW, 7, 24), device=img.device)
    ref[..., :21] = cols.permute(0, 4, 5, 2, 3, 1).reshape(n, OH, OW, 7, 21)
    got = pl.float()[..., :168].reshape(n, OH, OW, 7, 24)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    print(f"stem im2col {H}x{W}: max abs err {err:.2e}")
    assert err < 2e-5                                                               # 16-bit split of values in [-1, 1]
    assert (pl.float()[..., 168:] == 0).all()


def test_conv_addend_residual_and_single_pass(ops):
    n, cin, cout, H, W = 2, 128, 128, 32, 40
    x = dev(det_uniform((n, cin, H, W), 121, -2.0, 2.0))
    w = dev(det_uniform((cout, cin, 3, 3), 122, -1.0, 1.0)) * (1.0 / np.sqrt(cin * 9))
    b = dev(det_uniform((cout,), 123, -0.5, 0.5))
    pre = dev(det_uniform((n, cout, H, W), 124, -1.0, 1.0))
    res = dev(det_uniform((n, cout, H, W), 125, -1.0, 1.0))
    conv = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    ref = torch.relu(torch.tanh(conv + pre.double()) + res.double()).float()
    got = ops.conv2d_f16x3(x, w, b, activation="tanh", pre=pre, res=res)
    assert (got - ref).abs().max().item() < 3e-5
    # single-pass bf16 (no compensation): only bf16-level agreement is expected
    got1 = ops.conv2d_f16x3(x, w, b, single_pass=True)
    err1 = (got1 - conv.float()).abs().max().item()
    print(f"single-pass bf16 conv: max abs err {err1:.2e}")
    assert 1e-5 < err1 < 5e-2


def test_residual_read_from_split_planes_in_place(ops):
    """The encoders keep their residual stream in fp16 split planes only: the block-output convolution (cnet, folded BN) and
    rpe_norm_act_split_res (fnet, instance norm) read the skip connection from planes and write the block output over them."""
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib, tc
    from rpe_b200.ops import _p, _stream, check
    n, C, H, W = 2, 96, 36, 44
    x = dev(det_uniform((n, H, W, C), 131, -2.0, 2.0))                       # conv input, NHWC
    skip = torch.relu(dev(det_uniform((n, H, W, C), 132, -1.0, 3.0)))        # residual stream (a relu output, like the encoder's)
    w = dev(det_uniform((C, C, 3, 3), 133, -1.0, 1.0)) * (1.0 / np.sqrt(C * 9))
    b = dev(det_uniform((C,), 134, -0.5, 0.5))
    xin, stream = tc.Planes(n, H, W, C, x.device), tc.Planes(n, H, W, C, x.device)
    xin.hi, xin.lo = tc.split_planes(x)
    stream.hi, stream.lo = tc.split_planes(skip)
    skip_seen = stream.float().double()                                      # what the planes hold (22 bits of `skip`)
    assert (skip_seen - skip.double()).abs().max().item() < 1e-6
    conv = F.conv2d(xin.float().double().permute(0, 3, 1, 2), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    # (1) convolution epilogue: out planes == residual planes
    plan = tc.ConvPlan("res_planes", [(xin, 0, C, tc.pack_weight(w, 0, C, C))], (n, H, W), 3, 3, C, "relu", bias=b, out_planes=stream,
                       res_planes=stream)
    plan.run()
    torch.cuda.synchronize()
    ref = torch.relu(torch.relu(conv) + skip_seen)
    err = (stream.float().double() - ref).abs().max().item()
    print(f"conv + residual from planes, in place: max abs err {err:.2e}")
    assert err < 3e-5
    # (2) norm_act: relu((a - mean) * rstd) + planes, in place
    stream.hi, stream.lo = tc.split_planes(skip)
    a = conv.float().contiguous()
    mean, var = a.double().mean((1, 2)), a.double().var((1, 2), unbiased=False)
    stats = torch.stack((mean, 1.0 / torch.sqrt(var + 1e-5)), -1).float().contiguous()        # (n, C, 2) = mean, rstd
    check(_lib.lib().rpe_norm_act_split_res(_p(a), _p(stats), 1, None, None, _p(stream.hi), _p(stream.lo), C, None, _p(stream.hi), _p(stream.lo),
                                            C, n, H * W, C, _stream()), "rpe_norm_act_split_res")
    torch.cuda.synchronize()
    ref = torch.relu(torch.relu((a.double() - stats[:, None, None, :, 0].double()) * stats[:, None, None, :, 1].double()) + skip_seen)
    err = (stream.float().double() - ref).abs().max().item()
    print(f"norm_act + residual from planes, in place: max abs err {err:.2e}")
    assert err < 3e-6
    # both an fp32 and a plane addend: rejected
    assert _lib.lib().rpe_norm_act_split_res(_p(a), _p(stats), 1, _p(a), None, _p(stream.hi), _p(stream.lo), C, None, _p(stream.hi),
                                             _p(stream.lo), C, n, H * W, C, _stream()) != 0


def test_update_operator_vs_torch_trunk(ops):
    """12 GRU iterations on the tensor-core path vs the torch fp32 trunk, trained weights, real feature maps."""
    if not os.path.isfile(CKPT) or not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "golden_full.npz")):
        pytest.skip("checkpoint / full golden not shipped")
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_net import PoseNet
    g = np.load(os.path.join(ROOT, "oracle", "_ref", "golden_full.npz"))
    ck = torch.load(CKPT, map_location="cpu", weights_only=False)
    outs = {}
    for prec in ("fp32", "fp16x3"):
        cfg = dict(ck["config"]["model"], image_shape=(512, 640), lbgfs_iters=20, use_weights=True, precision=prec)
        model = PoseNet(cfg)
        model.load_state_dict(ck["state_dict"])
        model = model.cuda().eval()
        i1 = dev(g["imgs_l"][1:3].astype(np.float32))
        i2 = torch.cat((dev(g["imgs_l"][2:3].astype(np.float32)), dev(g["imgs_r"][2:3].astype(np.float32))))
        preds, net, inp = model.flow(i1, i2)
        outs[prec] = (preds[-1].clone(), net.clone())
    epe = (outs["fp32"][0] - outs["fp16x3"][0]).pow(2).sum(1).sqrt()
    print(f"fp16x3 update operator vs fp32 trunk: flow EPE mean {epe.mean().item():.2e} max {epe.max().item():.2e}; "
          f"hidden state max diff {(outs['fp32'][1] - outs['fp16x3'][1]).abs().max().item():.2e}")
    assert epe.mean().item() < 2e-4 and epe.max().item() < 5e-3
    ref = dev(np.stack((g["s_time_flow"],)))
    epe_ref = (outs["fp16x3"][0][0:1] - ref).pow(2).sum(1).sqrt()
    assert epe_ref.mean().item() < 1e-2                                    # north-star flow gate vs the reference itself


@pytest.mark.parametrize("H,W", [(64, 80), (44, 52)])
def test_flow_head_tap_projection_vs_torch_fp32(ops, H, W):
    """FlowHead conv2(relu(conv1(x))) with conv2 evaluated as a per-pixel projection in conv1's epilogue (conv.cu mode 3)
    plus rpe_tap_gather3x3, against torch fp32 convolutions (update.py:6-13).  44x52: ragged tiles and an odd tile count."""
    import ctypes as C
    from rpe_b200 import _lib, tc
    from rpe_b200.ops import _p, _stream, check
    n = 3
    x = dev(det_uniform((n, 128, H, W), 31, -1, 1))
    w1 = dev(det_uniform((256, 128, 3, 3), 32, -1, 1)) / (128 * 9) ** 0.5
    b1 = dev(det_uniform((256,), 33, -0.5, 0.5))
    w2 = dev(det_uniform((2, 256, 3, 3), 34, -1, 1)) / (256 * 9) ** 0.5
    b2 = dev(det_uniform((2,), 35, -0.5, 0.5))
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ref = F.conv2d(torch.relu(F.conv2d(x.double(), w1.double(), b1.double(), padding=1)), w2.double(), b2.double(), padding=1).float()
    planes = tc.Planes(n, H, W, 128, x.device)
    tc.nchw_to_planes(x, planes)
    part = torch.full((n, H, W, 36), float("nan"), dtype=torch.float32, device=x.device)
    proj = w2.permute(1, 2, 3, 0).reshape(256, 18).contiguous()
    plan = tc.ConvPlan("fh1", [(planes, 0, 128, tc.pack_weight(w1, 0, 128, 256))], (n, H, W), 3, 3, 256, "relu", bias=b1.contiguous(),
                       out_f32=part, mode=3, aux2=proj)
    plan.run()
    delta = torch.zeros((n, H, W, 4), dtype=torch.float32, device=x.device)
    check(_lib.lib().rpe_tap_gather3x3(_p(part), 36, _p(b2.contiguous()), _p(delta), 4, n, H, W, _stream()), "rpe_tap_gather3x3")
    torch.cuda.synchronize()
    assert not torch.isnan(part).any()
    got = delta[..., :2].permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    print(f"fused flow head {H}x{W}: max abs err {err:.2e} (|ref| max {ref.abs().max().item():.2f})")
    assert err < 5e-5 and float(delta[..., 2:].abs().max()) == 0.0


def test_update_operator_fused_vs_plain_flow_head(ops, monkeypatch):
    """The whole 12-iteration update operator with the fused flow head equals the plain conv1 -> conv2 plans."""
    if not os.path.isfile(CKPT) or not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "golden_full.npz")):
        pytest.skip("checkpoint / full golden not shipped")
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_net import PoseNet
    g = np.load(os.path.join(ROOT, "oracle", "_ref", "golden_full.npz"))
    ck = torch.load(CKPT, map_location="cpu", weights_only=False)
    i1 = dev(g["imgs_l"][1:3].astype(np.float32))
    i2 = torch.cat((dev(g["imgs_l"][2:3].astype(np.float32)), dev(g["imgs_r"][2:3].astype(np.float32))))
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("RPE_FUSED_FLOW_HEAD", flag)
        cfg = dict(ck["config"]["model"], image_shape=(512, 640), lbgfs_iters=20, use_weights=True, precision="fp16x3")
        model = PoseNet(cfg)
        model.load_state_dict(ck["state_dict"])
        model = model.cuda().eval()
        preds, net, inp = model.flow(i1, i2)
        assert model.flow._tc.fused_flow_head == (flag == "1")
        outs[flag] = preds[-1].clone()
    epe = (outs["1"] - outs["0"]).pow(2).sum(1).sqrt()
    print(f"fused vs plain flow head: flow EPE mean {epe.mean().item():.2e} max {epe.max().item():.2e}")
    assert epe.mean().item() < 1e-4 and epe.max().item() < 5e-3


@pytest.mark.parametrize("which", ["fnet", "cnet"])
def test_encoder_tc_vs_torch_fp32(ops, which):
    """Feature / context encoder on the tensor-core path (instance-norm statistics, folded batch norm, strided TMA
    convolutions, im2col stem) vs the torch fp32 encoder, trained weights, real 640x512 images."""
    if not os.path.isfile(CKPT) or not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "golden_full.npz")):
        pytest.skip("checkpoint / full golden not shipped")
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_net import PoseNet
    g = np.load(os.path.join(ROOT, "oracle", "_ref", "golden_full.npz"))
    ck = torch.load(CKPT, map_location="cpu", weights_only=False)
    imgs = torch.cat((dev(g["imgs_l"][0:2].astype(np.float32)), dev(g["imgs_r"][0:1].astype(np.float32))))
    outs = {}
    for prec in ("fp32", "fp16x3"):
        cfg = dict(ck["config"]["model"], image_shape=(512, 640), lbgfs_iters=20, use_weights=True, precision=prec)
        model = PoseNet(cfg)
        model.load_state_dict(ck["state_dict"])
        raft = model.cuda().eval().flow
        with torch.no_grad():
            outs[prec] = (raft.features(imgs),) if which == "fnet" else raft.context(imgs)
    for a, b in zip(outs["fp32"], outs["fp16x3"]):
        assert a.shape == b.shape
        err = (a - b).abs().max().item()
        print(f"{which}: max abs diff {err:.2e} (|ref| max {a.abs().max().item():.2f}, mean {a.abs().mean().item():.3f})")
        assert err < 2e-4 * max(1.0, a.abs().max().item())
