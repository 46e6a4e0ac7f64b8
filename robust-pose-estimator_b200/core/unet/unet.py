"""TinyUNet confidence head (reference: /root/reference/core/unet/unet.py:8-82): encoder (C_in,16,32,64)
and decoder (64,32,16) of un-padded 3x3 convolutions, max-pooling, transposed-conv up-sampling with
centre-cropped skips, 1x1 head, bilinear resize to the image size.  Functional forward over a weight table.
DownBlock = conv, BN, ReLU, conv;  UpBlock = conv, ReLU, BN, conv."""
import torch
import torch.nn.functional as F

from ..utils.param_tree import bn_entries, conv_entries


def tiny_unet_entries(prefix, in_channels):
    e = []
    enc = (in_channels, 16, 32, 64)
    for i in range(3):
        p = f"{prefix}encoder.enc_blocks.{i}."
        e += conv_entries(p + "conv1", enc[i], enc[i + 1], 3) + bn_entries(p + "norm", enc[i + 1])
        e += conv_entries(p + "conv2", enc[i + 1], enc[i + 1], 3)
    dec = (64, 32, 16)
    for i in range(2):
        e += [(f"{prefix}decoder.upconvs.{i}.weight", (dec[i], dec[i + 1], 2, 2), "conv"),
              (f"{prefix}decoder.upconvs.{i}.bias", (dec[i + 1],), "zeros")]
    for i in range(2):
        p = f"{prefix}decoder.dec_blocks.{i}."
        e += conv_entries(p + "conv1", dec[i], dec[i + 1], 3) + bn_entries(p + "norm", dec[i + 1])
        e += conv_entries(p + "conv2", dec[i + 1], dec[i + 1], 3)
    e += conv_entries(prefix + "head", 16, 1, 1)
    return e


def _bn(x, W, n):
    return F.batch_norm(x, W[n + ".running_mean"], W[n + ".running_var"], W[n + ".weight"], W[n + ".bias"], False, 0.0, 1e-5)


def _conv(x, W, n):
    return F.conv2d(x, W[n + ".weight"], W[n + ".bias"])


def tiny_unet_forward(x, W, prefix, out_size):
    feats = []
    for i in range(3):
        p = f"{prefix}encoder.enc_blocks.{i}."
        x = _conv(F.relu(_bn(_conv(x, W, p + "conv1"), W, p + "norm")), W, p + "conv2")
        feats.append(x)
        if i < 2:
            x = F.max_pool2d(x, 2)
    x = feats[2]
    for i in range(2):
        x = F.conv_transpose2d(x, W[f"{prefix}decoder.upconvs.{i}.weight"], W[f"{prefix}decoder.upconvs.{i}.bias"], stride=2)
        skip = feats[1 - i]
        H, Wd = x.shape[-2:]
        dh, dw = (skip.shape[-2] - H) // 2, (skip.shape[-1] - Wd) // 2
        skip = skip[..., dh:skip.shape[-2] - dh, dw:skip.shape[-1] - dw]
        p = f"{prefix}decoder.dec_blocks.{i}."
        x = torch.cat((x, skip), 1)
        x = _conv(_bn(F.relu(_conv(x, W, p + "conv1")), W, p + "norm"), W, p + "conv2")
    x = _conv(x, W, prefix + "head")
    return F.interpolate(x, out_size, mode="bilinear")
