"""RAFT feature / context encoder, written as a weight table + functional forward.

Architecture of the reference's ``BasicEncoder`` (/root/reference/core/RAFT/core/extractor.py:118-192;
``ResidualBlock`` :6-56): 7x7/2 conv 3->64, norm, relu; three stages of two residual blocks
(64 /1, 96 /2, 128 /2; a strided block adds a 1x1/stride projection + norm on the skip); 1x1 conv to
``output_dim``.  fnet uses InstanceNorm2d (no parameters), cnet uses eval-mode BatchNorm2d."""
import torch.nn.functional as F

from ...utils.param_tree import bn_entries, conv_entries

_STAGES = (("layer1", 64, 1), ("layer2", 96, 2), ("layer3", 128, 2))


def encoder_entries(prefix, norm, output_dim):
    e = []
    if norm == "batch":
        e += bn_entries(prefix + "norm1", 64)
    e += conv_entries(prefix + "conv1", 3, 64, 7)
    cin = 64
    for layer, dim, stride in _STAGES:
        for blk in (0, 1):
            p = f"{prefix}{layer}.{blk}."
            s = stride if blk == 0 else 1
            e += conv_entries(p + "conv1", cin if blk == 0 else dim, dim, 3)
            e += conv_entries(p + "conv2", dim, dim, 3)
            if norm == "batch":
                e += bn_entries(p + "norm1", dim) + bn_entries(p + "norm2", dim)
                if s != 1:
                    e += bn_entries(p + "norm3", dim)
            if s != 1:
                e += conv_entries(p + "downsample.0", cin, dim, 1)
                if norm == "batch":
                    e += bn_entries(p + "downsample.1", dim)       # the reference registers norm3 twice
        cin = dim
    e += conv_entries(prefix + "conv2", 128, output_dim, 1)
    return e


def _norm(x, W, name, norm):
    if norm == "instance":
        return F.instance_norm(x, eps=1e-5)
    if norm == "folded":
        return x
    return F.batch_norm(x, W[name + ".running_mean"], W[name + ".running_var"], W[name + ".weight"], W[name + ".bias"],
                        False, 0.0, 1e-5)


def _conv(x, W, name, stride=1, padding=0):
    return F.conv2d(x, W[name + ".weight"], W[name + ".bias"], stride, padding)


def encoder_forward(x, W, prefix, norm):
    """x (N,3,H,W) already scaled to [-1,1] -> (N,output_dim,H/8,W/8)."""
    x = F.relu(_norm(_conv(x, W, prefix + "conv1", 2, 3), W, prefix + "norm1", norm))
    for layer, _, stride in _STAGES:
        for blk in (0, 1):
            p = f"{prefix}{layer}.{blk}."
            s = stride if blk == 0 else 1
            y = F.relu(_norm(_conv(x, W, p + "conv1", s, 1), W, p + "norm1", norm))
            y = F.relu(_norm(_conv(y, W, p + "conv2", 1, 1), W, p + "norm2", norm))
            if s != 1:
                x = _norm(_conv(x, W, p + "downsample.0", s, 0), W, p + "norm3", norm)
            x = F.relu(x + y)
    return _conv(x, W, prefix + "conv2")
