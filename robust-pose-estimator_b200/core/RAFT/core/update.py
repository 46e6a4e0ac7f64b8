"""RAFT update operator as a weight table + functional forward.

Reference: ``BasicUpdateBlock`` = ``BasicMotionEncoder`` + ``SepConvGRU`` + ``FlowHead`` + mask head
(/root/reference/core/RAFT/core/update.py:79-97, 33-60, 6-14, 114-136).

Exact restructurings (same arithmetic per output element, fewer / larger convolutions):
  * the z and r gates of each GRU half read the same input -> one conv with 256 output channels;
  * the convex-upsampling mask head only runs when its output is consumed (last iteration; the
    reference evaluates it 12x and discards 11 results, pose_net.py:66-67)."""
import torch
import torch.nn.functional as F

from ...utils.param_tree import conv_entries


def update_entries(prefix, cor_planes=324, hidden=128):
    e = []
    p = prefix + "encoder."
    e += conv_entries(p + "convc1", cor_planes, 256, 1) + conv_entries(p + "convc2", 256, 192, 3)
    e += conv_entries(p + "convf1", 2, 128, 7) + conv_entries(p + "convf2", 128, 64, 3)
    e += conv_entries(p + "conv", 64 + 192, 128 - 2, 3)
    g = prefix + "gru."
    for n in ("convz1", "convr1", "convq1"):
        e += conv_entries(g + n, hidden + 128 + hidden, hidden, 1, 5)
    for n in ("convz2", "convr2", "convq2"):
        e += conv_entries(g + n, hidden + 128 + hidden, hidden, 5, 1)
    e += conv_entries(prefix + "flow_head.conv1", hidden, 256, 3) + conv_entries(prefix + "flow_head.conv2", 256, 2, 3)
    e += conv_entries(prefix + "mask.0", 128, 256, 3) + conv_entries(prefix + "mask.2", 256, 64 * 9, 1)
    return e


def prepare_update_weights(W, prefix):
    """Derived tensors: concatenated z|r gate weights.  Call again after loading new weights."""
    g = prefix + "gru."
    for half in ("1", "2"):
        W[g + "convzr" + half + ".weight"] = torch.cat((W[g + "convz" + half + ".weight"], W[g + "convr" + half + ".weight"]), 0)
        W[g + "convzr" + half + ".bias"] = torch.cat((W[g + "convz" + half + ".bias"], W[g + "convr" + half + ".bias"]), 0)
    return W


def _conv(x, W, name, padding=0):
    return F.conv2d(x, W[name + ".weight"], W[name + ".bias"], 1, padding)


def _gru_half(h, x, W, g, half, pad):
    hx = torch.cat((h, x), 1)
    zr = torch.sigmoid(_conv(hx, W, g + "convzr" + half, pad))
    z, r = zr[:, :128], zr[:, 128:]
    q = torch.tanh(_conv(torch.cat((r * h, x), 1), W, g + "convq" + half, pad))
    return (1 - z) * h + z * q


def update_forward(net, inp, corr, flow, W, prefix, want_mask):
    """-> (net, up_mask or None, delta_flow); tensors (B,C,h,w)."""
    p = prefix + "encoder."
    cor = F.relu(_conv(corr, W, p + "convc1"))
    cor = F.relu(_conv(cor, W, p + "convc2", 1))
    flo = F.relu(_conv(flow, W, p + "convf1", 3))
    flo = F.relu(_conv(flo, W, p + "convf2", 1))
    out = F.relu(_conv(torch.cat((cor, flo), 1), W, p + "conv", 1))
    x = torch.cat((inp, out, flow), 1)
    g = prefix + "gru."
    net = _gru_half(net, x, W, g, "1", (0, 2))
    net = _gru_half(net, x, W, g, "2", (2, 0))
    delta = _conv(F.relu(_conv(net, W, prefix + "flow_head.conv1", 1)), W, prefix + "flow_head.conv2", 1)
    mask = None
    if want_mask:
        mask = 0.25 * _conv(F.relu(_conv(net, W, prefix + "mask.0", 1)), W, prefix + "mask.2")
    return net, mask, delta
