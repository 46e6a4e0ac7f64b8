"""Parameter container whose ``state_dict`` keys are arbitrary dotted paths.

The network trunk of this package is written functionally (weight table + F.conv2d) so that layers can
be fused, re-laid-out and cast per precision mode; it still has to load the reference's checkpoints
(trained/*.pth, 282 tensors keyed like ``flow.fnet.layer2.0.downsample.0.weight``) unchanged and to
emit the same keys from ``state_dict()``.  ``ParamTree`` builds the nested module skeleton from the
key list alone."""
import math

import torch
from torch import nn


class ParamTree(nn.Module):
    def add(self, key, tensor, buffer=False):
        head, _, rest = key.partition(".")
        if rest:
            if head not in self._modules:
                self.add_module(head, ParamTree())
            self._modules[head].add(rest, tensor, buffer)
        elif buffer:
            self.register_buffer(head, tensor)
        else:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=False))
        return self

    def table(self, prefix=""):
        """Flat {dotted key: tensor} view (no copies)."""
        out = {}
        for k, v in self.named_parameters(prefix=prefix.rstrip("."), recurse=True):
            out[k] = v.data
        for k, v in self.named_buffers(prefix=prefix.rstrip("."), recurse=True):
            out[k] = v
        return out


def conv_entries(name, cin, cout, kh, kw=None):
    kw = kh if kw is None else kw
    return [(name + ".weight", (cout, cin, kh, kw), "conv"), (name + ".bias", (cout,), "zeros")]


def bn_entries(name, c):
    return [(name + ".weight", (c,), "ones"), (name + ".bias", (c,), "zeros"),
            (name + ".running_mean", (c,), "buf_zeros"), (name + ".running_var", (c,), "buf_ones"),
            (name + ".num_batches_tracked", (), "buf_count")]


def build_tree(entries, generator=None):
    """entries: [(key, shape, kind)] with kind in conv|zeros|ones|buf_zeros|buf_ones|buf_count."""
    tree = ParamTree()
    for key, shape, kind in entries:
        if kind == "conv":
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=generator) * math.sqrt(2.0 / fan_out)   # kaiming_normal_(fan_out, relu)
            tree.add(key, t)
        elif kind == "zeros":
            tree.add(key, torch.zeros(shape))
        elif kind == "ones":
            tree.add(key, torch.ones(shape))
        elif kind == "buf_zeros":
            tree.add(key, torch.zeros(shape), buffer=True)
        elif kind == "buf_ones":
            tree.add(key, torch.ones(shape), buffer=True)
        elif kind == "buf_count":
            tree.add(key, torch.zeros(shape, dtype=torch.int64), buffer=True)
        else:
            raise ValueError(kind)
    return tree
