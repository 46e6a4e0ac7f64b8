"""Flow warping with the reference's function names (/root/reference/core/interpol/flow_utils.py:4-26),
backed by the sm_100a warp kernel (bit-exact sampling coordinates, see csrc/geometry.cu)."""
from ... import ops


def remap_from_flow(x, flow):
    """Bilinear backward warp of x (n,C,H,W) by flow (n,2,H,W); returns (x_warped, valid)."""
    out = ops.remap_bilinear(x.float().contiguous(), flow.float().contiguous())
    valid = (out > 0).any(dim=1).unsqueeze(1)
    return out, valid


def remap_from_flow_nearest(x, flow):
    """Nearest-neighbour backward warp; returns (x_warped float, valid)."""
    out = ops.remap_nearest(x.float().contiguous(), flow.float().contiguous())
    valid = (out > 0).any(dim=1).unsqueeze(1)
    return out, valid
