"""CorrBlock with the reference's interface (/root/reference/core/RAFT/core/corr.py:12-60), backed by
the sm_100a kernels: tcgen05/TMA all-pairs GEMM + pyramid pooling (rpe_corr_build) and the
warp-cooperative window lookup (rpe_corr_lookup)."""
from .... import ops


class CorrBlock:
    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, precision=ops.CORR_TF32):
        self.num_levels = num_levels
        self.radius = radius
        self._pyr = ops.CorrPyramid(fmap1.float().contiguous(), fmap2.float().contiguous(), num_levels, radius, precision)
        # same attribute the reference exposes: list of (B*h*w, 1, h_l, w_l) volumes
        self.corr_pyramid = [self._pyr.level(l) for l in range(num_levels)]

    def __call__(self, coords):
        """coords (B,2,h,w), channel 0 = x  ->  (B, num_levels*(2r+1)^2, h, w) float32 contiguous."""
        return self._pyr(coords.float().contiguous())

    @staticmethod
    def corr(fmap1, fmap2):
        """All-pairs volume (B,h,w,1,h,w) / sqrt(C)  (corr.py:52-60)."""
        B, _, h, w = fmap1.shape
        pyr = ops.CorrPyramid(fmap1.float().contiguous(), fmap2.float().contiguous(), 1, 4, ops.CORR_TF32X3)
        return pyr.level(0).view(B, h, w, 1, h, w).clone()
