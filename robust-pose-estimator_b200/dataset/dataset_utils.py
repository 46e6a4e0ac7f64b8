"""Input contract of the tracker (reference: /root/reference/dataset/dataset_utils.py:10-55).  The
reference's file-based datasets (OpenCV rectification, video decoding) are CPU preprocessing outside the
hot path; this module provides the same ``get_data`` entry point for the synthetic sequences used by the
tests and the benchmark (``synthetic:<n_frames>[:seed]`` as input path) and the sequential sub-sampler."""
import torch
from torch.utils.data import Dataset, Sampler

from .synthetic import SyntheticStereoSequence


class SyntheticStereoDataset(Dataset):
    def __init__(self, seq):
        self.seq = seq

    def __len__(self):
        return len(self.seq)

    def __getitem__(self, i):
        limg, rimg, mask, n = self.seq[i]
        return torch.from_numpy(limg), torch.from_numpy(rimg), torch.from_numpy(mask), n


def mask_specularities(img, mask=None, spec_thr=0.96):
    """The reference datasets' specularity mask (/root/reference/dataset/stereo_dataset.py:12-16) on frames that are already on
    the device: img (n,3,H,W) uint8 RGB, mask (n,1,H,W) bool or None -> bool (n,1,H,W).  rpe_mask_specularities, bit-exact."""
    from .. import ops
    return ops.mask_specularities(img, mask, spec_thr=spec_thr, radius=5)


class SequentialSubSampler(Sampler):
    def __init__(self, data_source, start=None, stop=None, step=1):
        n = len(data_source)
        self.start = 0 if start is None else max(start, 0)
        self.stop = n if stop is None else min(stop, n)
        self.step = step

    def __iter__(self):
        return iter(range(self.start, self.stop, self.step))

    def __len__(self):
        return max(0, (self.stop - self.start + self.step - 1) // self.step)


def get_data(input_path, img_size, sample_video=1, rect_mode="conventional", force_video=False):
    """-> (dataset, calib) with calib = {'intrinsics': {'left': 3x3}, 'bf': float}."""
    if isinstance(input_path, SyntheticStereoSequence):
        seq = input_path
    elif isinstance(input_path, str) and input_path.startswith("synthetic:"):
        parts = input_path.split(":")
        seq = SyntheticStereoSequence(int(parts[1]), tuple(img_size), seed=int(parts[2]) if len(parts) > 2 else 0,
                                      smooth_walk=True)
    else:
        raise NotImplementedError("file-based StereoMIS/SCARED datasets are CPU preprocessing outside the f2f hot path; "
                                  "use 'synthetic:<frames>[:seed]'")
    return SyntheticStereoDataset(seq), seq.calib
