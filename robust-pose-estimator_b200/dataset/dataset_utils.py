"""Input contract of the tracker (reference: /root/reference/dataset/dataset_utils.py:10-55): ``get_data`` resolves an input
folder to (dataset, rectified calibration) -- a folder of stereo PNG frames (``video_frames*/*l.png`` + ``masks*``) or a
top/bottom stereo ``*.mp4`` with ``groundtruth.txt`` -- plus the synthetic sequences used by the tests and the benchmark
(``synthetic:<n_frames>[:seed]`` as input path), and the sequential sub-sampler.  OpenCV only decodes; the per-frame
preprocessing runs on the device (dataset/stereo_dataset.py)."""
import glob
import os

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


def get_data(input_path, img_size, sample_video=1, rect_mode="conventional", force_video=False, raw=False):
    """-> (dataset, calib) with calib = {'intrinsics': {'left': 3x3, ...}, 'bf': float, ...}.  raw=True: the file datasets
    return decoded uint8 host frames and leave ``dataset.preprocess`` (device) to the consumer."""
    if isinstance(input_path, SyntheticStereoSequence):
        seq = input_path
    elif isinstance(input_path, str) and input_path.startswith("synthetic:"):
        parts = input_path.split(":")
        seq = SyntheticStereoSequence(int(parts[1]), tuple(img_size), seed=int(parts[2]) if len(parts) > 2 else 0,
                                      smooth_walk=True)
    else:
        return _file_data(input_path, tuple(img_size), sample_video, rect_mode, force_video, raw)
    return SyntheticStereoDataset(seq), seq.calib


def _file_data(input_path, img_size, sample_video, rect_mode, force_video, raw):
    from .rectification import StereoRectifier, find_calibration_file
    from .stereo_dataset import StereoDataset
    from .video_dataset import StereoVideoDataset
    rect = StereoRectifier(find_calibration_file(input_path), img_size_new=img_size, mode=rect_mode)
    calib = rect.get_rectified_calib()
    if not force_video:
        try:
            return StereoDataset(input_path, img_size=calib["img_size"], raw=raw), calib
        except AssertionError:
            pass                                     # no frame folder: fall through to the video, like the reference
    videos = glob.glob(os.path.join(input_path, "*.mp4"))
    if not videos:
        raise RuntimeError(f"neither video_frames*/*l.png nor an .mp4 found in {input_path}")
    dataset = StereoVideoDataset(videos[0], os.path.join(input_path, "groundtruth.txt"), img_size=calib["img_size"], sample=sample_video,
                                 rectify=rect, raw=raw)
    return dataset, calib
