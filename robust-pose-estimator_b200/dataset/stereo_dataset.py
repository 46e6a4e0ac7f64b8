"""File datasets of the reference's input pipeline: a folder of rectified stereo PNG frames (reference:
/root/reference/dataset/stereo_dataset.py:19-44) and a top/bottom stereo video (/root/reference/dataset/video_dataset.py:14-78).

What stays on the host is what only the host can do -- PNG / video decoding with OpenCV.  Everything the reference then does
per frame on the CPU runs on the device from the decoded uint8 frames (``DevicePreprocessor``): the specularity mask with its
11x11 erosion (rpe_mask_specularities, bit-exact), the aspect-preserving anti-aliased resize + centre crop of both images and
the nearest-neighbour resize of the mask (rpe_resize_crop), and for videos the nearest-neighbour rectification remap.  The
uint8 -> float conversion of stereo_dataset.py:36-37 is folded into the resize kernel.

``dataset[i]`` returns what the reference returns (float images (3,H,W) on 0..255, bool mask (1,H,W), frame id), as device
tensors; with ``raw=True`` it returns the decoded uint8 host tensors instead, so that DataLoader workers only decode and the
consumer calls ``dataset.preprocess`` on the uploaded batch (scripts/infer_trajectory.py does)."""
import glob
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from .transforms import ResizeStereo


def _cv2():
    import cv2
    return cv2


class DevicePreprocessor:
    """(n,3,H0,W0) uint8 RGB left / right + optional (n,1,H0,W0) bool mask, on the device -> float images + bool mask at img_size."""

    def __init__(self, img_size, spec_thr=0.96):
        self.transform = ResizeStereo(img_size)
        self.spec_thr = spec_thr

    def __call__(self, left_u8, right_u8, mask=None):
        from .. import ops
        if mask is not None and mask.dtype != torch.bool:
            mask = mask > 0
        valid = ops.mask_specularities(left_u8.contiguous(), None if mask is None else mask.contiguous(), spec_thr=self.spec_thr, radius=5)
        return self.transform(left_u8, right_u8, valid)


def _chw(img_bgr):
    """decoded BGR HWC uint8 -> RGB CHW uint8 host tensor"""
    return torch.from_numpy(np.ascontiguousarray(img_bgr[:, :, ::-1].transpose(2, 0, 1)))


class StereoDataset(Dataset):
    def __init__(self, input_folder, img_size, raw=False, device="cuda"):
        super().__init__()
        self.imgs = sorted(glob.glob(os.path.join(input_folder, "video_frames*", "*l.png")))
        assert len(self.imgs) > 0
        self.raw = raw
        self.device = torch.device(device)
        self.preprocess = DevicePreprocessor(img_size)

    def decode(self, item):
        """-> (left, right) uint8 (3,H0,W0) RGB, mask uint8 (1,H0,W0) in {0,1}, frame id"""
        cv2 = _cv2()
        path = self.imgs[item]
        left, right = cv2.imread(path), cv2.imread(path.replace("l.png", "r.png"))
        mask = cv2.imread(path.replace("video_frames", "masks"), cv2.IMREAD_GRAYSCALE)
        mask = cv2.resize(mask, dsize=(left.shape[1], left.shape[0]), interpolation=cv2.INTER_NEAREST) > 0
        return _chw(left), _chw(right), torch.from_numpy(mask.astype(np.uint8))[None], os.path.basename(path).split("l.png")[0]

    def __getitem__(self, item):
        left, right, mask, number = self.decode(item)
        if self.raw:
            return left, right, mask, number
        dev = self.device
        left, right, mask = self.preprocess(left[None].to(dev), right[None].to(dev), mask[None].to(dev))
        return left[0], right[0], mask[0], number

    def __len__(self):
        return len(self.imgs)
