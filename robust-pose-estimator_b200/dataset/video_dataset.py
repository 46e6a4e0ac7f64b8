"""Top/bottom stereo video dataset (reference: /root/reference/dataset/video_dataset.py:14-78): OpenCV decodes on the host, the
device does the rest -- specularity mask, resize + crop, and the rectification as a nearest-neighbour gather through the
integer maps OpenCV itself derives from the rectifier's float maps (``cv2.convertMaps(..., nninterpolation=True)``, which is what
``cv2.remap(INTER_NEAREST)`` evaluates).  Yields (left, right, mask, pose 7-vector, frame id) like the reference."""
import json
import os

import numpy as np
import torch
from torch.utils.data import IterableDataset

from ..core.utils.trajectory import read_freiburg
from ..lie import SE3
from .stereo_dataset import DevicePreprocessor, _chw, _cv2


class DeviceRemap:
    """cv2.remap(img, map1, map2, INTER_NEAREST) (constant zero border) for (n,C,H,W) device tensors."""

    def __init__(self, map1, map2, device):
        cv2 = _cv2()
        xy, _ = cv2.convertMaps(map1, map2, cv2.CV_16SC2, nninterpolation=True)
        x, y = torch.from_numpy(xy[..., 0].astype(np.int64)), torch.from_numpy(xy[..., 1].astype(np.int64))
        h, w = map1.shape
        self.inside = ((x >= 0) & (x < w) & (y >= 0) & (y < h)).to(device)
        self.index = (y.clamp(0, h - 1) * w + x.clamp(0, w - 1)).to(device).reshape(-1)

    def __call__(self, img):
        n, c, h, w = img.shape
        out = img.reshape(n, c, h * w).index_select(2, self.index).reshape(n, c, h, w)
        return out * self.inside.to(out.dtype)


class StereoVideoDataset(IterableDataset):
    def __init__(self, video_file, pose_file=None, img_size=None, rectify=None, sample=1, raw=False, device="cuda"):
        super().__init__()
        assert os.path.isfile(video_file)
        self.video_file = video_file
        self.rectify = rectify
        self.timestamps = None
        stamp_file = video_file.replace(".mp4", ".json")
        if os.path.isfile(stamp_file):
            with open(stamp_file, "r") as f:
                self.timestamps = [s["timestamp"] for s in json.load(f)]
        grabber = _cv2().VideoCapture(video_file)
        self.length = int(grabber.get(_cv2().CAP_PROP_FRAME_COUNT) / sample)
        grabber.release()
        self.sample = sample
        self.poses = read_freiburg(pose_file) if pose_file is not None and os.path.isfile(pose_file) else None
        self.raw = raw
        self.device = torch.device(device)
        self._resize = DevicePreprocessor(img_size) if img_size is not None else None
        self._remaps = None

    def preprocess(self, left_u8, right_u8, mask=None):
        """decoded uint8 halves on the device -> (left, right, mask) as the reference yields them"""
        if self._resize is None:
            from .. import ops
            left, right, mask = left_u8.float(), right_u8.float(), ops.mask_specularities(left_u8.contiguous(), None, radius=5)
        else:
            left, right, mask = self._resize(left_u8, right_u8, mask)
        if self.rectify is not None:
            if self.rectify.mode == "pseudo":          # a sub-pixel affine shift of the right view: OpenCV on the host, like the reference
                pairs = [self.rectify(l.cpu(), r.cpu()) for l, r in zip(left, right)]
                left = torch.stack([p[0] for p in pairs]).to(left.device)
                right = torch.stack([p[1] for p in pairs]).to(left.device)
            else:
                if self._remaps is None:
                    m = self.rectify.maps
                    self._remaps = (DeviceRemap(m["lmap1"], m["lmap2"], left.device), DeviceRemap(m["rmap1"], m["rmap2"], left.device))
                left, right = self._remaps[0](left), self._remaps[1](right)
        return left, right, mask

    def __iter__(self):
        cv2 = _cv2()
        grabber = cv2.VideoCapture(self.video_file)
        counter = 0
        try:
            while True:
                ok, img = grabber.read()
                counter += 1
                if not ok:
                    break
                if (counter - 1) % self.sample != 0:
                    continue
                if self.poses is not None and len(self.poses) <= counter - 1:
                    break                                   # the sequence ends with its ground truth (video_dataset.py:53-56)
                pose = self.poses[counter - 1] if self.poses is not None else SE3.Identity(1)[0]
                half = img.shape[0] // 2
                left, right = _chw(img[:half]), _chw(img[half:])          # upper half = left view
                number = str(self.timestamps[counter - 1] if self.timestamps is not None else counter)
                if self.raw:
                    yield left, right, pose.vec(), number
                else:
                    l, r, m = self.preprocess(left[None].to(self.device), right[None].to(self.device))
                    yield l[0], r[0], m[0], pose.vec(), number
        finally:
            grabber.release()

    def __len__(self):
        return self.length
