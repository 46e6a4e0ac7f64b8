"""Stereo transforms of the reference's input pipeline (/root/reference/dataset/transforms.py:20-39) on frames that already are
on the device: ``ResizeStereo`` = resize that conserves the aspect ratio + centre crop, one kernel per tensor (rpe_resize_crop).
Images may be uint8 (the uint8 -> float conversion of dataset/stereo_dataset.py:36-37 is folded in) or float; masks are resized
with nearest-neighbour like the reference."""
import torch

from .. import _lib
from ..ops import _p, _stream, check


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, *args):
        for tr in self.transforms:
            args = tr(*args)
        return args


class StereoTransform:
    def __call__(self, left, right, mask):
        return left, right, mask


class ResizeStereo(StereoTransform):
    def __init__(self, size):
        """size: (W, H) like the reference's ``img_size`` config key."""
        self.size = [int(size[1]), int(size[0])]

    def __call__(self, left, right, mask=None):
        h, w = left.shape[-2:]
        scale = max(self.size[0] / h, self.size[1] / w)
        size = [int(scale * h), int(scale * w)]                    # resize with cropping to conserve the aspect ratio
        return self._resize_with_crop(left, size), self._resize_with_crop(right, size), self._resize_with_crop(mask, size, nearest=True)

    def _resize_with_crop(self, img, size, nearest=False):
        if img is None:
            return None
        if not img.is_cuda:
            raise _lib.RpeError("ResizeStereo: expected CUDA tensors (rpe_b200 has no CPU path)")
        lead = img.shape[:-2]
        Hi, Wi = img.shape[-2:]
        x = img.reshape(-1, Hi, Wi).contiguous()
        if x.dtype == torch.bool:
            x = x.view(torch.uint8)
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        H, W = self.size
        rh, rw = size
        top, left = int(round((rh - H) / 2.0)), int(round((rw - W) / 2.0))      # torchvision center_crop
        out = torch.empty((x.shape[0], H, W), dtype=torch.uint8 if nearest else torch.float32, device=img.device)
        check(_lib.lib().rpe_resize_crop(_p(x), 1 if x.dtype == torch.uint8 else 0, _p(out), x.shape[0], 1, Hi, Wi, rh, rw, top, left, H, W,
                                         1 if nearest else 0, _stream()), "rpe_resize_crop")
        out = out.reshape(*lead, H, W)
        if nearest:
            return out.bool() if img.dtype == torch.bool else out.to(img.dtype)
        return out
