"""Frame container with the reference's fields (/root/reference/core/utils/frame_class.py:5-84)."""
import torch


class Frame:
    def __init__(self, img, rimg=None, depth=None, mask=None, confidence=None, flow=None):
        assert img.ndim == 4
        self.img = img.contiguous()
        self.rimg = self.img if rimg is None else rimg.contiguous()
        hw = tuple(self.img.shape[-2:])
        dev = self.img.device
        self.mask = torch.ones((1, 1, *hw), dtype=torch.bool, device=dev) if mask is None else mask.bool()
        self.depth = torch.ones((1, 1, *hw), device=dev) if depth is None else depth.contiguous()
        self.confidence = torch.ones((1, 1, *hw), device=dev) if confidence is None else confidence.contiguous()
        self.flow = torch.zeros((1, 2, *hw), device=dev) if flow is None else flow.contiguous()
        assert self.rimg.shape == self.img.shape
        for t in (self.depth, self.mask, self.confidence, self.flow):
            assert tuple(t.shape[-2:]) == hw

    def to(self, dev_or_type):
        for k in ("img", "rimg", "depth", "mask", "confidence"):
            setattr(self, k, getattr(self, k).to(dev_or_type))
        return self

    @property
    def shape(self):
        return self.img.shape[-2:]

    @property
    def device(self):
        return self.img.device
