"""ORACLE / TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's ResizeStereo
(/root/reference/dataset/transforms.py:20-39): torchvision.transforms.functional.resize on float tensors (= ATen
_upsample_bilinear2d_aa: anti-aliased bilinear, align_corners=False) or with NEAREST (= ATen upsample_nearest2d), then
torchvision center_crop.  Pinned by tests/golden/resize_stereo.npz (outputs of the reference class itself)."""
import numpy as np

F32 = np.float32


def _aa_weights(in_size, out_size):
    scale = F32(in_size) / F32(out_size)
    support = scale if scale >= 1 else F32(1)
    invs = F32(1) / scale if scale >= 1 else F32(1)
    taps = []
    for i in range(out_size):
        center = scale * (F32(i) + F32(0.5))
        xmin = max(0, int(center - support + F32(0.5)))
        xsize = min(in_size, int(center + support + F32(0.5))) - xmin
        w = np.array([max(F32(0), F32(1) - abs((F32(j + xmin) - center + F32(0.5)) * invs)) for j in range(xsize)], dtype=F32)
        taps.append((xmin, (w / w.sum(dtype=F32)).astype(F32)))
    return taps


def resize_bilinear_aa(img, rh, rw):
    """img (C,H,W) float32 -> (C,rh,rw): horizontal pass, then vertical pass, fp32 accumulation in tap order."""
    C, H, W = img.shape
    tx, ty = _aa_weights(W, rw), _aa_weights(H, rh)
    tmp = np.zeros((C, H, rw), F32)
    for x, (x0, w) in enumerate(tx):
        acc = np.zeros((C, H), F32)
        for j, wj in enumerate(w):
            acc = (acc + img[:, :, x0 + j] * wj).astype(F32)
        tmp[:, :, x] = acc
    out = np.zeros((C, rh, rw), F32)
    for y, (y0, w) in enumerate(ty):
        acc = np.zeros((C, rw), F32)
        for j, wj in enumerate(w):
            acc = (acc + tmp[:, y0 + j, :] * wj).astype(F32)
        out[:, y, :] = acc
    return out


def resize_nearest(img, rh, rw):
    C, H, W = img.shape
    ys = np.minimum(np.floor(np.arange(rh, dtype=F32) * (F32(H) / F32(rh))).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(rw, dtype=F32) * (F32(W) / F32(rw))).astype(np.int64), W - 1)
    return img[:, ys][:, :, xs]


def resize_stereo(img, size_wh, nearest=False):
    """-> (C, H, W): resize conserving the aspect ratio, then centre crop (transforms.py:25-39)."""
    W_t, H_t = int(size_wh[0]), int(size_wh[1])
    h, w = img.shape[-2:]
    scale = max(H_t / h, W_t / w)
    rh, rw = int(scale * h), int(scale * w)
    r = resize_nearest(img, rh, rw) if nearest else resize_bilinear_aa(img.astype(F32), rh, rw)
    top, left = int(round((rh - H_t) / 2.0)), int(round((rw - W_t) / 2.0))
    return r[:, top:top + H_t, left:left + W_t]
