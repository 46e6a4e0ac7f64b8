"""ORACLE / TEST INFRASTRUCTURE ONLY -- deterministic on-disk fixtures for the file datasets (a StereoMIS-style frame folder and a
SCARED-style top/bottom video folder), written the same way by oracle/make_golden.py (which runs the unmodified reference's
``get_data`` on them) and by the tests (which run the product's).  PNG is lossless, so the frame folder is regenerated from
seeds; the video is lossy, so the encoded ``video.mp4`` itself is committed under tests/golden/file_video/."""
import json
import os

import numpy as np

try:
    from .detrand import det_uniform
except ImportError:                      # oracle/make_golden.py runs with oracle/ itself on sys.path
    from detrand import det_uniform

ORIG_W, ORIG_H = 200, 170            # decoded frame size; the working size (160, 128) needs a 0.8 scale and a 4-row crop
IMG_SIZE = (160, 128)                # (W, H) like the reference's ``img_size`` config key
N_FRAMES = 4


def _frame(seed):
    """smooth-ish RGB uint8 frame with saturated highlights (some touching the border) -> HWC"""
    coarse = det_uniform((ORIG_H // 5 + 2, ORIG_W // 5 + 2, 3), seed, 20.0, 235.0)
    img = np.kron(coarse, np.ones((5, 5, 1), np.float32))[:ORIG_H, :ORIG_W]
    img = img + det_uniform((ORIG_H, ORIG_W, 3), seed + 1, -12.0, 12.0)
    yy, xx = np.mgrid[0:ORIG_H, 0:ORIG_W]
    for k, (cy, cx, rad) in enumerate(((0, 30, 9), (ORIG_H - 1, 120, 11), (60, ORIG_W - 1, 8), (90, 70, 6), (40, 150, 14))):
        img[(yy - cy) ** 2 + (xx - cx - 3 * (seed % 5)) ** 2 <= rad * rad] = 250 + k % 3
    return np.clip(img, 0, 255).astype(np.uint8)


def _tool_mask(seed, h, w):
    """instrument mask at its own resolution: 255 = tissue, 0 = tool (a wedge entering from the bottom)"""
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.full((h, w), 255, np.uint8)
    m[(yy > h * 0.55) & (np.abs(xx - w * (0.3 + 0.05 * (seed % 4))) < (yy - h * 0.55) * 0.6)] = 0
    return m


INI = """[StereoLeft]
res_x = 200
res_y = 170
fc_x = 181.5
fc_y = 180.25
cc_x = 101.25
cc_y = 83.5
kc_0 = -0.0125
kc_1 = 0.004
kc_2 = 0.0002
kc_3 = -0.0003
kc_4 = 0.0
kc_5 = 0.0
kc_6 = 0.0
kc_7 = 0.0

[StereoRight]
res_x = 200
res_y = 170
fc_x = 182.0
fc_y = 180.75
cc_x = 98.5
cc_y = 84.25
kc_0 = -0.0105
kc_1 = 0.0035
kc_2 = -0.0001
kc_3 = 0.0002
kc_4 = 0.0
kc_5 = 0.0
kc_6 = 0.0
kc_7 = 0.0
R_0 = 0.99995
R_1 = -0.0012
R_2 = 0.0099
R_3 = 0.00125
R_4 = 0.999987
R_5 = -0.005
R_6 = -0.009894
R_7 = 0.005012
R_8 = 0.999938
T_0 = -4.21
T_1 = 0.015
T_2 = 0.022
"""

JSON_CAL = {"data": {"width": 200, "height": 170,
                     "intrinsics": [{"f": [181.5, 180.25], "c": [101.25, 83.5], "k": [-0.0125, 0.004, 0.0002, -0.0003, 0.0]},
                                    {"f": [182.0, 180.75], "c": [98.5, 84.25], "k": [-0.0105, 0.0035, -0.0001, 0.0002, 0.0]}],
                     "extrinsics": {"T": [-4.21, 0.015, 0.022], "om": [0.005006, 0.009897, 0.001225]}}}

YAML_CAL = """%YAML:1.0
---
Camera.width: 200
Camera.height: 170
M1: !!opencv-matrix
   rows: 3
   cols: 3
   dt: d
   data: [ 181.5, 0., 101.25, 0., 180.25, 83.5, 0., 0., 1. ]
D1: !!opencv-matrix
   rows: 1
   cols: 5
   dt: d
   data: [ -0.0125, 0.004, 0.0002, -0.0003, 0. ]
M2: !!opencv-matrix
   rows: 3
   cols: 3
   dt: d
   data: [ 182., 0., 98.5, 0., 180.75, 84.25, 0., 0., 1. ]
D2: !!opencv-matrix
   rows: 1
   cols: 5
   dt: d
   data: [ -0.0105, 0.0035, -0.0001, 0.0002, 0. ]
R: !!opencv-matrix
   rows: 3
   cols: 3
   dt: d
   data: [ 0.99995, -0.0012, 0.0099, 0.00125, 0.999987, -0.005, -0.009894, 0.005012, 0.999938 ]
T: !!opencv-matrix
   rows: 1
   cols: 3
   dt: d
   data: [ -4.21, 0.015, 0.022 ]
"""


def write_frame_folder(root):
    """<root>/video_frames/00000kl.png|r.png, <root>/masks/00000kl.png (half resolution), StereoCalibration.ini"""
    import cv2
    os.makedirs(os.path.join(root, "video_frames"), exist_ok=True)
    os.makedirs(os.path.join(root, "masks"), exist_ok=True)
    for k in range(N_FRAMES):
        for side, seed in (("l", 1000 + 10 * k), ("r", 1005 + 10 * k)):
            cv2.imwrite(os.path.join(root, "video_frames", f"{k:06d}{side}.png"), _frame(seed)[:, :, ::-1])
        cv2.imwrite(os.path.join(root, "masks", f"{k:06d}l.png"), _tool_mask(k, ORIG_H // 2, ORIG_W // 2))
    with open(os.path.join(root, "StereoCalibration.ini"), "w") as f:
        f.write(INI)
    return root


def video_frames():
    """the 5 top/bottom frames that were encoded into tests/golden/file_video/video.mp4 (BGR for cv2.VideoWriter)"""
    return [np.concatenate((_frame(2000 + 10 * k), _frame(2005 + 10 * k)), axis=0)[:, :, ::-1].copy() for k in range(5)]


def write_video_folder(root, encode=True):
    """<root>/video.mp4 (only with encode=True: the committed file is the fixture), video.json, groundtruth.txt (4 poses for 5
    frames: the sequence must stop with its ground truth), endoscope_calibration.yaml"""
    import cv2
    os.makedirs(root, exist_ok=True)
    if encode:
        wr = cv2.VideoWriter(os.path.join(root, "video.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 10, (ORIG_W, 2 * ORIG_H))
        assert wr.isOpened()
        for fr in video_frames():
            wr.write(fr)
        wr.release()
    with open(os.path.join(root, "video.json"), "w") as f:
        json.dump([{"timestamp": 1700000000 + 40 * k} for k in range(5)], f)
    with open(os.path.join(root, "groundtruth.txt"), "w") as f:
        f.write("# timestamp tx ty tz qx qy qz qw\n")
        for k in range(4):
            a = 0.01 * k
            f.write(f"{k} {0.001 * k} {-0.0005 * k} {0.002 * k} 0.0 {np.sin(a / 2)} 0.0 {np.cos(a / 2)}\n")
    with open(os.path.join(root, "endoscope_calibration.yaml"), "w") as f:
        f.write(YAML_CAL)
    return root


def write_json_calibration(root):
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, "camcal.json"), "w") as f:
        json.dump(JSON_CAL, f)
    return os.path.join(root, "camcal.json")
