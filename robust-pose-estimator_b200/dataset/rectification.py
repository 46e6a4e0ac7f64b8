"""Stereo calibration files and rectified calibration (reference: /root/reference/dataset/rectification.py:10-184 and
dataset/preprocess/stereo_rectify.py:5-64).  Host-side setup that runs once per sequence: it parses the three calibration
formats of the reference's datasets (StereoMIS ``StereoCalibration.ini``, SCARED ``endoscope_calibration.yaml``, ``camcal.json`` /
``camera_calibration.json``), rescales the intrinsics to the working resolution and asks OpenCV for the rectifying
projections, from which the tracker takes ``intrinsics['left']`` and ``bf`` (baseline x focal length).  The per-frame remap of
the video dataset stays an OpenCV call on the decoded host frame, as in the reference (nearest-neighbour ``cv2.remap``)."""
import configparser
import json
import os
import warnings

import numpy as np
import torch


def _cv2():
    import cv2          # only the file datasets need OpenCV
    return cv2


def _camera_matrix(fx, fy, cx, cy):
    return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)


def _read_json(path):
    with open(path, "rb") as f:
        data = json.load(f)["data"]
    left, right = data["intrinsics"][0], data["intrinsics"][1]
    return dict(lkmat=_camera_matrix(left["f"][0], left["f"][1], left["c"][0], left["c"][1]),
                rkmat=_camera_matrix(right["f"][0], right["f"][1], right["c"][0], right["c"][1]),
                ld=np.array(left["k"]), rd=np.array(right["k"]), T=np.array(data["extrinsics"]["T"]),
                R=_cv2().Rodrigues(np.array(data["extrinsics"]["om"]))[0], img_size=(data["width"], data["height"]))


def _read_ini(path):
    cfg = configparser.ConfigParser()
    cfg.read(path)
    left, right = cfg["StereoLeft"], cfg["StereoRight"]

    def kmat(sec):
        return _camera_matrix(float(sec["fc_x"]), float(sec["fc_y"]), float(sec["cc_x"]), float(sec["cc_y"]))

    def dist(sec):
        return np.array([float(sec[f"kc_{k}"]) for k in range(8)])

    return dict(lkmat=kmat(left), rkmat=kmat(right), ld=dist(left), rd=dist(right),
                T=np.array([float(right[f"T_{k}"]) for k in range(3)]),
                R=np.array([float(right[f"R_{k}"]) for k in range(9)]).reshape(3, 3),
                img_size=(float(left["res_x"]), float(left["res_y"])))


def _read_yaml(path):
    fs = _cv2().FileStorage(path, _cv2().FILE_STORAGE_READ)
    mat = lambda key: fs.getNode(key).mat()
    return dict(lkmat=mat("M1"), rkmat=mat("M2"), ld=mat("D1"), rd=mat("D2"), T=mat("T"), R=mat("R"),
                img_size=(int(fs.getNode("Camera.width").real()), int(fs.getNode("Camera.height").real())))


_READERS = {".json": _read_json, ".ini": _read_ini, ".yaml": _read_yaml}
CALIBRATION_FILES = ("camcal.json", "camera_calibration.json", "StereoCalibration.ini", "endoscope_calibration.yaml")


def find_calibration_file(folder):
    """The reference's search order (dataset_utils.py:14-23)."""
    for name in CALIBRATION_FILES:
        path = os.path.join(folder, name)
        if os.path.isfile(path):
            return path
    raise RuntimeError(f"no valid calibration file found in {folder}")


def rectifying_projections(cal, mode="conventional"):
    """-> (maps, P1, P2).  'conventional': cv2.stereoRectify(alpha=0) + undistort-rectify maps (the right map is built with the
    LEFT distortion coefficients, as the reference does, stereo_rectify.py:30); 'pseudo': the camera matrices themselves."""
    if mode == "pseudo":
        return {}, cal["lkmat"].astype("float64"), cal["rkmat"].astype("float64")
    if mode != "conventional":
        raise NotImplementedError(mode)
    cv2 = _cv2()
    size = tuple(cal["img_size"])
    r1, r2, p1, p2, *_ = cv2.stereoRectify(cameraMatrix1=cal["lkmat"].astype("float64"), distCoeffs1=cal["ld"].astype("float64"),
                                           cameraMatrix2=cal["rkmat"].astype("float64"), distCoeffs2=cal["rd"].astype("float64"),
                                           imageSize=size, R=cal["R"].astype("float64"), T=cal["T"].T.astype("float64"), alpha=0)
    maps = {}
    for side, kmat, rot, proj in (("l", cal["lkmat"], r1, p1), ("r", cal["rkmat"], r2, p2)):
        maps[side + "map1"], maps[side + "map2"] = cv2.initUndistortRectifyMap(cameraMatrix=kmat, distCoeffs=cal["ld"], R=rot,
                                                                               newCameraMatrix=proj, size=size, m1type=cv2.CV_32FC1)
    return maps, p1, p2


class StereoRectifier:
    """Same constructor, call and ``get_rectified_calib`` as the reference class."""

    def __init__(self, calib_file, img_size_new=None, mode="conventional"):
        ext = os.path.splitext(calib_file)[1]
        if ext not in _READERS:
            raise NotImplementedError(f"calibration format {ext}")
        cal = _READERS[ext](calib_file)
        assert mode in ("conventional", "pseudo")
        self.mode = mode
        if mode == "pseudo":
            warnings.warn("pseudo rectification used", UserWarning)
        self.scale = 1.0
        if img_size_new is not None:
            # intrinsics at the working resolution: uniform scale to the new width, symmetric vertical crop
            self.scale = img_size_new[0] / cal["img_size"][0]
            h_crop = int((cal["img_size"][1] * self.scale - img_size_new[1]) / 2)
            assert h_crop >= 0, "only vertical crop implemented"
            for k in ("lkmat", "rkmat"):
                cal[k][:2] *= self.scale
                cal[k][1, 2] -= h_crop
            cal["img_size"] = img_size_new
        self.img_size = cal["img_size"]
        self.cal = cal
        self.maps, self.l_intr, self.r_intr = rectifying_projections(cal, mode)

    def __call__(self, img_left, img_right):
        """(3,H,W) host tensors -> rectified (3,H,W) host tensors."""
        cv2 = _cv2()
        left, right = img_left.permute(1, 2, 0).numpy(), img_right.permute(1, 2, 0).numpy()
        if self.mode == "pseudo":
            shift = np.array([[1, 0, self.cal["lkmat"][0][-1] - self.cal["rkmat"][0][-1]],
                              [0, 1, self.cal["lkmat"][1][-1] - self.cal["rkmat"][1][-1]]], dtype=np.float32)
            right = cv2.warpAffine(right, shift, (right.shape[1], right.shape[0]))
        else:
            left = cv2.remap(np.copy(left), self.maps["lmap1"], self.maps["lmap2"], interpolation=cv2.INTER_NEAREST)
            right = cv2.remap(np.copy(right), self.maps["rmap1"], self.maps["rmap2"], interpolation=cv2.INTER_NEAREST)
        return torch.tensor(left).permute(2, 0, 1), torch.tensor(right).permute(2, 0, 1)

    def get_rectified_calib(self):
        extrinsics = np.eye(4)
        if self.mode == "conventional":
            extrinsics[:3, 3] = [self.r_intr[0, 3] / self.r_intr[0, 0], 0.0, 0.0]        # P2[0,3] = Tx * f
        else:
            extrinsics[:3, 3] = self.cal["T"]
        bf = np.sqrt(np.sum(extrinsics[:3, 3] ** 2)) * self.l_intr[0, 0]
        return {"intrinsics": {"left": self.l_intr[:3, :3], "right": self.r_intr[:3, :3]}, "extrinsics": extrinsics, "bf": bf,
                "bf_orig": bf / self.scale, "img_size": self.img_size}
