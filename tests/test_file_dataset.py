"""File datasets of the input pipeline (reference dataset/dataset_utils.py:10-35, stereo_dataset.py:19-44,
video_dataset.py:14-78, rectification.py:10-184) against tests/golden/file_dataset.npz, which oracle/make_golden.py --filedata
produced by running the UNMODIFIED reference's ``get_data`` on the fixtures of oracle/file_fixture.py.

CPU: calibration parsing + rectified calibration for the three file formats, folder discovery, decoding.
GPU: what the datasets yield (device specularity mask + resize/crop + rectification remap) against the reference's items."""
import hashlib
import os
import types

import numpy as np
import pytest

from oracle import file_fixture as ff
from oracle.detrand import unpack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
VIDEO_DIR = os.path.join(GOLDEN, "file_video")
CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
W, H = ff.IMG_SIZE


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "file_dataset.npz"))


@pytest.fixture(scope="module")
def frame_folder(tmp_path_factory):
    return ff.write_frame_folder(str(tmp_path_factory.mktemp("frames")))


def _check_calib(calib, g, prefix):
    np.testing.assert_allclose(calib["intrinsics"]["left"], g[prefix + "K_left"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(calib["intrinsics"]["right"], g[prefix + "K_right"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(calib["extrinsics"], g[prefix + "extrinsics"], rtol=1e-12, atol=1e-12)
    assert abs(calib["bf"] - float(g[prefix + "bf"])) <= 1e-12 * abs(float(g[prefix + "bf"]))
    assert abs(calib["bf_orig"] - float(g[prefix + "bf_orig"])) <= 1e-12 * abs(float(g[prefix + "bf_orig"]))
    assert tuple(calib["img_size"]) == tuple(int(v) for v in g[prefix + "img_size"])


def _video_decodes_like_the_golden(g):
    import cv2
    cap = cv2.VideoCapture(os.path.join(VIDEO_DIR, "video.mp4"))
    sha = hashlib.sha1()
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        sha.update(fr.tobytes())
    return sha.hexdigest() == str(g["video_decoded_sha1"])


# ---------------------------------------------------------------------------------------------------------------
# CPU: host logic
# ---------------------------------------------------------------------------------------------------------------
def test_rectified_calibration_of_the_three_formats(golden, frame_folder, tmp_path):
    import rpe_b200  # noqa: F401
    from rpe_b200.dataset.rectification import StereoRectifier, find_calibration_file
    ini = find_calibration_file(frame_folder)
    assert ini.endswith("StereoCalibration.ini")
    _check_calib(StereoRectifier(ini, img_size_new=ff.IMG_SIZE).get_rectified_calib(), golden, "frames_")
    yaml_file = find_calibration_file(VIDEO_DIR)
    assert yaml_file.endswith("endoscope_calibration.yaml")
    for mode in ("conventional", "pseudo"):
        with pytest.warns(UserWarning) if mode == "pseudo" else _nullcontext():
            _check_calib(StereoRectifier(yaml_file, img_size_new=ff.IMG_SIZE, mode=mode).get_rectified_calib(), golden, f"video_{mode}_")
            json_file = ff.write_json_calibration(str(tmp_path / "json"))
            _check_calib(StereoRectifier(json_file, img_size_new=ff.IMG_SIZE, mode=mode).get_rectified_calib(), golden, f"json_{mode}_")
    with pytest.raises(RuntimeError):
        find_calibration_file(str(tmp_path))


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def test_get_data_resolves_folders_and_decodes(golden, frame_folder):
    import rpe_b200  # noqa: F401
    from rpe_b200.dataset.dataset_utils import get_data
    from rpe_b200.dataset.stereo_dataset import StereoDataset
    from rpe_b200.dataset.video_dataset import StereoVideoDataset
    ds, calib = get_data(frame_folder, ff.IMG_SIZE, raw=True)
    assert isinstance(ds, StereoDataset) and len(ds) == ff.N_FRAMES
    _check_calib(calib, golden, "frames_")
    left, right, mask, number = ds[2]
    assert number == "000002" and left.dtype.is_floating_point is False and tuple(left.shape) == (3, ff.ORIG_H, ff.ORIG_W)
    assert np.array_equal(left.permute(1, 2, 0).numpy(), ff._frame(1020)) and np.array_equal(right.permute(1, 2, 0).numpy(), ff._frame(1025))
    assert tuple(mask.shape) == (1, ff.ORIG_H, ff.ORIG_W) and set(np.unique(mask.numpy())) <= {0, 1} and 0 < mask.sum() < mask.numel()
    vd, vcalib = get_data(VIDEO_DIR, ff.IMG_SIZE, raw=True)
    assert isinstance(vd, StereoVideoDataset) and len(vd) == int(golden["video_conventional_len"])
    _check_calib(vcalib, golden, "video_conventional_")
    items = list(vd)
    assert len(items) == 4                                   # 5 frames, 4 ground-truth poses: the sequence ends with its ground truth
    assert [it[3] for it in items] == list(golden["video_conventional_number"])
    np.testing.assert_allclose(np.stack([it[2].numpy() for it in items]), golden["video_conventional_pose"], rtol=1e-12, atol=1e-12)
    assert tuple(items[0][0].shape) == (3, ff.ORIG_H, ff.ORIG_W)
    forced, _ = get_data(VIDEO_DIR, ff.IMG_SIZE, force_video=True, raw=True)
    assert isinstance(forced, StereoVideoDataset)
    with pytest.raises(RuntimeError):
        get_data(os.path.join(GOLDEN), ff.IMG_SIZE)          # no calibration file


# ---------------------------------------------------------------------------------------------------------------
# GPU: device preprocessing against the reference's dataset items
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_stereo_dataset_items_match_the_reference(golden, frame_folder):
    import torch
    import rpe_b200  # noqa: F401
    from rpe_b200.dataset.dataset_utils import get_data
    ds, _ = get_data(frame_folder, ff.IMG_SIZE)
    want_mask = unpack(golden["frames_mask"], (ff.N_FRAMES, 1, H, W))
    for k in range(ff.N_FRAMES):
        left, right, mask, number = ds[k]
        assert left.is_cuda and left.dtype == torch.float32 and tuple(left.shape) == (3, H, W) and mask.dtype == torch.bool
        assert number == str(golden["frames_number"][k])
        # anti-aliased resize: within 2e-3 on the 0..255 scale of the reference's torchvision path; masks bit-exact
        assert np.abs(left.cpu().numpy()[..., 1::3, 2::3] - golden["frames_left"][k]).max() < 2e-3
        assert np.abs(right.cpu().numpy()[..., 1::3, 2::3] - golden["frames_right"][k]).max() < 2e-3
        assert np.array_equal(mask.cpu().numpy(), want_mask[k])
    assert 0.3 < want_mask.mean() < 0.98                       # the fixture exercises tool mask, highlights and erosion


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["conventional", "pseudo"])
def test_video_dataset_items_match_the_reference(golden, mode):
    import warnings
    import rpe_b200  # noqa: F401
    from rpe_b200.dataset.dataset_utils import get_data
    if not _video_decodes_like_the_golden(golden):
        pytest.skip("this OpenCV build decodes the committed mp4 differently from the one the golden was made with")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ds, _ = get_data(VIDEO_DIR, ff.IMG_SIZE, rect_mode=mode)
        items = list(ds)
    assert len(items) == 4
    want_mask = unpack(golden[f"video_{mode}_mask"], (4, 1, H, W))
    for k, (left, right, mask, pose, number) in enumerate(items):
        assert number == str(golden[f"video_{mode}_number"][k])
        np.testing.assert_allclose(pose.numpy(), golden[f"video_{mode}_pose"][k], rtol=1e-12, atol=1e-12)
        assert np.abs(left.cpu().numpy()[..., 1::3, 2::3] - golden[f"video_{mode}_left"][k]).max() < 2e-3
        assert np.abs(right.cpu().numpy()[..., 1::3, 2::3] - golden[f"video_{mode}_right"][k]).max() < 2e-3
        assert np.array_equal(mask.cpu().numpy(), want_mask[k])


@pytest.mark.gpu
def test_infer_trajectory_main_on_a_frame_folder(frame_folder, tmp_path):
    """the CLI entry point end to end on a real folder: worker decodes, device preprocesses, one line per frame, frame ids kept"""
    if not os.path.isfile(CKPT):
        pytest.skip("reference checkpoint not shipped (oracle/_ref/trained is created by oracle/make_golden.py)")
    import yaml
    import rpe_b200  # noqa: F401
    from rpe_b200.scripts import infer_trajectory
    with open(os.path.join(ROOT, "robust-pose-estimator_b200", "configuration", "infer_f2f_nw.yaml")) as f:
        config = yaml.load(f, Loader=yaml.SafeLoader)
    config["img_size"] = list(ff.IMG_SIZE)
    args = types.SimpleNamespace(input=frame_folder, checkpoint=CKPT, outpath=str(tmp_path), device="gpu", start=0, stop=10000000000,
                                 step=1, log=None, force_video=False, viewer="none", block_viewer=False)
    trajectory = infer_trajectory.main(args, config)
    assert len(trajectory) == ff.N_FRAMES + 1
    lines = open(tmp_path / "trajectory.freiburg").read().splitlines()
    assert [ln.split(" ")[0] for ln in lines] == ["0", "000000", "000001", "000002", "000003"]
    assert np.isfinite(np.array([[float(v) for v in ln.split(" ")[1:]] for ln in lines])).all()
