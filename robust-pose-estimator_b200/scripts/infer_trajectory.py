"""infer_trajectory with the reference's ``main(args, config)`` signature
(/root/reference/scripts/infer_trajectory.py:23-116): runs the f2f tracker over a dataset and writes
``<outpath>/trajectory.freiburg``.  wandb logging and the Open3D viewers of the reference are outside the
pose path and are not provided (``--log`` / ``--viewer`` other than 'none' raise)."""
import os
import warnings

import torch
from torch.utils.data import DataLoader

from ..core.pose.pose_estimator import PoseEstimator
from ..core.utils.trajectory import read_freiburg, save_trajectory
from ..dataset.dataset_utils import SequentialSubSampler, get_data
from ..dataset.video_dataset import StereoVideoDataset
from ..lie import SE3


def main(args, config):
    if getattr(args, "device", "gpu") != "gpu" or not torch.cuda.is_available():
        raise RuntimeError("rpe_b200 runs on a CUDA device only (no CPU fallback)")
    device = torch.device("cuda")
    if getattr(args, "log", None) is not None:
        raise NotImplementedError("wandb logging is outside the pose path")
    if getattr(args, "viewer", "none") != "none":
        raise NotImplementedError("viewers are outside the pose path")
    if args.outpath is None:
        args.outpath = os.path.join(os.getcwd(), "infer_trajectory")
    os.makedirs(args.outpath, exist_ok=True)

    # file datasets: the loader worker only decodes (raw uint8 frames); mask, resize and rectification run on the device
    dataset, calib = get_data(args.input, config["img_size"], rect_mode=config.get("rect_mode", "conventional"),
                              force_video=getattr(args, "force_video", False), raw=True)
    is_video = isinstance(dataset, StereoVideoDataset)
    gt_file = os.path.join(args.input, "groundtruth.txt") if isinstance(args.input, str) else ""
    gt_trajectory = read_freiburg(gt_file) if os.path.isfile(gt_file) else None
    init_pose = gt_trajectory[None, args.start] if gt_trajectory is not None else SE3.Identity(1)

    estimator = PoseEstimator(config["slam"], torch.tensor(calib["intrinsics"]["left"]).to(device), baseline=calib["bf"],
                              checkpoint=args.checkpoint, img_shape=config["img_size"], init_pose=init_pose).to(device)
    if is_video:
        warnings.warn("start/stop arguments not supported for video dataset. ignored.", UserWarning)
        sampler = None
    else:
        sampler = SequentialSubSampler(dataset, args.start, args.stop, args.step)
    loader = DataLoader(dataset, num_workers=0 if config["slam"].get("debug", False) else 1, pin_memory=True, sampler=sampler)

    trajectory = [{"camera-pose": init_pose, "timestamp": args.start}]
    with torch.no_grad():
        for data in loader:
            if is_video:
                limg, rimg, _, img_number = data
                mask = None
            else:
                limg, rimg, mask, img_number = data
            limg, rimg = limg.to(device, non_blocking=True), rimg.to(device, non_blocking=True)
            mask = mask.to(device, non_blocking=True) if mask is not None else None
            if hasattr(dataset, "preprocess"):
                limg, rimg, mask = dataset.preprocess(limg, rimg, mask)
            pose, _, _, _ = estimator(limg, rimg, mask)
            stamp = img_number[0]
            trajectory.append({"camera-pose": pose, "timestamp": stamp if isinstance(stamp, str) else int(stamp)})
    failed = estimator.check_failures()
    if failed:
        warnings.warn(f"{len(failed)} pairs did not converge", RuntimeWarning)
    save_trajectory(trajectory, args.outpath)
    print("finished")
    return trajectory


def build_parser():
    import argparse
    p = argparse.ArgumentParser(description="script to run pose estimation")
    p.add_argument("input", type=str, help="Path to input folder, or synthetic:<frames>[:seed].")
    p.add_argument("--checkpoint", type=str, default="../trained/poseNet_2xf8up4b.pth")
    p.add_argument("--outpath", type=str)
    p.add_argument("--config", type=str, default="../configuration/infer_f2f.yaml")
    p.add_argument("--device", choices=["cpu", "gpu"], default="gpu")
    p.add_argument("--stop", type=int, default=10000000000)
    p.add_argument("--start", type=int, default=0)
    p.add_argument("--step", type=int, default=1)
    p.add_argument("--log", default=None)
    p.add_argument("--force_video", action="store_true")
    p.add_argument("--viewer", default="none", choices=["none", "2d", "3d", "video"])
    p.add_argument("--block_viewer", action="store_true")
    return p


if __name__ == "__main__":
    import yaml
    a = build_parser().parse_args()
    with open(a.config, "r") as f:
        cfg = yaml.load(f, Loader=yaml.SafeLoader)
    assert os.path.isfile(a.checkpoint), "no valid checkpoint file"
    main(a, cfg)
