"""CPU: the end-to-end oracle port (oracle/pipeline_ref.py) against the reference's own tracker outputs
(tests/golden/e2e_384x352.npz).  Needs the reference checkpoint (oracle/_ref/trained, git-ignored data)."""
import os

import numpy as np
import pytest
import torch

from oracle import pipeline_ref, se3_np
from oracle.detrand import unpack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")


@pytest.mark.skipif(not os.path.isfile(CKPT), reason="reference checkpoint not present")
def test_oracle_tracker_matches_reference(golden_dir):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = np.load(os.path.join(golden_dir, "e2e_384x352.npz"))
    W, H = [int(v) for v in g["size"]]
    sd = torch.load(CKPT, map_location="cpu", weights_only=False)["state_dict"]
    trk = pipeline_ref.RefTracker(sd, g["K"], float(g["bf"]))
    poses = []
    for i in range(3):
        poses.append(trk.step(torch.from_numpy(g["imgs_l"][i].astype(np.float32))[None],
                              torch.from_numpy(g["imgs_r"][i].astype(np.float32))[None],
                              torch.from_numpy(unpack(g["masks_in"][i], (1, 1, H, W)))))
    for k in range(3):
        d = se3_np.mul(se3_np.inv(poses[k].astype(np.float64)), g["traj"][k].astype(np.float64))
        assert np.linalg.norm(se3_np.log(d)[3:]) < 1e-6
        assert np.linalg.norm(poses[k][:3] - g["traj"][k][:3]) <= 1e-5 * max(1.0, np.linalg.norm(g["traj"][k][:3]))
    last = trk.last
    assert last["n_evals"] == len(g["pair1_eval_pose"])
    assert np.abs(last["time_flow"][0].numpy() - g["s_time_flow"]).max() < 1e-4
    assert np.abs(last["stereo_flow2"][0].numpy() - g["s_stereo_flow2"]).max() < 1e-4
    assert np.array_equal(np.packbits(last["mask2w"].numpy().reshape(-1)), g["s_mask2w"])
    assert np.array_equal(np.packbits(last["mask2_valid"].numpy().reshape(-1)), g["s_mask2_valid"])
    assert np.abs(last["conf1"][0].numpy() - g["s_conf1"].astype(np.float32)).max() < 1e-3
