"""CPU, world_size 2 over gloo: host-side sharding logic of the multi-GPU path (shard ranges, halo frames,
all_gather of per-pair records in global order, trajectory composition)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_records(a, b):
    """Deterministic per-pair record [pose7 | log6] as a function of the global pair index."""
    import rpe_b200  # noqa: F401
    from rpe_b200.lie import SE3
    idx = torch.arange(a, b, dtype=torch.float32)
    xi = torch.stack([0.01 * torch.sin(idx + k) for k in range(6)], 1)
    if a <= 3 < b:
        xi[3 - a, 0] = 0.3                                  # pair 3 trips the |log| > 0.1 guard
    X = SE3.exp(xi)
    return torch.cat((X.data, X.log()), 1)


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rpe_b200  # noqa: F401
    from rpe_b200 import parallel
    a, b = parallel.shard_ranges(n_pairs, world)[rank]
    rec = parallel.gather_pair_records(_fake_records(a, b), n_pairs)
    traj, failed = parallel.compose_trajectory(rec, [0, 0, 0, 0, 0, 0, 1.0], 250.0)
    np.save(os.path.join(out_dir, f"traj{rank}.npy"), traj.numpy())
    np.save(os.path.join(out_dir, f"failed{rank}.npy"), failed.numpy())
    dist.destroy_process_group()


def test_shard_ranges_cover_all_pairs():
    import rpe_b200  # noqa: F401
    from rpe_b200.parallel import frames_of, shard_ranges
    for n in (0, 1, 7, 64, 999):
        for world in (1, 2, 4, 8):
            r = shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert frames_of((8, 16)) == (8, 17) and frames_of((5, 5)) == (5, 5)


@pytest.mark.parametrize("n_pairs", [1, 9, 64])       # 1: the second rank holds an empty shard
def test_two_rank_gather_and_compose_equals_single_process(tmp_path, n_pairs):
    import rpe_b200  # noqa: F401
    from rpe_b200 import parallel
    mp.spawn(_worker, args=(2, _free_port(), n_pairs, str(tmp_path)), nprocs=2, join=True)
    ref_traj, ref_failed = parallel.compose_trajectory(_fake_records(0, n_pairs), [0, 0, 0, 0, 0, 0, 1.0], 250.0)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"traj{r}.npy"), ref_traj.numpy())
        assert np.array_equal(np.load(tmp_path / f"failed{r}.npy"), ref_failed.numpy())
    if n_pairs > 3:
        assert ref_failed.sum() == 1 and ref_failed[3]
        # the failed pair contributes identity: pose after pair 3 == pose after pair 2
        assert np.array_equal(ref_traj[4].numpy(), ref_traj[3].numpy())
    else:
        assert ref_traj.shape == (n_pairs + 1, 7) and not ref_failed.any()


def test_compose_matches_se3_class():
    import rpe_b200  # noqa: F401
    from rpe_b200 import parallel
    from rpe_b200.lie import SE3
    rec = _fake_records(0, 6)
    traj, failed = parallel.compose_trajectory(rec, [1.0, 2.0, 3.0, 0, 0, 0, 1.0], 250.0)
    last = SE3(torch.tensor([[1.0, 2.0, 3.0, 0, 0, 0, 1.0]]))
    for k in range(6):
        rel = SE3.Identity(1) if failed[k] else SE3(rec[k:k + 1, :7])
        last = last * rel.scale(250.0).inv()
        np.testing.assert_allclose(traj[k + 1].numpy(), last.data[0].numpy(), atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# The sharded public entry point itself (parallel.infer_sequence_sharded) fed with FRAMES: a stand-in estimator turns frame
# content into pair records on the CPU, so that halo frames, the sequence_start rule and the frame offsets of every shard are
# exercised without a GPU (the GPU test of the same call is tests/test_gpu_configs.py::test_sharded_halo_run_...).
# ---------------------------------------------------------------------------------------------------------------
class _FrameEstimator:
    """infer_pairs() -> record of pair k = f(frame k, frame k+1, mask rule of frame k): the first frame of a shard contributes
    its un-and-ed mask only when it is the first frame of the SEQUENCE."""

    def __init__(self):
        self.baseline = torch.zeros(1)
        self.scale = torch.tensor(1 / 250)
        from rpe_b200.lie import SE3
        self.last_pose = SE3.Identity(1)
        self.calls = []

    def infer_pairs(self, limgs, rimgs, masks, chunk=8, use_graphs=False, sequence_start=True):
        from rpe_b200.lie import SE3
        self.calls.append((int(limgs.shape[0]), bool(sequence_start)))
        f = limgs.float().mean((1, 2, 3)) + 0.5 * rimgs.float().mean((1, 2, 3))              # one scalar per frame
        m = masks.float().mean((1, 2, 3))
        m[1:] = m[1:] * 0.5                                                                 # "and-ed with stereo validity"
        if not sequence_start:
            m[0] = m[0] * 0.5                                                               # halo frame: and-ed like any other
        xi = torch.stack([1e-3 * (f[1:] - f[:-1]) * (k + 1) + 1e-3 * m[:-1] for k in range(6)], 1)
        X = SE3.exp(xi)
        return X.data, X.log(), torch.full((xi.shape[0],), 12.0)


def _frames(n):
    g = torch.Generator().manual_seed(5)
    return (torch.randint(0, 255, (n, 3, 8, 8), generator=g, dtype=torch.uint8), torch.randint(0, 255, (n, 3, 8, 8), generator=g, dtype=torch.uint8),
            torch.rand((n, 1, 8, 8), generator=g) > 0.2)


def _worker_frames(rank, world, port, n_frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rpe_b200  # noqa: F401
    from rpe_b200 import parallel
    L, R, M = _frames(n_frames)
    est = _FrameEstimator()
    loaded = []

    def load(a, b):
        loaded.append((a, b))
        return L[a:b], R[a:b], M[a:b]
    traj, failed = parallel.infer_sequence_sharded(est, load, n_frames, chunk=4)
    np.save(os.path.join(out_dir, f"ftraj{rank}.npy"), traj.numpy())
    meta = [loaded[0][0], loaded[0][1], est.calls[0][0], int(est.calls[0][1])] if loaded else [-1, -1, 0, 0]      # idle rank: no load, no call
    np.save(os.path.join(out_dir, f"fmeta{rank}.npy"), np.array(meta))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [2, 10, 65])     # 2 frames: one pair, the second rank has nothing to do
def test_sharded_entry_point_on_frames_equals_single_process(tmp_path, n_frames):
    import rpe_b200  # noqa: F401
    from rpe_b200 import parallel
    mp.spawn(_worker_frames, args=(2, _free_port(), n_frames, str(tmp_path)), nprocs=2, join=True)
    L, R, M = _frames(n_frames)
    ref, _ = parallel.infer_sequence_sharded(_FrameEstimator(), lambda a, b: (L[a:b], R[a:b], M[a:b]), n_frames, chunk=4)   # world size 1
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"ftraj{r}.npy"), ref.numpy())
    n_pairs = n_frames - 1
    (a0, b0), (a1, b1) = parallel.shard_ranges(n_pairs, 2)
    m0, m1 = np.load(tmp_path / "fmeta0.npy"), np.load(tmp_path / "fmeta1.npy")
    assert list(m0) == [a0, b0 + 1, b0 - a0 + 1, 1]                      # rank 0: frames [0, b0], sequence start
    if b1 > a1:
        assert list(m1) == [a1, b1 + 1, b1 - a1 + 1, 0]                  # rank 1: halo frame a1 = b0, NOT a sequence start
    else:
        assert list(m1) == [-1, -1, 0, 0]                                # empty shard: the rank only takes part in the gather
