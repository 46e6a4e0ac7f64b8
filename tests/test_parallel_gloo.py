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


@pytest.mark.parametrize("n_pairs", [9, 64])
def test_two_rank_gather_and_compose_equals_single_process(tmp_path, n_pairs):
    import rpe_b200  # noqa: F401
    from rpe_b200 import parallel
    mp.spawn(_worker, args=(2, _free_port(), n_pairs, str(tmp_path)), nprocs=2, join=True)
    ref_traj, ref_failed = parallel.compose_trajectory(_fake_records(0, n_pairs), [0, 0, 0, 0, 0, 0, 1.0], 250.0)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"traj{r}.npy"), ref_traj.numpy())
        assert np.array_equal(np.load(tmp_path / f"failed{r}.npy"), ref_failed.numpy())
    assert ref_failed.sum() == 1 and ref_failed[3]
    # the failed pair contributes identity: pose after pair 3 == pose after pair 2
    assert np.array_equal(ref_traj[4].numpy(), ref_traj[3].numpy())


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
