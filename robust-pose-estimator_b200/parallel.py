"""Multi-GPU sharding of the f2f path: one process per GPU, contiguous frame ranges, no data-path
collective; NCCL (or gloo in the CPU tests) is used only to gather the per-pair poses.

Pair k depends only on frames k and k+1 (the reference overwrites ``last_frame`` every call and a failed pair
contributes identity, pose_estimator.py:62,81-85), so rank r owns pairs [r*P, (r+1)*P) and frames
[r*P, (r+1)*P] (one halo frame).  Exactness caveat handled here: only the first frame of the SEQUENCE keeps its
input mask; the first frame of every other shard is and-ed with its stereo validity like any non-initial frame
(SURVEY.md A.6)."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def shard_ranges(n_pairs, world):
    """[(first_pair, last_pair_exclusive)] per rank: contiguous, sizes differ by at most one, earlier ranks larger."""
    base, extra = divmod(n_pairs, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < extra else 0)
        out.append((a, b))
        a = b
    return out


def frames_of(pair_range):
    """Frame indices a rank must load for its pair range (pairs [a,b) need frames [a, b])."""
    a, b = pair_range
    return (a, b + 1) if b > a else (a, a)


class PendingRecords:
    """Handle of an in-flight ``gather_pair_records(..., async_op=True)``: the collective runs on the backend's own stream
    while the caller keeps launching the next sequence; ``wait()`` orders the current stream after it and returns the
    (n_pairs, D) tensor in global pair order."""

    def __init__(self, work, bufs, ranges, local):
        self._work, self._bufs, self._ranges, self._local = work, bufs, ranges, local

    def wait(self):
        if self._work is None:
            return self._local
        self._work.wait()
        return torch.cat([self._bufs[r][:b - a] for r, (a, b) in enumerate(self._ranges)], 0)


def gather_pair_records(local, n_pairs, group=None, async_op=False):
    """all_gather of per-pair records -- the only exchange step of the path.  local: (P_r, D) tensor of this rank's pairs
    (P_r from shard_ranges); returns the (n_pairs, D) tensor in global pair order on every rank, or with ``async_op`` a
    ``PendingRecords`` handle (the gather then overlaps whatever the caller launches next)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return PendingRecords(None, None, None, local) if async_op else local
    ranges = shard_ranges(n_pairs, world)
    pmax = max(b - a for a, b in ranges)
    padded = torch.zeros((pmax, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    work = dist.all_gather(bufs, padded, group=group, async_op=True)
    pending = PendingRecords(work, bufs, ranges, local)
    return pending if async_op else pending.wait()


def compose_trajectory(records, init_pose, inv_scale):
    """Host composition of the gathered (n,13) [pose7 | log6] records through rpe_compose_trajectory_host.
    -> (abs poses (n+1,7) float32 CPU, failed (n,) bool)."""
    rec = records.detach().float().cpu().contiguous()
    n = rec.shape[0]
    rel, log = rec[:, :7].contiguous(), rec[:, 7:13].contiguous()
    init = torch.as_tensor(init_pose, dtype=torch.float32).reshape(7).contiguous()
    out = torch.empty((n + 1, 7), dtype=torch.float32)
    failed = torch.zeros((max(n, 1),), dtype=torch.uint8)
    _lib.check(_lib.lib().rpe_compose_trajectory_host(C.c_void_p(rel.data_ptr()), C.c_void_p(log.data_ptr()), n,
                                                      C.c_void_p(init.data_ptr()), float(inv_scale),
                                                      C.c_void_p(out.data_ptr()), C.c_void_p(failed.data_ptr())),
               "rpe_compose_trajectory_host")
    return out, failed[:n].bool()


def infer_sequence_sharded(estimator, load_frames, n_frames, chunk=8, use_graphs=False, group=None):
    """Sharded ``PoseEstimator.infer_sequence``: every rank calls this with a ``load_frames(a, b) -> (limgs, rimgs, masks)``
    callback returning frames [a, b) of the GLOBAL sequence, either device tensors or (pinned) host tensors -- host
    frames are uploaded chunk by chunk behind the compute, like ``infer_sequence`` does.  One all_gather of the (P,13) pair
    records, then every rank composes the trajectory on its host.  Returns (trajectory (n_frames,7) float32 CPU, failed)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_pairs = n_frames - 1
    pr = shard_ranges(n_pairs, world)[rank]
    fa, fb = frames_of(pr)
    if fb > fa:
        limgs, rimgs, masks = load_frames(fa, fb)
        rel, log, _ = estimator.infer_pairs(limgs, rimgs, masks, chunk=chunk, use_graphs=use_graphs, sequence_start=(pr[0] == 0))
        local = torch.cat((rel, log), 1)
    else:
        local = torch.zeros((0, 13), device=estimator.baseline.device)
    rec = gather_pair_records(local, n_pairs, group)
    inv_scale = float((1 / estimator.scale).float().cpu())
    return compose_trajectory(rec, estimator.last_pose.data.reshape(7).float().cpu(), inv_scale)
