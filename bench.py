#!/usr/bin/env python
"""Benchmark of the f2f pose path (BASELINE.json metric: f2f frame-pairs/sec @640x512 stereo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp16x3|fp32|tf32|fp16|bf16]

A step = one pass of the hot path over one batch of synthetic input: a 65-frame 640x512 stereo sequence
(64 frame pairs) per rank (weak scaling; pairs are independent in f2f).
  value  pairs/s with the frames resident in HBM (uint8 RGB as decoded), device-timed.
  e2e    pairs/s through PoseEstimator.infer_sequence with HOST buffers: pinned uint8 frames copied to the device
         and the poses copied back + composed into the trajectory inside the timed region.
For N > 1 (torchrun) each rank owns a contiguous shard; the per-pair poses are all-gathered over NCCL inside the
step and rank 0 composes the trajectory on the host.  One JSON line is printed by rank 0.
`--impl reference` times the CPU port of the reference path (oracle/pipeline_ref.py; the reference itself is
Python + an uninstallable third-party lietorch and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CKPT = os.path.join(ROOT, "oracle", "_ref", "trained", "poseNet_2xf8up4b.pth")
SLAM = {"frame2frame": True, "dist_thr": 0.05, "depth_clipping": [1, 250], "debug": False, "conf_weighing": True,
        "average_pts": False, "lbgfs_iters": 20}
W_IMG, H_IMG = 640, 512
METRIC = "f2f_frame_pairs_per_sec_640x512"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def _rows(self):
        try:
            return [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        except OSError:
            return []

    def mark(self):
        """Start of the timed region: nvidia-smi was started earlier (its start-up can exceed a short timed region on a
        multi-GPU box); only the rows written from here on are reported."""
        self.begin = len(self._rows())

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = self._rows()
        os.unlink(self.f.name)
        begin = getattr(self, "begin", 0)
        if len(rows) > begin:
            rows = rows[begin:]
        elif rows:                                   # no sample landed inside a very short timed region: last warm-up samples
            rows = rows[-3:]
            out["window"] = "last warm-up steps (same load); the timed region was shorter than the sampling period"
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except (ValueError, IndexError):
                continue
            for k, n in enumerate(names):
                if len(r) > 5 + k and r[5 + k].strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def base_frames(rank=0, world=1, dist=None):
    """The 65 rendered base frames (uint8 L, R, bool M) of ``bench_sequence()`` -- the frames the committed 64-pair reference
    golden (tests/golden/bench64_poses.npz) was produced from -- rendered once per box by a pool of host processes and
    cached under the temp dir.  Under torchrun rank 0 renders, the other ranks read its cache."""
    import rpe_b200  # noqa: F401
    from rpe_b200.dataset.synthetic import bench_sequence
    seq = bench_sequence((W_IMG, H_IMG))
    cache = tempfile.gettempdir()
    if world > 1 and rank != 0:
        dist.barrier()
        return (*seq.frames_u8(cache_dir=cache), seq)
    out = seq.frames_u8(cache_dir=cache)
    if world > 1:
        dist.barrier()
    return (*out, seq)


def frame_indices(a, b):
    """Base-frame indices of frames [a, b) of the (arbitrarily long) bench sequence: the 65 base frames walked back and forth."""
    from rpe_b200.dataset.synthetic import triangle_index
    return [triangle_index(i) for i in range(a, b)]


def _ref_tracker(seq, device="cpu", gpu=False, conf=True):
    import torch
    from oracle import pipeline_ref
    return pipeline_ref.RefTracker(_state_dict(torch), seq.calib["intrinsics"]["left"], seq.calib["bf"], conf_weighing=conf,
                                   device=device, autocast=gpu, solver="torch" if gpu else "numpy")


def _time_tracker(trk, L, R, M, warmup, steps, sync=None):
    """Feed frames 0..warmup+steps to a RefTracker; wall-clock the last `steps` pairs."""
    import torch
    f = lambda a: torch.from_numpy(a.astype(np.float32))[None]
    for i in range(0, 1 + warmup):                                              # frame 0: stereo depth only
        trk.step(f(L[i]), f(R[i]), torch.from_numpy(M[i])[None])
    for k in trk.timing:
        trk.timing[k] = 0.0
    if sync:
        sync()
    t0 = time.perf_counter()
    for i in range(1 + warmup, 1 + warmup + steps):
        trk.step(f(L[i]), f(R[i]), torch.from_numpy(M[i])[None])
    if sync:
        sync()
    return time.perf_counter() - t0


def run_reference(args, rank):
    """Reference arm: the CPU port of the reference's f2f path on the host cores (oracle/pipeline_ref.py), or with
    ``--impl reference-gpu`` the same port executed the way the stock reference runs on a GPU (torch cuDNN kernels under fp16
    autocast, batch 1, torch.optim.LBFGS with a host sync per iteration)."""
    if rank != 0:
        return
    import torch
    gpu = args.impl == "reference-gpu"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bL, bR, bM, seq = base_frames()
    idx = frame_indices(0, args.steps + args.warmup + 1)
    L, R, M = bL[idx], bR[idx], bM[idx]
    if gpu:
        if not torch.cuda.is_available():
            _emit({"impl": "reference-gpu", "unavailable": "no CUDA device"})
            return
        trk = _ref_tracker(seq, "cuda", gpu=True)
        dt = _time_tracker(trk, L, R, M, args.warmup, args.steps, torch.cuda.synchronize)
    else:
        trk = _ref_tracker(seq)
        dt = _time_tracker(trk, L, R, M, args.warmup, args.steps)
    val = args.steps / dt
    what = ("oracle/pipeline_ref.py on cuda:0 (torch cuDNN/cuBLAS/ATen, fp16 autocast, batch 1, torch.optim.LBFGS + host sync per iteration)"
            if gpu else "oracle/pipeline_ref.py (torch CPU fp32 + numpy fp64 L-BFGS)")
    line = {"impl": args.impl, "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 autocast + f64 solver" if gpu else "f32", "data": "synthetic",
            "config": {"workload": "f2f_640x512_seq65", "sample": "1 frame pair per step (bounded sample of the 64-pair batch)",
                       "lbgfs_iters": 20, "conf_weighing": True},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} consecutive pairs, {what}"},
            "split_s_per_pair": {k: v / args.steps for k, v in trk.timing.items()},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def _state_dict(torch):
    if os.path.isfile(CKPT):
        return torch.load(CKPT, map_location="cpu", weights_only=False)["state_dict"]
    import rpe_b200  # noqa: F401
    from rpe_b200.core.pose.pose_net import PoseNet
    torch.manual_seed(0)
    return PoseNet({"image_shape": (H_IMG, W_IMG), "use_weights": True, "lbgfs_iters": 20, "small": False,
                    "dropout": 0.0}).state_dict()


def cpu_baseline(L, R, M, seq, budget_pairs=8):          # ~10-15 s of host work on 16 cores
    """The oracle port timed on the host cores over a bounded sample of the same workload."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    trk = _ref_tracker(seq)
    dt = _time_tracker(trk, L, R, M, 1, budget_pairs)
    return {"value": budget_pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{budget_pairs} pairs of the same sequence after 1 warm-up pair ({dt:.1f} s)",
            "split_s_per_pair": {k: v / budget_pairs for k, v in trk.timing.items()}}


def gpu_reference(L, R, M, seq, pairs=10, warmup=3):
    """The reference's GPU behaviour (SURVEY D5 / BASELINE target ">= 50x the reference's single-GPU PyTorch f2f throughput"):
    the oracle port on cuda:0 with torch's own kernels, fp16 autocast, batch 1 per frame and the host-synchronised
    torch.optim.LBFGS.  A reported baseline; nothing of the product runs in it."""
    import torch
    trk = _ref_tracker(seq, "cuda", gpu=True)
    dt = _time_tracker(trk, L, R, M, warmup, pairs, torch.cuda.synchronize)
    return {"value": pairs / dt, "unit": "pairs/s", "kind": "port on the GPU (oracle/pipeline_ref.py, device=cuda)",
            "how": "torch cuDNN/cuBLAS/ATen kernels, fp16 autocast around fnet / cnet / update block (raft.py:92,100,117), batch 1, "
                   "torch.optim.LBFGS over the pure-torch lietorch stand-in with float(loss) host sync per iteration (pose_head.py:60-79)",
            "sample": f"{pairs} pairs after {warmup} warm-up pairs ({dt:.2f} s)",
            "split_s_per_pair": {k: v / pairs for k, v in trk.timing.items()}}


ALG_BYTES = {
    # algorithmic bytes per unit (SURVEY.md section 8d / DESIGN.md): unit = one RAFT sample or one image / pair
    "corr_lookup": 5120 * 4 * 100 * 4 + 5120 * 8 + 5120 * 324 * 4,            # 14.87 MB per sample and iteration
    "corr_build": 5120 * 5120 * 4 * (1 + 1 / 4 + 1 / 16 + 1 / 64) + 2 * 5120 * 256 * 4,   # pyramid write + fmaps read
    "warp8_mask": 74 * H_IMG * W_IMG,
    "depth_proj": 25 * H_IMG * W_IMG,
    "proj": 16 * H_IMG * W_IMG,
    "convex_upsample8": 5120 * (576 + 2) * 4 + 2 * H_IMG * W_IMG * 4,
    "pose_solve": 42 * H_IMG * W_IMG,                                            # per objective evaluation
}


def _emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else that libraries print to fd 1 (the NCCL
    version banner, torchrun notices) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


_REAL_STDOUT = None


def _timed_steps(fn, steps, sync_all, torch):
    """K calls of `fn` between two CUDA events on the current stream, barrier + synchronize on both sides -> (ms, last result)."""
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(steps):
        last = fn()
    e1.record()
    sync_all()
    return e0.elapsed_time(e1), last


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                  # fd 1 -> stderr for the whole run (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): 64 pairs per GPU per step; strong: BASELINE config 3 as worded, 64 pairs per step in total, "
                         "64/N per GPU.  With N > 1 the other mode is always measured too and reported under 'strong' / 'weak'.")
    ap.add_argument("--precision", default="fp16x3", choices=["fp32", "fp16x3", "tf32", "fp16", "bf16"],
                    help="fp16x3 (default): whole trunk on the hand-written tcgen05 kernels, fp32-equivalent split arithmetic "
                         "(parity-gated); fp32: cuDNN fp32 trunk; tf32 / fp16 / bf16: cuDNN reduced precision (not parity-gated)")
    ap.add_argument("--chunk", type=int, default=32, help="frames per engine chunk (measured on B200: 11 -> 318, 22 -> 330, 32 -> 334 pairs/s; fewer, larger launches)")
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--graphs", action="store_true")
    ap.add_argument("--pose-groups", type=int, default=0, help="concurrently solved pairs in rpe_pose_solve (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference-on-GPU leg (N = 1 only)")
    ap.add_argument("--latency-pairs", type=int, default=200, help="batch-1 latency leg (config 2): pairs timed one by one after 20 warm-ups; 0 = off")
    ap.add_argument("--config5-frames", type=int, default=1000, help="infer_f2f_nw end-to-end leg (config 5): frames of the sequence; 0 = off")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl != "b200":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib, build, ops, parallel
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
    from rpe_b200.core.utils.trajectory import save_trajectory_array
    from rpe_b200.engine import F2FEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (rpe_b200 has no CPU path); use --impl reference for the CPU port")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (the NCCL banner goes to stdout)
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    lib = _lib.lib()
    if args.pose_groups > 0:
        _lib.check(lib.rpe_pose_set_groups(args.pose_groups), "rpe_pose_set_groups")

    bL, bR, bM, seq = base_frames(rank, world, dist)
    trained = os.path.isfile(CKPT)
    cfg = dict(SLAM, precision=args.precision)
    K_t = torch.tensor(seq.calib["intrinsics"]["left"])
    est = PoseEstimator(cfg, K_t, seq.calib["bf"], CKPT if trained else None, (W_IMG, H_IMG)).to(dev)
    inv_scale = float((1 / est.scale).float().cpu())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Shard:
        """This rank's frames of a `total_pairs`-pair sequence: pinned uint8 host copies and uint8 device copies (the kernels
        that read the image convert on load)."""

        def __init__(self, total_pairs):
            self.total = total_pairs
            self.a, self.b = parallel.shard_ranges(total_pairs, world)[rank]
            fa, fb = parallel.frames_of((self.a, self.b))
            idx = frame_indices(fa, fb)
            self.h = [torch.from_numpy(np.ascontiguousarray(x[idx])).pin_memory() for x in (bL, bR, bM)]
            self.d = [t.to(dev) for t in self.h]
            self.h2d = sum(t.numel() for t in self.h)
            self.pairs = self.b - self.a
            self.pending = None

        def load_host(self, fa, fb):                     # frames [fa, fb) of the global sequence (this rank's shard only)
            o = parallel.frames_of((self.a, self.b))[0]
            return tuple(t[fa - o: fb - o] for t in self.h)

    def mode_runner(total_pairs, chunk):
        sh = Shard(total_pairs)
        engine = F2FEngine(est, chunk=chunk, use_graphs=args.graphs)

        def device_step():
            """Frames resident in HBM.  The pose records of step k are all-gathered asynchronously (the only exchange step);
            the gather of step k-1 is awaited here, so ranks never wait for each other inside a step."""
            engine.reset()
            rel, log, evals = engine.infer_sequence(*sh.d, sequence_start=(sh.a == 0))
            if sh.pending is not None:
                sh.pending.wait()
            sh.pending = parallel.gather_pair_records(torch.cat((rel, log), 1).contiguous(), sh.total, async_op=True)
            return evals

        def finish():
            rec = sh.pending.wait() if sh.pending is not None else None
            sh.pending = None
            return rec

        def e2e_step():
            """The public API fed from pinned HOST frames: H2D copies, solve, (all-gather,) D2H, host composition."""
            if world == 1:
                est.last_pose = est.last_pose.__class__.Identity(1, device=dev)
                return est.infer_sequence(*sh.h, chunk=chunk, use_graphs=args.graphs)
            return parallel.infer_sequence_sharded(est, sh.load_host, sh.total + 1, chunk=chunk, use_graphs=args.graphs)

        return sh, device_step, finish, e2e_step

    def measure(total_pairs, chunk, with_timers):
        sh, device_step, finish, e2e_step = mode_runner(total_pairs, chunk)
        for _ in range(args.warmup):
            device_step()
        finish()
        timers = ops.enable_timers(True) if with_timers else None
        launches0 = lib.rpe_launch_count()
        if with_timers:
            sampler.mark()

        def timed():
            ev = device_step()
            return ev
        ms, evals = _timed_steps(timed, args.steps, sync_all, torch)
        # the last step's gather is awaited inside the timed region's closing synchronize (sync_all synchronises the device)
        rec = finish()
        clocks = sampler.stop() if with_timers else None
        launches = lib.rpe_launch_count() - launches0
        stage = {}
        if with_timers:
            for name, evs in timers.items():
                t = [a.elapsed_time(b) for a, b, _ in evs]
                units = sum(u for _, _, u in evs)
                stage[name] = {"launches": len(evs), "total_ms": float(np.sum(t)), "avg_us": 1e3 * float(np.mean(t)),
                               "min_us": 1e3 * float(np.min(t)), "max_us": 1e3 * float(np.max(t)), "units": units}
            ops.enable_timers(False)
        for _ in range(max(1, min(2, args.warmup))):
            e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            traj, failed = e2e_step()
        sync_all()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"ms": float(t[0]), "e2e_ms": float(t[1]), "evals": float(evals.sum().item()), "pairs_rank": sh.pairs,
                "total": total_pairs, "h2d": sh.h2d, "failed": int(failed.sum()), "stage": stage, "clocks": clocks,
                "launches": int(launches), "rec": rec, "shard": sh, "chunk": chunk}

    sampler = ClockSampler(local_rank)
    sampler.start()                                  # nvidia-smi polls from the warm-up on; mark() opens the reported window
    weak_total, strong_total = world * args.pairs, args.pairs
    primary_total = weak_total if args.scaling == "weak" else strong_total
    # strong mode: 64 / N pairs per GPU -> one engine chunk holds the whole shard
    chunk_of = lambda total: min(args.chunk, max(1, -(-total // world)))
    main_res = measure(primary_total, chunk_of(primary_total), True)
    other = None
    if world > 1:
        other_total = strong_total if args.scaling == "weak" else weak_total
        other = measure(other_total, chunk_of(other_total), False)
    ms, e2e_ms, stage, clocks = main_res["ms"], main_res["e2e_ms"], main_res["stage"], main_res["clocks"]
    evals_total = main_res["evals"]
    sh = main_res["shard"]
    dL, dR, dM = sh.d

    # ---- pose parity of the timed inputs against the committed reference golden (rank 0 holds pairs 0.. of the sequence)
    parity = None
    gpath = os.path.join(ROOT, "tests", "golden", "bench64_poses.npz")
    if rank == 0 and trained and os.path.isfile(gpath) and main_res["rec"] is not None:
        g = np.load(gpath)
        rec = main_res["rec"][: min(64, main_res["total"])].double().cpu().numpy()
        ref = g["rel_pose"][: rec.shape[0]].astype(np.float64)
        tr = np.linalg.norm(rec[:, :3] - ref[:, :3], axis=1) / np.maximum(np.linalg.norm(ref[:, :3], axis=1), 1e-12)
        sgn = np.sign(np.sum(rec[:, 3:7] * ref[:, 3:7], axis=1))[:, None]
        rot = 2 * np.linalg.norm(sgn * rec[:, 3:7] - ref[:, 3:7], axis=1)      # angle between two nearby unit quaternions
        parity = {"pairs": int(rec.shape[0]), "max_rel_translation": float(tr.max()), "max_rotation_rad": float(rot.max()),
                  "gate": 1e-4, "ok": bool(tr.max() < 1e-4 and rot.max() < 1e-4),
                  "golden": "tests/golden/bench64_poses.npz (the unmodified reference on the same 65 frames, CPU fp32)"}

    # ---- GN-mode residual reduction (BASELINE metric "GN GB/s") and the L-BFGS one on identical inputs, one chunk of pairs
    gn = None
    if rank == 0:
        probe = F2FEngine(est, chunk=min(args.chunk, 32), use_graphs=False)
        probe.keep_solve_inputs = True
        probe.infer_sequence(dL[: min(33, dL.shape[0])], dR[: min(33, dL.shape[0])], dM[: min(33, dL.shape[0])])
        a = probe.last_solve_inputs
        gn = {}
        for name, mode, iters, hess in (("lbfgs_ref", ops.SOLVER_LBFGS_REF, 20, False), ("gn", ops.SOLVER_GN, 10, True)):
            for _ in range(2):
                sol = ops.pose_solve(*a, mode=mode, max_iter=iters, with_hessian=hess)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                sol = ops.pose_solve(*a, mode=mode, max_iter=iters, with_hessian=hess)
            e1.record()
            torch.cuda.synchronize()
            ev = float(sol.n_evals.sum().item())
            secs = e0.elapsed_time(e1) * 1e-3 / 5
            gn[name] = {"pairs": int(a[0].shape[0]), "evals_per_pair": ev / a[0].shape[0], "ms": secs * 1e3,
                        "GBps": 42 * H_IMG * W_IMG * ev / secs / 1e9}
        del probe

    # ---- latency path (BASELINE config 2: batch 1, the per-frame tracker call of the reference API), rank 0, outside `value`
    latency = None
    if rank == 0 and args.latency_pairs > 0:
        nwarm = 20
        order = frame_indices(0, args.latency_pairs + nwarm + 1)
        fL, fR, fM = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (bL, bR, bM))
        fL, fR = fL.float(), fR.float()                        # the per-frame API takes the reference's float32 0..255 tensors

        def run_leg():
            est.frame = est.last_frame = None
            est.failure_flags = []
            evs = []
            for k in order:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                est(fL[k:k + 1], fR[k:k + 1], fM[k:k + 1].clone())
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            lat = np.array([a.elapsed_time(b) for a, b in evs[1 + nwarm:]])      # frame 0 has no pair; 20 warm-up pairs
            return {"p50_ms_per_pair": float(np.percentile(lat, 50)), "p90_ms_per_pair": float(np.percentile(lat, 90)),
                    "pairs": int(lat.size), "warmup_pairs": nwarm}
        eager = dict(run_leg(), call="PoseEstimator.forward, batch 1, device-resident frame, CUDA events")
        latency = dict(eager)
        try:
            est.config = dict(est.config, cuda_graph=True)     # the per-frame device work replayed as one captured CUDA graph
            latency = dict(run_leg(), call="PoseEstimator.forward with config cuda_graph=True, batch 1, device-resident frame, CUDA events",
                           eager=eager)
        except Exception as exc:                                   # keep the bench line; report why the graph leg is missing
            latency["cuda_graph_error"] = f"{type(exc).__name__}: {exc}"[:300]
        est.config = dict(est.config, cuda_graph=False)
        del fL, fR, fM

    # ---- BASELINE config 5: infer_f2f_nw (no confidence heads), long sequence, end to end on all N GPUs: pinned host frames ->
    #      sharded solve -> all_gather -> host composition -> trajectory.freiburg on rank 0
    config5 = None
    if args.config5_frames > 1:
        nf = args.config5_frames
        est5 = PoseEstimator(dict(cfg, conf_weighing=False), K_t, seq.calib["bf"], CKPT if trained else None, (W_IMG, H_IMG)).to(dev)
        a5, b5 = parallel.frames_of(parallel.shard_ranges(nf - 1, world)[rank])
        idx5 = frame_indices(a5, b5)
        h5 = [torch.from_numpy(np.ascontiguousarray(x[idx5])).pin_memory() for x in (bL, bR, bM)]
        load5 = lambda fa, fb: tuple(t[fa - a5: fb - a5] for t in h5)
        outdir = tempfile.mkdtemp(prefix="rpe_cfg5_")

        def run5():
            traj, failed = parallel.infer_sequence_sharded(est5, load5, nf, chunk=args.chunk, use_graphs=args.graphs)
            if rank == 0:
                save_trajectory_array(traj, list(range(nf)), outdir)
            return failed, traj
        run5()                                               # warm-up (plans of the shard shapes)
        sync_all()
        t0 = time.perf_counter()
        failed5, traj5 = run5()
        sync_all()
        dt5 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt5, op=dist.ReduceOp.MAX)
        ate = None
        if rank == 0:                                        # accuracy of the written trajectory against the renderer's ground truth
            from rpe_b200.core.metrics.trajectory_metrics import absolute_trajectory_error, relative_pose_error
            from rpe_b200.lie import SE3
            gt = np.stack([np.block([[seq._extr[b][0].T, -seq._extr[b][0].T @ seq._extr[b][1][:, None]], [np.zeros((1, 3)), np.ones((1, 1))]])
                           for b in frame_indices(0, nf)])
            pred = SE3(traj5.double()).matrix().numpy()
            ate_rmse, _ = absolute_trajectory_error(gt, pred)
            rpe_t, rpe_r = relative_pose_error(gt, pred)
            ate = {"ate_rmse_mm": float(ate_rmse), "rpe_trans_mm": float(rpe_t.mean()), "rpe_rot_rad": float(rpe_r.mean()),
                   "path_length_mm": float(np.linalg.norm(np.diff(gt[:, :3, 3], axis=0), axis=1).sum()),
                   "against": "camera poses of the synthetic renderer (core/metrics/trajectory_metrics.py, the reference's ATE / RPE)"}
        config5 = {"workload": "infer_f2f_nw (conf_weighing False, lbgfs_iters 20)", "frames": nf, "pairs": nf - 1, "n_gpus": world, "accuracy": ate,
                   "seconds": float(dt5[0]), "pairs_per_s": (nf - 1) / float(dt5[0]), "failed_pairs": int(failed5.sum()),
                   "h2d_bytes_per_rank": int(sum(t.numel() for t in h5)),
                   "includes": "H2D of pinned uint8 frames, sharded solve, all_gather of (P,13) records, host composition, trajectory.freiburg write"}
        del est5, h5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pairs_total = main_res["total"] * args.steps
    value = pairs_total / (ms / 1e3)
    hbm_peak, tc_peak, peak_src = 6650.0, 1400.0, "fallback"     # B200_PROFILING.md fallbacks: copy GB/s, sustained cuBLAS bf16 TFLOP/s
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(mp["hbm_gbs"]), "measured"
        tc_peak = float(mp.get("bf16_tflops_sustained", mp.get("bf16_tflops", tc_peak)))
    except (OSError, KeyError, ValueError):
        pass
    # per-kernel rooflines from the in-run CUDA-event timers.  Convolution stages carry their executed tensor-core flops
    # as "units" (3 fp16 tensor-core products per multiply-add in the fp16x3 split); the algorithmic (fp32-equivalent) flops are a third.
    split = 3.0 if args.precision == "fp16x3" else 1.0
    kernels = {}
    for name, d in stage.items():
        secs = d["total_ms"] * 1e-3
        if name.startswith("conv_tc"):
            kernels[name] = {"bound": "tensor", "unit": "TFLOP/s", "achieved": d["units"] / split / secs / 1e12,
                             "executed": d["units"] / secs / 1e12, "peak": tc_peak}
        elif name in ALG_BYTES:
            units = evals_total * args.steps if name == "pose_solve" else d["units"]
            kernels[name] = {"bound": "hbm", "unit": "GB/s", "achieved": ALG_BYTES[name] * units / secs / 1e9, "peak": hbm_peak}
        else:
            continue
        kernels[name]["frac"] = kernels[name]["achieved"] / kernels[name]["peak"]
        kernels[name]["share_of_step"] = d["total_ms"] / ms
    if gn:
        for name, d in gn.items():
            kernels["pose_solve_" + name + "_probe"] = {"bound": "hbm", "unit": "GB/s", "achieved": d["GBps"], "peak": hbm_peak,
                                                        "frac": d["GBps"] / hbm_peak, "pairs": d["pairs"], "evals_per_pair": d["evals_per_pair"],
                                                        "ms_per_launch": d["ms"], "note": "42 B/px per objective evaluation"}
    conv = [k for k in stage if k.startswith("conv_tc")]
    own = {k: v for k, v in stage.items() if k in kernels and not k.startswith("conv_tc")}
    conv_ms = sum(stage[k]["total_ms"] for k in conv)
    if conv and conv_ms >= max([v["total_ms"] for v in own.values()] + [0.0]):
        # dominant kernel = conv_f16x3_kernel (tcgen05 implicit-GEMM convolution; update operator + encoders)
        flops_exec = sum(stage[k]["units"] for k in conv)
        n_launch = sum(stage[k]["launches"] for k in conv)
        achieved = flops_exec / split / (conv_ms * 1e-3) / 1e12
        traffic = None                              # DRAM bytes per conv launch from the committed ncu capture of this configuration
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json")))
            if tr["config"]["chunk"] == main_res["chunk"] and tr["config"]["precision"] == args.precision and args.pairs >= args.chunk:
                traffic = float(tr["dram_bytes_per_launch"])
        except (OSError, KeyError, ValueError):
            pass
        roofline = {"kernel": "conv_f16x3_pair_kernel", "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": achieved / tc_peak, "traffic": traffic, "traffic_source": "profiles/conv_traffic.json (ncu, launch-weighted mean)" if traffic else None,
                    "peak_source": peak_src,
                    "avg_launch_us": 1e3 * conv_ms / n_launch, "algorithmic_flops_per_launch": flops_exec / split / n_launch,
                    "executed_tflops": flops_exec / (conv_ms * 1e-3) / 1e12, "executed_frac": flops_exec / (conv_ms * 1e-3) / 1e12 / tc_peak,
                    "note": "algorithmic = fp32-equivalent convolution flops; the fp16x3 split executes 3 fp16 tensor-core products per multiply-add, "
                            "so executed_frac is the tensor-pipe figure and frac <= 1/3 by construction"}
    else:
        dom = max(own, key=lambda k: own[k]["total_ms"])
        d = own[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac"], "traffic": None, "peak_source": peak_src, "avg_launch_us": d["avg_us"],
                    "algorithmic_bytes_per_launch": kernels[dom]["achieved"] * 1e9 * d["avg_us"] * 1e-6}
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16x3": "fp16x3 (fp32-equivalent split, fp32 accumulate)", "tf32": "tf32", "fp16": "f16", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": {"workload": "f2f_640x512_seq65", "pairs_per_gpu_per_step": main_res["pairs_rank"], "pairs_per_step": main_res["total"],
                       "chunk": main_res["chunk"], "frames": "65 distinct rendered frames (64 distinct pairs) walked back and forth",
                       "precision": args.precision, "solver": "lbfgs_ref", "lbgfs_iters": 20, "conf_weighing": True,
                       "weights": "poseNet_2xf8up4b.pth" if trained else "random-init (checkpoint not shipped)",
                       "cuda_graphs": bool(args.graphs), "l2": "inputs larger than L2 (511 MB of frames per step at 64 pairs per GPU)",
                       "exchange": "one asynchronous all_gather of the (P,13) pose records per step, awaited one step later" if world > 1 else "none (N = 1)",
                       "ms_per_pair": ms / args.steps / max(main_res["pairs_rank"], 1)},
            "clocks": clocks, "gpu_launches": main_res["launches"],
            "e2e": {"value": pairs_total / (e2e_ms / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": int(main_res["h2d"]),
                    "d2h_bytes_per_step": int(main_res["total"] * 13 * 4), "input": "pinned uint8 frames + bool masks",
                    "call": "PoseEstimator.infer_sequence" if world == 1 else "parallel.infer_sequence_sharded",
                    "failed_pairs": main_res["failed"]},
            "roofline": roofline, "latency": latency, "kernels": kernels, "stages": stage,
            "lbfgs_evals_per_pair": evals_total / max(main_res["pairs_rank"], 1), "pose_parity": parity, "config5": config5}
    if other is not None:
        line["strong" if args.scaling == "weak" else "weak"] = {
            "value": other["total"] * args.steps / (other["ms"] / 1e3), "unit": "pairs/s", "ms_per_step": other["ms"] / args.steps,
            "pairs_per_step": other["total"], "pairs_per_gpu_per_step": other["pairs_rank"], "chunk": other["chunk"],
            "e2e": other["total"] * args.steps / (other["e2e_ms"] / 1e3),
            "note": "the other scaling mode, same run, same timing rules (device events, max over ranks)"}
    if world == 1:
        idx = frame_indices(0, 16)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(bL[idx], bR[idx], bM[idx], seq)
        if not args.no_gpu_reference:
            try:
                line["gpu_reference"] = gpu_reference(bL[idx], bR[idx], bM[idx], seq)
                line["gpu_reference"]["e2e_speedup"] = line["e2e"]["value"] / line["gpu_reference"]["value"]
            except Exception as exc:
                line["gpu_reference"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
