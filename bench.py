#!/usr/bin/env python
"""Benchmark of the f2f pose path (BASELINE.json metric: f2f frame-pairs/sec @640x512 stereo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16x3|fp32|tf32|fp16|bf16]

A step = one pass of the hot path over one batch of synthetic input: a 65-frame 640x512 stereo sequence
(64 frame pairs) per rank (weak scaling; pairs are independent in f2f).
  value  pairs/s with the frames resident in HBM (float32 0..255 like the reference's tensors), device-timed.
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


def synthetic_sequence(n_frames, n_base=6, seed=0):
    """(limg, rimg, mask) uint8/bool arrays of an n_frames sequence: n_base rendered frames of a smooth camera
    walk, traversed back and forth so consecutive frames always differ by one small motion."""
    import rpe_b200  # noqa: F401
    from rpe_b200.dataset.synthetic import SyntheticStereoSequence
    seq = SyntheticStereoSequence(n_base, (W_IMG, H_IMG), seed=seed, smooth_walk=True, motion_sigma=0.02, holes=2)
    base = [seq.frame_u8(i) for i in range(n_base)]
    period = 2 * (n_base - 1)
    idx = [(i % period) if (i % period) < n_base else period - (i % period) for i in range(n_frames)]
    L = np.stack([base[i][0] for i in idx])
    R = np.stack([base[i][1] for i in idx])
    M = np.stack([base[i][2] for i in idx])
    return L, R, M, seq


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


def run_reference(args, rank):
    """Reference arm: the CPU port of the reference's f2f path on the host cores (oracle/pipeline_ref.py)."""
    if rank != 0:
        return
    import torch
    from oracle import pipeline_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    L, R, M, seq = synthetic_sequence(args.steps + args.warmup + 1)
    sd = _state_dict(torch)
    trk = pipeline_ref.RefTracker(sd, seq.calib["intrinsics"]["left"], seq.calib["bf"])
    f = lambda a: torch.from_numpy(a.astype(np.float32))[None]
    trk.step(f(L[0]), f(R[0]), torch.from_numpy(M[0])[None])                    # first frame: stereo depth only
    for i in range(1, 1 + args.warmup):
        trk.step(f(L[i]), f(R[i]), torch.from_numpy(M[i])[None])
    t0 = time.perf_counter()
    for i in range(1 + args.warmup, 1 + args.warmup + args.steps):
        trk.step(f(L[i]), f(R[i]), torch.from_numpy(M[i])[None])
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "f2f_640x512_seq65", "sample": "1 frame pair per step (bounded sample of the 64-pair batch)",
                       "lbgfs_iters": 20, "conf_weighing": True},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} consecutive pairs, oracle/pipeline_ref.py (torch CPU fp32 + numpy fp64 L-BFGS)"},
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


def cpu_baseline(L, R, M, seq, budget_s=25.0, max_pairs=3):
    """The oracle port timed on the host cores over a bounded sample of the same workload."""
    import torch
    from oracle import pipeline_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    trk = pipeline_ref.RefTracker(_state_dict(torch), seq.calib["intrinsics"]["left"], seq.calib["bf"])
    f = lambda a: torch.from_numpy(a.astype(np.float32))[None]
    trk.step(f(L[0]), f(R[0]), torch.from_numpy(M[0])[None])
    trk.step(f(L[1]), f(R[1]), torch.from_numpy(M[1])[None])                     # warm-up pair
    for k in trk.timing:
        trk.timing[k] = 0.0
    n, t0 = 0, time.perf_counter()
    while n < max_pairs and time.perf_counter() - t0 < budget_s:
        trk.step(f(L[2 + n]), f(R[2 + n]), torch.from_numpy(M[2 + n])[None])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{n} pairs of the same sequence after 1 warm-up pair ({dt:.1f} s)",
            "split_s_per_pair": {k: v / n for k, v in trk.timing.items()}}, trk


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


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                  # fd 1 -> stderr for the whole run (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "tf32", "fp16", "bf16"],
                    help="bf16x3 (default): whole trunk on the hand-written tcgen05 kernels, fp32-equivalent split arithmetic "
                         "(parity-gated); fp32: cuDNN fp32 trunk; tf32 / fp16 / bf16: cuDNN reduced precision (not parity-gated)")
    ap.add_argument("--chunk", type=int, default=32, help="frames per engine chunk (measured on B200: 11 -> 318, 22 -> 330, 32 -> 334 pairs/s; fewer, larger launches)")
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--graphs", action="store_true")
    ap.add_argument("--pose-groups", type=int, default=0, help="concurrently solved pairs in rpe_pose_solve (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-frames", type=int, default=40, help="batch-1 latency leg (config 2): pairs timed one by one; 0 = off")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib, build, ops
    from rpe_b200.core.pose.pose_estimator import PoseEstimator
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

    # rank r owns pairs [r*P, (r+1)*P) of one global (N*P + 1)-frame sequence: frames [r*P, (r+1)*P] (one halo frame)
    T = args.pairs + 1
    Lg, Rg, Mg, seq = synthetic_sequence(world * args.pairs + 1, seed=0)
    L, R, M = (a[rank * args.pairs: rank * args.pairs + T] for a in (Lg, Rg, Mg))
    del Lg, Rg, Mg
    trained = os.path.isfile(CKPT)
    cfg = dict(SLAM, precision=args.precision)
    est = PoseEstimator(cfg, torch.tensor(seq.calib["intrinsics"]["left"]), seq.calib["bf"], CKPT if trained else None,
                        (W_IMG, H_IMG)).to(dev)
    # host (pinned) and device copies of the step input
    hL, hR, hM = (torch.from_numpy(a).pin_memory() for a in (L, R, M))
    dL, dR, dM = hL.to(dev).float(), hR.to(dev).float(), hM.to(dev)
    h2d = hL.numel() + hR.numel() + hM.numel()
    from rpe_b200 import parallel
    engine = F2FEngine(est, chunk=args.chunk, use_graphs=args.graphs)
    inv_scale = float((1 / est.scale).float().cpu())

    def device_step(l=dL, r=dR, m=dM):
        engine.reset()
        rel, log, evals = engine.infer_sequence(l, r, m, sequence_start=(rank == 0))
        rec = parallel.gather_pair_records(torch.cat((rel, log), 1).contiguous(), world * args.pairs)   # only exchange step
        return rel, log, evals, rec

    def e2e_step():
        if world == 1:                                       # the public API call a user makes, fed from pinned host frames
            est.last_pose = est.last_pose.__class__.Identity(1, device=dev)   # (uploads chunk k+1 while chunk k is solved)
            return est.infer_sequence(hL, hR, hM, chunk=args.chunk, use_graphs=args.graphs)
        l = hL.to(dev, non_blocking=True).float()
        r = hR.to(dev, non_blocking=True).float()
        m = hM.to(dev, non_blocking=True)
        rec = device_step(l, r, m)[3]
        if rank == 0:                                                             # host composition of the gathered poses
            return parallel.compose_trajectory(rec, [0, 0, 0, 0, 0, 0, 1.0], inv_scale)
        return None, torch.zeros(1)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                                  # nvidia-smi polls from the warm-up on; mark() opens the reported window
    for _ in range(args.warmup):
        device_step()
    # ---- timed region (device-resident inputs), CUDA events, per-stage timers on
    timers = ops.enable_timers(True)
    sync_all()
    launches0 = lib.rpe_launch_count()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evals_total = 0
    last = None
    for _ in range(args.steps):
        last = device_step()
    e1.record()
    sync_all()
    clocks = sampler.stop()
    launches = lib.rpe_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    evals_total = float(last[2].sum().item())
    stage = {}
    for name, evs in timers.items():
        t = [a.elapsed_time(b) for a, b, _ in evs]
        units = sum(u for _, _, u in evs)
        stage[name] = {"launches": len(evs), "total_ms": float(np.sum(t)), "avg_us": 1e3 * float(np.mean(t)), "units": units}
    ops.enable_timers(False)
    # ---- e2e (host buffers in, host trajectory out), wall clock bracketed by synchronize
    for _ in range(max(1, min(2, args.warmup))):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        traj, failed = e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    # ---- latency path (BASELINE config 2: batch 1, the per-frame tracker call of the reference API), rank 0, outside `value`
    latency = None
    if rank == 0 and args.latency_frames > 0:
        est.frame = est.last_frame = None
        est.failure_flags = []
        nlat = min(args.latency_frames, T - 1)
        evs = []
        for k in range(nlat + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            est(dL[k:k + 1], dR[k:k + 1], dM[k:k + 1].clone())
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        lat = np.array([a.elapsed_time(b) for a, b in evs[1 + min(5, nlat // 4):]])   # frame 0 has no pair; skip warm-up pairs
        eager = {"p50_ms_per_pair": float(np.percentile(lat, 50)), "p90_ms_per_pair": float(np.percentile(lat, 90)),
                 "pairs": int(lat.size), "call": "PoseEstimator.forward, batch 1, device-resident frame, CUDA events"}
        latency = dict(eager)
        try:
            # the same call with config['cuda_graph']: the per-frame device work replayed as one captured CUDA graph
            est.config = dict(est.config, cuda_graph=True)
            est.frame = est.last_frame = None
            evs = []
            for k in range(nlat + 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                est(dL[k:k + 1], dR[k:k + 1], dM[k:k + 1].clone())
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            lat = np.array([a.elapsed_time(b) for a, b in evs[1 + min(5, nlat // 4):]])   # pair 1 captures the graph
            latency = {"p50_ms_per_pair": float(np.percentile(lat, 50)), "p90_ms_per_pair": float(np.percentile(lat, 90)),
                       "pairs": int(lat.size),
                       "call": "PoseEstimator.forward with config cuda_graph=True, batch 1, device-resident frame, CUDA events",
                       "eager": eager}
        except Exception as exc:                                   # keep the bench line; report why the graph leg is missing
            latency["cuda_graph_error"] = f"{type(exc).__name__}: {exc}"[:300]
        est.config = dict(est.config, cuda_graph=False)
    t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pairs_total = world * args.pairs * args.steps
    value = pairs_total / (ms / 1e3)
    hbm_peak, tc_peak, peak_src = 6650.0, 1400.0, "fallback"     # B200_PROFILING.md fallbacks: copy GB/s, sustained cuBLAS bf16 TFLOP/s
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(mp["hbm_gbs"]), "measured"
        tc_peak = float(mp.get("bf16_tflops_sustained", mp.get("bf16_tflops", tc_peak)))
    except (OSError, KeyError, ValueError):
        pass
    # per-kernel rooflines from the in-run CUDA-event timers.  Convolution stages carry their executed tensor-core flops
    # as "units" (3 bf16 MMAs per multiply-add in the bf16x3 split); the algorithmic (fp32-equivalent) flops are a third.
    split = 3.0 if args.precision == "bf16x3" else 1.0
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
    conv = [k for k in stage if k.startswith("conv_tc")]
    own = {k: v for k, v in stage.items() if k in kernels and not k.startswith("conv_tc")}
    conv_ms = sum(stage[k]["total_ms"] for k in conv)
    if conv and conv_ms >= max([v["total_ms"] for v in own.values()] + [0.0]):
        # dominant kernel = conv_bf16_kernel (tcgen05 implicit-GEMM convolution; update operator + encoders)
        flops_exec = sum(stage[k]["units"] for k in conv)
        n_launch = sum(stage[k]["launches"] for k in conv)
        achieved = flops_exec / split / (conv_ms * 1e-3) / 1e12
        traffic = None                              # DRAM bytes per conv launch from the committed ncu capture of this configuration
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json")))
            if tr["config"]["chunk"] == args.chunk and tr["config"]["precision"] == args.precision and args.pairs >= args.chunk:
                traffic = float(tr["dram_bytes_per_launch"])
        except (OSError, KeyError, ValueError):
            pass
        roofline = {"kernel": "conv_bf16_pair_kernel", "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": achieved / tc_peak, "traffic": traffic, "traffic_source": "profiles/conv_traffic.json (ncu, launch-weighted mean)" if traffic else None,
                    "peak_source": peak_src,
                    "avg_launch_us": 1e3 * conv_ms / n_launch, "algorithmic_flops_per_launch": flops_exec / split / n_launch,
                    "executed_tflops": flops_exec / (conv_ms * 1e-3) / 1e12, "executed_frac": flops_exec / (conv_ms * 1e-3) / 1e12 / tc_peak,
                    "note": "algorithmic = fp32-equivalent convolution flops; the bf16x3 split executes 3 bf16 MMAs per multiply-add, "
                            "so executed_frac is the tensor-pipe figure and frac <= 1/3 by construction"}
    else:
        dom = max(own, key=lambda k: own[k]["total_ms"])
        d = own[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac"], "traffic": None, "peak_source": peak_src, "avg_launch_us": d["avg_us"],
                    "algorithmic_bytes_per_launch": kernels[dom]["achieved"] * 1e9 * d["avg_us"] * 1e-6}
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (fp32-equivalent split, fp32 accumulate)", "tf32": "tf32", "fp16": "f16", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": {"workload": "f2f_640x512_seq65", "pairs_per_gpu_per_step": args.pairs, "chunk": args.chunk,
                       "precision": args.precision, "solver": "lbfgs_ref", "lbgfs_iters": 20, "conf_weighing": True,
                       "weights": "poseNet_2xf8up4b.pth" if trained else "random-init (checkpoint not shipped)",
                       "cuda_graphs": bool(args.graphs), "l2": "inputs larger than L2 (511 MB of frames per step)",
                       "ms_per_pair": ms / args.steps / args.pairs},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": pairs_total / (e2e_ms / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(world * args.pairs * 13 * 4), "input": "pinned uint8 frames + bool masks",
                    "failed_pairs": int(failed.sum())},
            "roofline": roofline, "latency": latency, "kernels": kernels, "stages": stage,
            "lbfgs_evals_per_pair": evals_total / args.pairs}
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline(L, R, M, seq)
        line["cpu_baseline"] = cb
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
