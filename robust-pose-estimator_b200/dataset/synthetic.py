"""Deterministic synthetic stereo sequences of the StereoMIS shape.

The reference has no synthetic generator; its dataset classes (dataset/stereo_dataset.py:19-44)
yield ``(limg, rimg, mask, img_number)`` with images float32 0..255 RGB ``(3,H,W)`` and a bool mask
``(1,H,W)``.  This module produces the same tuple from a ray-cast textured height field so that
stereo disparity, temporal flow and the camera motion are geometrically consistent and the pose
ground truth is known.  numpy only (host side); images are quantised to uint8 like a real camera.

Geometry (mm): world = camera-1 frame.  Surface z = Z(x, y) (plane + smooth bumps), texture
T(x, y) a sum of sine octaves.  A camera with extrinsics (R, t) (p_cam = R p_world + t) is rendered
by Newton ray/height-field intersection.  Pixel centres at +0.5 (pinhole_transforms.py:15-17).
"""
import numpy as np

__all__ = ["SyntheticStereoSequence", "se3_exp_np", "se3_log_np", "default_intrinsics", "bench_sequence", "triangle_index"]

BENCH_FRAMES = 65       # BASELINE config 3: 64 frame pairs


def bench_sequence(size=(640, 512)):
    """The 65-frame sequence (64 distinct pairs) behind bench.py, the 64-pair reference golden (tests/golden/bench64_poses.npz,
    --bench64) and the config-5 leg: tethered walk, seed 0, two mask holes per frame."""
    return SyntheticStereoSequence(BENCH_FRAMES, size, seed=0, motion_sigma=0.012, holes=2, tether=0.05)


def triangle_index(i, n_base=BENCH_FRAMES):
    """Frame i of an arbitrarily long sequence that walks the n_base rendered frames back and forth (0..n-1, n-2..0, 1..):
    consecutive frames are always rendered neighbours, so every pair is one small camera motion."""
    period = 2 * (n_base - 1)
    r = i % period
    return r if r < n_base else period - r


def default_intrinsics(width=640, height=512):
    f = 0.625 * width
    return np.array([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def se3_exp_np(xi):
    """xi = [tau, phi] -> (R, t); same convention as lietorch (translation first, t = V tau)."""
    xi = np.asarray(xi, dtype=np.float64)
    tau, phi = xi[:3], xi[3:]
    th = np.linalg.norm(phi)
    P = _hat(phi)
    if th < 1e-8:
        R = np.eye(3) + P + 0.5 * P @ P
        V = np.eye(3) + 0.5 * P + P @ P / 6.0
    else:
        R = np.eye(3) + np.sin(th) / th * P + (1 - np.cos(th)) / th ** 2 * P @ P
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * P + (th - np.sin(th)) / th ** 3 * P @ P
    return R, V @ tau


def se3_log_np(R, t):
    """Inverse of ``se3_exp_np``: (R, t) -> xi = [tau, phi]."""
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    phi = w if th < 1e-8 else w * th / np.sin(th)
    P = _hat(phi)
    if th < 1e-8:
        V = np.eye(3) + 0.5 * P + P @ P / 6.0
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * P + (th - np.sin(th)) / th ** 3 * P @ P
    return np.concatenate((np.linalg.solve(V, t), phi))


class SyntheticStereoSequence:
    """Indexable like the reference's StereoDataset: ``seq[i] -> (limg, rimg, mask, i)``.

    :param n_frames: number of stereo frames
    :param size: (W, H) like the reference's ``img_size`` config key (configuration/infer_f2f.yaml:13-15)
    :param seed: RNG seed (numpy ``default_rng``)
    :param depth_scale_mm: the tracker's ``depth_clipping[1]`` (250 mm); poses are generated in
           normalised units (translation / depth_scale) like unit_test_pose_head.py:28
    :param motion_sigma: per-frame ``xi ~ sigma * N(0, I)`` in normalised units (config 2), or a smooth
           random walk ``xi_t = 0.9 xi_{t-1} + 0.3 sigma N`` when ``smooth_walk`` (config 5)
    :param holes: number of rectangular invalid regions in each mask (specularity stand-ins)
    :param tether: > 0: the ABSOLUTE camera pose follows a mean-reverting walk ``a_{k+1} = (1 - tether) a_k + sigma N`` in
           the tangent space, so a long sequence stays in front of the surface (the relative walks drift without bound);
           consecutive frames still differ by one small motion of size ~sigma
    """

    def __init__(self, n_frames=2, size=(640, 512), seed=0, bf=2200.0, depth_scale_mm=250.0,
                 motion_sigma=0.01, smooth_walk=False, holes=0, intrinsics=None, tether=0.0):
        self.n_frames = int(n_frames)
        self.W, self.H = int(size[0]), int(size[1])
        self.bf = float(bf)
        self.depth_scale_mm = float(depth_scale_mm)
        self.K = default_intrinsics(self.W, self.H) if intrinsics is None else np.asarray(intrinsics, np.float64)
        self.baseline_mm = self.bf / self.K[0, 0]
        rng = np.random.default_rng(seed)
        # --- surface: tilted plane + bumps, depth kept inside (0.2, 0.9) * depth_scale
        self._z0 = depth_scale_mm * rng.uniform(0.45, 0.6)
        self._slope = rng.uniform(-0.25, 0.25, size=2)
        nb = 6
        self._bump_amp = depth_scale_mm * rng.uniform(0.01, 0.03, size=nb)
        self._bump_w = rng.uniform(0.008, 0.025, size=(nb, 2)) * rng.choice([-1.0, 1.0], size=(nb, 2))
        self._bump_p = rng.uniform(0, 2 * np.pi, size=nb)
        # --- texture: octaves with wavelengths 2.5..80 mm, amplitude ~ 1/f^0.35
        no = 40
        lam = np.exp(rng.uniform(np.log(2.5), np.log(80.0), size=no))
        ang = rng.uniform(0, 2 * np.pi, size=no)
        self._tex_w = (2 * np.pi / lam)[:, None] * np.stack((np.cos(ang), np.sin(ang)), -1)
        self._tex_p = rng.uniform(0, 2 * np.pi, size=(no, 3))
        amp = lam ** 0.35
        self._tex_a = 330.0 * amp / amp.sum()
        # --- camera trajectory: T_k maps camera-k coordinates to camera-(k+1) coordinates
        self.rel_xi = np.zeros((max(self.n_frames - 1, 0), 6))
        xi = np.zeros(6)
        for k in range(self.n_frames - 1):
            if smooth_walk:
                xi = 0.9 * xi + 0.3 * motion_sigma * rng.standard_normal(6)
            else:
                xi = motion_sigma * rng.standard_normal(6)
            self.rel_xi[k] = xi
        self._extr = [(np.eye(3), np.zeros(3))]
        if tether > 0.0:
            a = np.zeros(6)
            for k in range(self.n_frames - 1):
                a = (1.0 - tether) * a + motion_sigma * rng.standard_normal(6)
                R, t = se3_exp_np(a)
                self._extr.append((R, t * depth_scale_mm))
                Rp, tp = self._extr[-2]
                Rr = R @ Rp.T                                              # T_k = P_{k+1} P_k^-1
                self.rel_xi[k] = se3_log_np(Rr, (self._extr[-1][1] - Rr @ tp) / depth_scale_mm)
        else:
            for k in range(self.n_frames - 1):
                R, t = se3_exp_np(self.rel_xi[k])
                t = t * depth_scale_mm
                Rp, tp = self._extr[-1]
                self._extr.append((R @ Rp, R @ tp + t))
        self._hole_rng_seed = seed * 7919 + 13
        self.holes = int(holes)
        v, u = np.meshgrid(np.arange(self.H) + 0.5, np.arange(self.W) + 0.5, indexing="ij")
        Kinv = np.linalg.inv(self.K)
        self._rays = np.stack((u, v, np.ones_like(u)), 0).reshape(3, -1)
        self._rays = Kinv @ self._rays                                    # (3, HW), z = 1

    # ---- scene functions ------------------------------------------------------------------
    def _Z(self, x, y, grad=False):
        z = self._z0 + self._slope[0] * x + self._slope[1] * y
        zx = np.full_like(x, self._slope[0]) if grad else None
        zy = np.full_like(x, self._slope[1]) if grad else None
        for a, w, p in zip(self._bump_amp, self._bump_w, self._bump_p):
            ph = w[0] * x + w[1] * y + p
            z = z + a * np.sin(ph)
            if grad:
                c = a * np.cos(ph)
                zx += c * w[0]
                zy += c * w[1]
        return (z, zx, zy) if grad else z

    def _tex(self, x, y):
        out = np.full((3,) + x.shape, 127.5)
        for a, w, p in zip(self._tex_a, self._tex_w, self._tex_p):
            ph = w[0] * x + w[1] * y
            s, c = np.sin(ph), np.cos(ph)
            for ch in range(3):                                   # sin(ph + p) = sin ph cos p + cos ph sin p
                out[ch] += (a * np.cos(p[ch])) * s + (a * np.sin(p[ch])) * c
        return out

    def _render(self, R, t):
        """Return (rgb uint8 (3,H,W), depth_mm (H,W)) for camera p_cam = R p_w + t."""
        o = -R.T @ t
        d = R.T @ self._rays                                               # world ray dirs, (3, HW)
        s = np.full(d.shape[1], self._z0)
        for _ in range(8):                                                 # Newton on o.z + s d.z = Z(x(s), y(s))
            x = o[0] + s * d[0]
            y = o[1] + s * d[1]
            z, zx, zy = self._Z(x, y, grad=True)
            s = s - (o[2] + s * d[2] - z) / (d[2] - zx * d[0] - zy * d[1])
        x = o[0] + s * d[0]
        y = o[1] + s * d[1]
        rgb = np.clip(np.rint(self._tex(x, y)), 0, 255).astype(np.uint8).reshape(3, self.H, self.W)
        return rgb, s.reshape(self.H, self.W)                              # rays have z_cam = 1 -> s is depth

    # ---- dataset protocol -----------------------------------------------------------------
    def __len__(self):
        return self.n_frames

    def frame_u8(self, i):
        R, t = self._extr[i]
        left, depth = self._render(R, t)
        right, _ = self._render(R, t - np.array([self.baseline_mm, 0.0, 0.0]))
        mask = np.ones((1, self.H, self.W), dtype=bool)
        if self.holes:
            rng = np.random.default_rng(self._hole_rng_seed + i)
            for _ in range(self.holes):
                h, w = rng.integers(8, max(9, self.H // 8)), rng.integers(8, max(9, self.W // 8))
                y0, x0 = rng.integers(0, self.H - h), rng.integers(0, self.W - w)
                mask[0, y0:y0 + h, x0:x0 + w] = False
        return left, right, mask, depth

    def __getitem__(self, i):
        left, right, mask, _ = self.frame_u8(i)
        return left.astype(np.float32), right.astype(np.float32), mask, i

    def gt_depth_mm(self, i):
        return self.frame_u8(i)[3]

    def gt_rel_pose(self, k):
        """(R, t_normalised) of T_k with p_{k+1} = T_k p_k, translation in normalised units."""
        return se3_exp_np(self.rel_xi[k])

    def frames_u8(self, indices=None, workers=None, cache_dir=None):
        """(L, R, M) uint8 / bool arrays of the frames ``indices`` (default: all), rendered by a pool of worker processes
        (one frame costs ~2.5 s of numpy) and cached as an .npz keyed by the generator parameters when ``cache_dir`` is given."""
        import hashlib
        import os
        idx = list(range(self.n_frames)) if indices is None else list(indices)
        path = None
        if cache_dir is not None:
            sig = repr((self.W, self.H, self.bf, self.depth_scale_mm, self.holes, self._hole_rng_seed, idx,
                        [(R.tobytes(), t.tobytes()) for R, t in (self._extr[i] for i in idx)], self._tex_p.tobytes(),
                        self._bump_p.tobytes())).encode()
            path = os.path.join(cache_dir, "rpe_synth_" + hashlib.sha1(sig).hexdigest()[:16] + ".npz")
            if os.path.isfile(path):
                try:
                    z = np.load(path)
                    return z["L"], z["R"], z["M"]
                except Exception:
                    pass
        workers = workers or min(len(idx), os.cpu_count() or 1)
        if workers > 1:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(workers) as pool:
                out = pool.map(self._frame_for_pool, idx)
        else:
            out = [self._frame_for_pool(i) for i in idx]
        L, R, M = (np.stack([o[k] for o in out]) for k in range(3))
        if path is not None:
            try:
                tmp = path + f".{os.getpid()}.tmp.npz"
                np.savez(tmp, L=L, R=R, M=M)
                os.replace(tmp, path)
            except OSError:
                pass
        return L, R, M

    def _frame_for_pool(self, i):
        return self.frame_u8(i)[:3]

    @property
    def calib(self):
        return {"intrinsics": {"left": self.K.copy()}, "bf": self.bf}
