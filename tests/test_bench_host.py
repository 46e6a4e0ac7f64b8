"""CPU: host-side logic of bench.py and of the host-fed throughput call (no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _sampler_with(rows_before, rows_after):
    import bench
    s = bench.ClockSampler(0)
    s.p = subprocess.Popen([sys.executable, "-c", "pass"])          # a finished stand-in for the nvidia-smi poller
    s.p.wait()
    open(s.f.name, "w").write("\n".join(rows_before) + ("\n" if rows_before else ""))
    s.mark()
    open(s.f.name, "a").write("\n".join(rows_after) + ("\n" if rows_after else ""))
    return s.stop()


ROW = "0, {sm}, 1965, 900.0, 0x4, Not Active, Not Active, Not Active, {cap}"


def test_clock_sampler_reports_only_the_timed_window():
    before = [ROW.format(sm=1000, cap="Not Active")] * 3
    after = [ROW.format(sm=1700, cap="Active"), ROW.format(sm=1600, cap="Active"), ROW.format(sm=1650, cap="Not Active")]
    out = _sampler_with(before, after)
    assert out["samples"] == 3 and out["sm_mhz"] == 1650.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and "window" not in out


def test_clock_sampler_falls_back_to_warmup_rows_when_the_window_is_empty():
    out = _sampler_with([ROW.format(sm=1500 + k, cap="Not Active") for k in range(5)], [])
    assert out["samples"] == 3 and out["sm_mhz"] == 1503.0 and "window" in out


def test_reference_arm_prints_one_json_line_on_stdout():
    """`bench.py --impl reference` (the CPU port of the reference path): exactly one JSON line on stdout with the contract keys."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "f2f_frame_pairs_per_sec_640x512" and d["unit"] == "pairs/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_host_fed_chunk_bounds():
    """The chunk boundaries of PoseEstimator._infer_from_host equal the engine's own batching (the first chunk carries the
    first frame as an extra image), so host-fed and device-resident calls see identical batches."""
    def host_bounds(T, chunk):
        bounds, a = [], 0
        while a < T:
            b = min(a + chunk + (1 if a == 0 else 0), T)
            bounds.append((a, b))
            a = b
        return bounds

    def engine_bounds(T, chunk):                      # F2FEngine.infer_sequence with prev = None
        if T == 1:
            return [(0, 1)]
        start = min(1 + chunk, T)
        return [(0, start)] + [(a, min(a + chunk, T)) for a in range(start, T, chunk)]

    for T in (1, 2, 3, 6, 33, 65, 66, 100):
        for chunk in (1, 2, 11, 32):
            assert host_bounds(T, chunk) == engine_bounds(T, chunk), (T, chunk)
    src = open(os.path.join(ROOT, "robust-pose-estimator_b200", "core", "pose", "pose_estimator.py")).read()
    assert "b = min(a + chunk + (1 if a == 0 else 0), T)" in src          # the formula under test is the one in the product code
