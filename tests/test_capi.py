"""CPU: the C-ABI shared library loads and exports every symbol declared in include/rpe_b200.h
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rpe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rpe_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib, build
    build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/rpe_b200.h but not exported"


def test_ctypes_binding_covers_header():
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib
    assert set(_declared_symbols()) == set(_lib.SIGNATURES)


def test_integration_doc_lists_every_entry_point():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared_symbols() if n not in doc and not any(g in doc for g in (n.rsplit("_", 1)[0] + "/",))]
    assert not missing, f"INTEGRATION.md does not mention {missing}"


def test_host_only_entry_points():
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib
    l = _lib.lib()
    assert l.rpe_version() >= 100
    assert l.rpe_status_string(0) == b"ok"
    assert l.rpe_status_string(-3) == b"workspace too small"
    # pyramid layout arithmetic is host side: 4 levels of a 64x80 grid, batch 2
    q = 64 * 80
    assert l.rpe_corr_level_offset(2, 64, 80, 0) == 0
    assert l.rpe_corr_level_offset(2, 64, 80, 1) == 2 * q * q * 4
    assert l.rpe_corr_pyramid_bytes(2, 64, 80, 4) >= 2 * q * (q + q // 4 + q // 16 + q // 64) * 4
    assert l.rpe_pose_workspace_bytes(64) > 0
    assert l.rpe_pose_set_groups(0) == -1 and l.rpe_pose_set_groups(8) == 0


def test_compose_trajectory_host_matches_the_oracle():
    """Stage a13 (pose_estimator.py:81-91) is host C: failure guard, rel.scale(1 / scale), last <- last * rel^-1 in fp32 --
    against the numpy restatement, including a failed pair (kept pose) and the empty sequence (n = 0: one frame, no pair)."""
    import numpy as np
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib
    from oracle import se3_np
    from oracle.detrand import det_uniform
    l = _lib.lib()
    n = 6
    xi = det_uniform((n, 6), 77, -0.02, 0.02).astype(np.float64)
    rel = np.stack([se3_np.exp(x) for x in xi]).astype(np.float32)
    log = xi.astype(np.float32)
    log[3, 1] = 0.2                                        # |log| > 0.1: the pair failed, the pose is kept
    rel[4, 0] = np.nan                                     # NaN pose: failed as well
    init = np.array([1.0, -2.0, 3.0, 0.0, 0.0, np.sin(0.05), np.cos(0.05)], np.float32)
    out, failed = np.zeros((n + 1, 7), np.float32), np.zeros((n,), np.uint8)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert l.rpe_compose_trajectory_host(ptr(rel), ptr(log), n, ptr(init), 250.0, ptr(out), ptr(failed)) == 0
    assert failed.tolist() == [0, 0, 0, 1, 1, 0]
    last = init.astype(np.float64)
    assert np.array_equal(out[0], init)
    for i in range(n):
        if not failed[i]:
            last = se3_np.mul(last, se3_np.inv(se3_np.scale(rel[i].astype(np.float64), 250.0)))
        assert np.abs(out[i + 1] - last).max() < 1e-4 * max(1.0, np.abs(last[:3]).max()), (i, out[i + 1], last)
    out0 = np.zeros((1, 7), np.float32)
    assert l.rpe_compose_trajectory_host(None, None, 0, ptr(init), 250.0, ptr(out0), None) == 0 and np.array_equal(out0[0], init)
    assert l.rpe_compose_trajectory_host(None, None, 2, ptr(init), 250.0, ptr(out0), None) != 0


def test_ops_refuse_cpu_tensors():
    import torch
    import rpe_b200  # noqa: F401
    from rpe_b200 import ops
    from rpe_b200._lib import RpeError
    with pytest.raises(RpeError):
        ops.proj(torch.ones(1, 1, 8, 8), torch.eye(3)[None])


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, call or open anything under oracle/
    (only tests/, __graft_entry__.smoke() and bench.py's CPU legs do)."""
    pkg = os.path.join(ROOT, "robust-pose-estimator_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/" in src.replace("oracle/_ref", ""):
                    offenders.append(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert not offenders, offenders


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback: if librpe_b200.so is missing, the first use raises instead of computing on the CPU."""
    import rpe_b200  # noqa: F401
    from rpe_b200 import _lib
    from rpe_b200._lib import RpeError
    loaded = _lib.lib()
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RpeError, match="no CPU fallback"):
        _lib.lib()
    monkeypatch.undo()
    assert _lib.lib() is loaded
