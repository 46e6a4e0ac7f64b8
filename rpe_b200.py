"""Import shim: exposes the package directory ``robust-pose-estimator_b200/`` (not a valid Python
identifier) as the importable package ``rpe_b200``.

    import rpe_b200
    from rpe_b200.core.pose.pose_net import PoseNet
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "robust-pose-estimator_b200")
_spec = importlib.util.spec_from_file_location(
    "rpe_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rpe_b200"] = _mod
_spec.loader.exec_module(_mod)
