"""ORACLE / TEST INFRASTRUCTURE ONLY.

Runs the reference's OWN unit tests (/root/reference/tests/unit_test_pose_head.py,
unit_test_pinhole_transforms.py) against the pure-torch lietorch stand-in in oracle/lietorch.
This is the acceptance test of the stand-in: the reference is imported unmodified from
/root/reference (read-only); nothing is copied.  Only runnable in the build container.
"""
import os
import sys
import unittest

REF = os.environ.get("RPE_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if not os.path.isdir(REF):
        print(f"reference not present at {REF}; nothing to do")
        return 0
    sys.path.insert(0, HERE)                       # -> `import lietorch` resolves to the stand-in
    sys.path.insert(0, REF)                        # -> `import core...`
    sys.path.insert(0, os.path.join(REF, "tests"))
    import torch
    torch.manual_seed(0)
    import unit_test_pose_head as t_head
    import unit_test_pinhole_transforms as t_pin
    suite = unittest.TestSuite()
    for case, names in ((t_pin.PinholeTransformTester, ("test_transform", "test_transform_backward")),
                        (t_head.PoseHeadTester, ("test_forward", "test_backward"))):
        for n in names:
            suite.addTest(case(n))
    res = unittest.TextTestRunner(verbosity=2).run(suite)
    return 0 if res.wasSuccessful() else 1


if __name__ == "__main__":
    sys.exit(main())
