"""ORACLE / TEST INFRASTRUCTURE ONLY.  numpy float64 restatement of the SE(3) arithmetic that the
reference obtains from the third-party ``lietorch`` package (un-vendored, un-pinned:
/root/reference/README.md:37).  Published lietorch semantics:

  pose  = [tx ty tz qx qy qz qw]             tangent xi = [tau | phi]
  Exp   : q = [sin(th/2)/th phi, cos(th/2)],  t = V(phi) tau
  Log   : inverse of Exp (tau = V^-1 t)
  X * Y : group product,  X * p = R(q) p + t,  retraction  X <- Exp(a) X  (left)

Pinned by oracle/lietorch (the stand-in the reference's own unit tests pass with) in
tests/test_oracle_se3.py.  Single poses (7,) and point sets (..., 3).
"""
import numpy as np

EPS = 1e-6


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_rotate(q, p):
    qv, qw = q[:3], q[3]
    uv = 2.0 * np.cross(qv, p)
    return p + qw * uv + np.cross(qv, uv)


def rotmat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def exp(xi):
    xi = np.asarray(xi, dtype=np.float64)
    tau, phi = xi[:3], xi[3:]
    th2 = float(phi @ phi)
    th = np.sqrt(th2)
    if th < EPS:
        imag = 0.5 - th2 / 48.0 + th2 * th2 / 3840.0
        real = 1.0 - th2 / 8.0 + th2 * th2 / 384.0
        c1 = 0.5 - th2 / 24.0
        c2 = 1.0 / 6.0 - th2 / 120.0
    else:
        imag = np.sin(0.5 * th) / th
        real = np.cos(0.5 * th)
        c1 = (1.0 - np.cos(th)) / th2
        c2 = (th - np.sin(th)) / (th2 * th)
    pxt = np.cross(phi, tau)
    t = tau + c1 * pxt + c2 * np.cross(phi, pxt)
    return np.concatenate((t, imag * phi, [real]))


def log(X):
    X = np.asarray(X, dtype=np.float64)
    t, qv, qw = X[:3], X[3:6], X[6]
    n2 = float(qv @ qv)
    if n2 < EPS * EPS:
        s = 2.0 / qw - (2.0 / 3.0) * n2 / (qw ** 3)
    else:
        n = np.sqrt(n2)
        s = 2.0 * np.arctan(n / qw) / n
    phi = s * qv
    th2 = float(phi @ phi)
    th = np.sqrt(th2)
    if th < EPS:
        c2 = 1.0 / 12.0
    else:
        c2 = (1.0 - th * np.cos(0.5 * th) / (2.0 * np.sin(0.5 * th))) / th2
    pxt = np.cross(phi, t)
    tau = t - 0.5 * pxt + c2 * np.cross(phi, pxt)
    return np.concatenate((tau, phi))


def mul(X, Y):
    return np.concatenate((X[:3] + quat_rotate(X[3:], Y[:3]), quat_mul(X[3:], Y[3:])))


def inv(X):
    qi = np.array([-X[3], -X[4], -X[5], X[6]])
    return np.concatenate((-quat_rotate(qi, X[:3]), qi))


def act(X, p):
    """p: (..., 3) -> R p + t."""
    return p @ rotmat(X[3:]).T + X[:3]


def scale(X, s):
    return np.concatenate((X[:3] * s, X[3:]))


def identity():
    return np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])


def retract(X, a):
    """lietorch LieGroupParameter update: X <- Exp(a) * X."""
    return mul(exp(a), X)
