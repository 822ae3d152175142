"""ORACLE -- TEST INFRASTRUCTURE ONLY.  SE3 group maths, geodesic loss, 3x3 SVD / E->(R,t).

PARITY UNPINNED: the arithmetic restated here lives in a third-party dependency that is absent
from /root/reference -- `lietorch` (pinned `lietorch==0.2`, /root/reference/environment.yml:20;
README.md:22 installs git master) -- and the reference holds no test, fixture or golden vector at
that boundary.  What is restated is lietorch's published algorithm (group formulas of its
so3.h / se3.h and the left-perturbation backward of its group_ops / lietorch_gpu.cu), anchored on
the reference's call sites: src/geom/losses.py:8-10, src/model.py:146-152, train.py:144-146.
The formulas are self-checked in tests/ (exp/log round trip, scipy.linalg.expm/logm, finite
differences of exp(delta)*X perturbations) but could not be compared with lietorch itself.

The SVD / essential->(R,t) stage of BASELINE.json config 3 has NO reference implementation at all
(src/geom/ holds only losses.py); its oracle is numpy.linalg.svd in float64 plus the textbook
decomposition (Hartley & Zisserman 9.6.2).  Also parity unpinned.

Data layout (lietorch): SE3 element = [tx,ty,tz, qx,qy,qz,qw]; tangent = [tau(3), phi(3)].
All functions are float64 numpy, vectorised over leading dims.
"""
import numpy as np

EPS = 1e-6  # lietorch common.h


# ------------------------------------------------------------------ quaternion helpers
def _normalize_q(q):
    """lietorch's SO3(const Scalar*) constructor normalises the quaternion on load."""
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def qmul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def qrot(q, v):
    """p + w*uv + qv x uv with uv = 2 qv x p (lietorch SO3::operator*(Point))."""
    qv, w = q[..., :3], q[..., 3:]
    uv = 2.0 * np.cross(qv, v)
    return v + w * uv + np.cross(qv, uv)


def qmat(q):
    x, y, z, w = np.moveaxis(q, -1, 0)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def hat(v):
    x, y, z = np.moveaxis(v, -1, 0)
    o = np.zeros_like(x)
    return np.stack([np.stack([o, -z, y], -1), np.stack([z, o, -x], -1), np.stack([-y, x, o], -1)], -2)


# ------------------------------------------------------------------ SE3 forward ops (A12)
def se3_split(X):
    X = np.asarray(X, np.float64)
    return X[..., :3], _normalize_q(X[..., 3:])


def se3_mul(X, Y):
    """(q1 q2, t1 + q1*t2)."""
    t1, q1 = se3_split(X)
    t2, q2 = se3_split(Y)
    return np.concatenate([t1 + qrot(q1, t2), qmul(q1, q2)], -1)


def se3_inv(X):
    """(q^-1, -(q^-1 * t))."""
    t, q = se3_split(X)
    qi = qconj(q)
    return np.concatenate([-qrot(qi, t), qi], -1)


def so3_log(q):
    """atan-based log with lietorch's branches (so3.h Log)."""
    v, w = q[..., :3], q[..., 3]
    n2 = (v * v).sum(-1)
    n = np.sqrt(n2)
    with np.errstate(divide="ignore", invalid="ignore"):
        small = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w)
        near_pi = np.where(w > 0, np.pi, -np.pi) / n
        gen = 2.0 * np.arctan(n / w) / n
    f = np.where(n2 < EPS * EPS, small, np.where(np.abs(w) < EPS, near_pi, gen))
    return f[..., None] * v


def so3_exp(phi):
    th2 = (phi * phi).sum(-1)
    th = np.sqrt(th2)
    th4 = th2 * th2
    with np.errstate(divide="ignore", invalid="ignore"):
        imag = np.where(th < EPS, 0.5 - th2 / 48.0 + th4 / 3840.0, np.sin(0.5 * th) / th)
        real = np.where(th < EPS, 1.0 - th2 / 8.0 + th4 / 384.0, np.cos(0.5 * th))
    return np.concatenate([imag[..., None] * phi, real[..., None]], -1)


def so3_left_jacobian(phi):
    th2 = (phi * phi).sum(-1)
    th = np.sqrt(th2)
    with np.errstate(divide="ignore", invalid="ignore"):
        c1 = np.where(th < EPS, 0.5 - th2 / 24.0, (1.0 - np.cos(th)) / th2)
        c2 = np.where(th < EPS, 1.0 / 6.0 - th2 / 120.0, (th - np.sin(th)) / (th2 * th))
    P = hat(phi)
    return np.eye(3) + c1[..., None, None] * P + c2[..., None, None] * (P @ P)


def so3_left_jacobian_inverse(phi):
    th2 = (phi * phi).sum(-1)
    th = np.sqrt(th2)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where(th < EPS, 1.0 / 12.0, (1.0 - th * np.cos(0.5 * th) / (2.0 * np.sin(0.5 * th))) / th2)
    P = hat(phi)
    return np.eye(3) - 0.5 * P + c[..., None, None] * (P @ P)


def se3_log(X):
    """[tau, phi] with phi = SO3.log(q), tau = J_l^-1(phi) t."""
    t, q = se3_split(X)
    phi = so3_log(q)
    tau = np.einsum("...ij,...j->...i", so3_left_jacobian_inverse(phi), t)
    return np.concatenate([tau, phi], -1)


def se3_exp(a):
    a = np.asarray(a, np.float64)
    tau, phi = a[..., :3], a[..., 3:]
    t = np.einsum("...ij,...j->...i", so3_left_jacobian(phi), tau)
    return np.concatenate([t, so3_exp(phi)], -1)


def se3_matrix(X):
    t, q = se3_split(X)
    T = np.zeros(X.shape[:-1] + (4, 4))
    T[..., :3, :3] = qmat(q)
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T


# ------------------------------------------------------------------ adjoint / jacobians (backward)
def se3_adjoint(X):
    """Ad = [[R, [t]x R],[0, R]] for tangent order [tau, phi]."""
    t, q = se3_split(X)
    R = qmat(q)
    A = np.zeros(X.shape[:-1] + (6, 6))
    A[..., :3, :3] = R
    A[..., :3, 3:] = hat(t) @ R
    A[..., 3:, 3:] = R
    return A


def _calcQ(tau, phi):
    T, P = hat(tau), hat(phi)
    th2 = (phi * phi).sum(-1)
    th = np.sqrt(th2)
    th4 = th2 * th2
    with np.errstate(divide="ignore", invalid="ignore"):
        c1 = np.where(th < EPS, 1 / 6.0 + th2 / 120.0, (th - np.sin(th)) / (th2 * th))
        c2 = np.where(th < EPS, 1 / 24.0 - th2 / 720.0, (th2 + 2 * np.cos(th) - 2) / (2 * th4))
        c3 = np.where(th < EPS, 1 / 120.0 - th2 / 2520.0, (2 * th - 3 * np.sin(th) + th * np.cos(th)) / (2 * th4 * th))
    e = lambda c: c[..., None, None]
    return (0.5 * T + e(c1) * (P @ T + T @ P + P @ T @ P)
            + e(c2) * (P @ P @ T + T @ P @ P - 3 * P @ T @ P)
            + e(c3) * (P @ T @ P @ P + P @ P @ T @ P))


def se3_left_jacobian(a):
    a = np.asarray(a, np.float64)
    tau, phi = a[..., :3], a[..., 3:]
    J = so3_left_jacobian(phi)
    out = np.zeros(a.shape[:-1] + (6, 6))
    out[..., :3, :3] = J
    out[..., :3, 3:] = _calcQ(tau, phi)
    out[..., 3:, 3:] = J
    return out


def se3_left_jacobian_inverse(a):
    a = np.asarray(a, np.float64)
    tau, phi = a[..., :3], a[..., 3:]
    Ji = so3_left_jacobian_inverse(phi)
    out = np.zeros(a.shape[:-1] + (6, 6))
    out[..., :3, :3] = Ji
    out[..., :3, 3:] = -Ji @ _calcQ(tau, phi) @ Ji
    out[..., 3:, 3:] = Ji
    return out


def _pad7(g6):
    return np.concatenate([g6, np.zeros(g6.shape[:-1] + (1,))], -1)


def se3_mul_backward(dZ, X, Y):
    """lietorch: dX = dZ, dY = dZ * Ad(X) (row vector); 7-wide buffers, slot 7 = 0."""
    g = np.asarray(dZ, np.float64)[..., :6]
    dY = np.einsum("...i,...ij->...j", g, se3_adjoint(X))
    return _pad7(g), _pad7(dY)


def se3_inv_backward(dY, X):
    """lietorch: dX = -dY * Ad(X^-1)."""
    g = np.asarray(dY, np.float64)[..., :6]
    return _pad7(-np.einsum("...i,...ij->...j", g, se3_adjoint(se3_inv(X))))


def se3_log_backward(da, X):
    """lietorch: dX = da * J_l^-1(log X)."""
    return _pad7(np.einsum("...i,...ij->...j", np.asarray(da, np.float64), se3_left_jacobian_inverse(se3_log(X))))


def se3_exp_backward(dX, a):
    """lietorch: da = dX[:6] * J_l(a)."""
    return np.einsum("...i,...ij->...j", np.asarray(dX, np.float64)[..., :6], se3_left_jacobian(a))


# ------------------------------------------------------------------ A11: geodesic loss
def geodesic_loss(Ps, Gs):
    """src/geom/losses.py:3-21.  Ps (ground truth), Gs (prediction) [B,2,7].
    Returns (loss_tr, loss_rot, d[B,2,6])."""
    Ps = np.asarray(Ps, np.float64)
    Gs = np.asarray(Gs, np.float64)
    ii, jj = [0, 1], [1, 0]
    dP = se3_mul(Ps[:, jj], se3_inv(Ps[:, ii]))
    dG = se3_mul(Gs[:, jj], se3_inv(Gs[:, ii]))
    d = se3_log(se3_mul(dG, se3_inv(dP)))
    tau, phi = d[..., :3], d[..., 3:]
    return np.linalg.norm(tau, axis=-1).mean(), np.linalg.norm(phi, axis=-1).mean(), d


def geodesic_loss_grad(Ps, Gs, w_tr=10.0, w_rot=10.0):
    """Gradient of w_tr*loss_tr + w_rot*loss_rot w.r.t. Gs *in lietorch's convention*
    (tangent-space gradient of a left perturbation written into the first 6 of 7 slots),
    chained by hand through log -> mul -> (mul, inv) exactly as autograd would chain lietorch's
    custom backward functions."""
    Ps = np.asarray(Ps, np.float64)
    Gs = np.asarray(Gs, np.float64)
    ii, jj = [0, 1], [1, 0]
    Pi_inv = se3_inv(Ps[:, ii]); dP = se3_mul(Ps[:, jj], Pi_inv); dP_inv = se3_inv(dP)
    Gi_inv = se3_inv(Gs[:, ii]); dG = se3_mul(Gs[:, jj], Gi_inv)
    E = se3_mul(dG, dP_inv)
    d = se3_log(E)
    tau, phi = d[..., :3], d[..., 3:]
    cnt = tau.shape[0] * tau.shape[1]
    g_d = np.concatenate([w_tr * tau / np.linalg.norm(tau, axis=-1, keepdims=True) / cnt,
                          w_rot * phi / np.linalg.norm(phi, axis=-1, keepdims=True) / cnt], -1)
    g_E = se3_log_backward(g_d, E)
    g_dG, _ = se3_mul_backward(g_E, dG, dP_inv)
    g_Gj, g_Giinv = se3_mul_backward(g_dG, Gs[:, jj], Gi_inv)
    g_Gi = se3_inv_backward(g_Giinv, Gs[:, ii])
    g = np.zeros_like(Gs)
    for s, (i, j) in enumerate(zip(ii, jj)):
        g[:, j] += g_Gj[:, s]
        g[:, i] += g_Gi[:, s]
    return g


# ------------------------------------------------------------------ config 3: SVD, E -> (R,t)
def svd3(E):
    """float64 LAPACK SVD; singular values descending.  Returns U, S, V (not V^T)."""
    U, S, Vt = np.linalg.svd(np.asarray(E, np.float64))
    return U, S, np.swapaxes(Vt, -1, -2)


def essential_to_rt(E):
    """Textbook decomposition: with E = U diag(s) V^T, det(U),det(V) forced to +1,
    R1 = U W V^T, R2 = U W^T V^T, t = U[:,2] (up to sign)."""
    U, S, V = svd3(E)
    U = U * np.where(np.linalg.det(U) < 0, -1.0, 1.0)[..., None, None]
    V = V * np.where(np.linalg.det(V) < 0, -1.0, 1.0)[..., None, None]
    Wm = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    R1 = U @ Wm @ np.swapaxes(V, -1, -2)
    R2 = U @ Wm.T @ np.swapaxes(V, -1, -2)
    return R1, R2, U[..., :, 2]
