"""Self-checks of the SE3 / SVD oracle (oracle/geom_oracle.py).  lietorch is not available, so
these are consistency checks of the restated group maths, not a comparison with lietorch:
parity unpinned (see the oracle header)."""
import numpy as np
import scipy.linalg

import geom_oracle as G


def _rand_tangent(n, seed=0, scale=1.0):
    r = np.random.RandomState(seed)
    return r.randn(n, 6) * np.array([1, 1, 1, 0.5, 0.5, 0.5]) * scale


def _hat6(a):
    T = np.zeros((4, 4))
    T[:3, :3] = G.hat(a[3:])
    T[:3, 3] = a[:3]
    return T


def test_exp_matches_expm_and_log_roundtrip():
    a = _rand_tangent(64, 1, 1.3)
    X = G.se3_exp(a)
    M = G.se3_matrix(X)
    for i in range(8):
        np.testing.assert_allclose(M[i], scipy.linalg.expm(_hat6(a[i])), atol=1e-12)
    np.testing.assert_allclose(G.se3_log(X), a, atol=1e-10)


def test_small_and_near_pi_angles():
    a = _rand_tangent(16, 2)
    a[:8, 3:] *= 1e-9
    np.testing.assert_allclose(G.se3_log(G.se3_exp(a)), a, atol=1e-12)
    phi = np.array([[np.pi - 1e-7, 0, 0], [0, -(np.pi - 1e-9), 0]])
    q = G.so3_exp(phi)
    np.testing.assert_allclose(np.abs(G.so3_log(q)), np.abs(phi), atol=1e-6)


def test_group_axioms():
    X = G.se3_exp(_rand_tangent(32, 3))
    Y = G.se3_exp(_rand_tangent(32, 4))
    ident = np.tile(np.array([0, 0, 0, 0, 0, 0, 1.0]), (32, 1))
    np.testing.assert_allclose(G.se3_matrix(G.se3_mul(X, G.se3_inv(X))), G.se3_matrix(ident), atol=1e-12)
    np.testing.assert_allclose(G.se3_matrix(G.se3_mul(X, Y)), G.se3_matrix(X) @ G.se3_matrix(Y), atol=1e-12)


def test_left_jacobian_by_finite_difference():
    a = _rand_tangent(6, 5)
    J = G.se3_left_jacobian(a)
    Ji = G.se3_left_jacobian_inverse(a)
    np.testing.assert_allclose(J @ Ji, np.tile(np.eye(6), (6, 1, 1)), atol=1e-10)
    h = 1e-6
    for k in range(6):
        d = np.zeros(6); d[k] = h
        # exp(a + d) = exp(J d) exp(a)   =>   log(exp(a+d) exp(a)^-1) / h  = J[:,k]
        lhs = G.se3_log(G.se3_mul(G.se3_exp(a + d), G.se3_inv(G.se3_exp(a)))) / h
        np.testing.assert_allclose(lhs, J[:, :, k], atol=5e-6)


def test_backward_rules_are_left_perturbation_gradients():
    """f(exp(delta) X) = f(X) + g . delta  for every op's lietorch-style gradient."""
    r = np.random.RandomState(7)
    X = G.se3_exp(_rand_tangent(5, 8)); Y = G.se3_exp(_rand_tangent(5, 9))
    w7 = r.randn(5, 7); w6 = r.randn(5, 6)
    h = 1e-6

    def pert(Z, k):
        d = np.zeros((5, 6)); d[:, k] = h
        return G.se3_mul(G.se3_exp(d), Z)

    def left_grad(f, Z):        # numeric tangent-space gradient of scalar-per-row f
        return np.stack([(f(pert(Z, k)) - f(Z)) / h for k in range(6)], -1)

    # a scalar test function of a group element must itself be expressed through log to be
    # chart independent: use <w6, log(.)>
    fl = lambda Z: (w6 * G.se3_log(Z)).sum(-1)
    gl = G.se3_log_backward(w6, X)
    np.testing.assert_allclose(left_grad(fl, X), gl[:, :6], atol=2e-5)
    assert np.all(gl[:, 6] == 0)
    # mul
    gz = G.se3_log_backward(w6, G.se3_mul(X, Y))
    gX, gY = G.se3_mul_backward(gz, X, Y)
    np.testing.assert_allclose(left_grad(lambda Z: fl(G.se3_mul(Z, Y)), X), gX[:, :6], atol=2e-5)
    np.testing.assert_allclose(left_grad(lambda Z: fl(G.se3_mul(X, Z)), Y), gY[:, :6], atol=2e-5)
    # inv
    gi = G.se3_inv_backward(G.se3_log_backward(w6, G.se3_inv(X)), X)
    np.testing.assert_allclose(left_grad(lambda Z: fl(G.se3_inv(Z)), X), gi[:, :6], atol=2e-5)
    # exp: d/da <w, log(exp(a))> chained = w
    a = _rand_tangent(5, 10)
    ge = G.se3_exp_backward(G.se3_log_backward(w6, G.se3_exp(a)), a)
    np.testing.assert_allclose(ge, w6, atol=1e-9)


def test_geodesic_loss_gradient_convention():
    r = np.random.RandomState(11)
    B = 4
    Ps = np.stack([np.tile([0, 0, 0, 0, 0, 0, 1.0], (B, 1)), G.se3_exp(_rand_tangent(B, 12))], 1)
    Gs = np.stack([np.tile([0, 0, 0, 0, 0, 0, 1.0], (B, 1)), G.se3_exp(_rand_tangent(B, 13))], 1)
    g = G.geodesic_loss_grad(Ps, Gs)
    h = 1e-6
    f = lambda Gx: 10 * G.geodesic_loss(Ps, Gx)[0] + 10 * G.geodesic_loss(Ps, Gx)[1]
    for b in range(B):
        for s in range(2):
            for k in range(6):
                d = np.zeros(6); d[k] = h
                Gp = Gs.copy()
                Gp[b, s] = G.se3_mul(G.se3_exp(d), Gs[b, s])
                np.testing.assert_allclose((f(Gp) - f(Gs)) / h, g[b, s, k], atol=5e-5)
    assert np.all(g[..., 6] == 0)


def test_svd_and_essential_decomposition():
    r = np.random.RandomState(3)
    E = r.randn(100, 3, 3)
    U, S, V = G.svd3(E)
    np.testing.assert_allclose(U * S[:, None, :] @ np.swapaxes(V, -1, -2), E, atol=1e-12)
    # true essential matrices [t]x R
    a = _rand_tangent(50, 5)
    X = G.se3_exp(a)
    t = X[:, :3] / np.linalg.norm(X[:, :3], axis=-1, keepdims=True)
    R = G.qmat(X[:, 3:])
    Es = G.hat(t) @ R
    R1, R2, tt = G.essential_to_rt(Es)
    for Rc in (R1, R2):
        np.testing.assert_allclose(np.linalg.det(Rc), 1.0, atol=1e-9)
    e1 = np.linalg.norm((R1 - R).reshape(50, -1), axis=-1)
    e2 = np.linalg.norm((R2 - R).reshape(50, -1), axis=-1)
    assert np.all(np.minimum(e1, e2) < 1e-8)
    assert np.all(np.minimum(np.linalg.norm(tt - t, axis=-1), np.linalg.norm(tt + t, axis=-1)) < 1e-8)
