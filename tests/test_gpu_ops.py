"""GPU parity tests, op by op: every C-ABI entry point against the numpy oracle on seeded inputs.
Integer / index work is bit-exact; floating point within the tolerance written at each check."""
import numpy as np
import pytest
import torch

import relpose_oracle as O
import geom_oracle as G
from rel_pose_b200 import ops, synthetic as S
from rel_pose_b200.lietorch import SE3

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rnd(seed, *shape, scale=1.0):
    return (S.hash_normal(seed, "t", int(np.prod(shape))).reshape(shape) * scale).astype(np.float32)


def report(name, got, ref, atol, rtol=0.0):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    lim = atol + rtol * np.abs(ref)
    print(f"[parity] {name}: max_abs_err={err.max():.3e} max_ref={np.abs(ref).max():.3e} worst_ratio={(err / lim).max():.3f}")
    assert np.isfinite(got).all(), name
    assert (err <= lim).all(), f"{name}: max err {err.max():.3e}"


def test_arch_is_blackwell():
    from rel_pose_b200 import _lib
    assert _lib.lib().rp_device_arch(0) == 100


@pytest.mark.parametrize("H,W", [(96, 128), (384, 384), (64, 80), (480, 640), (224, 224), (17, 501)])
def test_preprocess_bit_exact(H, W):
    img = S.make_images_numpy(5, 2, H, W, True)
    ref = O.preprocess(img)
    got = ops.preprocess(cu(img)).cpu().numpy()
    assert np.array_equal(got, ref)
    got8 = ops.preprocess(cu(img.astype(np.uint8))).cpu().numpy()
    assert np.array_equal(got8, ref)
    img2 = S.make_images_numpy(6, 1, H, W, False)          # non-integer pixels
    assert np.array_equal(ops.preprocess(cu(img2)).cpu().numpy(), O.preprocess(img2))


@pytest.mark.parametrize("kind,H,W", [("matterport", 384, 512), ("square", 256, 256), ("varied", 64, 80)])
def test_intrinsics_and_posenc_bit_exact(kind, H, W):
    B = 3
    k = S.make_intrinsics_numpy(B, kind, 7)
    ref_k = O.update_intrinsics(k, H, W)
    dk = cu(k)
    kxy, flags = ops.intrinsics_prepare(dk, H, W)
    assert np.array_equal(dk.cpu().numpy(), ref_k)            # in-place side effect, bit exact
    assert int(flags.item()) == 0
    pos = ops.posenc(B, kxy, dk.device).cpu().numpy()
    lin = ops.lin24().numpy()
    # oracle with the same 24-entry table the product uses
    O_lin = O.linspace_pm1()
    ref = O.positional_encodings(B, ref_k)
    if np.array_equal(lin, O_lin):
        assert np.array_equal(pos, ref)
    else:  # ATen's linspace differs by 1 ulp on this host (vector width); compare with slack
        np.testing.assert_allclose(pos, ref, rtol=0, atol=3e-7)
    pos0 = ops.posenc(2, None, dk.device).cpu().numpy()
    np.testing.assert_allclose(pos0, O.positional_encodings(2, None), rtol=0, atol=1.2e-7)


def test_intrinsics_flags():
    k = S.make_intrinsics_numpy(2, "matterport")
    k[1, 1, 0] += 1.0
    _, flags = ops.intrinsics_prepare(cu(k), 384, 512)
    assert int(flags.item()) & 1
    k = S.make_intrinsics_numpy(2, "matterport")
    k[0, :, 2] = 0.0
    _, flags = ops.intrinsics_prepare(cu(k), 384, 512)
    assert int(flags.item()) & 2


def test_tokens_posembed():
    fm = rnd(1, 3, 192, 24, 24)
    pe = rnd(2, 1, 576, 192)
    got = ops.tokens_posembed(cu(fm), cu(pe)).cpu().numpy()
    ref = O.tokens_from_feature_map(fm) + pe
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("rows", [1, 7, 70, 1152])
def test_layernorm(rows):
    x = rnd(3, rows, 192, scale=3.0) + 0.5
    g = 1 + 0.1 * rnd(4, 192); b = 0.1 * rnd(5, 192)
    got = ops.layernorm(cu(x), cu(g), cu(b), 1e-6).cpu().numpy()
    ref = O.layernorm(x.astype(np.float64), g.astype(np.float64), b.astype(np.float64))
    report(f"layernorm rows={rows}", got, ref, atol=2e-6, rtol=2e-6)


@pytest.mark.parametrize("M,N,K,act,res", [
    (1152, 576, 192, 0, False), (1152, 192, 192, 0, True), (1152, 768, 192, 1, False), (1152, 192, 768, 0, True),
    (140, 768, 192, 1, False), (2, 512, 26880, 2, False), (64, 512, 26880, 2, False), (3, 512, 512, 2, False),
    (3, 14, 512, 0, False), (1, 96, 16, 0, False), (129, 97, 20, 1, True), (500, 200, 148, 2, True)])
def test_linear(M, N, K, act, res):
    a = rnd(6, M, K); w = rnd(7, N, K, scale=1.0 / np.sqrt(K)); b = rnd(8, N, scale=0.1)
    r = rnd(9, M, N) if res else None
    got = ops.linear(cu(a), cu(w), cu(b), act=act, residual=cu(r) if res else None).cpu().numpy()
    y = a.astype(np.float64) @ w.astype(np.float64).T + b
    if act == 1:
        y = O.gelu(y)
    elif act == 2:
        y = np.maximum(y, 0)
    if res:
        y = y + r
    report(f"linear {M}x{N}x{K} act={act}", got, y, atol=1e-5, rtol=1e-5)
    # no bias
    got2 = ops.linear(cu(a), cu(w)).cpu().numpy()
    report(f"linear-nobias {M}x{N}x{K}", got2, a.astype(np.float64) @ w.astype(np.float64).T, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("B", [1, 3, 64, 130])
def test_regressor_tail(B):
    """rp_regressor_tail_f32 = pose_regressor[2:5] (src/model.py:93-97) in one launch, against float64."""
    h = np.maximum(rnd(40, B, 512), 0); w1 = rnd(41, 512, 512, scale=1.0 / np.sqrt(512)); b1 = rnd(42, 512, scale=0.1)
    w2 = rnd(43, 14, 512, scale=1.0 / np.sqrt(512)); b2 = rnd(44, 14, scale=0.1)
    got = ops.regressor_tail(cu(h), cu(np.ascontiguousarray(w1.T)), cu(b1), cu(w2), cu(b2)).cpu().numpy()
    h2 = np.maximum(h.astype(np.float64) @ w1.astype(np.float64).T + b1, 0)
    ref = h2 @ w2.astype(np.float64).T + b2
    report(f"regressor_tail B={B}", got, ref, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("B", [1, 7, 8, 9, 64, 256])
def test_regressor_tail_cluster_kernel_against_first_version(B, monkeypatch):
    """The cluster kernel (8 CTAs per 8 rows, rows exchanged through distributed shared memory; default) against the
    one-row-per-CTA kernel (RELPOSE_REGRESSOR_TAIL_V1=1) and float64: partial clusters, exact multiples, several clusters."""
    h = np.maximum(rnd(45, B, 512), 0); w1 = rnd(46, 512, 512, scale=1.0 / np.sqrt(512)); b1 = rnd(47, 512, scale=0.1)
    w2 = rnd(48, 14, 512, scale=1.0 / np.sqrt(512)); b2 = rnd(49, 14, scale=0.1)
    args = (cu(h), cu(np.ascontiguousarray(w1.T)), cu(b1), cu(w2), cu(b2))
    monkeypatch.delenv("RELPOSE_REGRESSOR_TAIL_V1", raising=False)
    new = ops.regressor_tail(*args).cpu().numpy()
    again = ops.regressor_tail(*args).cpu().numpy()
    monkeypatch.setenv("RELPOSE_REGRESSOR_TAIL_V1", "1")
    old = ops.regressor_tail(*args).cpu().numpy()
    monkeypatch.delenv("RELPOSE_REGRESSOR_TAIL_V1", raising=False)
    assert np.array_equal(new, again)                                    # fixed summation order: deterministic
    h2 = np.maximum(h.astype(np.float64) @ w1.astype(np.float64).T + b1, 0)
    ref = h2 @ w2.astype(np.float64).T + b2
    report(f"regressor_tail v2 B={B}", new, ref, atol=1e-5, rtol=1e-5)
    report(f"regressor_tail v2 vs v1 B={B}", new, old, atol=1e-5, rtol=1e-5)


def test_linear_residual_may_alias_output():
    a = rnd(1, 300, 192); w = rnd(2, 192, 192, scale=0.07); b = rnd(3, 192, scale=0.1); r = rnd(4, 300, 192)
    x = cu(r)
    ops.linear(cu(a), cu(w), cu(b), residual=x, out=x)
    report("linear alias", x.cpu().numpy(), a.astype(np.float64) @ w.astype(np.float64).T + b + r, 1e-5, 1e-5)


@pytest.mark.parametrize("n,scale", [(1, 1.0), (3, 3.0)])
def test_self_attention(n, scale):
    qkv = rnd(11, n, 576, 576, scale=scale)
    got = ops.self_attention(cu(qkv)).cpu().numpy()
    q, k, v = O.split_qkv(qkv.astype(np.float64))
    p = O.softmax(q @ k.transpose(0, 1, 3, 2) * 0.125, -1)
    ref = (p @ v).transpose(0, 2, 1, 3).reshape(n, 576, 192)
    # logits reach |S| ~ 40 at scale 3: one float32 ulp of S is already a 4e-6 relative change of exp(S)
    report(f"self_attention n={n}", got, ref, atol=2e-5 * scale, rtol=1e-5 * scale)


@pytest.mark.parametrize("B,scale", [(1, 1.0), (2, 2.5)])
def test_essential_module(B, scale):
    x = rnd(12, 2 * B, 576, 192)
    w = rnd(13, 576, 192, scale=scale / np.sqrt(192)); bq = rnd(14, 576, scale=0.1)
    pw = rnd(15, 192, 210, scale=1 / np.sqrt(210)); pb = rnd(16, 192, scale=0.1)
    k = O.update_intrinsics(S.make_intrinsics_numpy(B, "varied", 3), 384, 512)
    p = {"c.qkv.weight": w.astype(np.float64), "c.qkv.bias": bq.astype(np.float64),
         "c.proj_fundamental.weight": pw.astype(np.float64), "c.proj_fundamental.bias": pb.astype(np.float64)}
    x64 = x.astype(np.float64).reshape(B, 2, 576, 192)
    (y2, y1), (f1, f2) = O.essential_matrix_module(x64[:, 0], x64[:, 1], p, "c", k, return_bilinear=True)
    qkv = ops.linear(cu(x), cu(w), cu(bq))
    kxy = cu(np.stack([1 / (k[:, 0, 0] / k[:, 0, 2]), 1 / (k[:, 0, 1] / k[:, 0, 3])], -1).astype(np.float32))
    pos = ops.posenc(B, kxy, qkv.device)
    bil = ops.essential(qkv, pos)
    # entries of one 70x70 form are sums with cancellation: tolerance relative to the form's scale
    report("bilinear1", bil[:, 0].cpu().numpy(), f1, atol=1e-5 * np.abs(f1).max(), rtol=2e-4)
    report("bilinear2", bil[:, 1].cpu().numpy(), f2, atol=1e-5 * np.abs(f2).max(), rtol=2e-4)
    out = ops.em_project(bil, cu(pw), cu(pb)).cpu().numpy().reshape(B, 2, 70, 192)
    report("em_project slot0 (=Y2)", out[:, 0], y2, atol=1e-5 * np.abs(y2).max(), rtol=2e-4)
    report("em_project slot1 (=Y1)", out[:, 1], y1, atol=1e-5 * np.abs(y1).max(), rtol=2e-4)


@pytest.mark.parametrize("B", [1, 5])
def test_em_project_versions_are_bit_identical(B, monkeypatch):
    """proj_fundamental (vision_transformer.py:225-231): the one-CTA-per-matrix kernel (default) accumulates every output
    over k in the same order as the first kernel (RELPOSE_EM_PROJECT_V1=1) -> identical bits; and it matches float64."""
    bil = rnd(31, B, 2, 3, 70, 70, scale=3.0)
    pw = rnd(32, 192, 210, scale=1 / np.sqrt(210)); pb = rnd(33, 192, scale=0.1)
    monkeypatch.delenv("RELPOSE_EM_PROJECT_V1", raising=False)
    new = ops.em_project(cu(bil), cu(pw), cu(pb)).cpu().numpy()
    monkeypatch.setenv("RELPOSE_EM_PROJECT_V1", "1")
    old = ops.em_project(cu(bil), cu(pw), cu(pb)).cpu().numpy()
    monkeypatch.delenv("RELPOSE_EM_PROJECT_V1", raising=False)
    assert new.shape == (2 * B, 70, 192) and np.array_equal(new, old)
    z = bil.astype(np.float64).reshape(B, 2, 210, 70).transpose(0, 1, 3, 2)            # [b, dir, c, h*70+a]
    ref = (z @ pw.astype(np.float64).T + pb.astype(np.float64))[:, ::-1].reshape(2 * B, 70, 192)   # slot = 1 - dir
    report("em_project vs float64", new, ref, atol=1e-5 * np.abs(ref).max(), rtol=2e-4)


class _BN:
    def __init__(self, seed, C):
        self.weight = cu(1 + 0.2 * rnd(seed, C)); self.bias = cu(0.1 * rnd(seed + 1, C))
        self.running_mean = cu(0.2 * rnd(seed + 2, C)); self.running_var = cu(0.6 + np.abs(rnd(seed + 3, C)))
        self.eps = 1e-5

    def params(self, prefix):
        return {prefix + ".weight": self.weight.cpu().numpy(), prefix + ".bias": self.bias.cpu().numpy(),
                prefix + ".running_mean": self.running_mean.cpu().numpy(), prefix + ".running_var": self.running_var.cpu().numpy()}


@pytest.mark.parametrize("n,H,W,C,O,k,stride,pad,bias,act,res", [
    (2, 30, 34, 3, 64, 7, 2, 3, False, 2, "none"), (3, 20, 20, 64, 64, 3, 1, 1, False, 2, "pre"),
    (2, 21, 19, 64, 128, 3, 2, 1, False, 2, "none"), (2, 21, 19, 64, 128, 1, 2, 0, False, 0, "none"),
    (2, 28, 28, 128, 192, 5, 1, 0, True, 2, "pre+post"), (1, 28, 28, 192, 192, 5, 1, 0, True, 2, "none"),
    (130, 6, 6, 8, 20, 3, 1, 1, True, 0, "none")])
def test_conv2d_bn_act_residual(n, H, W, C, O, k, stride, pad, bias, act, res):
    x = rnd(31, n, C, H, W)
    w = rnd(32, O, C, k, k, scale=1.0 / np.sqrt(C * k * k))
    b = rnd(33, O, scale=0.1) if bias else None
    bn = _BN(40, O)
    ref = O_conv(x, w, b, stride, pad, bn)
    Ho, Wo = ref.shape[2], ref.shape[3]
    rp = rnd(34, n, O, Ho, Wo) if "pre" in res else None
    rq = rnd(35, Ho * Wo, O) if "post" in res else None
    if rp is not None:
        ref = ref + rp
    if act == 2:
        ref = np.maximum(ref, 0)
    if rq is not None:
        ref = ref + rq.reshape(Ho, Wo, O).transpose(2, 0, 1)[None]
    Cp = (C + 3) // 4 * 4
    xh = np.zeros((n, H, W, Cp), np.float32); xh[..., :C] = x.transpose(0, 2, 3, 1)
    wp = ops.permute_conv_weight(cu(w))
    assert tuple(wp.shape) == (O, k, k, Cp)
    scale, shift = ops.bn_fold(bn, cu(b) if bias else None)
    y = ops.conv2d_nhwc(cu(xh), wp, scale, shift, stride, pad, act,
                        cu(rp.transpose(0, 2, 3, 1)) if rp is not None else None,
                        cu(rq) if rq is not None else None, Ho * Wo if rq is not None else 0)
    report(f"conv {k}x{k}/s{stride} {C}->{O}", y.cpu().numpy().transpose(0, 3, 1, 2), ref, atol=2e-5, rtol=2e-5)


def O_conv(x, w, b, stride, pad, bn):
    y = O.conv2d(x.astype(np.float64), w.astype(np.float64), None if b is None else b.astype(np.float64), stride, pad)
    return O.batchnorm_eval(y, {k: v.astype(np.float64) for k, v in bn.params("bn").items()}, "bn")


def test_maxpool_and_preprocess_nhwc4():
    x = rnd(36, 3, 64, 23, 29)
    got = ops.maxpool3x3s2_nhwc(cu(x.transpose(0, 2, 3, 1))).cpu().numpy().transpose(0, 3, 1, 2)
    assert np.array_equal(got, O.maxpool3x3s2(x))
    img = S.make_images_numpy(7, 2, 96, 128, True)
    got = ops.preprocess_nhwc4(cu(img)).cpu().numpy()
    assert np.array_equal(got[..., :3].transpose(0, 3, 1, 2), O.preprocess(img)) and np.all(got[..., 3] == 0)
    got8 = ops.preprocess_nhwc4(cu(img.astype(np.uint8))).cpu().numpy()
    assert np.array_equal(got8, got)


def test_normalize_pose():
    raw = rnd(17, 5, 2, 7)
    raw[3, 1, 3:] *= 1e-3                    # ||q|| < 0.01 branch
    Gs = rnd(18, 5, 2, 7)
    got = ops.normalize_pose(cu(raw), cu(Gs)).cpu().numpy()
    ref = O.normalize_preds(Gs, raw)
    report("normalize_pose", got, ref, atol=1e-6, rtol=1e-6)
    assert np.array_equal(got[:, 0], Gs[:, 0]) and np.array_equal(got[:, 1, :3], raw[:, 1, :3])


# ------------------------------------------------------------------------------------------ SE3
def _tangents(n, seed, scale=1.0):
    a = S.hash_normal(seed, "tan", n * 6).reshape(n, 6) * np.array([1, 1, 1, 0.5, 0.5, 0.5]) * scale
    return a


@pytest.mark.parametrize("n", [1, 127, 128, 1000, 4099])
def test_se3_forward_ops(n):
    a = _tangents(n, 1, 1.2); b = _tangents(n, 2, 1.2)
    if n >= 1000:
        a[:50, 3:] *= 1e-8                                   # near identity
        a[50:60, 3:] = a[50:60, 3:] / np.linalg.norm(a[50:60, 3:], axis=-1, keepdims=True) * (np.pi - 1e-3)
    X = G.se3_exp(a); Y = G.se3_exp(b)
    Xd, Yd = cu(X.astype(np.float32)), cu(Y.astype(np.float32))
    X32, Y32 = X.astype(np.float32).astype(np.float64), Y.astype(np.float32).astype(np.float64)
    report("se3_exp", ops.se3_exp_fwd(cu(a.astype(np.float32))).cpu().numpy(), G.se3_exp(a.astype(np.float32)), 3e-6, 3e-6)
    report("se3_mul", ops.se3_mul_fwd(Xd, Yd).cpu().numpy(), G.se3_mul(X32, Y32), 3e-6, 3e-6)
    report("se3_inv", ops.se3_inv_fwd(Xd).cpu().numpy(), G.se3_inv(X32), 3e-6, 3e-6)
    got = ops.se3_log_fwd(Xd).cpu().numpy()
    ref = G.se3_log(X32)
    big = np.linalg.norm(ref[:, 3:], axis=-1) > 3.0       # near pi: log is ill-conditioned in float32
    report("se3_log", got[~big], ref[~big], 1e-5, 1e-5)
    report("se3_log(near pi)", got[big], ref[big], 5e-3, 0) if big.any() else None


@pytest.mark.parametrize("n", [5, 300])
def test_se3_backward_ops(n):
    a = _tangents(n, 3); b = _tangents(n, 4)
    a[: n // 5, 3:] *= 1e-4
    X = G.se3_exp(a).astype(np.float32); Y = G.se3_exp(b).astype(np.float32)
    g7 = S.hash_normal(5, "g7", n * 7).reshape(n, 7).astype(np.float32)
    g6 = np.ascontiguousarray(g7[:, :6])
    dX, dY = ops.se3_mul_bwd(cu(g7), cu(X), cu(Y))
    rX, rY = G.se3_mul_backward(g7, X, Y)
    report("mul_bwd dX", dX.cpu().numpy(), rX, 1e-5, 1e-5); report("mul_bwd dY", dY.cpu().numpy(), rY, 1e-5, 1e-5)
    report("inv_bwd", ops.se3_inv_bwd(cu(g7), cu(X)).cpu().numpy(), G.se3_inv_backward(g7, X), 1e-5, 1e-5)
    report("log_bwd", ops.se3_log_bwd(cu(g6), cu(X)).cpu().numpy(), G.se3_log_backward(g6, X), 3e-5, 3e-5)
    a32 = a.astype(np.float32)
    report("exp_bwd", ops.se3_exp_bwd(cu(g7), cu(a32)).cpu().numpy(), G.se3_exp_backward(g7, a32), 3e-5, 3e-5)


def test_geodesic_loss_and_lietorch_gradient():
    from rel_pose_b200.losses import geodesic_loss
    B = 6
    Ps = S.make_poses_numpy(1, B)
    Gn = S.make_poses_numpy(2, B)
    Gt = cu(Gn).requires_grad_(True)
    ltr, lrot, metrics = geodesic_loss(SE3(cu(Ps)), [SE3(Gt)])
    rtr, rrot, _ = G.geodesic_loss(Ps, Gn)
    report("geo_loss_tr", ltr.item(), rtr, 1e-5, 1e-5); report("geo_loss_rot", lrot.item(), rrot, 1e-5, 1e-5)
    assert set(metrics) == {"train_geo_loss_tr", "train_geo_loss_rot"}
    (10.0 * ltr + 10.0 * lrot).backward()
    report("geodesic grad (lietorch convention)", Gt.grad.cpu().numpy(), G.geodesic_loss_grad(Ps, Gn), 2e-4, 2e-4)
    assert torch.all(Gt.grad[..., 6] == 0)


def test_se3_broadcast_mul_and_exp_log_roundtrip():
    a = cu(_tangents(64, 9).astype(np.float32))
    X = SE3.exp(a)
    report("log(exp(a))", X.log().cpu().numpy(), a.cpu().numpy(), 2e-5, 2e-5)
    one = SE3(X.data[:1])
    Z = one * X                                             # [1] x [64] broadcast
    ref = G.se3_mul(np.tile(X.data[:1].cpu().numpy().astype(np.float64), (64, 1)), X.data.cpu().numpy().astype(np.float64))
    report("broadcast mul", Z.data.cpu().numpy(), ref, 3e-6, 3e-6)
    e = SE3(torch.zeros(0, 7, device=DEV))
    assert e.inv().data.shape == (0, 7)                      # empty input is legal


# ------------------------------------------------------------------------------------------ config 3
def test_svd3_random_and_rank_deficient():
    n = 5000
    E = rnd(21, n, 3, 3).astype(np.float64)
    E[:100, :, 2] = 0.0                                       # rank 2
    E[100:150] = np.einsum("ni,nj->nij", E[100:150, :, 0], E[100:150, 0, :])   # rank 1
    E[150:160] = 0.0                                          # zero matrix
    E = E.astype(np.float32)
    U, Sg, V = (t.cpu().numpy().astype(np.float64) for t in ops.svd3(cu(E)))
    _, Sref, _ = G.svd3(E)
    scale = np.maximum(Sref[:, :1], 1e-30)
    print("[parity] svd3 sigma max rel err", (np.abs(Sg - Sref) / scale).max())
    assert (np.abs(Sg - Sref) <= 2e-6 * scale + 1e-30).all()
    assert (np.diff(Sg, axis=-1) <= 1e-7 * scale).all() and (Sg >= 0).all()
    rec = U * Sg[:, None, :] @ np.swapaxes(V, -1, -2)
    assert (np.abs(rec - E).reshape(n, -1).max(-1) <= 3e-6 * scale[:, 0] + 1e-30).all()
    eye = np.eye(3)
    assert np.abs(np.swapaxes(U, -1, -2) @ U - eye).max() < 5e-6
    assert np.abs(np.swapaxes(V, -1, -2) @ V - eye).max() < 5e-6


def test_essential_to_rt_recovers_pose():
    n = 3000
    X = G.se3_exp(_tangents(n, 5, 1.0))
    t = X[:, :3] / np.linalg.norm(X[:, :3], axis=-1, keepdims=True)
    R = G.qmat(X[:, 3:])
    E = (G.hat(t) @ R).astype(np.float32)
    R1, R2, tt = (v.cpu().numpy().astype(np.float64) for v in ops.essential_to_rt(cu(E)))
    for Rc in (R1, R2):
        assert np.abs(np.linalg.det(Rc) - 1).max() < 1e-5
        assert np.abs(np.swapaxes(Rc, -1, -2) @ Rc - np.eye(3)).max() < 1e-5
    e = np.minimum(np.abs(R1 - R).reshape(n, -1).max(-1), np.abs(R2 - R).reshape(n, -1).max(-1))
    et = np.minimum(np.abs(tt - t).max(-1), np.abs(tt + t).max(-1))
    print("[parity] essential_to_rt max R err", e.max(), "max t err", et.max())
    assert e.max() < 2e-5 and et.max() < 2e-5
    # oracle agreement up to the (u_k,v_k) sign ambiguity: {R1,R2} as a set
    o1, o2, ot = G.essential_to_rt(E.astype(np.float64))
    d = np.minimum(np.abs(R1 - o1).reshape(n, -1).max(-1), np.abs(R1 - o2).reshape(n, -1).max(-1))
    assert d.max() < 5e-5
