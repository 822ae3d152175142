"""GPU parity of the tensor-core (tcgen05 + TMA + TMEM) GEMM on split-bf16 planes against float64."""
import numpy as np
import pytest
import torch

import relpose_oracle as O
from rel_pose_b200 import ops, synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rnd(seed, *shape, scale=1.0):
    return (S.hash_normal(seed, "tc", int(np.prod(shape))).reshape(shape) * scale).astype(np.float32)


def planes_to_f64(p):
    return p.float().double().sum(0).cpu().numpy()


def test_split_planes_roundtrip():
    x = rnd(1, 1000, 192, scale=3.0)
    p = ops.split_planes(cu(x), 2)
    assert p.dtype == torch.bfloat16 and tuple(p.shape) == (2, 1000, 192)
    p0 = torch.from_numpy(x).to(DEV).bfloat16()
    assert torch.equal(p[0], p0)                                   # plane 0 is RNE bf16(x)
    rec = planes_to_f64(p)
    assert np.abs(rec - x).max() <= 2.0 ** -16 * np.abs(x).max()   # 16 mantissa bits survive
    p1 = ops.split_planes(cu(x), 1)
    assert torch.equal(p1[0], p0)


def test_layernorm_planes():
    x = rnd(2, 333, 192, scale=2.0) + 0.3
    g = 1 + 0.1 * rnd(3, 192); b = 0.1 * rnd(4, 192)
    ref = O.layernorm(x.astype(np.float64), g.astype(np.float64), b.astype(np.float64))
    got = planes_to_f64(ops.layernorm_planes(cu(x), cu(g), cu(b), 1e-6, 2))
    assert np.abs(got - ref).max() < 3e-5 * np.abs(ref).max()
    got1 = planes_to_f64(ops.layernorm_planes(cu(x), cu(g), cu(b), 1e-6, 1))
    assert np.abs(got1 - ref).max() < 5e-3 * np.abs(ref).max()


def _diagnose(name, got, ref):
    e = np.abs(got - ref)
    i, j = np.unravel_index(np.argmax(e), e.shape)
    print(f"[tc-diag] {name}: max err {e.max():.3e} at ({i},{j}) got {got[i, j]:.5f} ref {ref[i, j]:.5f}; "
          f"row-err-profile {e.max(1)[:8]} ... col-err-profile {e.max(0)[:8]}")
    bad_rows = np.nonzero(e.max(1) > 1e-2 * np.abs(ref).max())[0]
    bad_cols = np.nonzero(e.max(0) > 1e-2 * np.abs(ref).max())[0]
    print(f"[tc-diag] bad rows {len(bad_rows)}/{e.shape[0]} first {bad_rows[:16]}  bad cols {len(bad_cols)}/{e.shape[1]} first {bad_cols[:16]}")


@pytest.mark.parametrize("P", [1, 2])
@pytest.mark.parametrize("M,N,K,act,res,pout", [
    (128, 192, 64, 0, False, 0), (128, 192, 192, 0, False, 0), (256, 384, 128, 0, False, 0),
    (1152, 576, 192, 0, False, 0), (1152, 192, 192, 0, True, 0), (1152, 768, 192, 1, False, 2),
    (1152, 192, 768, 0, True, 0), (140, 768, 192, 1, False, 1), (2, 512, 26880, 2, False, 2),
    (3, 14, 512, 0, False, 0), (300, 200, 72, 2, True, 2), (20000, 576, 192, 0, False, 0)])
def test_linear_tc(P, M, N, K, act, res, pout):
    a = rnd(5, M, K); w = rnd(6, N, K, scale=1.0 / np.sqrt(K)); b = rnd(7, N, scale=0.1)
    r = rnd(8, M, N) if res else None
    ap = ops.split_planes(cu(a), P); wp = ops.split_planes(cu(w), P)
    out, outp = ops.linear_tc(ap, wp, cu(b), act=act, residual=cu(r) if res else None, want_f32=True, planes_out=pout)
    torch.cuda.synchronize()
    y = a.astype(np.float64) @ w.astype(np.float64).T + b
    if act == 1:
        y = O.gelu(y)
    elif act == 2:
        y = np.maximum(y, 0)
    if res:
        y = y + r
    got = out.cpu().numpy().astype(np.float64)
    # tcgen05 accumulates in fp32 with truncation: the error grows with the number of chained MMAs (K/16),
    # which is why the K=26880 regressor layer stays on the fp32 SIMT path in the model
    tol = (3e-5 if P == 2 else 2e-2) * np.abs(y).max() * max(1.0, (K / 192.0) ** 0.5)
    err = np.abs(got - y).max()
    print(f"[parity] linear_tc P={P} {M}x{N}x{K} act={act}: max_abs_err={err:.3e} max_ref={np.abs(y).max():.3e} ratio={err / tol:.3f}")
    if not err <= tol:
        _diagnose(f"P={P} {M}x{N}x{K}", got, y)
    assert np.isfinite(got).all() and err <= tol
    if pout:
        gp = planes_to_f64(outp)
        lim = (2.0 ** -15 if pout == 2 else 2.0 ** -7) * np.abs(got).max()
        assert np.abs(gp - got).max() <= lim


def test_linear_tc_only_planes_output():
    a = rnd(9, 256, 192); w = rnd(10, 768, 192, scale=0.07)
    out, outp = ops.linear_tc(ops.split_planes(cu(a), 2), ops.split_planes(cu(w), 2), None, act=1, want_f32=False, planes_out=2)
    assert out is None
    y = O.gelu(a.astype(np.float64) @ w.astype(np.float64).T)
    assert np.abs(planes_to_f64(outp) - y).max() < 5e-5 * np.abs(y).max()
