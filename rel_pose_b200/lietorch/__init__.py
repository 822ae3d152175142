"""`lietorch`-compatible SE3 type backed by the relpose_b200 CUDA kernels.

The reference imports `from lietorch import SE3` (src/model.py:9, src/geom/losses.py via its
callers, train.py:11, demo.py:22) but lietorch (C++/CUDA, pinned 0.2 in environment.yml:20) is a
third-party package.  This module provides the API subset the reference touches -- constructor
from a [...,7] tensor, `.data` (a real tensor, slice-assignable: src/model.py:151), `__getitem__`,
`IdentityLike`, `*`, `.inv()`, `.log()`, `.exp()`, `.detach()` -- with lietorch's autograd
convention (tangent-space gradients of a left perturbation in the first 6 of 7 slots).

`install_as_lietorch()` registers it under the name `lietorch` so the reference's scripts run
unchanged.  Group arithmetic runs on CUDA only (no CPU fallback); indexing / construction work
anywhere.
"""
import sys

import torch

from .. import ops


def _bcast(a, b):
    if a.shape == b.shape:
        return a.contiguous(), b.contiguous()
    shape = torch.broadcast_shapes(a.shape[:-1], b.shape[:-1])
    return a.expand(shape + (7,)).contiguous(), b.expand(shape + (7,)).contiguous()


class _Mul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, Y):
        X = X.contiguous(); Y = Y.contiguous()
        ctx.save_for_backward(X, Y)
        return ops.se3_mul_fwd(X, Y)

    @staticmethod
    def backward(ctx, dZ):
        X, Y = ctx.saved_tensors
        dX, dY = ops.se3_mul_bwd(dZ.contiguous(), X, Y)
        return dX, dY


class _Inv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        X = X.contiguous()
        ctx.save_for_backward(X)
        return ops.se3_inv_fwd(X)

    @staticmethod
    def backward(ctx, dY):
        (X,) = ctx.saved_tensors
        return ops.se3_inv_bwd(dY.contiguous(), X)


class _Log(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        X = X.contiguous()
        ctx.save_for_backward(X)
        return ops.se3_log_fwd(X)

    @staticmethod
    def backward(ctx, da):
        (X,) = ctx.saved_tensors
        return ops.se3_log_bwd(da.contiguous(), X)


class _Exp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        a = a.contiguous()
        ctx.save_for_backward(a)
        return ops.se3_exp_fwd(a)

    @staticmethod
    def backward(ctx, dX):
        (a,) = ctx.saved_tensors
        return ops.se3_exp_bwd(dX.contiguous(), a)


class SE3:
    group_name = "SE3"
    manifold_dim = 6
    embedded_dim = 7

    def __init__(self, data):
        if isinstance(data, SE3):
            data = data.data
        if data.shape[-1] != 7:
            raise ValueError(f"SE3 expects [...,7] data (tx,ty,tz,qx,qy,qz,qw), got {tuple(data.shape)}")
        self.data = data

    # ---- container behaviour -------------------------------------------------------------
    def __repr__(self):
        return f"SE3: size={tuple(self.shape)}, device={self.device}, dtype={self.dtype}"

    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def device(self):
        return self.data.device

    @property
    def dtype(self):
        return self.data.dtype

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, index):
        return SE3(self.data[index])

    def view(self, *dims):
        dims = dims[0] if len(dims) == 1 and isinstance(dims[0], (tuple, list)) else dims
        return SE3(self.data.view(tuple(dims) + (7,)))

    def detach(self):
        return SE3(self.data.detach())

    def to(self, *a, **k):
        return SE3(self.data.to(*a, **k))

    def cuda(self, *a, **k):
        return SE3(self.data.cuda(*a, **k))

    def cpu(self):
        return SE3(self.data.cpu())

    def float(self):
        return SE3(self.data.float())

    def vec(self):
        return self.data

    def translation(self):
        return self.data[..., :3]

    def quaternion(self):
        return self.data[..., 3:]

    @classmethod
    def Identity(cls, *batch_shape, **kwargs):
        shape = batch_shape[0] if len(batch_shape) == 1 and isinstance(batch_shape[0], (tuple, list)) else batch_shape
        d = torch.zeros(tuple(shape) + (7,), **kwargs)
        d[..., 6] = 1.0
        return cls(d)

    @classmethod
    def IdentityLike(cls, G):
        d = torch.zeros_like(G.data)
        d[..., 6] = 1.0
        return cls(d)

    @classmethod
    def InitFromVec(cls, data):
        return cls(data)

    # ---- group arithmetic (CUDA kernels) -------------------------------------------------
    def _f32(self):
        if self.data.dtype != torch.float32:
            raise TypeError("SE3 group ops run in float32 (lietorch itself is fp32/fp64 only)")
        return self.data

    def mul(self, other):
        if not isinstance(other, SE3):
            raise TypeError("SE3 * " + type(other).__name__ + " is not supported (only SE3 * SE3)")
        a, b = _bcast(self._f32(), other._f32())
        return SE3(_Mul.apply(a, b))

    __mul__ = mul

    def inv(self):
        return SE3(_Inv.apply(self._f32()))

    def log(self):
        return _Log.apply(self._f32())

    @classmethod
    def exp(cls, a):
        return cls(_Exp.apply(a))

    def matrix(self):
        """4x4 homogeneous matrices (plain torch indexing; not on the hot path)."""
        t, q = self.data[..., :3], self.data[..., 3:]
        q = q / q.norm(dim=-1, keepdim=True)
        x, y, z, w = q.unbind(-1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)
        R = R.reshape(q.shape[:-1] + (3, 3))
        T = torch.zeros(q.shape[:-1] + (4, 4), dtype=q.dtype, device=q.device)
        T[..., :3, :3] = R
        T[..., :3, 3] = t
        T[..., 3, 3] = 1.0
        return T


def cat(group_list, dim):
    return SE3(torch.cat([g.data for g in group_list], dim=dim))


def stack(group_list, dim):
    return SE3(torch.stack([g.data for g in group_list], dim=dim))


def install_as_lietorch(force=False):
    """Make `import lietorch` / `from lietorch import SE3` resolve to this module."""
    if "lietorch" in sys.modules and not force:
        return sys.modules["lietorch"]
    sys.modules["lietorch"] = sys.modules[__name__]
    return sys.modules[__name__]
