"""Training path of `ViTEss` (BASELINE.json config 5): train-mode forward (batch-statistics
BatchNorm, saved softmax probabilities) and the backward pass, every FLOP in librelpose_b200.so (csrc/train_ops.cu,
fp32, deterministic).  PyTorch contributes what it contributes to the reference: the autograd graph, gradient
accumulation at fan-outs, views / re-layout copies, and the optimizer the reference's train.py constructs itself.

What is differentiated (reference lines): src/model.py:114-191; vision_transformer.py:188-238 (Essential Matrix Module),
285-296 (CrossBlock), 321-354 (Attention, Block); vit_layers/mlp.py:20-26; extractor.py:51-65; torchvision BasicBlock.
The lietorch SE3 ops of the loss have their own kernels (rel_pose_b200/lietorch).
"""
import ctypes
import os

import torch

from . import _lib, ops

NTOK, EMBED, HEADS, HDIM, EMW = ops.NTOK, ops.EMBED, ops.HEADS, ops.HDIM, ops.EMW
IMG = NTOK * 3 * EMBED                      # elements of one image's qkv block
SS = NTOK * NTOK


def _ctx(t):
    return ops._ctx(t)


def _p(t, offset=0):
    return ctypes.c_void_p(t.data_ptr() + 4 * offset)


def _new(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


def _ws(nbytes, like):
    return torch.empty((max(nbytes, 16) // 4,), dtype=torch.float32, device=like.device)


# ------------------------------------------------------------------------------------------ raw kernels
def gemm(ta, tb, M, N, K, A, a_off, lda, B, b_off, ldb, C, c_off, ldc, alpha=1.0, beta=0.0, bo=1, bi=1,
         sA=(0, 0), sB=(0, 0), sC=(0, 0)):
    """C = alpha op(A) op(B) + beta C on sub-blocks of contiguous float32 tensors (element offsets / strides)."""
    dev, st = _ctx(C)
    _lib.check(_lib.lib().rp_gemm_f32(int(ta), int(tb), M, N, K, float(alpha), _p(A, a_off), lda, _p(B, b_off), ldb,
                                      float(beta), _p(C, c_off), ldc, bo, bi, sA[0], sA[1], sB[0], sB[1], sC[0], sC[1],
                                      dev, st), "rp_gemm")
    ops._count()


def mm(a, b, ta=False, tb=False, alpha=1.0):
    """2-D product of contiguous matrices."""
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[0] if tb else b.shape[1]
    assert (b.shape[1] if tb else b.shape[0]) == K
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    if ta and not tb and K >= 4096 and tiles < 296:
        # Weight gradients dW = dY^T X contract over the (long) row dimension and have a small output: a
        # tile-per-block grid is a handful of blocks (9 for a 64x576 conv filter) each walking 37 632 rows -- 76 % of
        # the training step by ncu.  Split the contraction into S slabs run as one batched launch (each slab's
        # partial product in its own buffer), then add the slabs in a fixed order: deterministic, no atomics.
        S = min((592 + tiles - 1) // tiles, K // 512)
        chunk = ((K + S - 1) // S + 15) // 16 * 16
        S = K // chunk
        rem = K - S * chunk
        part = _new((S, M * N), a)
        gemm(True, False, M, N, chunk, a, 0, a.shape[1], b, 0, b.shape[1], part, 0, N, alpha, bo=S, bi=1,
             sA=(chunk * a.shape[1], 0), sB=(chunk * b.shape[1], 0), sC=(M * N, 0))
        if rem:
            gemm(True, False, M, N, rem, a, S * chunk * a.shape[1], a.shape[1], b, S * chunk * b.shape[1], b.shape[1],
                 part, 0, N, alpha, beta=1.0)
        return colsum(part).reshape(M, N)
    c = _new((M, N), a)
    gemm(ta, tb, M, N, K, a, 0, a.shape[1], b, 0, b.shape[1], c, 0, N, alpha)
    return c


def colsum(a2d, b2d=None):
    rows, cols = a2d.shape
    L = _lib.lib()
    nb = L.rp_colsum_workspace_bytes(rows, cols)
    ws = _ws(nb, a2d)
    out = _new((cols,), a2d)
    dev, st = _ctx(a2d)
    _lib.check(L.rp_colsum_f32(_p(a2d), _p(b2d) if b2d is not None else None, _p(out), rows, cols, _p(ws), nb, dev, st), "rp_colsum")
    ops._count(2)
    return out


def _ew(name, out, *args):
    dev, st = _ctx(out)
    _lib.check(getattr(_lib.lib(), name)(*args, dev, st), name)
    ops._count()
    return out


# ------------------------------------------------------------------------------------------ tensor-core GEMMs of the step
# The three products of every nn.Linear / nn.Conv2d in the step -- y = x W^T, dX = dY W, dW = dY^T X -- run on the tcgen05
# engine of the inference path (rp_linear_tc / rp_linear_tc_splitk / rp_conv2d_tc, split-bf16 operands = fp32 class) whenever
# the shape qualifies; all three are brought to its K-major form C = A B^T:
#   y  = x W^T            A = planes(x)      [M,K]    B = planes(W)     [N,K]
#   dX = dY W             A = planes(dY)     [M,N]    B = planes(W^T)   [K,N]     (re-layout copy of the small weight)
#   dW = dY^T X           A = planes(dY^T)   [N,M]    B = planes(X^T)   [K,M]     (rp_transpose_split_planes, split-K: the
#                                                                                 contraction runs over the M rows)
# Convolutions: forward = the implicit-GEMM kernel; dX of a stride-1 convolution = the same kernel on dY with the
# spatially flipped, channel-transposed filter and padding K-1-p; dW = dY^T im2col(X).  RELPOSE_TRAIN_TC=0 keeps the
# fp32 SIMT kernels (A/B measurements); tiny problems (the pose regressor at 6 rows) always stay there.
TRAIN_TC = os.environ.get("RELPOSE_TRAIN_TC", "1") != "0"
TRAIN_DW_TC = TRAIN_TC and os.environ.get("RELPOSE_TRAIN_DW_TC", "1") != "0"         # implicit-GEMM conv weight gradients
TRAIN_LIN_DW_TC = TRAIN_DW_TC and os.environ.get("RELPOSE_TRAIN_LIN_DW_TC", "1") != "0"   # ... and nn.Linear weight gradients on the same kernel
TRAIN_FLASH = TRAIN_TC and os.environ.get("RELPOSE_TRAIN_FLASH", "1") != "0"      # attention gradients without 576 x 576 tensors
_TCP = 2                                   # bf16x3
TOKENS_GRAD_HOOK = None                    # callable(grad) -> None, see forward_train


def _lin_fwd(x2, w, b, act=ops.ACT_NONE):
    M, K = x2.shape
    if TRAIN_TC and M >= 64 and K % 8 == 0 and (M * K) % 8 == 0 and (w.numel() % 8) == 0:
        return ops.linear_tc(ops.split_planes(x2, _TCP), ops.split_planes(w, _TCP), b, act=act)[0]
    if TRAIN_TC and K >= 4096 and K % 64 == 0 and w.shape[0] % 64 == 0:
        # pose_regressor.0 at a handful of rows (26 880 -> 512): weight-bandwidth bound, the split-K kernel of the inference path
        return ops.linear_tc_splitk(ops.split_planes(x2, _TCP), ops.split_planes(w, _TCP), b, act=act)
    return ops.linear(x2, w, b, act=act)


def _tc_rows(M, K, w):
    return TRAIN_TC and M >= 64 and K % 8 == 0 and (M * K) % 8 == 0 and (w.numel() % 8) == 0


def _lin_dx(dy2, w, dyp=None):
    """dy2 [M,N], w [N,K] -> dy2 @ w  [M,K]   (dyp: the split planes of dy2 when the caller already has them)"""
    M, N = dy2.shape
    if _tc_rows(M, N, w):
        return ops.linear_tc(dyp if dyp is not None else ops.split_planes(dy2, _TCP), ops.split_planes(w.t().contiguous(), _TCP), None)[0]
    return mm(dy2, w)


def _lin_dw(dy2, x2, dyp=None, xp=None):
    """dy2 [M,N], x2 [M,K] -> dy2^T @ x2  [N,K]   (x2 may be None when its planes xp are given)"""
    M, N = dy2.shape
    K = x2.shape[1] if x2 is not None else xp.shape[-1]
    if TRAIN_LIN_DW_TC and ops.linear_dw_tc_supported(M, N, K):
        # implicit-GEMM kernel of the convolution weight gradients: X and dY read in place as MN-major operands
        return ops.linear_dw_tc(xp if xp is not None else ops.split_planes(x2, _TCP), dyp if dyp is not None else ops.split_planes(dy2, _TCP))
    if TRAIN_TC and M >= 512 and M % 8 == 0 and K % 4 == 0:
        return ops.linear_tc_splitk(ops.transpose_split_planes(dy2, _TCP), ops.transpose_split_planes(x2, _TCP))
    return mm(dy2, x2, ta=True)


# ------------------------------------------------------------------------------------------ autograd functions
class LinearFn(torch.autograd.Function):
    """y = x W^T + b   (vision_transformer.py:323,331; mlp.py:21,24; model.py:91-98).  When the weight gradient runs on
    the implicit-GEMM kernel the forward keeps the bf16 planes of x it built for its own product (same bytes as x) and
    the backward splits dy once for both of its products."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        N, K = w.shape
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        wd = w.detach().contiguous()
        bd = b.detach().contiguous() if b is not None else None
        ctx.planes = _tc_rows(M, K, w) and TRAIN_LIN_DW_TC and ops.linear_dw_tc_supported(M, N, K)
        ctx.xshape = tuple(x.shape)
        if ctx.planes:
            xp = ops.split_planes(x2, _TCP)
            y = ops.linear_tc(xp, ops.split_planes(wd, _TCP), bd)[0]
            ctx.save_for_backward(xp, w)
        else:
            y = _lin_fwd(x2, wd, bd)
            ctx.save_for_backward(x, w)
        return y.reshape(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        N, K = w.shape
        dy2 = dy.contiguous().reshape(-1, N)
        M = dy2.shape[0]
        dyp = ops.split_planes(dy2, _TCP) if (ctx.planes or (ctx.needs_input_grad[0] and _tc_rows(M, N, w))) else None
        dx = _lin_dx(dy2, w.detach().contiguous(), dyp).reshape(ctx.xshape) if ctx.needs_input_grad[0] else None
        if not ctx.needs_input_grad[1]:
            dw = None
        elif ctx.planes:
            dw = _lin_dw(dy2, None, dyp, x)
        else:
            dw = _lin_dw(dy2, x.reshape(-1, K), dyp)
        db = colsum(dy2) if ctx.needs_input_grad[2] else None
        return dx, dw, db


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        z = z.contiguous()
        ctx.save_for_backward(z)
        y = torch.empty_like(z)
        return _ew("rp_gelu_fwd_f32", y, _p(z), _p(y), z.numel())

    @staticmethod
    def backward(ctx, dy):
        (z,) = ctx.saved_tensors
        dy = dy.contiguous()
        dz = torch.empty_like(z)
        return _ew("rp_gelu_bwd_f32", dz, _p(dy), _p(z), _p(dz), z.numel())


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a = a.contiguous(); b = b.contiguous()
        out = torch.empty_like(a)
        return _ew("rp_axpby_f32", out, ctypes.c_float(1.0), _p(a), ctypes.c_float(1.0), _p(b), _p(out), a.numel())

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class AddPosFn(torch.autograd.Function):
    """x [n,576,192] + pos_embed [1,576,192]   (model.py:172)"""

    @staticmethod
    def forward(ctx, x, pos):
        x = x.contiguous()
        out = torch.empty_like(x)
        ctx.reps = x.shape[0]
        return _ew("rp_add_bcast_rows_f32", out, _p(x), _p(pos.detach().contiguous()), _p(out), x.shape[0] * NTOK, EMBED, NTOK)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dpos = _new((1, NTOK, EMBED), dy)
        _ew("rp_sum_over_period_f32", dpos, _p(dy), _p(dpos), ctx.reps, NTOK, EMBED)
        return dy, dpos


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g, b, eps):
        x = x.contiguous()
        cols = x.shape[-1]
        rows = x.numel() // cols
        y = torch.empty_like(x)
        mean = _new((rows,), x); rstd = _new((rows,), x)
        dev, st = _ctx(x)
        _lib.check(_lib.lib().rp_layernorm_train_fwd_f32(_p(x), _p(g.detach()), _p(b.detach()), _p(y), _p(mean), _p(rstd), rows, cols,
                                                         float(eps), dev, st), "rp_layernorm_train_fwd")
        ops._count()
        ctx.save_for_backward(x, g, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        cols = x.shape[-1]
        rows = x.numel() // cols
        dx = torch.empty_like(x)
        dg = _new((cols,), x); db = _new((cols,), x)
        L = _lib.lib()
        nb = L.rp_layernorm_bwd_workspace_bytes(rows, cols)
        ws = _ws(nb, x)
        dev, st = _ctx(x)
        _lib.check(L.rp_layernorm_bwd_f32(_p(dy), _p(x), _p(g.detach()), _p(mean), _p(rstd), _p(dx), _p(dg), _p(db), rows, cols,
                                          _p(ws), nb, dev, st), "rp_layernorm_bwd")
        ops._count(2)
        return dx, dg, db, None


class AttentionMaterialisedFn(torch.autograd.Function):
    """softmax(q k^T / 8) v per image and head on qkv [n,576,576]   (vision_transformer.py:323-329), fp32 SIMT with the
    probabilities materialised (12 images = 48 MB per block): the A/B partner of AttentionFn (RELPOSE_TRAIN_FLASH=0)."""

    @staticmethod
    def forward(ctx, qkv):
        qkv = qkv.contiguous()
        n = qkv.shape[0]
        S = _new((n, HEADS, NTOK, NTOK), qkv)
        gemm(False, True, NTOK, NTOK, HDIM, qkv, 0, 3 * EMBED, qkv, EMBED, 3 * EMBED, S, 0, NTOK,
             bo=n, bi=HEADS, sA=(IMG, HDIM), sB=(IMG, HDIM), sC=(HEADS * SS, SS))
        P = torch.empty_like(S)
        dev, st = _ctx(qkv)
        _lib.check(_lib.lib().rp_softmax_rows_fwd_f32(_p(S), _p(P), n * HEADS * NTOK, NTOK, 0.125, dev, st), "rp_softmax_rows_fwd")
        ops._count()
        out = _new((n, NTOK, EMBED), qkv)
        gemm(False, False, NTOK, HDIM, NTOK, P, 0, NTOK, qkv, 2 * EMBED, 3 * EMBED, out, 0, EMBED,
             bo=n, bi=HEADS, sA=(HEADS * SS, SS), sB=(IMG, HDIM), sC=(NTOK * EMBED, HDIM))
        ctx.save_for_backward(qkv, P)
        return out

    @staticmethod
    def backward(ctx, dO):
        qkv, P = ctx.saved_tensors
        dO = dO.contiguous()
        n = qkv.shape[0]
        dqkv = torch.empty_like(qkv)
        # dV = P^T dO
        gemm(True, False, NTOK, HDIM, NTOK, P, 0, NTOK, dO, 0, EMBED, dqkv, 2 * EMBED, 3 * EMBED,
             bo=n, bi=HEADS, sA=(HEADS * SS, SS), sB=(NTOK * EMBED, HDIM), sC=(IMG, HDIM))
        # dP = dO V^T ; dS = softmax'(dP)
        dP = torch.empty_like(P)
        gemm(False, True, NTOK, NTOK, HDIM, dO, 0, EMBED, qkv, 2 * EMBED, 3 * EMBED, dP, 0, NTOK,
             bo=n, bi=HEADS, sA=(NTOK * EMBED, HDIM), sB=(IMG, HDIM), sC=(HEADS * SS, SS))
        dev, st = _ctx(qkv)
        _lib.check(_lib.lib().rp_softmax_rows_bwd_f32(_p(dP), _p(P), _p(dP), n * HEADS * NTOK, NTOK, 0.125, 0, dev, st), "rp_softmax_rows_bwd")
        ops._count()
        dS = dP                                                      # in place: dS = 0.125 * P .* (dP - <dP,P>)
        # dQ = dS K ; dK = dS^T Q
        gemm(False, False, NTOK, HDIM, NTOK, dS, 0, NTOK, qkv, EMBED, 3 * EMBED, dqkv, 0, 3 * EMBED,
             bo=n, bi=HEADS, sA=(HEADS * SS, SS), sB=(IMG, HDIM), sC=(IMG, HDIM))
        gemm(True, False, NTOK, HDIM, NTOK, dS, 0, NTOK, qkv, 0, 3 * EMBED, dqkv, EMBED, 3 * EMBED,
             bo=n, bi=HEADS, sA=(HEADS * SS, SS), sB=(IMG, HDIM), sC=(IMG, HDIM))
        return dqkv


class AttentionFn(torch.autograd.Function):
    """softmax(q k^T / 8) v per image and head on qkv [n,576,576]   (vision_transformer.py:323-329), flash style in both
    directions: the forward is the inference kernel (tcgen05, split-bf16 operands) which also saves the log-sum-exp of
    every score row; the backward recomputes the probabilities tile by tile (csrc/attention_bwd_tc.cu).  Saved for the
    backward: the bf16 planes of qkv, the output and 3 x 576 floats per image -- no 576 x 576 tensor exists."""

    @staticmethod
    def forward(ctx, qkv):
        planes = ops.split_planes(qkv.contiguous(), _TCP)
        out, lse = ops.self_attention_tc_lse(planes)
        ctx.save_for_backward(planes, out, lse)
        return out

    @staticmethod
    def backward(ctx, dO):
        planes, out, lse = ctx.saved_tensors
        return ops.attention_bwd_tc(planes, dO.contiguous(), out, lse)


def attention(qkv):
    return (AttentionFn if TRAIN_FLASH else AttentionMaterialisedFn).apply(qkv)


class EssentialMaterialisedFn(torch.autograd.Function):
    """Essential Matrix Module core on materialised 576 x 576 matrices, fp32 SIMT: the A/B partner of EssentialFn
    (RELPOSE_TRAIN_FLASH=0).  (vision_transformer.py:198-223): qkv [2B,576,576], pos [B,576,6] -> F [B,2,3,70,70].
    dir 0: S = q2 k1^T/8, V' = [v1|pos];  dir 1: S = q1 k2^T/8, V' = [v2|pos];  A = softmax(S,-1)*softmax(S,-2);
    F = V'^T A V'.  The positional columns receive no gradient (9.2)."""

    @staticmethod
    def forward(ctx, qkv, pos):
        qkv = qkv.contiguous()
        n = qkv.shape[0]
        B = n // 2
        L = _lib.lib()
        dev, st = _ctx(qkv)
        S = _new((B, 2, HEADS, NTOK, NTOK), qkv)
        for d in range(2):          # queries from image 2b+1-d, keys from image 2b+d
            gemm(False, True, NTOK, NTOK, HDIM, qkv, (1 - d) * IMG, 3 * EMBED, qkv, d * IMG + EMBED, 3 * EMBED,
                 S, d * HEADS * SS, NTOK, bo=B, bi=HEADS, sA=(2 * IMG, HDIM), sB=(2 * IMG, HDIM), sC=(2 * HEADS * SS, SS))
        R = torch.empty_like(S); C = torch.empty_like(S)
        _lib.check(L.rp_softmax_rows_fwd_f32(_p(S), _p(R), B * 2 * HEADS * NTOK, NTOK, 0.125, dev, st), "rp_softmax_rows_fwd")
        _lib.check(L.rp_softmax_cols_fwd_f32(_p(S), _p(C), B * 2 * HEADS, NTOK, 0.125, dev, st), "rp_softmax_cols_fwd")
        A = S                                                        # S is dead: reuse its storage for A = R .* C
        _lib.check(L.rp_mul_f32(_p(R), _p(C), _p(A), A.numel(), dev, st), "rp_mul")
        Vp = _new((n, HEADS, NTOK, EMW), qkv)
        _lib.check(L.rp_concat_vpos_f32(_p(qkv), _p(pos.contiguous()), _p(Vp), n, dev, st), "rp_concat_vpos")
        ops._count(4)
        T = _new((B, 2, HEADS, NTOK, EMW), qkv)
        F = _new((B, 2, HEADS, EMW, EMW), qkv)
        VP = HEADS * NTOK * EMW                                      # elements of one image's V'
        for d in range(2):          # V' of image 2b+d
            gemm(False, False, NTOK, EMW, NTOK, A, d * HEADS * SS, NTOK, Vp, d * VP, EMW, T, d * HEADS * NTOK * EMW, EMW,
                 bo=B, bi=HEADS, sA=(2 * HEADS * SS, SS), sB=(2 * VP, NTOK * EMW), sC=(2 * HEADS * NTOK * EMW, NTOK * EMW))
            gemm(True, False, EMW, EMW, NTOK, Vp, d * VP, EMW, T, d * HEADS * NTOK * EMW, EMW, F, d * HEADS * EMW * EMW, EMW,
                 bo=B, bi=HEADS, sA=(2 * VP, NTOK * EMW), sB=(2 * HEADS * NTOK * EMW, NTOK * EMW),
                 sC=(2 * HEADS * EMW * EMW, EMW * EMW))
        ctx.save_for_backward(qkv, R, C, A, Vp, T)
        return F

    @staticmethod
    def backward(ctx, dF):
        qkv, R, C, A, Vp, T = ctx.saved_tensors
        dF = dF.contiguous()
        n = qkv.shape[0]
        B = n // 2
        L = _lib.lib()
        dev, st = _ctx(qkv)
        VP = HEADS * NTOK * EMW
        TT = HEADS * NTOK * EMW
        FF = HEADS * EMW * EMW
        dT = torch.empty_like(T)
        dVp = _new((n, HEADS, NTOK, EMW), qkv)
        dA = torch.empty_like(A)
        for d in range(2):
            # F = V'^T T:  dT = V' dF ;  dV' = T dF^T
            gemm(False, False, NTOK, EMW, EMW, Vp, d * VP, EMW, dF, d * FF, EMW, dT, d * TT, EMW,
                 bo=B, bi=HEADS, sA=(2 * VP, NTOK * EMW), sB=(2 * FF, EMW * EMW), sC=(2 * TT, NTOK * EMW))
            gemm(False, True, NTOK, EMW, EMW, T, d * TT, EMW, dF, d * FF, EMW, dVp, d * VP, EMW,
                 bo=B, bi=HEADS, sA=(2 * TT, NTOK * EMW), sB=(2 * FF, EMW * EMW), sC=(2 * VP, NTOK * EMW))
            # T = A V':  dA = dT V'^T ;  dV' += A^T dT
            gemm(False, True, NTOK, NTOK, EMW, dT, d * TT, EMW, Vp, d * VP, EMW, dA, d * HEADS * SS, NTOK,
                 bo=B, bi=HEADS, sA=(2 * TT, NTOK * EMW), sB=(2 * VP, NTOK * EMW), sC=(2 * HEADS * SS, SS))
            gemm(True, False, NTOK, EMW, NTOK, A, d * HEADS * SS, NTOK, dT, d * TT, EMW, dVp, d * VP, EMW, beta=1.0,
                 bo=B, bi=HEADS, sA=(2 * HEADS * SS, SS), sB=(2 * TT, NTOK * EMW), sC=(2 * VP, NTOK * EMW))
        # A = R .* C:  dR = dA .* C, dC = dA .* R;  dS = rowsoftmax'(dR) + colsoftmax'(dC)   (same scaled S in both)
        dR = torch.empty_like(A)
        _lib.check(L.rp_mul_f32(_p(dA), _p(C), _p(dR), dA.numel(), dev, st), "rp_mul")
        _lib.check(L.rp_mul_f32(_p(dA), _p(R), _p(dA), dA.numel(), dev, st), "rp_mul")          # dA now holds dC
        dS = _new(A.shape, qkv)
        _lib.check(L.rp_softmax_rows_bwd_f32(_p(dR), _p(R), _p(dS), B * 2 * HEADS * NTOK, NTOK, 0.125, 0, dev, st), "rp_softmax_rows_bwd")
        _lib.check(L.rp_softmax_cols_bwd_f32(_p(dA), _p(C), _p(dS), B * 2 * HEADS, NTOK, 0.125, 1, dev, st), "rp_softmax_cols_bwd")
        ops._count(4)
        dqkv = torch.empty_like(qkv)
        for d in range(2):
            # S = q_(1-d) k_d^T:  dq = dS k ;  dk = dS^T q   (each (image, q/k/v) slot is written exactly once)
            gemm(False, False, NTOK, HDIM, NTOK, dS, d * HEADS * SS, NTOK, qkv, d * IMG + EMBED, 3 * EMBED, dqkv, (1 - d) * IMG, 3 * EMBED,
                 bo=B, bi=HEADS, sA=(2 * HEADS * SS, SS), sB=(2 * IMG, HDIM), sC=(2 * IMG, HDIM))
            gemm(True, False, NTOK, HDIM, NTOK, dS, d * HEADS * SS, NTOK, qkv, (1 - d) * IMG, 3 * EMBED, dqkv, d * IMG + EMBED, 3 * EMBED,
                 bo=B, bi=HEADS, sA=(2 * HEADS * SS, SS), sB=(2 * IMG, HDIM), sC=(2 * IMG, HDIM))
        _lib.check(L.rp_scatter_dv_f32(_p(dVp), _p(dqkv), n, dev, st), "rp_scatter_dv")
        ops._count()
        return dqkv, None


class EssentialFn(torch.autograd.Function):
    """Essential Matrix Module core (vision_transformer.py:198-223), flash style in both directions: qkv [2B,576,576],
    pos [B,576,6] -> F [B,2,3,70,70].  Forward = the inference kernels (rp_essential_tc); saved for the backward: the bf16
    planes of qkv, pos and the row / column log-sum-exp vectors; csrc/em_bwd_tc.cu recomputes S and the dual-softmax matrix
    per tile.  The positional columns receive no gradient."""

    @staticmethod
    def forward(ctx, qkv, pos):
        planes = ops.split_planes(qkv.contiguous(), _TCP)
        pos = pos.contiguous()
        bil, lse2 = ops.essential_tc_train(planes, pos)
        ctx.save_for_backward(planes, pos, lse2)
        return bil

    @staticmethod
    def backward(ctx, dF):
        planes, pos, lse2 = ctx.saved_tensors
        return ops.em_bwd_tc(planes, pos, lse2, dF.contiguous()), None


def essential(qkv, pos):
    return (EssentialFn if TRAIN_FLASH else EssentialMaterialisedFn).apply(qkv, pos)


class EmProjectFn(torch.autograd.Function):
    """proj_fundamental on Z[b,c,h*70+a] = F[b,h,a,c] and the (fundamental_2, fundamental_1) flip
    (vision_transformer.py:229-238,292-294): bil [B,2,3,70,70] -> [2B,70,192], slot 2b + (1 - dir)."""

    @staticmethod
    def forward(ctx, bil, w, b):
        bil = bil.contiguous()
        ctx.save_for_backward(bil, w)
        return ops.em_project(bil, w.detach().contiguous(), b.detach().contiguous())

    @staticmethod
    def backward(ctx, dout):
        bil, w = ctx.saved_tensors
        B = bil.shape[0]
        KK = HEADS * EMW                                             # 210
        w = w.detach().contiguous()
        # (b, dir) order of the incoming gradient: a re-layout copy, no arithmetic
        dz = dout.reshape(B, 2, EMW, EMBED).flip(1).contiguous()
        dbil = torch.empty_like(bil)
        # dbil_mat[k][c] = sum_o W[o][k] dz[c][o]      (bil_mat = bil[b,dir] viewed as [210,70])
        gemm(True, True, KK, EMW, EMBED, w, 0, KK, dz, 0, EMBED, dbil, 0, EMW,
             bo=B * 2, bi=1, sA=(0, 0), sB=(EMW * EMBED, 0), sC=(KK * EMW, 0))
        # dW[o][k] = sum_{b,dir} sum_c dz[c][o] bil_mat[k][c]: per-(b,dir) products, then a fixed-order sum
        part = _new((B * 2, EMBED, KK), bil)
        gemm(True, True, EMBED, KK, EMW, dz, 0, EMBED, bil, 0, EMW, part, 0, KK,
             bo=B * 2, bi=1, sA=(EMW * EMBED, 0), sB=(KK * EMW, 0), sC=(EMBED * KK, 0))
        dw = _new((EMBED, KK), bil)
        _ew("rp_sum_over_period_f32", dw, _p(part), _p(dw), B * 2, EMBED, KK)
        db = colsum(dz.reshape(-1, EMBED))
        return dbil, dw, db


class NormalizePoseFn(torch.autograd.Function):
    """normalize_preds (model.py:145-152): quaternion / max(|q|, 0.01) on pose 1, pose 0 := Gs[:, :1]."""

    @staticmethod
    def forward(ctx, raw, gs):
        raw = raw.contiguous()
        ctx.save_for_backward(raw)
        return ops.normalize_pose(raw, gs.contiguous())

    @staticmethod
    def backward(ctx, dout):
        (raw,) = ctx.saved_tensors
        dout = dout.contiguous()
        draw = torch.empty_like(raw)
        dev, st = _ctx(raw)
        _lib.check(_lib.lib().rp_normalize_pose_bwd_f32(_p(dout), _p(raw), _p(draw), raw.shape[0], dev, st), "rp_normalize_pose_bwd")
        ops._count()
        return draw, None


class LinearReluFn(torch.autograd.Function):
    """relu(x W^T + b)   (pose regressor, model.py:91-95)"""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        N, K = w.shape
        y = _lin_fwd(x.reshape(-1, K), w.detach().contiguous(), b.detach().contiguous(), act=ops.ACT_RELU).reshape(x.shape[:-1] + (N,))
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        dz = torch.empty_like(dy)
        _ew("rp_relu_bwd_f32", dz, _p(dy), _p(y), _p(dz), dy.numel())
        N, K = w.shape
        dz2 = dz.reshape(-1, N)
        return _lin_dx(dz2, w.detach().contiguous()).reshape(x.shape), _lin_dw(dz2, x.reshape(-1, K)), colsum(dz2)


class ConvFn(torch.autograd.Function):
    """nn.Conv2d on NHWC activations; the weight argument is the nn.Conv2d parameter [O,C,KH,KW] (re-laid out to
    [O,KH,KW,Cp] by a kernel in the forward, back by a copy in the backward)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad):
        x = x.contiguous()
        C = x.shape[-1]
        wp = ops.permute_conv_weight(weight.detach(), C)
        O, KH, KW, Cp = wp.shape
        b = bias.detach().contiguous() if bias is not None else None
        use_tc = TRAIN_TC and C % 64 == 0 and O in (64, 128, 192)
        xp = ops.split_planes(x, _TCP) if use_tc else None
        if use_tc:
            y = ops.conv2d_tc(xp, ops.split_planes(wp.reshape(O, -1), _TCP), KH, KW, None, b, stride, pad,
                              ops.ACT_NONE, want_f32=True, planes_out=0)[0]
        else:
            y = ops.conv2d_nhwc(x, wp, None, b, stride, pad, ops.ACT_NONE)
        # the implicit-GEMM weight gradient reads the bf16 planes the forward made anyway: the float32 input is not kept
        dw_tc = use_tc and TRAIN_DW_TC and ops.conv_dw_tc_supported(C, O, KH, KW, stride)
        ctx.save_for_backward(xp if dw_tc else x, wp)
        ctx.geom = (stride, pad, tuple(weight.shape), bias is not None, dw_tc, tuple(x.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wp = ctx.saved_tensors
        stride, pad, wshape, has_bias, dw_tc, xshape = ctx.geom
        dy = dy.contiguous()
        n, H, W, C = xshape
        O, KH, KW, Cp = wp.shape
        M = dy.numel() // O
        K = KH * KW * Cp
        dy2 = dy.reshape(M, O)
        dev, st = _ctx(dy)
        L = _lib.lib()
        dx = dwt = db = None
        dx_tc = ctx.needs_input_grad[0] and TRAIN_TC and stride == 1 and O % 64 == 0 and C in (64, 128, 192) and KH - 1 - pad >= 0
        dyp = ops.split_planes(dy, _TCP) if (dx_tc or (dw_tc and ctx.needs_input_grad[1])) else None
        if ctx.needs_input_grad[1]:
            if dw_tc:
                dwp = ops.conv_dw_tc(x, dyp, KH, KW, pad, stride)                         # [O,KH,KW,C], no im2col
            elif TRAIN_TC and M >= 512 and M % 8 == 0:
                # the stem (4 input channels): im2col written directly as the transposed bf16 planes the split-K GEMM reads
                colsT = torch.empty((_TCP, K, M), dtype=torch.bfloat16, device=x.device)
                _lib.check(L.rp_im2col_t_planes_bf16(_p(x), _p(colsT), _TCP, n, H, W, C, KH, KW, stride, pad, dev, st), "rp_im2col_t_planes")
                ops._count()
                dwp = ops.linear_tc_splitk(ops.transpose_split_planes(dy2, _TCP), colsT).reshape(O, KH, KW, Cp)
                del colsT
            else:
                cols = _new((M, K), x)
                _lib.check(L.rp_im2col_nhwc_f32(_p(x), _p(cols), n, H, W, C, KH, KW, stride, pad, dev, st), "rp_im2col")
                ops._count()
                dwp = _lin_dw(dy2, cols).reshape(O, KH, KW, Cp)                   # [O, K]
                del cols
            dwt = dwp[..., :wshape[1]].permute(0, 3, 1, 2).contiguous()          # re-layout copy back to [O,C,KH,KW]
        if ctx.needs_input_grad[0]:
            if dx_tc:
                # dX = "full" correlation of dY with the flipped filter, input and output channels swapped
                wf = wp.flip(1, 2).permute(3, 1, 2, 0).contiguous()              # [C][KH][KW][O]: re-layout copy of the filter
                dx = ops.conv2d_tc(dyp, ops.split_planes(wf.reshape(C, -1), _TCP), KH, KW, None, None, 1,
                                   KH - 1 - pad, ops.ACT_NONE, want_f32=True, planes_out=0)[0]
            else:
                dcols = _lin_dx(dy2, wp.reshape(O, K))                           # [M, K]
                dx = torch.empty(xshape, dtype=torch.float32, device=dy.device)
                _lib.check(L.rp_col2im_nhwc_f32(_p(dcols), _p(dx), n, H, W, C, KH, KW, stride, pad, dev, st), "rp_col2im")
                ops._count()
        if has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2)
        return dx, dwt, db, None, None


_stem_idx_cache = {}


def _stem_window_index(device):
    """position of conv1.weight[o, c, ky, kx] inside the [4 x 64] space-to-depth window filter of rp_stem_weight_windows_f32
    (conv_aux.cu): row a = (ky + 1) // 2, column b * 16 + (dy * 2 + dx) * 3 + c with b = (kx + 1) // 2"""
    key = str(device)
    if key not in _stem_idx_cache:
        c, ky, kx = torch.meshgrid(torch.arange(3), torch.arange(7), torch.arange(7), indexing="ij")
        a, dy, b, dx = (ky + 1) // 2, (ky + 1) % 2, (kx + 1) // 2, (kx + 1) % 2
        _stem_idx_cache[key] = (a * 64 + b * 16 + (dy * 2 + dx) * 3 + c).reshape(-1).to(device)
    return _stem_idx_cache[key]


class StemConvFn(torch.autograd.Function):
    """resnet.conv1 (7x7 / 2 on 3 channels, model.py:127) on the tensor cores in both directions: the images go straight
    into the space-to-depth window planes of the inference path (A1 fused, rp_preprocess_stem_windows_*), where the stem is
    a 4 x 1 convolution over 64 'channels'; its weight gradient is the implicit-GEMM kernel on the same planes, gathered
    back to [64,3,7,7].  No gradient flows to the images."""

    @staticmethod
    def forward(ctx, images, weight):
        O = weight.shape[0]
        xw = ops.preprocess_stem_windows(images, _TCP)                                   # [P, 2B, 115, 112, 64]
        w2 = ops.stem_weight_windows(weight.detach())                                    # [O, 4, 1, 64]
        y = ops.conv2d_tc(xw, ops.split_planes(w2.reshape(O, -1), _TCP), 4, 1, None, None, 1, 0, ops.ACT_NONE,
                          want_f32=True, planes_out=0)[0]
        ctx.save_for_backward(xw)
        ctx.O = O
        return y

    @staticmethod
    def backward(ctx, dy):
        (xw,) = ctx.saved_tensors
        O = ctx.O
        dw2 = ops.conv_dw_tc(xw, ops.split_planes(dy.contiguous(), _TCP), 4, 1, 0, 1)    # [O, 4, 1, 64]
        dw = dw2.reshape(O, 256).index_select(1, _stem_window_index(dy.device)).reshape(O, 3, 7, 7)
        return None, dw


class BatchNormTrainFn(torch.autograd.Function):
    """y = act(BN_train(x) + residual) on NHWC [.., C]; updates the running statistics like nn.BatchNorm2d.train()."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, residual, relu):
        x = x.contiguous()
        C = x.shape[-1]
        M = x.numel() // C
        L = _lib.lib()
        dev, st = _ctx(x)
        mean = _new((C,), x); var = _new((C,), x)
        nb = L.rp_bn_workspace_bytes(M, C)
        ws = _ws(nb, x)
        _lib.check(L.rp_bn_train_stats_f32(_p(x), _p(mean), _p(var), _p(running_mean), _p(running_var), float(momentum), M, C,
                                           _p(ws), nb, dev, st), "rp_bn_train_stats")
        ops.bump_param_generation()          # running statistics changed behind torch's version counters
        y = torch.empty_like(x)
        res = residual.contiguous() if residual is not None else None
        _lib.check(L.rp_bn_apply_f32(_p(x), _p(mean), _p(var), _p(gamma.detach()), _p(beta.detach()), _p(res) if res is not None else None,
                                     _p(y), M, C, float(eps), int(relu), dev, st), "rp_bn_apply")
        ops._count(3)
        ctx.save_for_backward(x, gamma, mean, var, y)
        ctx.cfg = (float(eps), bool(relu), residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, var, y = ctx.saved_tensors
        eps, relu, has_res = ctx.cfg
        dy = dy.contiguous()
        C = x.shape[-1]
        M = x.numel() // C
        L = _lib.lib()
        dev, st = _ctx(x)
        # dz = dy .* (y > 0) is formed inside the two BatchNorm passes; it is only written out for the residual branch
        dz = torch.empty_like(dy) if (relu and has_res) else (dy if has_res else None)
        dx = torch.empty_like(x)
        dg = _new((C,), x); db = _new((C,), x)
        nb = L.rp_bn_workspace_bytes(M, C)
        ws = _ws(nb, x)
        _lib.check(L.rp_bn_bwd_f32(_p(dy), _p(y) if relu else None, _p(x), _p(mean), _p(var), _p(gamma.detach()), eps, _p(dx),
                                   _p(dz) if (relu and has_res) else None, _p(dg), _p(db), M, C, _p(ws), nb, dev, st), "rp_bn_bwd")
        ops._count(3)
        return dx, dg, db, None, None, None, None, dz, None


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        ctx.save_for_backward(x)
        return ops.maxpool3x3s2_nhwc(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        n, H, W, C = x.shape
        dx = torch.empty_like(x)
        dev, st = _ctx(x)
        L = _lib.lib()
        nb = L.rp_maxpool3x3s2_bwd_workspace_bytes(n, H, W, C)
        ws = _ws(nb, x)
        _lib.check(L.rp_maxpool3x3s2_bwd_f32(_p(dy), _p(x), _p(dx), n, H, W, C, _p(ws), nb, dev, st), "rp_maxpool3x3s2_bwd")
        ops._count(2)
        return dx


# ------------------------------------------------------------------------------------------ the training forward
def _bn(x, bn, residual=None, relu=True):
    y = BatchNormTrainFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                               bn.momentum if bn.momentum is not None else 0.1, bn.eps, residual, relu)
    bn.num_batches_tracked += 1
    return y


def _conv(x, conv):
    return ConvFn.apply(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0])


def _basic_block(x, blk):
    y = _bn(_conv(x, blk.conv1), blk.bn1)
    ident = x if blk.downsample is None else _bn(_conv(x, blk.downsample[0]), blk.downsample[1], relu=False)
    return _bn(_conv(y, blk.conv2), blk.bn2, residual=ident, relu=True)


def _mlp(x, mlp):
    return LinearFn.apply(GeluFn.apply(LinearFn.apply(x, mlp.fc1.weight, mlp.fc1.bias)), mlp.fc2.weight, mlp.fc2.bias)


def _ln(x, ln):
    return LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps)


def forward_train(model, images, Gs, intrinsics):
    """ViTEss.forward in training mode with gradients (train.py:155): returns the [B,2,7] pose tensor."""
    B = images.shape[0]
    images = images.contiguous()
    if images.dtype != torch.uint8:
        images = images.float()
    kxy = None
    if intrinsics is not None:
        intrinsics, kxy, flags = model.update_intrinsics(images.shape, intrinsics)
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture (train_synthetic.GraphedStep): no host read inside the graph; the caller checks the flag
            # word after each replay (GraphedStep.check_flags)
            model.__dict__["_train_flags"] = flags
            f = 0
        else:
            f = int(flags.item())
        if f & 1:
            raise AssertionError("intrinsics must be identical for both views of a pair (vision_transformer.py:117)")
        if f & 2:
            raise ValueError("principal point is in upper left, not setup for this right now (vision_transformer.py:124-126)")
    r, e, vt = model.resnet, model.extractor_final_conv, model.fusion_transformer
    if TRAIN_DW_TC and tuple(r.conv1.weight.shape) == (64, 3, 7, 7) and r.conv1.bias is None:
        x = StemConvFn.apply(images, r.conv1.weight)                  # A1 + conv1 (no gradient w.r.t. the images)
    else:
        x = _conv(ops.preprocess_nhwc4(images), r.conv1)
    x = _bn(x, r.bn1)                                                 # A2
    x = MaxPoolFn.apply(x)
    for blk in (r.layer1[0], r.layer1[1], r.layer2[0], r.layer2[1]):
        x = _basic_block(x, blk)
    y = _bn(_conv(x, e.conv1), e.norm1)                               # A3 (extractor.py:51-65)
    y = _bn(_conv(y, e.conv2), e.norm2)
    x = _bn(_conv(x, e.downsample[0]), e.norm3, residual=y, relu=True)
    x = x.reshape(2 * B, NTOK, EMBED)                                 # A4: NHWC output is already [2B,576,192]
    if TOKENS_GRAD_HOOK is not None and x.requires_grad:
        # fires when the backward pass reaches the CNN: every gradient of the transformer, the Essential Matrix Module and
        # the pose regressor exists by then (train_synthetic.py starts their all-reduce here, under the CNN's backward)
        x.register_hook(TOKENS_GRAD_HOOK)
    x = AddPosFn.apply(x, vt.pos_embed)
    depth = model.transformer_depth
    for i in range(depth - 1):                                        # A5
        blk = vt.blocks[i]
        qkv = LinearFn.apply(_ln(x, blk.norm1), blk.attn.qkv.weight, blk.attn.qkv.bias)
        a = attention(qkv)
        x = AddFn.apply(x, LinearFn.apply(a, blk.attn.proj.weight, blk.attn.proj.bias))
        x = AddFn.apply(x, _mlp(_ln(x, blk.norm2), blk.mlp))
    blk = vt.blocks[depth - 1]                                        # A6-A8
    ca = blk.cross_attn
    qkv = LinearFn.apply(_ln(x, blk.norm1), ca.qkv.weight, ca.qkv.bias)
    pos = ops.posenc(B, kxy, x.device)
    bil = essential(qkv, pos)
    f = EmProjectFn.apply(bil, ca.proj_fundamental.weight, ca.proj_fundamental.bias)
    f = AddFn.apply(f, _mlp(_ln(f, blk.norm2), blk.mlp))
    feat = _ln(f, vt.norm).reshape(B, -1)                             # A9
    reg = model.pose_regressor
    h = LinearReluFn.apply(feat, reg[0].weight, reg[0].bias)
    h = LinearReluFn.apply(h, reg[2].weight, reg[2].bias)
    raw = LinearFn.apply(h, reg[4].weight, reg[4].bias).reshape(B, 2, 7)
    return NormalizePoseFn.apply(raw, Gs.data)                        # A10
