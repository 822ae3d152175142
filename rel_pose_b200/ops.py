"""Tensor-level wrappers over the C ABI.  PyTorch is plumbing only here: it owns device memory and
streams; every computation below is a call into librelpose_b200.so.  Inputs must be CUDA tensors --
there is no CPU or eager-PyTorch fallback.
"""
import ctypes
import os

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
CONV_HALO = os.environ.get("RELPOSE_CONV_HALO", "1") != "0"      # halo variant of the 3x3 / 64-channel convolutions
NTOK, EMBED, HEADS, HDIM, NPOS, EMW = 576, 192, 3, 64, 6, 70

_launch_count = 0     # number of kernels launched through the library (bench.py reports it)
_param_generation = 0   # bumped whenever the library writes parameters / buffers through raw pointers (see below)


def param_generation():
    """Torch's `tensor._version` does not see writes made through `data_ptr()`: the fused optimizer (optim.py) and the
    train-mode BatchNorm kernels (train_path.py) update parameters / running statistics that way.  Every such write
    bumps this counter, and every cache of derived parameter data (folded BatchNorm, bf16 planes, transposes) carries
    it in its tag, so an eval forward after a training step never reuses stale weights."""
    return _param_generation


def bump_param_generation():
    global _param_generation
    _param_generation += 1

_timer = None         # optional StageTimer: CUDA-event timing of every library call (bench.py)


def launches():
    return _launch_count


class StageTimer:
    """Collects (name, start_event, end_event, flops, bytes) for calls made while installed.
    Events are recorded on the stream the kernels are launched on (torch's current stream)."""

    def __init__(self):
        self.records = []
        self._open = None

    def begin(self, name, flops=0.0, nbytes=0.0):
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        self._open = (name, ev0, ev1, flops, nbytes)

    def end(self):
        name, ev0, ev1, flops, nbytes = self._open
        ev1.record()
        self.records.append((name, ev0, ev1, flops, nbytes))
        self._open = None

    def summary(self):
        """name -> dict(calls, ms, flops, bytes); call after torch.cuda.synchronize()."""
        out = {}
        for name, ev0, ev1, flops, nbytes in self.records:
            d = out.setdefault(name, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
            d["calls"] += 1
            d["ms"] += ev0.elapsed_time(ev1)
            d["flops"] += flops
            d["bytes"] += nbytes
        return out


def set_timer(timer):
    global _timer
    _timer = timer


def _count(n=1):
    global _launch_count
    _launch_count += n
    if _timer is not None and _timer._open is not None:
        _timer.end()


def _tbegin(name, flops=0.0, nbytes=0.0):
    if _timer is not None:
        _timer.begin(name, flops, nbytes)


def _tend():
    if _timer is not None and _timer._open is not None:
        _timer.end()


def _req(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise _lib.RelposeLibraryError(
            f"{name}: tensor is on {t.device}; rel_pose_b200 runs on CUDA devices only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _ctx(t):
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    return dev, ctypes.c_void_p(stream)


# ------------------------------------------------------------------------------------------ A1
def preprocess(images):
    """[B,2,3,H,W] (float32 BGR 0..255, or uint8) -> [2B,3,224,224] float32."""
    if images.dtype == torch.uint8:
        _req(images, "images", torch.uint8)
        fn = _lib.lib().rp_preprocess_u8
    else:
        _req(images, "images")
        fn = _lib.lib().rp_preprocess_f32
    B, V, C, H, W = images.shape
    assert C == 3
    out = torch.empty((B * V, 3, 224, 224), dtype=torch.float32, device=images.device)
    dev, st = _ctx(images)
    _tbegin("preprocess", 0.0, float(images.numel() * images.element_size()) + 4.0 * B * V * 3 * 224 * 224)
    _lib.check(fn(_p(images), _p(out), B * V, H, W, dev, st), "rp_preprocess")
    _count()
    return out


def intrinsics_prepare(intrinsics, H, W):
    """In place rescale of [B,2,4] intrinsics to the 24x24 grid; returns (kxy [B,2], flags int32[1])."""
    _req(intrinsics, "intrinsics")
    B = intrinsics.shape[0]
    assert tuple(intrinsics.shape[1:]) == (2, 4)
    kxy = torch.empty((B, 2), dtype=torch.float32, device=intrinsics.device)
    flags = torch.zeros((1,), dtype=torch.int32, device=intrinsics.device)
    dev, st = _ctx(intrinsics)
    _lib.check(_lib.lib().rp_intrinsics_prepare_f32(_p(intrinsics), _p(kxy), _p(flags), B, H, W, dev, st),
               "rp_intrinsics_prepare")
    _count()
    return kxy, flags


# ------------------------------------------------------------------------------------------ A2/A3
def preprocess_nhwc4(images):
    """[B,2,3,H,W] (float32 BGR 0..255, or uint8) -> [2B,224,224,4] float32 NHWC (4th channel 0)."""
    if images.dtype == torch.uint8:
        _req(images, "images", torch.uint8)
        fn = _lib.lib().rp_preprocess_nhwc4_u8
    else:
        _req(images, "images")
        fn = _lib.lib().rp_preprocess_nhwc4_f32
    B, V, C, H, W = images.shape
    assert C == 3
    out = torch.empty((B * V, 224, 224, 4), dtype=torch.float32, device=images.device)
    dev, st = _ctx(images)
    _tbegin("preprocess", 0.0, float(images.numel() * images.element_size()) + 4.0 * B * V * 4 * 224 * 224)
    _lib.check(fn(_p(images), _p(out), B * V, H, W, dev, st), "rp_preprocess_nhwc4")
    _count()
    return out


def preprocess_stem_windows(images, P):
    """A1 fused with the stem's space-to-depth window layout: [B,2,3,H,W] -> bf16 planes [P,2B,115,112,64]."""
    if images.dtype == torch.uint8:
        _req(images, "images", torch.uint8)
        fn = _lib.lib().rp_preprocess_stem_windows_u8
    else:
        _req(images, "images")
        fn = _lib.lib().rp_preprocess_stem_windows_f32
    B, V, C, H, W = images.shape
    assert C == 3
    out = torch.empty((P, B * V, 115, 112, 64), dtype=torch.bfloat16, device=images.device)
    dev, st = _ctx(images)
    _tbegin("preprocess_stem_windows", 0.0, float(images.numel() * images.element_size()) + 2.0 * out.numel())
    _lib.check(fn(_p(images), _p(out), B * V, H, W, P, dev, st), "rp_preprocess_stem_windows")
    _count()
    return out


def stem_compact_supported(device=0):
    return bool(_lib.lib().rp_stem_compact_supported(int(device)))


def preprocess_stem_compact(images, P):
    """A1 into the compact space-to-depth layout: [B,2,3,H,W] -> bf16 planes [P,2B,115,116,16] (csrc/stem_pool_tc.cu)."""
    if images.dtype == torch.uint8:
        _req(images, "images", torch.uint8)
        fn = _lib.lib().rp_preprocess_stem_compact_u8
    else:
        _req(images, "images")
        fn = _lib.lib().rp_preprocess_stem_compact_f32
    B, V, C, H, W = images.shape
    assert C == 3
    out = torch.empty((P, B * V, 115, 116, 16), dtype=torch.bfloat16, device=images.device)
    dev, st = _ctx(images)
    _tbegin("preprocess_stem_compact", 0.0, float(images.numel() * images.element_size()) + 2.0 * out.numel())
    _lib.check(fn(_p(images), _p(out), B * V, H, W, P, dev, st), "rp_preprocess_stem_compact")
    _count()
    return out


def stem_pool_tc(z_planes, w_planes, scale, shift, planes_out, want_f32=True):
    """conv1 + bn1 + relu + maxpool in one launch: z_planes = window tensor [P,n,115,112,64] or compact image
    [P,n,115,116,16]; w_planes [P,64,256] -> (float32 [n,56,56,64] | None, bf16 planes [planes_out,n,56,56,64] | None)."""
    _req(z_planes, "z_planes", torch.bfloat16); _req(w_planes, "w_planes", torch.bfloat16)
    _req(scale, "scale"); _req(shift, "shift")
    P, n = z_planes.shape[0], z_planes.shape[1]
    compact = tuple(z_planes.shape[2:]) == (115, 116, 16)
    assert compact or tuple(z_planes.shape[2:]) == (115, 112, 64), z_planes.shape
    assert tuple(w_planes.shape) == (P, 64, 256)
    out = torch.empty((n, 56, 56, 64), dtype=torch.float32, device=z_planes.device) if want_f32 else None
    outp = torch.empty((planes_out, n, 56, 56, 64), dtype=torch.bfloat16, device=z_planes.device) if planes_out else None
    dev, st = _ctx(z_planes)
    M = n * 112 * 112
    _tbegin(f"stem_pool_tc{'x3' if P == 2 else ''}", 2.0 * M * 64 * 256,
            2.0 * z_planes.numel() + (4.0 if want_f32 else 0.0) * n * 56 * 56 * 64 + 2.0 * planes_out * n * 56 * 56 * 64)
    _lib.check(_lib.lib().rp_stem_pool_tc(_p(z_planes), int(compact), _p(w_planes), _p(scale), _p(shift), _p(out), _p(outp), n, P,
                                          int(planes_out), dev, st), "rp_stem_pool_tc")
    _count()
    return out, outp


def stem_weight_windows(w):
    """conv1.weight [O,3,7,7] -> [O,4,1,64] (the 4x4 space-to-depth kernel, KW folded into C); parameter preparation."""
    w = _req(w.detach().contiguous(), "w")
    O = w.shape[0]
    assert tuple(w.shape[1:]) == (3, 7, 7)
    out = torch.empty((O, 4, 1, 64), dtype=torch.float32, device=w.device)
    dev, st = _ctx(w)
    _lib.check(_lib.lib().rp_stem_weight_windows_f32(_p(w), _p(out), O, dev, st), "rp_stem_weight_windows")
    _count()
    return out


def conv2d_nhwc(x, w, scale=None, shift=None, stride=1, pad=0, act=ACT_NONE, res_pre=None, res_post=None,
                res_post_rows=0):
    """x [n,H,W,C] NHWC, w [O,KH,KW,C] -> [n,Ho,Wo,O]; y = act(conv*scale+shift+res_pre)+res_post."""
    _req(x, "x"); _req(w, "w")
    n, H, W, C = x.shape
    O, KH, KW, C2 = w.shape
    assert C == C2, (x.shape, w.shape)
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    for t, nm in ((scale, "scale"), (shift, "shift"), (res_pre, "res_pre"), (res_post, "res_post")):
        if t is not None:
            _req(t, nm)
    y = torch.empty((n, Ho, Wo, O), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    ws_bytes = L.rp_conv2d_workspace_bytes(n, H, W, C, O, KH, KW, stride, pad)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=x.device) if ws_bytes else None
    dev, st = _ctx(x)
    M, K = n * Ho * Wo, KH * KW * C
    _tbegin(f"conv[{O}x{KH}x{KW}x{C}/s{stride}]", 2.0 * M * O * K, 4.0 * (x.numel() + w.numel() + M * O))
    _lib.check(L.rp_conv2d_nhwc_f32(_p(x), _p(w), _p(scale), _p(shift), _p(res_pre), _p(res_post), int(res_post_rows),
                                    _p(y), n, H, W, C, O, KH, KW, stride, pad, int(act), _p(ws), ws_bytes, dev, st),
               "rp_conv2d_nhwc")
    _count(2 if ws_bytes else 1)
    return y


def maxpool3x3s2_nhwc(x):
    _req(x, "x")
    n, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((n, Ho, Wo, C), dtype=torch.float32, device=x.device)
    dev, st = _ctx(x)
    _tbegin("maxpool", 0.0, 4.0 * (x.numel() + y.numel()))
    _lib.check(_lib.lib().rp_maxpool3x3s2_nhwc_f32(_p(x), _p(y), n, H, W, C, dev, st), "rp_maxpool3x3s2")
    _count()
    return y


def permute_conv_weight(w, Cp=None):
    """nn.Conv2d weight [O,C,KH,KW] -> [O,KH,KW,Cp] (zero padded channels); parameter preparation."""
    w = _req(w.detach().contiguous(), "w")
    O, C, KH, KW = w.shape
    Cp = Cp or ((C + 3) // 4) * 4
    out = torch.empty((O, KH, KW, Cp), dtype=torch.float32, device=w.device)
    dev, st = _ctx(w)
    _lib.check(_lib.lib().rp_permute_conv_weight_f32(_p(w), _p(out), O, C, KH, KW, Cp, dev, st), "rp_permute_conv_weight")
    _count()
    return out


def bn_fold(bn, conv_bias=None):
    """(scale, shift) of an eval-mode nn.BatchNorm2d folded with the preceding conv bias."""
    g = _req(bn.weight.detach().contiguous(), "bn.weight")
    C = g.numel()
    scale = torch.empty((C,), dtype=torch.float32, device=g.device)
    shift = torch.empty_like(scale)
    cb = conv_bias.detach().contiguous() if conv_bias is not None else None
    dev, st = _ctx(g)
    _lib.check(_lib.lib().rp_bn_fold_f32(_p(g), _p(bn.bias.detach().contiguous()), _p(bn.running_mean.contiguous()),
                                         _p(bn.running_var.contiguous()), _p(cb), float(bn.eps), _p(scale), _p(shift),
                                         C, dev, st), "rp_bn_fold")
    _count()
    return scale, shift


# ------------------------------------------------------------------------------------------ A4
def tokens_posembed(fmap, pos_embed):
    """[n,192,24,24] -> [n,576,192] + pos_embed."""
    _req(fmap, "fmap"); _req(pos_embed, "pos_embed")
    n = fmap.shape[0]
    assert fmap.numel() == n * EMBED * NTOK and pos_embed.numel() == NTOK * EMBED
    x = torch.empty((n, NTOK, EMBED), dtype=torch.float32, device=fmap.device)
    dev, st = _ctx(fmap)
    _tbegin("tokens_posembed", 0.0, 8.0 * n * NTOK * EMBED)
    _lib.check(_lib.lib().rp_tokens_posembed_f32(_p(fmap), _p(pos_embed), _p(x), n, dev, st), "rp_tokens_posembed")
    _count()
    return x


def layernorm(x, gamma, beta, eps=1e-6):
    _req(x, "x"); _req(gamma, "gamma"); _req(beta, "beta")
    cols = x.shape[-1]
    rows = x.numel() // cols
    y = torch.empty_like(x)
    dev, st = _ctx(x)
    _tbegin("layernorm", 0.0, 8.0 * rows * cols)
    _lib.check(_lib.lib().rp_layernorm_f32(_p(x), _p(gamma), _p(beta), _p(y), rows, cols, float(eps), dev, st),
               "rp_layernorm")
    _count()
    return y


def linear(x, weight, bias=None, act=ACT_NONE, residual=None, out=None):
    """act(x @ weight.T + bias) + residual through rp_linear_f32."""
    _req(x, "x"); _req(weight, "weight")
    N, K = weight.shape
    assert x.shape[-1] == K
    M = x.numel() // K
    if bias is not None:
        _req(bias, "bias")
    if residual is not None:
        _req(residual, "residual")
        assert residual.numel() == M * N
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    ws_bytes = L.rp_linear_workspace_bytes(M, N, K)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=x.device) if ws_bytes else None
    dev, st = _ctx(x)
    _tbegin(f"linear[{N}x{K}]", 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N))
    _lib.check(L.rp_linear_f32(_p(x), _p(weight), _p(bias), _p(residual), _p(out), M, N, K, int(act),
                               _p(ws), ws_bytes, dev, st), "rp_linear")
    _count(2 if ws_bytes else 1)
    return out


# ------------------------------------------------------------------------------------------ tensor-core path
def split_planes(x, P=2):
    """float32 [...] -> bf16 planes [P, ...] (plane 0 = bf16(x), plane 1 = bf16(x - plane0))."""
    _req(x, "x")
    n = x.numel()
    out = torch.empty((P,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    dev, st = _ctx(x)
    _tbegin("split_planes", 0.0, 4.0 * n + 2.0 * P * n)
    _lib.check(_lib.lib().rp_split_planes_bf16(_p(x), _p(out), n, P, dev, st), "rp_split_planes")
    _count()
    return out


def transpose_split_planes(x, P=2):
    """float32 [R, C] -> bf16 planes of the transpose [P, C, R] (R even)."""
    _req(x, "x")
    assert x.dim() == 2
    R, C = x.shape
    out = torch.empty((P, C, R), dtype=torch.bfloat16, device=x.device)
    dev, st = _ctx(x)
    _tbegin("transpose_split_planes", 0.0, 4.0 * R * C + 2.0 * P * R * C)
    _lib.check(_lib.lib().rp_transpose_split_planes_bf16(_p(x), _p(out), R, C, P, dev, st), "rp_transpose_split_planes")
    _count()
    return out


def layernorm_planes(x, gamma, beta, eps=1e-6, P=2):
    _req(x, "x"); _req(gamma, "gamma"); _req(beta, "beta")
    cols = x.shape[-1]
    rows = x.numel() // cols
    out = torch.empty((P,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    dev, st = _ctx(x)
    _tbegin("layernorm_planes", 0.0, 4.0 * rows * cols + 2.0 * P * rows * cols)
    _lib.check(_lib.lib().rp_layernorm_planes_bf16(_p(x), _p(gamma), _p(beta), _p(out), rows, cols, float(eps), P, dev, st),
               "rp_layernorm_planes")
    _count()
    return out


def linear_tc(a_planes, w_planes, bias=None, act=ACT_NONE, residual=None, want_f32=True, planes_out=0):
    """tcgen05 GEMM: a_planes [P,...,K] bf16, w_planes [P,N,K] bf16 -> (float32 [...,N] | None, planes | None)."""
    _req(a_planes, "a_planes", torch.bfloat16); _req(w_planes, "w_planes", torch.bfloat16)
    P = a_planes.shape[0]
    assert w_planes.shape[0] == P and w_planes.dim() == 3
    N, K = w_planes.shape[1], w_planes.shape[2]
    assert a_planes.shape[-1] == K
    lead = tuple(a_planes.shape[1:-1])
    M = a_planes[0].numel() // K
    if bias is not None:
        _req(bias, "bias")
    if residual is not None:
        _req(residual, "residual")
        assert residual.numel() == M * N
    out = torch.empty(lead + (N,), dtype=torch.float32, device=a_planes.device) if want_f32 else None
    outp = torch.empty((planes_out,) + lead + (N,), dtype=torch.bfloat16, device=a_planes.device) if planes_out else None
    dev, st = _ctx(a_planes)
    _tbegin(f"linear_tc{'x3' if P == 2 else ''}[{N}x{K}]", 2.0 * M * N * K,
            2.0 * P * (M * K + N * K) + (4.0 * M * N if want_f32 else 0.0) + 2.0 * planes_out * M * N)
    _lib.check(_lib.lib().rp_linear_tc(_p(a_planes), _p(w_planes), _p(bias), _p(residual), _p(out), _p(outp), M, N, K, P,
                                       int(planes_out), int(act), dev, st), "rp_linear_tc")
    _count()
    return out, outp


def regressor_tail(h, w1t, b1, w2, b2):
    """pose_regressor[2:5]: h [B,512] (post-ReLU layer 0) -> [B,14]; w1t = layer-1 weight transposed [in,out]."""
    for t, nm in ((h, "h"), (w1t, "w1t"), (b1, "b1"), (w2, "w2"), (b2, "b2")):
        _req(t, nm)
    B, hidden = h.shape
    n_out = w2.shape[0]
    assert tuple(w1t.shape) == (hidden, hidden) and w2.shape[1] == hidden
    out = torch.empty((B, n_out), dtype=torch.float32, device=h.device)
    dev, st = _ctx(h)
    _tbegin("regressor_tail", 2.0 * B * hidden * (hidden + n_out), 4.0 * (hidden * hidden + n_out * hidden + B * hidden))
    _lib.check(_lib.lib().rp_regressor_tail_f32(_p(h), _p(w1t), _p(b1), _p(w2), _p(b2), _p(out), B, hidden, n_out, dev, st),
               "rp_regressor_tail")
    _count()
    return out


def linear_tc_splitk(a_planes, w_planes, bias=None, act=ACT_NONE):
    """Split-K tcgen05 GEMM for skinny weight-bandwidth-bound layers: a_planes [P,M,K] bf16, w_planes [P,N,K] -> float32 [M,N]."""
    _req(a_planes, "a_planes", torch.bfloat16); _req(w_planes, "w_planes", torch.bfloat16)
    P, M, K = a_planes.shape
    N = w_planes.shape[1]
    assert w_planes.shape[0] == P and w_planes.shape[2] == K
    if bias is not None:
        _req(bias, "bias")
    L = _lib.lib()
    ws_bytes = L.rp_linear_tc_splitk_workspace_bytes(M, N, K, None)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=a_planes.device)
    out = torch.empty((M, N), dtype=torch.float32, device=a_planes.device)
    dev, st = _ctx(a_planes)
    _tbegin(f"linear_tc_splitk{'x3' if P == 2 else ''}[{N}x{K}]", 2.0 * M * N * K, 2.0 * P * (M * K + N * K) + 4.0 * M * N)
    _lib.check(L.rp_linear_tc_splitk(_p(a_planes), _p(w_planes), _p(bias), _p(out), M, N, K, P, int(act), _p(ws), ws_bytes, dev, st),
               "rp_linear_tc_splitk")
    _count(2)
    return out


def ln_linear_tc(x, gamma, beta, eps, w_planes, bias=None, want_f32=False, planes_out=0):
    """Fused `linear(layernorm(x))` on tcgen05 (csrc/ln_linear_tc.cu): x float32 [...,192], w_planes bf16 [P,N,192]
    -> (float32 [...,N] | None, bf16 planes [planes_out,...,N] | None)."""
    _req(x, "x"); _req(gamma, "gamma"); _req(beta, "beta")
    _req(w_planes, "w_planes", torch.bfloat16)
    P, N, K = w_planes.shape
    assert x.shape[-1] == K and (want_f32 or planes_out)
    if bias is not None:
        _req(bias, "bias")
    lead = tuple(x.shape[:-1])
    M = x.numel() // K
    out = torch.empty(lead + (N,), dtype=torch.float32, device=x.device) if want_f32 else None
    outp = torch.empty((planes_out,) + lead + (N,), dtype=torch.bfloat16, device=x.device) if planes_out else None
    dev, st = _ctx(x)
    _tbegin(f"ln_linear_tc{'x3' if P == 2 else ''}[{N}x{K}]", 2.0 * M * N * K,
            4.0 * M * K + 2.0 * P * N * K + (4.0 * M * N if want_f32 else 0.0) + 2.0 * planes_out * M * N)
    _lib.check(_lib.lib().rp_ln_linear_tc(_p(x), _p(gamma), _p(beta), float(eps), _p(w_planes), _p(bias), _p(out), _p(outp),
                                          M, N, K, P, int(planes_out), dev, st), "rp_ln_linear_tc")
    _count()
    return out, outp


def planes_linear_tc(xn_planes, w_planes, bias=None, planes_out=0, residual=None, ln_next=None):
    """The K = 192 GEMM of csrc/ln_linear_tc.cu fed with LayerNorm planes its producer wrote (rp_ln_linear_tc_ex):
    xn_planes bf16 [P,...,192], w_planes bf16 [P,N,192].
      residual is None:  -> bf16 planes [planes_out,...,N]                 (the QKV projection)
      residual given:    -> (float32 [...,192] = A W^T + bias + residual,  (the attention projection + skip)
                             bf16 planes [P,...,192] of LayerNorm(out; *ln_next) | None)"""
    _req(xn_planes, "xn_planes", torch.bfloat16); _req(w_planes, "w_planes", torch.bfloat16)
    P, N, K = w_planes.shape
    assert xn_planes.shape[0] == P and xn_planes.shape[-1] == K
    lead = tuple(xn_planes.shape[1:-1])
    M = xn_planes[0].numel() // K
    if bias is not None:
        _req(bias, "bias")
    dev, st = _ctx(xn_planes)
    L = _lib.lib()
    if residual is None:
        assert planes_out
        outp = torch.empty((planes_out,) + lead + (N,), dtype=torch.bfloat16, device=xn_planes.device)
        _tbegin(f"planes_linear_tc{'x3' if P == 2 else ''}[{N}x{K}]", 2.0 * M * N * K,
                2.0 * P * M * K + 2.0 * P * N * K + 2.0 * planes_out * M * N)
        _lib.check(L.rp_ln_linear_tc_ex(None, _p(xn_planes), None, None, 0.0, _p(w_planes), _p(bias), None, None, _p(outp), None,
                                        None, None, 0.0, M, N, K, P, int(planes_out), dev, st), "rp_ln_linear_tc_ex")
        _count()
        return outp
    _req(residual, "residual")
    assert N == K and residual.numel() == M * N and bias is not None
    out = torch.empty(lead + (N,), dtype=torch.float32, device=xn_planes.device)
    lnp, g2, b2, e2 = None, None, None, 0.0
    if ln_next is not None:
        g2, b2, e2 = ln_next
        _req(g2, "ln_next gamma"); _req(b2, "ln_next beta")
        lnp = torch.empty((P,) + lead + (N,), dtype=torch.bfloat16, device=xn_planes.device)
    _tbegin(f"proj_ln_tc{'x3' if P == 2 else ''}[{N}x{K}]", 2.0 * M * N * K,
            2.0 * P * M * K + 2.0 * P * N * K + 8.0 * M * N + (2.0 * P * M * N if lnp is not None else 0.0))
    _lib.check(L.rp_ln_linear_tc_ex(None, _p(xn_planes), None, None, 0.0, _p(w_planes), _p(bias), _p(residual), _p(out), None,
                                    _p(lnp), _p(g2), _p(b2), float(e2), M, N, K, P, 0, dev, st), "rp_ln_linear_tc_ex")
    _count()
    return out, lnp


def mlp_tc(x, gamma, beta, eps, w1_planes, b1, w2_planes, b2, xn_planes=None, ln_next=None, inplace=False):
    """Fused `x + fc2(gelu(fc1(layernorm(x))))` on tcgen05 (csrc/mlp_tc.cu): x float32 [...,192],
    w1_planes bf16 [P,768,192], w2_planes bf16 [P,192,768] -> float32 [...,192].
    xn_planes: layernorm(x) already available as bf16 planes [P,...,192] (written by the producer of x).
    ln_next = (gamma, beta, eps): also return LayerNorm(out) with these weights as bf16 planes -> (out, planes).
    inplace: x is overwritten with the result and returned (the kernel then adds fc2's tile to x at the memory side with
    a TMA reduce-add instead of loading the residual rows and storing the sums); same bits as the out-of-place call."""
    _req(x, "x"); _req(b1, "b1"); _req(b2, "b2")
    _req(w1_planes, "w1_planes", torch.bfloat16); _req(w2_planes, "w2_planes", torch.bfloat16)
    P, hidden, dim = w1_planes.shape
    assert tuple(w2_planes.shape) == (P, dim, hidden) and x.shape[-1] == dim
    M = x.numel() // dim
    if xn_planes is not None:
        _req(xn_planes, "xn_planes", torch.bfloat16)
        assert xn_planes.shape[0] == P and xn_planes.numel() == P * M * dim
    else:
        _req(gamma, "gamma"); _req(beta, "beta")
    out = x if (inplace and ln_next is None) else torch.empty_like(x)
    lnp, g2, bt2, e2 = None, None, None, 0.0
    if ln_next is not None:
        g2, bt2, e2 = ln_next
        _req(g2, "ln_next gamma"); _req(bt2, "ln_next beta")
        lnp = torch.empty((P,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    dev, st = _ctx(x)
    _tbegin(f"mlp_fused_tc{'x3' if P == 2 else ''}", 4.0 * M * dim * hidden,
            8.0 * M * dim + 4.0 * P * dim * hidden + (2.0 * P * M * dim if xn_planes is not None else 0.0) +
            (2.0 * P * M * dim if lnp is not None else 0.0))
    _lib.check(_lib.lib().rp_mlp_tc_ex(_p(x), _p(xn_planes), _p(gamma), _p(beta), float(eps), _p(w1_planes), _p(b1),
                                       _p(w2_planes), _p(b2), _p(out), _p(lnp), _p(g2), _p(bt2), float(e2), M, dim, hidden, P,
                                       dev, st), "rp_mlp_tc")
    _count()
    return out if ln_next is None else (out, lnp)


def conv2d_tc(x_planes, w_planes, KH, KW, scale=None, shift=None, stride=1, pad=0, act=ACT_NONE, res_pre=None,
              res_post=None, res_post_rows=0, want_f32=True, planes_out=0):
    """Tensor-core conv: x_planes [P,n,H,W,C] bf16, w_planes [P,O,KH*KW*C] bf16 -> (f32 [n,Ho,Wo,O] | None, planes | None)."""
    _req(x_planes, "x_planes", torch.bfloat16); _req(w_planes, "w_planes", torch.bfloat16)
    P, n, H, W, C = x_planes.shape
    O = w_planes.shape[1]
    assert w_planes.shape[0] == P and w_planes.shape[2] == KH * KW * C
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    for t, nm in ((scale, "scale"), (shift, "shift"), (res_pre, "res_pre"), (res_post, "res_post")):
        if t is not None:
            _req(t, nm)
    out = torch.empty((n, Ho, Wo, O), dtype=torch.float32, device=x_planes.device) if want_f32 else None
    outp = torch.empty((planes_out, n, Ho, Wo, O), dtype=torch.bfloat16, device=x_planes.device) if planes_out else None
    dev, st = _ctx(x_planes)
    M, K = n * Ho * Wo, KH * KW * C
    L = _lib.lib()
    if (CONV_HALO and res_post is None and act in (ACT_NONE, ACT_RELU)
            and L.rp_conv3x3_halo_supported(H, W, C, O, KH, KW, stride, pad)):
        # layers bound by L2 -> SM traffic: one halo box per tile, taps as shifted descriptors (csrc/conv_halo_tc.cu)
        _tbegin(f"conv_halo_tc{'x3' if P == 2 else ''}[{O}x{KH}x{KW}x{C}/s{stride}]", 2.0 * M * O * K,
                2.0 * P * (x_planes[0].numel() + O * K) + (4.0 * M * O if want_f32 else 0.0) + 2.0 * planes_out * M * O)
        _lib.check(L.rp_conv3x3_halo_tc(_p(x_planes), _p(w_planes), _p(scale), _p(shift), _p(res_pre), _p(out), _p(outp),
                                        n, H, W, C, O, P, int(planes_out), int(act), dev, st), "rp_conv3x3_halo_tc")
        _count()
        return out, outp
    _tbegin(f"conv_tc{'x3' if P == 2 else ''}[{O}x{KH}x{KW}x{C}/s{stride}]", 2.0 * M * O * K,
            2.0 * P * (x_planes[0].numel() + O * K) + (4.0 * M * O if want_f32 else 0.0) + 2.0 * planes_out * M * O)
    _lib.check(_lib.lib().rp_conv2d_tc(_p(x_planes), _p(w_planes), _p(scale), _p(shift), _p(res_pre), _p(res_post),
                                       int(res_post_rows), _p(out), _p(outp), n, H, W, C, O, KH, KW, stride, pad, P,
                                       int(planes_out), int(act), dev, st), "rp_conv2d_tc")
    _count()
    return out, outp


def maxpool3x3s2_planes(x, P, want_f32=True):
    """NHWC float32 -> (float32 | None, bf16 planes [P,n,Ho,Wo,C])."""
    _req(x, "x")
    n, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((n, Ho, Wo, C), dtype=torch.float32, device=x.device) if want_f32 else None
    yp = torch.empty((P, n, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
    dev, st = _ctx(x)
    _tbegin("maxpool", 0.0, 4.0 * x.numel() + (4.0 if want_f32 else 0.0) * n * Ho * Wo * C + 2.0 * P * n * Ho * Wo * C)
    _lib.check(_lib.lib().rp_maxpool3x3s2_planes(_p(x), _p(y), _p(yp), P, n, H, W, C, dev, st), "rp_maxpool3x3s2_planes")
    _count()
    return y, yp


def self_attention(qkv, cross=False):
    """qkv [n,576,576] -> [n,576,192].  cross=True: image i attends to the keys/values of image i^1 (--noess)."""
    _req(qkv, "qkv")
    n = qkv.shape[0]
    assert tuple(qkv.shape[1:]) == (NTOK, 3 * EMBED)
    out = torch.empty((n, NTOK, EMBED), dtype=torch.float32, device=qkv.device)
    dev, st = _ctx(qkv)
    _tbegin("self_attention", 4.0 * n * HEADS * NTOK * NTOK * HDIM, 4.0 * n * NTOK * 4 * EMBED)
    fn = _lib.lib().rp_cross_attention_f32 if cross else _lib.lib().rp_self_attention_f32
    _lib.check(fn(_p(qkv), _p(out), n, dev, st), "rp_cross_attention" if cross else "rp_self_attention")
    _count()
    return out


def conv_dw_tc_supported(C, O, KH, KW, stride):
    return bool(_lib.lib().rp_conv_dw_tc_supported(int(C), int(O), int(KH), int(KW), int(stride)))


def conv_dw_tc(x_planes, dy_planes, KH, KW, pad, stride=1):
    """Weight gradient of a convolution (stride 1 or 2), implicit GEMM on tcgen05: x_planes bf16 [2,n,H,W,C], dy_planes bf16
    [2,n,OH,OW,O] -> float32 [O,KH,KW,C]."""
    _req(x_planes, "x_planes", torch.bfloat16); _req(dy_planes, "dy_planes", torch.bfloat16)
    P, n, H, W, C = x_planes.shape
    O = dy_planes.shape[-1]
    assert P == 2 and tuple(dy_planes.shape) == (2, n, (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1, O)
    dev, st = _ctx(x_planes)
    L = _lib.lib()
    nb = L.rp_conv_dw_tc_workspace_bytes(n, H, W, C, O, KH, KW, pad, stride, dev)
    assert nb > 0
    ws = torch.empty((nb // 4,), dtype=torch.float32, device=x_planes.device)
    dw = torch.empty((O, KH, KW, C), dtype=torch.float32, device=x_planes.device)
    M = dy_planes[0].numel() // O
    _tbegin(f"conv_dw_tcx3[{O}x{KH}x{KW}x{C}/s{stride}]", 2.0 * M * O * KH * KW * C, 2.0 * 2 * (x_planes[0].numel() + dy_planes[0].numel()) + 2.0 * nb)
    _lib.check(L.rp_conv_dw_tc(_p(x_planes), _p(dy_planes), _p(dw), n, H, W, C, O, KH, KW, pad, stride, _p(ws), nb, dev, st), "rp_conv_dw_tc")
    _count(2)
    return dw


def linear_dw_tc_supported(M, N, K):
    return bool(_lib.lib().rp_linear_dw_tc_supported(int(M), int(N), int(K)))


def linear_dw_tc(x_planes, dy_planes):
    """Weight gradient of nn.Linear on the implicit-GEMM kernel: x_planes bf16 [2,M,K], dy_planes bf16 [2,M,N] -> float32
    [N,K] = dY^T X, both operands read in place (no transposed copies)."""
    _req(x_planes, "x_planes", torch.bfloat16); _req(dy_planes, "dy_planes", torch.bfloat16)
    P, M, K = x_planes.shape
    N = dy_planes.shape[-1]
    assert P == 2 and tuple(dy_planes.shape) == (2, M, N)
    dev, st = _ctx(x_planes)
    L = _lib.lib()
    nb = L.rp_linear_dw_tc_workspace_bytes(M, N, K, dev)
    assert nb > 0
    ws = torch.empty((nb // 4,), dtype=torch.float32, device=x_planes.device)
    dw = torch.empty((N, K), dtype=torch.float32, device=x_planes.device)
    _tbegin(f"linear_dw_tcx3[{N}x{K}]", 2.0 * M * N * K, 2.0 * 2 * (x_planes[0].numel() + dy_planes[0].numel()) + 2.0 * nb)
    _lib.check(L.rp_linear_dw_tc(_p(x_planes), _p(dy_planes), _p(dw), M, N, K, _p(ws), nb, dev, st), "rp_linear_dw_tc")
    _count(2)
    return dw


def self_attention_tc_lse(qkv_planes):
    """Training forward: qkv_planes bf16 [P,n,576,576] -> (float32 out [n,576,192], float32 lse [n,3,576])."""
    _req(qkv_planes, "qkv_planes", torch.bfloat16)
    P, n = qkv_planes.shape[0], qkv_planes.shape[1]
    assert tuple(qkv_planes.shape[2:]) == (NTOK, 3 * EMBED)
    out = torch.empty((n, NTOK, EMBED), dtype=torch.float32, device=qkv_planes.device)
    lse = torch.empty((n, HEADS, NTOK), dtype=torch.float32, device=qkv_planes.device)
    dev, st = _ctx(qkv_planes)
    _tbegin(f"self_attention_tc{'x3' if P == 2 else ''}", 4.0 * n * HEADS * NTOK * NTOK * HDIM,
            2.0 * P * n * NTOK * 3 * EMBED + 4.0 * n * NTOK * EMBED)
    _lib.check(_lib.lib().rp_self_attention_tc_lse(_p(qkv_planes), _p(out), _p(lse), n, P, dev, st), "rp_self_attention_tc_lse")
    _count()
    return out, lse


def attention_bwd_tc(qkv_planes, d_out, out, lse):
    """Flash-style backward of the attention core: -> d_qkv float32 [n,576,576] (two launches, no 576x576 tensor)."""
    _req(qkv_planes, "qkv_planes", torch.bfloat16); _req(d_out, "d_out"); _req(out, "out"); _req(lse, "lse")
    P, n = qkv_planes.shape[0], qkv_planes.shape[1]
    assert P == 2 and tuple(d_out.shape) == (n, NTOK, EMBED) == tuple(out.shape) and tuple(lse.shape) == (n, HEADS, NTOK)
    dop = torch.empty((2, n, NTOK, EMBED), dtype=torch.bfloat16, device=d_out.device)
    delta = torch.empty((n, HEADS, NTOK), dtype=torch.float32, device=d_out.device)
    dqkv = torch.empty((n, NTOK, 3 * EMBED), dtype=torch.float32, device=d_out.device)
    dev, st = _ctx(d_out)
    L = _lib.lib()
    _lib.check(L.rp_attention_bwd_prep(_p(d_out), _p(out), _p(dop), _p(delta), n, dev, st), "rp_attention_bwd_prep")
    _tbegin("attention_bwd_tcx3", 14.0 * n * HEADS * NTOK * NTOK * HDIM, 2.0 * 2 * n * NTOK * 4 * EMBED + 4.0 * n * NTOK * 3 * EMBED)
    _lib.check(L.rp_attention_bwd_tc(_p(qkv_planes), _p(dop), _p(lse), _p(delta), _p(dqkv), n, dev, st), "rp_attention_bwd_tc")
    _count(2)
    return dqkv


def self_attention_tc(qkv_planes, want_f32=False, planes_out=0, cross=False):
    """qkv_planes bf16 [P,n,576,576] -> (float32 [n,576,192] | None, bf16 planes [planes_out,n,576,192] | None).
    cross=True: image i attends to the keys/values of image i^1 (--noess)."""
    _req(qkv_planes, "qkv_planes", torch.bfloat16)
    P, n = qkv_planes.shape[0], qkv_planes.shape[1]
    assert tuple(qkv_planes.shape[2:]) == (NTOK, 3 * EMBED)
    assert want_f32 or planes_out
    out = torch.empty((n, NTOK, EMBED), dtype=torch.float32, device=qkv_planes.device) if want_f32 else None
    outp = torch.empty((planes_out, n, NTOK, EMBED), dtype=torch.bfloat16, device=qkv_planes.device) if planes_out else None
    dev, st = _ctx(qkv_planes)
    _tbegin(f"self_attention_tc{'x3' if P == 2 else ''}", 4.0 * n * HEADS * NTOK * NTOK * HDIM,
            2.0 * P * n * NTOK * 3 * EMBED + (4.0 if want_f32 else 0.0) * n * NTOK * EMBED + 2.0 * planes_out * n * NTOK * EMBED)
    fn = _lib.lib().rp_cross_attention_tc if cross else _lib.lib().rp_self_attention_tc
    _lib.check(fn(_p(qkv_planes), _p(out), _p(outp), n, P, int(planes_out), dev, st),
               "rp_cross_attention_tc" if cross else "rp_self_attention_tc")
    _count()
    return out, outp


_LIN24 = None


def lin24():
    """torch.linspace(-1,1,24) evaluated on the host exactly like the reference does
    (vision_transformer.py:108-109) -- 24 floats handed to the kernel by value."""
    global _LIN24
    if _LIN24 is None:
        _LIN24 = torch.linspace(-1, 1, steps=24, dtype=torch.float32).contiguous()
    return _LIN24


EM_SINGLE_SOFTMAX, EM_CROSS_FEATURES = 1, 2      # RP_EM_* flags of rp_essential_ex_f32


def posenc(B, kxy, device, l1=False):
    """[B,576,6] positional monomials; kxy [B,2] or None (intrinsics=None).  l1: --l1_pos_encoding ([1,1,1,p3,p4,1])."""
    pos = torch.empty((B, NTOK, NPOS), dtype=torch.float32, device=device)
    if kxy is not None:
        _req(kxy, "kxy")
    t = lin24()
    dev, st = _ctx(pos)
    _lib.check(_lib.lib().rp_posenc_ex_f32(_p(kxy), ctypes.c_void_p(t.data_ptr()), _p(pos), B, int(bool(l1)), dev, st), "rp_posenc")
    _count()
    return pos


def essential(qkv, pos, flags=0):
    """qkv [2B,576,576], pos [B,576,6] or None -> bilinear forms [B,2,3,W,W].  flags: EM_SINGLE_SOFTMAX | EM_CROSS_FEATURES."""
    _req(qkv, "qkv")
    B = qkv.shape[0] // 2
    assert qkv.shape[0] == 2 * B and tuple(qkv.shape[1:]) == (NTOK, 3 * EMBED)
    width = EMW if pos is not None else HDIM
    if pos is not None:
        _req(pos, "pos")
        assert tuple(pos.shape) == (B, NTOK, NPOS)
    bil = torch.empty((B, 2, HEADS, width, width), dtype=torch.float32, device=qkv.device)
    L = _lib.lib()
    ws_bytes = L.rp_essential_workspace_bytes(B)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=qkv.device)
    dev, st = _ctx(qkv)
    # per (pair, head, direction): scores 2*N*N*64 (x3: two stat passes + recompute), A*V 2*N*N*W, V^T*T 2*N*W*W
    _tbegin("essential", B * 2.0 * HEADS * (2.0 * NTOK * NTOK * HDIM + 2.0 * NTOK * NTOK * width + 2.0 * NTOK * width * width),
            4.0 * (2 * B * NTOK * 3 * EMBED + B * 2 * HEADS * width * width))
    _lib.check(L.rp_essential_ex_f32(_p(qkv), _p(pos), _p(bil), B, int(flags), _p(ws), ws_bytes, dev, st), "rp_essential")
    _count(3)
    return bil


def essential_tc(qkv_planes, pos, flags=0):
    """qkv_planes bf16 [P,2B,576,576], pos [B,576,6] or None -> bilinear forms [B,2,3,W,W] (tensor cores).
    flags: EM_SINGLE_SOFTMAX | EM_CROSS_FEATURES (ablation branches)."""
    _req(qkv_planes, "qkv_planes", torch.bfloat16)
    P, n = qkv_planes.shape[0], qkv_planes.shape[1]
    B = n // 2
    assert n == 2 * B and tuple(qkv_planes.shape[2:]) == (NTOK, 3 * EMBED)
    width = EMW if pos is not None else HDIM
    if pos is not None:
        _req(pos, "pos")
        assert tuple(pos.shape) == (B, NTOK, NPOS)
    bil = torch.empty((B, 2, HEADS, width, width), dtype=torch.float32, device=qkv_planes.device)
    L = _lib.lib()
    ws_bytes = L.rp_essential_tc_workspace_bytes(B, P)
    ws = torch.empty((ws_bytes // 4 + 4,), dtype=torch.float32, device=qkv_planes.device)
    dev, st = _ctx(qkv_planes)
    _tbegin(f"essential_tc{'x3' if P == 2 else ''}",
            B * 2.0 * HEADS * (3 * 2.0 * NTOK * NTOK * HDIM + 2.0 * NTOK * NTOK * width + 2.0 * NTOK * width * width),
            2.0 * P * n * NTOK * 3 * EMBED + 4.0 * B * 2 * HEADS * width * width)
    _lib.check(L.rp_essential_ex_tc(_p(qkv_planes), _p(pos), _p(bil), B, P, int(flags), _p(ws), ws_bytes, dev, st), "rp_essential_tc")
    _count(3 if pos is not None else 2)
    return bil


def essential_tc_train(qkv_planes, pos):
    """Training forward of the module core: (bil [B,2,3,70,70], lse2 [B,2,2,3,576]) -- the row / column log2-sum-exp
    vectors are the head of rp_essential_tc's workspace and all the flash-style backward needs besides q, k, v."""
    _req(qkv_planes, "qkv_planes", torch.bfloat16); _req(pos, "pos")
    P, n = qkv_planes.shape[0], qkv_planes.shape[1]
    B = n // 2
    assert n == 2 * B and tuple(qkv_planes.shape[2:]) == (NTOK, 3 * EMBED) and tuple(pos.shape) == (B, NTOK, NPOS)
    bil = torch.empty((B, 2, HEADS, EMW, EMW), dtype=torch.float32, device=qkv_planes.device)
    L = _lib.lib()
    ws_bytes = L.rp_essential_tc_workspace_bytes(B, P)
    ws = torch.empty((ws_bytes // 4 + 4,), dtype=torch.float32, device=qkv_planes.device)
    dev, st = _ctx(qkv_planes)
    _tbegin(f"essential_tc{'x3' if P == 2 else ''}",
            B * 2.0 * HEADS * (3 * 2.0 * NTOK * NTOK * HDIM + 2.0 * NTOK * NTOK * EMW + 2.0 * NTOK * EMW * EMW),
            2.0 * P * n * NTOK * 3 * EMBED + 4.0 * B * 2 * HEADS * EMW * EMW)
    _lib.check(L.rp_essential_ex_tc(_p(qkv_planes), _p(pos), _p(bil), B, P, 0, _p(ws), ws_bytes, dev, st), "rp_essential_tc")
    _count(3)
    return bil, ws[:B * 2 * 2 * HEADS * NTOK].clone().reshape(B, 2, 2, HEADS, NTOK)


def em_bwd_tc(qkv_planes, pos, lse2, d_bil):
    """Flash-style backward of the module core -> d_qkv float32 [2B,576,576]."""
    _req(qkv_planes, "qkv_planes", torch.bfloat16); _req(pos, "pos"); _req(lse2, "lse2"); _req(d_bil, "d_bil")
    P, n = qkv_planes.shape[0], qkv_planes.shape[1]
    B = n // 2
    assert P == 2 and tuple(d_bil.shape) == (B, 2, HEADS, EMW, EMW) and tuple(lse2.shape) == (B, 2, 2, HEADS, NTOK)
    L = _lib.lib()
    nb = L.rp_em_bwd_tc_workspace_bytes(B)
    ws = torch.empty((nb // 4 + 4,), dtype=torch.float32, device=d_bil.device)
    dqkv = torch.empty((n, NTOK, 3 * EMBED), dtype=torch.float32, device=d_bil.device)
    dev, st = _ctx(d_bil)
    _tbegin("em_bwd_tcx3", B * 2.0 * HEADS * (7 * 2.0 * NTOK * NTOK * HDIM), 2.0 * P * n * NTOK * 3 * EMBED + 4.0 * n * NTOK * 3 * EMBED)
    _lib.check(L.rp_em_bwd_tc(_p(qkv_planes), _p(pos), _p(lse2), _p(d_bil), _p(dqkv), B, _p(ws), nb, dev, st), "rp_em_bwd_tc")
    _count(4)
    return dqkv


def em_project(bil, weight, bias):
    """bil [B,2,3,70,70] -> [2B,70,192] (proj_fundamental + the reference's output flip)."""
    _req(bil, "bil"); _req(weight, "weight"); _req(bias, "bias")
    B = bil.shape[0]
    assert tuple(bil.shape[1:]) == (2, HEADS, EMW, EMW) and tuple(weight.shape) == (EMBED, HEADS * EMW)
    out = torch.empty((2 * B, EMW, EMBED), dtype=torch.float32, device=bil.device)
    dev, st = _ctx(bil)
    _tbegin("em_project", 2.0 * B * 2 * EMW * HEADS * EMW * EMBED, 4.0 * (bil.numel() + 2 * B * EMW * EMBED))
    _lib.check(_lib.lib().rp_em_project_f32(_p(bil), _p(weight), _p(bias), _p(out), B, dev, st), "rp_em_project")
    _count()
    return out


def normalize_pose(raw, Gs):
    _req(raw, "raw"); _req(Gs, "Gs")
    B = raw.shape[0]
    assert tuple(raw.shape) == (B, 2, 7) and tuple(Gs.shape) == (B, 2, 7)
    out = torch.empty_like(raw)
    dev, st = _ctx(raw)
    _lib.check(_lib.lib().rp_normalize_pose_f32(_p(raw), _p(Gs), _p(out), B, dev, st), "rp_normalize_pose")
    _count()
    return out


# ------------------------------------------------------------------------------------------ A12
def _se3_unary(fn_name, x, in_w, out_w):
    _req(x, fn_name)
    assert x.shape[-1] == in_w
    n = x.numel() // in_w
    out = torch.empty(x.shape[:-1] + (out_w,), dtype=torch.float32, device=x.device)
    if n:
        dev, st = _ctx(x)
        _lib.check(getattr(_lib.lib(), fn_name)(_p(x), _p(out), n, dev, st), fn_name)
        _count()
    return out


def se3_inv_fwd(X):
    return _se3_unary("rp_se3_inv_fwd_f32", X, 7, 7)


def se3_log_fwd(X):
    return _se3_unary("rp_se3_log_fwd_f32", X, 7, 6)


def se3_exp_fwd(a):
    return _se3_unary("rp_se3_exp_fwd_f32", a, 6, 7)


def se3_mul_fwd(X, Y):
    _req(X, "X"); _req(Y, "Y")
    assert X.shape == Y.shape and X.shape[-1] == 7
    Z = torch.empty_like(X)
    n = X.numel() // 7
    if n:
        dev, st = _ctx(X)
        _lib.check(_lib.lib().rp_se3_mul_fwd_f32(_p(X), _p(Y), _p(Z), n, dev, st), "rp_se3_mul_fwd")
        _count()
    return Z


def se3_mul_bwd(dZ, X, Y):
    _req(dZ, "dZ"); _req(X, "X"); _req(Y, "Y")
    dX, dY = torch.empty_like(X), torch.empty_like(Y)
    n = X.numel() // 7
    if n:
        dev, st = _ctx(X)
        _lib.check(_lib.lib().rp_se3_mul_bwd_f32(_p(dZ), _p(X), _p(Y), _p(dX), _p(dY), n, dev, st), "rp_se3_mul_bwd")
        _count()
    return dX, dY


def _se3_bwd(fn_name, g, x, out_w):
    _req(g, "grad"); _req(x, "x")
    n = x.numel() // x.shape[-1]
    out = torch.empty(x.shape[:-1] + (out_w,), dtype=torch.float32, device=x.device)
    if n:
        dev, st = _ctx(x)
        _lib.check(getattr(_lib.lib(), fn_name)(_p(g), _p(x), _p(out), n, dev, st), fn_name)
        _count()
    return out


def se3_inv_bwd(dY, X):
    return _se3_bwd("rp_se3_inv_bwd_f32", dY, X, 7)


def se3_log_bwd(da, X):
    return _se3_bwd("rp_se3_log_bwd_f32", da, X, 7)


def se3_exp_bwd(dX, a):
    return _se3_bwd("rp_se3_exp_bwd_f32", dX, a, 6)


# ------------------------------------------------------------------------------------------ config 3
def svd3(E):
    _req(E, "E")
    assert tuple(E.shape[-2:]) == (3, 3)
    n = E.numel() // 9
    U = torch.empty_like(E)
    V = torch.empty_like(E)
    S = torch.empty(E.shape[:-2] + (3,), dtype=torch.float32, device=E.device)
    dev, st = _ctx(E)
    _lib.check(_lib.lib().rp_svd3_f32(_p(E), _p(U), _p(S), _p(V), n, dev, st), "rp_svd3")
    _count()
    return U, S, V


def essential_to_rt(E):
    _req(E, "E")
    n = E.numel() // 9
    R1 = torch.empty_like(E)
    R2 = torch.empty_like(E)
    t = torch.empty(E.shape[:-2] + (3,), dtype=torch.float32, device=E.device)
    dev, st = _ctx(E)
    _lib.check(_lib.lib().rp_essential_to_rt_f32(_p(E), _p(R1), _p(R2), _p(t), n, dev, st), "rp_essential_to_rt")
    _count()
    return R1, R2, t
