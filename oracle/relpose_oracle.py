"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU (numpy) restatement of the reference hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this module.  The product path (rel_pose_b200/) never does: it calls the CUDA library
through the C-ABI and fails loudly if that library is missing.

What is restated: crockwell/rel_pose @35d1352, `ViTEss.forward` with the default flags
(`--fusion_transformer --transformer_depth 6`), i.e. SURVEY.md section 8(a) rows A1-A11, plus
the lietorch SE3 group maths (A12; lietorch itself is NOT in the reference tree -- pinned
`lietorch==0.2` in /root/reference/environment.yml:20 -- so its published group formulas are
restated and *its backward convention is parity-unpinned*).

Pinning: this restatement is checked against golden vectors produced by the *real* reference
imported from /root/reference (oracle/make_golden.py -> tests/golden/*.npz;
tests/test_oracle_golden.py).  The SE3 part and the SVD / E->(R,t) part (oracle/geom_oracle.py)
have no reference counterpart to pin against: "parity unpinned" for those.

Every function cites the reference lines it follows.  `dtype` selects float32 (the
reference's arithmetic) or float64 (a tighter "truth" used to measure noise floors).
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

try:  # exact erf for GELU
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    import math
    _erf = np.vectorize(math.erf)

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
GRID = 24           # feature_resolution, src/model.py:20
NTOK = GRID * GRID  # 576
EMBED = 192
HEADS = 3
HDIM = 64
NPOS = 6
EMW = HDIM + NPOS   # 70


# ------------------------------------------------------------------ A1: preprocessing
def nearest_src_index(out_size, in_size):
    """Legacy `nearest` of F.interpolate (src/model.py:125): src = floor(dst * in/out), the
    scale being computed in float32 exactly as ATen does (scale = (float)in / out)."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def preprocess(images, dtype=np.float32, out_size=224):
    """src/model.py:114-125.  images [B,2,3,H,W] BGR 0..255 -> [2B,3,224,224] normalised RGB.
    Operation order kept: /255, -mean, /std (true divisions)."""
    x = np.asarray(images, dtype=dtype)[:, :, ::-1]            # BGR -> RGB (index [2,1,0])
    x = x / dtype(255.0)
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32).astype(dtype)[:, None, None]
    std = np.asarray(IMAGENET_STD, dtype=np.float32).astype(dtype)[:, None, None]
    x = (x - mean) / std
    B, V, C, H, W = x.shape
    x = x.reshape(B * V, C, H, W)                              # nn.Flatten(0,1)
    iy = nearest_src_index(out_size, H)
    ix = nearest_src_index(out_size, W)
    return np.ascontiguousarray(x[:, :, iy][:, :, :, ix])


def update_intrinsics(intrinsics, H, W):
    """src/model.py:100-109.  Returns the rescaled copy (the reference mutates in place;
    the in-place side effect is reproduced by the host wrapper, not here).
    scalex/scaley are Python doubles multiplied into a float32 tensor."""
    k = np.array(intrinsics, dtype=np.float32, copy=True)
    sx = GRID / W
    sy = GRID / H
    k[:, :, [0, 2]] = (np.float32(sx) * k[:, :, [0, 2]]).astype(np.float32)
    k[:, :, [1, 3]] = (np.float32(sy) * k[:, :, [1, 3]]).astype(np.float32)
    return k


# ------------------------------------------------------------------ A2/A3: CNN front end
def conv2d(x, w, b=None, stride=1, pad=0):
    """NCHW cross-correlation, as nn.Conv2d."""
    N, C, H, W = x.shape
    O, _, kh, kw = w.shape
    if pad:
        x = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    win = sliding_window_view(x, (kh, kw), axis=(2, 3))[:, :, ::stride, ::stride]  # N,C,Ho,Wo,kh,kw
    Ho, Wo = win.shape[2], win.shape[3]
    cols = win.transpose(0, 2, 3, 1, 4, 5).reshape(N * Ho * Wo, C * kh * kw)
    y = cols @ w.reshape(O, -1).T
    if b is not None:
        y = y + b
    return np.ascontiguousarray(y.reshape(N, Ho, Wo, O).transpose(0, 3, 1, 2))


def batchnorm_eval(x, p, prefix, eps=1e-5):
    """nn.BatchNorm2d in eval mode: (x-mean)/sqrt(var+eps)*w+b."""
    dt = x.dtype.type
    m = p[prefix + ".running_mean"].astype(dt)[None, :, None, None]
    v = p[prefix + ".running_var"].astype(dt)[None, :, None, None]
    g = p[prefix + ".weight"].astype(dt)[None, :, None, None]
    b = p[prefix + ".bias"].astype(dt)[None, :, None, None]
    return (x - m) / np.sqrt(v + dt(eps)) * g + b


def relu(x):
    return np.maximum(x, 0)


def maxpool3x3s2(x):
    """nn.MaxPool2d(3, stride=2, padding=1) (torchvision resnet stem)."""
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), constant_values=-np.inf)
    win = sliding_window_view(xp, (3, 3), axis=(2, 3))[:, :, ::2, ::2]
    return win.max(axis=(-1, -2))


def _w(p, k, dt):
    return p[k].astype(dt)


def basic_block(x, p, prefix, stride, dt):
    """torchvision BasicBlock: conv3x3-bn-relu-conv3x3-bn (+1x1/stride shortcut) add relu."""
    y = relu(batchnorm_eval(conv2d(x, _w(p, prefix + ".conv1.weight", dt), None, stride, 1), p, prefix + ".bn1"))
    y = batchnorm_eval(conv2d(y, _w(p, prefix + ".conv2.weight", dt), None, 1, 1), p, prefix + ".bn2")
    if (prefix + ".downsample.0.weight") in p:
        x = batchnorm_eval(conv2d(x, _w(p, prefix + ".downsample.0.weight", dt), None, stride, 0), p,
                           prefix + ".downsample.1")
    return relu(x + y)


def cnn_front_end(x, p):
    """src/model.py:127-134 + src/modules/extractor.py:51-65.  x [2B,3,224,224] -> [2B,192,24,24]."""
    dt = x.dtype.type
    x = conv2d(x, _w(p, "resnet.conv1.weight", dt), None, 2, 3)
    x = relu(batchnorm_eval(x, p, "resnet.bn1"))
    x = maxpool3x3s2(x)
    x = basic_block(x, p, "resnet.layer1.0", 1, dt)
    x = basic_block(x, p, "resnet.layer1.1", 1, dt)
    x = basic_block(x, p, "resnet.layer2.0", 2, dt)
    x = basic_block(x, p, "resnet.layer2.1", 1, dt)
    e = "extractor_final_conv"
    y = relu(batchnorm_eval(conv2d(x, _w(p, e + ".conv1.weight", dt), _w(p, e + ".conv1.bias", dt), 1, 1), p, e + ".norm1"))
    y = relu(batchnorm_eval(conv2d(y, _w(p, e + ".conv2.weight", dt), _w(p, e + ".conv2.bias", dt), 1, 0), p, e + ".norm2"))
    s = batchnorm_eval(conv2d(x, _w(p, e + ".downsample.0.weight", dt), _w(p, e + ".downsample.0.bias", dt), 1, 0), p, e + ".norm3")
    return relu(s + y)


def tokens_from_feature_map(fm):
    """src/model.py:136-141: [2B,192,24,24] -> [2B,576,192], token i = row*24+col."""
    n = fm.shape[0]
    return np.ascontiguousarray(fm.reshape(n, EMBED, NTOK).transpose(0, 2, 1))


# ------------------------------------------------------------------ A5: transformer pieces
def layernorm(x, g, b, eps=1e-6):
    """nn.LayerNorm(eps=1e-6) -- vision_transformer.py:396 (biased variance)."""
    dt = x.dtype.type
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + dt(eps)) * g.astype(dt) + b.astype(dt)


def linear(x, w, b):
    dt = x.dtype.type
    return x @ w.astype(dt).T + b.astype(dt)


def gelu(x):
    """nn.GELU() exact (erf) form."""
    dt = x.dtype.type
    return (dt(0.5) * x * (dt(1.0) + _erf(x * dt(0.7071067811865476)))).astype(x.dtype)


def softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def split_qkv(qkv):
    """vision_transformer.py:323-324 / :191-195: column = s*192 + h*64 + d."""
    Bn, N, _ = qkv.shape
    t = qkv.reshape(Bn, N, 3, HEADS, HDIM).transpose(2, 0, 3, 1, 4)
    return t[0], t[1], t[2]          # each [Bn, H, N, 64]


def mlp(x, p, prefix):
    """vit_layers/mlp.py:20-26 (dropout p=0)."""
    h = gelu(linear(x, p[prefix + ".fc1.weight"], p[prefix + ".fc1.bias"]))
    return linear(h, p[prefix + ".fc2.weight"], p[prefix + ".fc2.bias"])


def self_attention(x, p, prefix):
    """Attention.forward, vision_transformer.py:321-333."""
    dt = x.dtype.type
    q, k, v = split_qkv(linear(x, p[prefix + ".qkv.weight"], p[prefix + ".qkv.bias"]))
    attn = softmax((q @ k.transpose(0, 1, 3, 2)) * dt(HDIM ** -0.5), -1)
    o = (attn @ v).transpose(0, 2, 1, 3).reshape(x.shape)
    return linear(o, p[prefix + ".proj.weight"], p[prefix + ".proj.bias"])


def block(x, p, prefix):
    """Block.forward, vision_transformer.py:349-354 (pre-LN, drop_path = identity)."""
    x = x + self_attention(layernorm(x, p[prefix + ".norm1.weight"], p[prefix + ".norm1.bias"]), p, prefix + ".attn")
    return x + mlp(layernorm(x, p[prefix + ".norm2.weight"], p[prefix + ".norm2.bias"]), p, prefix + ".mlp")


# ------------------------------------------------------------------ A6: positional monomials
def linspace_pm1(steps=GRID):
    """torch.linspace(-1,1,steps) in float32: ATen fills the lower half as start+step*i and the
    upper half as end-step*(steps-1-i) (symmetric), step=(end-start)/(steps-1) in float32, each
    element with ONE rounding (the vectorised CPU kernel contracts to an FMA).  Bit-identical to
    torch on the build container; ATen's result can differ by 1 ulp on hosts with another
    vector width, which is why tests allow 1 ulp here."""
    step = np.float64((np.float32(1.0) - np.float32(-1.0)) / np.float32(steps - 1))
    out = np.empty(steps, np.float32)
    half = steps // 2
    for i in range(steps):
        out[i] = np.float32(-1.0 + step * i) if i < half else np.float32(1.0 - step * (steps - 1 - i))
    return out


def positional_encodings(B, intrinsics=None, l1=False):
    """get_positional_encodings, vision_transformer.py:90-158, N=576, float32 like the reference
    (it builds the table on the CPU in float32 irrespective of the model dtype).

    The reference builds K=diag(fx/cx, fy/cy, 1) (the principal point normalises to exactly 0:
    (cx/(2cx))*2-1), inverts it and applies it to [xs[k], ys[j], 1]; token k*24+j receives
    p3 = ys[j]*(1/(fy/cy)), p4 = xs[k]*(1/(fx/cx))  -- i.e. the grid is TRANSPOSED w.r.t. the
    row-major token order.  Output [B,576,6] = [p3^2, p4^2, p3*p4, p3, p4, 1]."""
    xs = linspace_pm1()
    ys = linspace_pm1()
    i = np.arange(NTOK)
    p3 = np.tile(ys[i % GRID][None], (B, 1)).astype(np.float32)
    p4 = np.tile(xs[i // GRID][None], (B, 1)).astype(np.float32)
    if intrinsics is not None:
        k = np.asarray(intrinsics, np.float32)
        assert np.all(k[:, 0] == k[:, 1]), "intrinsics must not change across the pair (vision_transformer.py:117)"
        fx, fy, cx, cy = (k[:, 0, j] for j in range(4))
        if cx[0] * cy[0] == 0:
            raise ValueError("principal point is in upper left (vision_transformer.py:124-126)")
        fxn = (fx / (cx * np.float32(2))) * np.float32(2)
        fyn = (fy / (cy * np.float32(2))) * np.float32(2)
        kx = (np.float32(1) / fxn).astype(np.float32)   # torch.inverse of a diagonal matrix
        ky = (np.float32(1) / fyn).astype(np.float32)
        p3 = (ky[:, None] * ys[i % GRID][None]).astype(np.float32)
        p4 = (kx[:, None] * xs[i // GRID][None]).astype(np.float32)
    pos = np.ones((B, NTOK, NPOS), np.float32)
    if not l1:          # get_l1_positional_encodings (vision_transformer.py:36-87) leaves the quadratic channels at 1
        pos[:, :, 0] = p3 * p3
        pos[:, :, 1] = p4 * p4
        pos[:, :, 2] = p3 * p4
    pos[:, :, 3] = p3
    pos[:, :, 4] = p4
    return pos


# ------------------------------------------------------------------ A7: Essential Matrix Module
def essential_matrix_module(x1, x2, p, prefix, intrinsics=None, return_bilinear=False, flags=()):
    """CrossAttention.forward, non-noess branch, vision_transformer.py:188-238.
    x1,x2 [B,576,192] (already norm1'ed).  Returns (Y2, Y1) -- note the flip at :238."""
    dt = x1.dtype.type
    B = x1.shape[0]
    q1, k1, v1 = split_qkv(linear(x1, p[prefix + ".qkv.weight"], p[prefix + ".qkv.bias"]))
    q2, k2, v2 = split_qkv(linear(x2, p[prefix + ".qkv.weight"], p[prefix + ".qkv.bias"]))
    scale = dt(HDIM ** -0.5)
    s1 = (q2 @ k1.transpose(0, 1, 3, 2)) * scale      # :198
    s2 = (q1 @ k2.transpose(0, 1, 3, 2)) * scale      # :199
    if "use_single_softmax" in flags:                 # :201-203
        a1, a2 = softmax(s1, -1), softmax(s2, -1)
    else:
        a1 = softmax(s1, -1) * softmax(s1, -2)        # :205
        a2 = softmax(s2, -1) * softmax(s2, -2)        # :206
    pos = positional_encodings(B, intrinsics, "l1_pos_encoding" in flags).astype(dt)      # :208-211
    posh = np.broadcast_to(pos[:, None], (B, HEADS, NTOK, NPOS))
    V1 = np.concatenate([v1, posh], 3)                # :215
    V2 = np.concatenate([v2, posh], 3)                # :216
    if "cross_features" in flags:                     # :219-220
        f1 = (V2.transpose(0, 1, 3, 2) @ a1) @ V1
        f2 = (V1.transpose(0, 1, 3, 2) @ a2) @ V2
    else:
        f1 = (V1.transpose(0, 1, 3, 2) @ a1) @ V1     # :222  [B,3,70,70]
        f2 = (V2.transpose(0, 1, 3, 2) @ a2) @ V2     # :223
    z1 = f1.reshape(B, HEADS * EMW, EMW).transpose(0, 2, 1)   # :229  Z[b,c,h*70+a] = F[b,h,a,c]
    z2 = f2.reshape(B, HEADS * EMW, EMW).transpose(0, 2, 1)   # :230
    y2 = linear(z2, p[prefix + ".proj_fundamental.weight"], p[prefix + ".proj_fundamental.bias"])
    y1 = linear(z1, p[prefix + ".proj_fundamental.weight"], p[prefix + ".proj_fundamental.bias"])
    if return_bilinear:
        return (y2, y1), (f1, f2)
    return y2, y1


def cross_attention_noess(x1, x2, p, prefix):
    """CrossAttention.forward, --noess branch, vision_transformer.py:239-262: plain cross attention
    (queries of one view, keys/values of the other), shared output projection.  Returns (X2, X1) -- the flip at :262."""
    dt = x1.dtype.type
    B, N, C = x1.shape
    q1, k1, v1 = split_qkv(linear(x1, p[prefix + ".qkv.weight"], p[prefix + ".qkv.bias"]))
    q2, k2, v2 = split_qkv(linear(x2, p[prefix + ".qkv.weight"], p[prefix + ".qkv.bias"]))
    scale = dt(HDIM ** -0.5)
    a1 = softmax((q2 @ k1.transpose(0, 1, 3, 2)) * scale, -1)         # :241-242
    o1 = (a1 @ v1).transpose(0, 2, 1, 3).reshape(B, N, C)             # :245
    a2 = softmax((q1 @ k2.transpose(0, 1, 3, 2)) * scale, -1)         # :248-249
    o2 = (a2 @ v2).transpose(0, 2, 1, 3).reshape(B, N, C)             # :252
    o1 = linear(o1, p[prefix + ".proj.weight"], p[prefix + ".proj.bias"])
    o2 = linear(o2, p[prefix + ".proj.weight"], p[prefix + ".proj.bias"])
    return o2, o1


def cross_block_noess(x, p, prefix):
    """CrossBlock.forward, --noess branch, vision_transformer.py:297-303.  x [2B,576,192] -> [2B,576,192]."""
    n2, N, C = x.shape
    xp = x.reshape(n2 // 2, 2, N, C)
    g, b = p[prefix + ".norm1.weight"], p[prefix + ".norm1.bias"]
    ya, yb = cross_attention_noess(layernorm(xp[:, 0], g, b), layernorm(xp[:, 1], g, b), p, prefix + ".cross_attn")
    x = x + np.stack([ya, yb], 1).reshape(n2, N, C)
    return x + mlp(layernorm(x, p[prefix + ".norm2.weight"], p[prefix + ".norm2.bias"]), p, prefix + ".mlp")


def pool_attn_noess(x, p, B):
    """src/model.py:183-187 with the pool_attn head of :71-80.  x [2B,576,192] is re-read as [B,24,24,384]
    (the reference's reshape: "pixel" j of a pair holds tokens 2j and 2j+1 of the pair's flat token list),
    1x1 conv 384->96 + BN + ReLU + 1x1 conv 96->43 + BN, flattened channel-major to [B, 43*576]."""
    f = x.reshape(B, GRID, GRID, -1).transpose(0, 3, 1, 2)
    h = relu(batchnorm_eval(conv2d(f, p["pool_attn.0.weight"], p["pool_attn.0.bias"]), p, "pool_attn.1"))
    h = batchnorm_eval(conv2d(h, p["pool_attn.3.weight"], p["pool_attn.3.bias"]), p, "pool_attn.4")
    return h.reshape(B, -1)


def pool_transformer_output_cnn_only(tok, p, B):
    """The model without --fusion_transformer (src/model.py:62-69,138-139,179-181,189): the first 96 channels of every
    token, re-read as [B,24,24,192] (two consecutive tokens of the pair's flat list per "pixel"), 1x1 conv 192->96 + BN
    + ReLU + 1x1 conv 96->pool_size + BN, flattened channel-major."""
    f = np.ascontiguousarray(tok[:, :, :EMBED // 2]).reshape(B, GRID, GRID, EMBED).transpose(0, 3, 1, 2)
    pre = "pool_transformer_output"
    h = relu(batchnorm_eval(conv2d(f, p[pre + ".0.weight"], p[pre + ".0.bias"]), p, pre + ".1"))
    h = batchnorm_eval(conv2d(h, p[pre + ".3.weight"], p[pre + ".3.bias"]), p, pre + ".4")
    return h.reshape(B, -1)


def cross_block(x, p, prefix, intrinsics=None, return_bilinear=False, flags=()):
    """CrossBlock.forward, vision_transformer.py:285-296.  x [2B,576,192] -> [2B,70,192]."""
    n2, N, C = x.shape
    xp = x.reshape(n2 // 2, 2, N, C)
    g, b = p[prefix + ".norm1.weight"], p[prefix + ".norm1.bias"]
    res = essential_matrix_module(layernorm(xp[:, 0], g, b), layernorm(xp[:, 1], g, b), p,
                                  prefix + ".cross_attn", intrinsics, return_bilinear, flags)
    (fa, fb), bil = (res if return_bilinear else (res, None))
    f = np.stack([fa, fb], 1).reshape(n2, EMW, C)
    out = f + mlp(layernorm(f, p[prefix + ".norm2.weight"], p[prefix + ".norm2.bias"]), p, prefix + ".mlp")
    return (out, bil) if return_bilinear else out


# ------------------------------------------------------------------ A9/A10: regressor + normalise
def pose_regressor(feat, p):
    """src/model.py:91-98,189.  feat [B,26880] -> [B,2,7]."""
    h = relu(linear(feat, p["pose_regressor.0.weight"], p["pose_regressor.0.bias"]))
    h = relu(linear(h, p["pose_regressor.2.weight"], p["pose_regressor.2.bias"]))
    return linear(h, p["pose_regressor.4.weight"], p["pose_regressor.4.bias"]).reshape(-1, 2, 7)


def normalize_preds(Gs, pose_preds):
    """src/model.py:145-152: q / max(||q||, 0.01); pose 0 := Gs[:, :1]."""
    dt = pose_preds.dtype.type
    out = np.array(pose_preds, copy=True)
    n = np.sqrt((out[:, :, 3:] ** 2).sum(-1, keepdims=True))
    out[:, :, 3:] = out[:, :, 3:] / np.maximum(n, dt(0.01))
    return np.concatenate([np.asarray(Gs, dtype=out.dtype)[:, :1], out[:, 1:]], 1)


# ------------------------------------------------------------------ the whole path
def vitess_forward(images, Gs, intrinsics, p, dtype=np.float32, depth=6, stages=None, flags=()):
    """ViTEss.forward (src/model.py:161-191), default flags.  `p` maps state-dict keys to
    numpy arrays.  If `stages` is a dict it is filled with named intermediate activations.
    Returns ([B,2,7] poses, rescaled intrinsics)."""
    images = np.asarray(images)
    B, _, _, H, W = images.shape
    x = preprocess(images, dtype)
    intr = update_intrinsics(intrinsics, H, W) if intrinsics is not None else None
    if stages is not None:
        stages["preprocessed"] = x
    fm = cnn_front_end(x, p)
    tok = tokens_from_feature_map(fm)
    if stages is not None:
        stages["tokens"] = tok
    if "cnn_only" in flags:
        feat = pool_transformer_output_cnn_only(tok, p, B)
        raw = pose_regressor(feat, p)
        if stages is not None:
            stages["features"], stages["raw_pose"] = feat, raw
        return normalize_preds(Gs, raw), intr
    x = tok + p["fusion_transformer.pos_embed"].astype(dtype)
    for i in range(depth - 1):
        x = block(x, p, f"fusion_transformer.blocks.{i}")
        if stages is not None:
            stages[f"block{i}"] = x
    if "noess" in flags:
        x = cross_block_noess(x, p, f"fusion_transformer.blocks.{depth - 1}")
    else:
        x, bil = cross_block(x, p, f"fusion_transformer.blocks.{depth - 1}", intr, return_bilinear=True, flags=flags)
        if stages is not None:
            stages["bilinear1"], stages["bilinear2"] = bil
    if stages is not None:
        stages["cross"] = x
    x = layernorm(x, p["fusion_transformer.norm.weight"], p["fusion_transformer.norm.bias"])
    feat = pool_attn_noess(x, p, B) if "noess" in flags else x.reshape(B, -1)
    if stages is not None:
        stages["features"] = feat
    raw = pose_regressor(feat, p)
    if stages is not None:
        stages["raw_pose"] = raw
    return normalize_preds(Gs, raw), intr


# ------------------------------------------------------------------ error metrics (eval scripts)
def rotation_error_rad(q_est, q_ref):
    """Angle between two rotations given as xyzw quaternions = 2*acos(|<q1,q2>|), the quantity of
    /root/reference/test_matterport.py:41 (there in degrees) -- evaluated here as
    2*atan2(|vec(q_ref^-1 q_est)|, |w|) in float64 on re-normalised inputs, because acos(1-eps)
    turns the 1e-7 rounding of float32 unit quaternions into a 5e-4 rad floor."""
    a = np.asarray(q_est, np.float64)
    b = np.asarray(q_ref, np.float64)
    a = a / np.linalg.norm(a, axis=-1, keepdims=True)
    b = b / np.linalg.norm(b, axis=-1, keepdims=True)
    av, aw = a[..., :3], a[..., 3]
    bv, bw = b[..., :3], b[..., 3]
    w = aw * bw + (av * bv).sum(-1)
    v = bw[..., None] * av - aw[..., None] * bv - np.cross(bv, av)
    return 2.0 * np.arctan2(np.linalg.norm(v, axis=-1), np.abs(w))


def translation_rel_error(t_est, t_ref):
    t_est = np.asarray(t_est, np.float64)
    t_ref = np.asarray(t_ref, np.float64)
    return np.linalg.norm(t_est - t_ref, axis=-1) / np.maximum(np.linalg.norm(t_ref, axis=-1), 1e-12)
