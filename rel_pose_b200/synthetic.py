"""Deterministic synthetic weights and inputs (no checkpoint / dataset is reachable offline).

Everything is derived from a counter-based integer hash (splitmix64 finaliser) using only
integer ops, float adds and multiplies -- no transcendental functions -- so the very same
bits come out on any host.  The golden fixtures under tests/golden/ store only *outputs*;
weights and inputs are regenerated from (seed, profile) at test time.

State-dict layout follows the reference ViTEss with the default flags
(`--fusion_transformer --transformer_depth 6`): 227 keys, see SURVEY.md section 8(b) and
/root/reference/src/model.py:12-98.
"""
from collections import OrderedDict

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _str_seed(name):
    h = np.uint64(0xCBF29CE484222325)
    with np.errstate(over="ignore"):
        for c in name.encode():
            h = ((h ^ np.uint64(c)) * np.uint64(0x100000001B3)) & _M64
    return h


def hash_uniform(seed, name, n, stream=0):
    """n doubles in [0,1), a pure function of (seed, name, stream, index)."""
    with np.errstate(over="ignore"):
        base = _splitmix(np.uint64(seed) ^ _str_seed(name)) + np.uint64(stream) * np.uint64(0xD1342543DE82EF95)
        idx = np.arange(n, dtype=np.uint64) + base
        z = _splitmix(idx)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hash_normal(seed, name, n):
    """Approximately N(0,1) (Irwin-Hall, 4 uniforms) -- arithmetic only, bit-reproducible."""
    s = np.zeros(n, np.float64)
    for k in range(4):
        s += hash_uniform(seed, name, n, stream=k + 1)
    return (s - 2.0) * 1.7320508075688772


# ----------------------------------------------------------------------------------------------
def state_dict_spec(transformer_depth=6, fc_hidden=512, noess=False, cnn_only=False, pool_size=60):
    """[(key, shape, kind)] in the reference's state_dict order.  kind drives the value profile.
    `noess` = the layout of the --noess ablation (model.py:71-88, vision_transformer.py:176-177); `cnn_only` = the
    model built without --fusion_transformer (model.py:62-69): no transformer, a pool_transformer_output head."""
    spec = []

    def conv(prefix, co, ci, k, bias):
        spec.append((prefix + ".weight", (co, ci, k, k), "conv"))
        if bias:
            spec.append((prefix + ".bias", (co,), "bias"))

    def bn(prefix, c):
        spec.append((prefix + ".weight", (c,), "gamma"))
        spec.append((prefix + ".bias", (c,), "beta"))
        spec.append((prefix + ".running_mean", (c,), "rmean"))
        spec.append((prefix + ".running_var", (c,), "rvar"))
        spec.append((prefix + ".num_batches_tracked", (), "count"))

    def lin(prefix, co, ci, kind="linear"):
        spec.append((prefix + ".weight", (co, ci), kind))
        spec.append((prefix + ".bias", (co,), "bias"))

    def ln(prefix, c):
        spec.append((prefix + ".weight", (c,), "gamma"))
        spec.append((prefix + ".bias", (c,), "beta"))

    # torchvision resnet18 minus fc (layer3/4 are unused by forward but must be present)
    conv("resnet.conv1", 64, 3, 7, False)
    bn("resnet.bn1", 64)
    cin = 64
    for li, cout in enumerate([64, 128, 256, 512], start=1):
        for bi in range(2):
            p = f"resnet.layer{li}.{bi}"
            conv(p + ".conv1", cout, cin if bi == 0 else cout, 3, False)
            bn(p + ".bn1", cout)
            conv(p + ".conv2", cout, cout, 3, False)
            bn(p + ".bn2", cout)
            if bi == 0 and li > 1:
                conv(p + ".downsample.0", cout, cin, 1, False)
                bn(p + ".downsample.1", cout)
        cin = cout
    # ResidualBlock(128,192,'batch',kernel_size=5) -- extractor.py:5-49; norm3 is downsample.1
    e = "extractor_final_conv"
    conv(e + ".conv1", 192, 128, 3, True)
    conv(e + ".conv2", 192, 192, 5, True)
    bn(e + ".norm1", 192)
    bn(e + ".norm2", 192)
    bn(e + ".norm3", 192)
    conv(e + ".downsample.0", 192, 128, 5, True)
    bn(e + ".downsample.1", 192)  # alias of norm3 (same tensors)
    if cnn_only:
        pf1 = min(96, 4 * pool_size)
        conv("pool_transformer_output.0", pf1, 192, 1, True)
        bn("pool_transformer_output.1", pf1)
        conv("pool_transformer_output.3", pool_size, pf1, 1, True)
        bn("pool_transformer_output.4", pool_size)
        lin("pose_regressor.0", fc_hidden, pool_size * 576)
        lin("pose_regressor.2", fc_hidden, fc_hidden)
        lin("pose_regressor.4", 14, fc_hidden)
        return spec
    f = "fusion_transformer"
    spec.append((f + ".pos_embed", (1, 576, 192), "posemb"))
    for i in range(transformer_depth):
        b = f"{f}.blocks.{i}"
        ln(b + ".norm1", 192)
        if i == transformer_depth - 1:
            lin(b + ".cross_attn.qkv", 576, 192, "qkv")
            if noess:
                lin(b + ".cross_attn.proj", 192, 192)
            else:
                lin(b + ".cross_attn.proj_fundamental", 192, 210)
        else:
            lin(b + ".attn.qkv", 576, 192, "qkv")
            lin(b + ".attn.proj", 192, 192)
        ln(b + ".norm2", 192)
        lin(b + ".mlp.fc1", 768, 192)
        lin(b + ".mlp.fc2", 192, 768)
    ln(f + ".norm", 192)
    H = 3 * 2 * (64 + 6) * 64
    if noess:
        H = 24 * 24 * 43
        conv("pool_attn.0", 96, 384, 1, True)
        bn("pool_attn.1", 96)
        conv("pool_attn.3", 43, 96, 1, True)
        bn("pool_attn.4", 43)
    lin("pose_regressor.0", fc_hidden, H)
    lin("pose_regressor.2", fc_hidden, fc_hidden)
    lin("pose_regressor.4", 14, fc_hidden)
    return spec


ALIASES = {"extractor_final_conv.downsample.1": "extractor_final_conv.norm3"}


def make_state_dict_numpy(seed=0, profile="stress", transformer_depth=6, fc_hidden=512, noess=False, cnn_only=False):
    """OrderedDict key -> np.ndarray (float32, int64 for num_batches_tracked).

    profile "init":   statistics close to the reference's random init (small ViT weights,
                      BN running stats 0/1).
    profile "stress": non-trivial BN statistics / affine terms, larger qkv weights so the
                      softmaxes are far from uniform; exercises every term of every formula.
    """
    assert profile in ("init", "stress")
    out = OrderedDict()
    for key, shape, kind in state_dict_spec(transformer_depth, fc_hidden, noess, cnn_only):
        gen_key = key
        for a, tgt in ALIASES.items():
            if key.startswith(a + "."):
                gen_key = tgt + key[len(a):]
        n = int(np.prod(shape)) if len(shape) else 1
        if kind == "count":
            out[key] = np.array(0 if profile == "init" else 7, dtype=np.int64)
            continue
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            v = hash_normal(seed, gen_key, n) * np.sqrt(2.0 / fan_in)
        elif kind == "linear":
            fan_in = shape[1]
            bound = 1.0 / np.sqrt(fan_in)
            if profile == "init" and "fusion_transformer" in key:
                v = hash_normal(seed, gen_key, n) * 0.02
            else:
                v = (hash_uniform(seed, gen_key, n) * 2.0 - 1.0) * bound
                if profile == "stress" and key.startswith("pose_regressor"):
                    v = v * 2.5   # make the pose depend strongly on the features, not on the biases
        elif kind == "qkv":
            std = 0.02 if profile == "init" else 0.09
            v = hash_normal(seed, gen_key, n) * std
        elif kind == "bias":
            amp = 0.0 if (profile == "init" and "fusion_transformer" in key) else 0.05
            if profile == "stress" and key.startswith("pose_regressor"):
                amp = 0.01
            v = (hash_uniform(seed, gen_key, n) * 2.0 - 1.0) * amp
        elif kind == "gamma":
            v = np.ones(n) if profile == "init" else 1.0 + 0.2 * (hash_uniform(seed, gen_key, n) - 0.5)
        elif kind == "beta":
            v = np.zeros(n) if profile == "init" else 0.1 * (hash_uniform(seed, gen_key, n) - 0.5)
        elif kind == "rmean":
            v = np.zeros(n) if profile == "init" else 0.2 * (hash_uniform(seed, gen_key, n) - 0.5)
        elif kind == "rvar":
            v = np.ones(n) if profile == "init" else 0.6 + 0.8 * hash_uniform(seed, gen_key, n)
        elif kind == "posemb":
            bound = np.sqrt(6.0 / (576 * 192 + 192))  # xavier_uniform on (1,576,192) -- model.py:54-56
            v = (hash_uniform(seed, gen_key, n) * 2.0 - 1.0) * (bound if profile == "init" else 0.3)
        else:
            raise AssertionError(kind)
        out[key] = v.astype(np.float32).reshape(shape)
    return out


def make_state_dict(seed=0, profile="stress", **kw):
    import torch
    sd = OrderedDict()
    for k, v in make_state_dict_numpy(seed, profile, **kw).items():
        sd[k] = torch.from_numpy(np.ascontiguousarray(v))
    return sd


def make_images_numpy(seed, B, H, W, integer_pixels=True):
    """[B,2,3,H,W] float32 BGR in 0..255 (cv2 convention, /root/reference/demo.py:65-76).
    The two views of a pair are correlated (shifted copies of a smooth-ish random field plus
    noise) so that the cross-attention affinities are not degenerate."""
    base = hash_uniform(seed, "images", B * 3 * (H + 8) * (W + 8)).reshape(B, 3, H + 8, W + 8)
    # cheap smoothing: average of shifted copies (arithmetic only)
    sm = (base + np.roll(base, 1, 2) + np.roll(base, 1, 3) + np.roll(base, (1, 1), (2, 3))) * 0.25
    img = np.empty((B, 2, 3, H, W), np.float64)
    img[:, 0] = sm[:, :, 0:H, 0:W]
    img[:, 1] = sm[:, :, 5:H + 5, 3:W + 3]
    img += 0.15 * (hash_uniform(seed, "noise", img.size).reshape(img.shape) - 0.5)
    img = np.clip(img, 0.0, 1.0) * 255.0
    if integer_pixels:
        img = np.floor(img)
    return img.astype(np.float32)


def make_intrinsics_numpy(B, kind="matterport", seed=0):
    """[B,2,4] = (fx,fy,cx,cy) identical for both views of a pair (vision_transformer.py:117)."""
    if kind == "matterport":          # demo.py:52-53
        k = np.array([517.97, 517.97, 320.0, 240.0], np.float32)
        return np.tile(k, (B, 2, 1)).astype(np.float32)
    if kind == "square":              # demo.py:55
        return np.full((B, 2, 4), 128.0, np.float32)
    if kind == "varied":              # different per pair, still equal within a pair
        u = hash_uniform(seed, "intrinsics", B * 4).reshape(B, 1, 4)
        k = np.array([400.0, 420.0, 300.0, 230.0]) + 120.0 * u
        return np.tile(k, (1, 2, 1)).astype(np.float32)
    raise ValueError(kind)


def make_poses_numpy(seed, B):
    """[B,2,7] ground-truth poses (tx,ty,tz,qx,qy,qz,qw): pose 0 identity, pose 1 random."""
    u = hash_uniform(seed, "poses", B * 7).reshape(B, 7) * 2.0 - 1.0
    q = u[:, 3:] + np.array([0, 0, 0, 1.5])
    q /= np.sqrt((q * q).sum(-1, keepdims=True))
    p = np.zeros((B, 2, 7), np.float64)
    p[:, 0, 6] = 1.0
    p[:, 1, :3] = u[:, :3] * 0.5
    p[:, 1, 3:] = q
    return p.astype(np.float32)
