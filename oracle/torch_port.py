"""ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.  PyTorch port of the reference forward path (CPU; also runs on a CUDA
device through torch's own eager kernels -- cuBLAS / cuDNN -- for bench.py's `gpu_eager_baseline` leg).

The reference's own CPU implementation of this path *is* PyTorch on CPU (ATen + oneDNN/MKL); the
reference tree does not exist on the GPU box, so `bench.py --impl reference` and the `cpu_baseline`
leg time this functional port instead ("kind": "port"): same library kernels the reference would
dispatch to, same arithmetic (float32), the state dict passed in explicitly.  It is pinned to the
golden vectors of the real reference in tests/test_oracle_golden.py.  One deliberate difference
makes the baseline FASTER than the reference: the positional table is evaluated in closed form
instead of the reference's 576-iteration host loop (vision_transformer.py:139-151).

Only tests/, smoke() and bench.py may import this; the product path never does.
Follows: src/model.py:114-191; src/modules/vision_transformer.py:188-238,285-296,321-354;
src/modules/vit_layers/mlp.py:20-26; src/modules/extractor.py:51-65.
"""
import numpy as np
import torch
import torch.nn.functional as F


TRAIN_BN = False     # set by forward(train=True): nn.BatchNorm2d.train() semantics (batch statistics, running-stat update)


def _bn(x, p, pre):
    return F.batch_norm(x, p[pre + ".running_mean"], p[pre + ".running_var"], p[pre + ".weight"], p[pre + ".bias"],
                        training=TRAIN_BN, momentum=0.1, eps=1e-5)


def _basic(x, p, pre, stride):
    y = F.relu(_bn(F.conv2d(x, p[pre + ".conv1.weight"], None, stride, 1), p, pre + ".bn1"))
    y = _bn(F.conv2d(y, p[pre + ".conv2.weight"], None, 1, 1), p, pre + ".bn2")
    if pre + ".downsample.0.weight" in p:
        x = _bn(F.conv2d(x, p[pre + ".downsample.0.weight"], None, stride, 0), p, pre + ".downsample.1")
    return F.relu(x + y)


def _ln(x, p, pre):
    return F.layer_norm(x, (x.shape[-1],), p[pre + ".weight"], p[pre + ".bias"], 1e-6)


def _mlp(x, p, pre):
    return F.linear(F.gelu(F.linear(x, p[pre + ".fc1.weight"], p[pre + ".fc1.bias"])), p[pre + ".fc2.weight"],
                    p[pre + ".fc2.bias"])


def _qkv(x, p, pre):
    n, N, C = x.shape
    t = F.linear(x, p[pre + ".qkv.weight"], p[pre + ".qkv.bias"]).reshape(n, N, 3, 3, 64).permute(2, 0, 3, 1, 4)
    return t[0], t[1], t[2]


def _posenc(B, intr, device="cpu"):
    lin = torch.linspace(-1, 1, 24).to(device)      # evaluated on the host like the reference, then moved
    i = torch.arange(576, device=device)
    p3 = lin[i % 24].repeat(B, 1)
    p4 = lin[i // 24].repeat(B, 1)
    if intr is not None:
        fx, fy, cx, cy = intr[:, 0].unbind(-1)
        kx = 1.0 / ((fx / (cx * 2)) * 2)
        ky = 1.0 / ((fy / (cy * 2)) * 2)
        p3 = ky[:, None] * lin[i % 24][None]
        p4 = kx[:, None] * lin[i // 24][None]
    return torch.stack([p3 * p3, p4 * p4, p3 * p4, p3, p4, torch.ones_like(p3)], -1)


def forward(images, Gs, intrinsics, p, depth=6, train=False):
    """images [B,2,3,H,W] float32 BGR, Gs [B,2,7], intrinsics [B,2,4] or None (NOT mutated),
    p: dict of float32 CPU tensors.  Returns [B,2,7].  train=True: BatchNorm in training mode (train.py:140-155;
    the running statistics in `p` are updated in place like the reference's buffers), differentiable by autograd."""
    global TRAIN_BN
    TRAIN_BN = bool(train)
    B, _, _, H, W = images.shape
    x = images[:, :, [2, 1, 0]] / 255.0
    dev = images.device       # CPU for the oracle / cpu_baseline; a CUDA device for bench.py's gpu_eager_baseline leg
    x = (x - torch.tensor([0.485, 0.456, 0.406], device=dev)[:, None, None]) / torch.tensor([0.229, 0.224, 0.225], device=dev)[:, None, None]
    intr = None
    if intrinsics is not None:
        intr = intrinsics.clone()
        intr[:, :, [0, 2]] *= 24 / W
        intr[:, :, [1, 3]] *= 24 / H
    x = F.interpolate(x.flatten(0, 1), size=224)
    x = F.relu(_bn(F.conv2d(x, p["resnet.conv1.weight"], None, 2, 3), p, "resnet.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    x = _basic(x, p, "resnet.layer1.0", 1)
    x = _basic(x, p, "resnet.layer1.1", 1)
    x = _basic(x, p, "resnet.layer2.0", 2)
    x = _basic(x, p, "resnet.layer2.1", 1)
    e = "extractor_final_conv"
    y = F.relu(_bn(F.conv2d(x, p[e + ".conv1.weight"], p[e + ".conv1.bias"], 1, 1), p, e + ".norm1"))
    y = F.relu(_bn(F.conv2d(y, p[e + ".conv2.weight"], p[e + ".conv2.bias"]), p, e + ".norm2"))
    s = _bn(F.conv2d(x, p[e + ".downsample.0.weight"], p[e + ".downsample.0.bias"]), p, e + ".norm3")
    x = F.relu(s + y).reshape(2 * B, 192, 576).permute(0, 2, 1) + p["fusion_transformer.pos_embed"]
    for i in range(depth - 1):
        b = f"fusion_transformer.blocks.{i}"
        q, k, v = _qkv(_ln(x, p, b + ".norm1"), p, b + ".attn")
        a = ((q @ k.transpose(-2, -1)) * 0.125).softmax(-1) @ v
        a = a.transpose(1, 2).reshape(2 * B, 576, 192)
        x = x + F.linear(a, p[b + ".attn.proj.weight"], p[b + ".attn.proj.bias"])
        x = x + _mlp(_ln(x, p, b + ".norm2"), p, b + ".mlp")
    b = f"fusion_transformer.blocks.{depth - 1}"
    xp = x.reshape(B, 2, 576, 192)
    q1, k1, v1 = _qkv(_ln(xp[:, 0], p, b + ".norm1"), p, b + ".cross_attn")
    q2, k2, v2 = _qkv(_ln(xp[:, 1], p, b + ".norm1"), p, b + ".cross_attn")
    s1 = (q2 @ k1.transpose(-2, -1)) * 0.125
    s2 = (q1 @ k2.transpose(-2, -1)) * 0.125
    a1 = s1.softmax(-1) * s1.softmax(-2)
    a2 = s2.softmax(-1) * s2.softmax(-2)
    pos = _posenc(B, intr, dev)[:, None].expand(B, 3, 576, 6)
    V1 = torch.cat([v1, pos], 3)
    V2 = torch.cat([v2, pos], 3)
    f1 = (V1.transpose(-2, -1) @ a1) @ V1
    f2 = (V2.transpose(-2, -1) @ a2) @ V2
    w, bb = p[b + ".cross_attn.proj_fundamental.weight"], p[b + ".cross_attn.proj_fundamental.bias"]
    y1 = F.linear(f1.reshape(B, 210, 70).transpose(-2, -1), w, bb)
    y2 = F.linear(f2.reshape(B, 210, 70).transpose(-2, -1), w, bb)
    f = torch.stack([y2, y1], 1).reshape(2 * B, 70, 192)
    f = f + _mlp(_ln(f, p, b + ".norm2"), p, b + ".mlp")
    feat = _ln(f, p, "fusion_transformer.norm").reshape(B, -1)
    h = F.relu(F.linear(feat, p["pose_regressor.0.weight"], p["pose_regressor.0.bias"]))
    h = F.relu(F.linear(h, p["pose_regressor.2.weight"], p["pose_regressor.2.bias"]))
    raw = F.linear(h, p["pose_regressor.4.weight"], p["pose_regressor.4.bias"]).reshape(B, 2, 7)
    qn = raw[:, :, 3:].norm(dim=-1, keepdim=True).clamp_min(0.01)
    out = torch.cat([raw[:, :, :3], raw[:, :, 3:] / qn], -1)
    return torch.cat([Gs[:, :1], out[:, 1:]], 1)


def forward_numpy(images, Gs, intrinsics, p_np, depth=6):
    p = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p_np.items() if v.dtype != np.int64}
    with torch.no_grad():
        out = forward(torch.from_numpy(np.ascontiguousarray(images)), torch.from_numpy(np.ascontiguousarray(Gs)),
                      None if intrinsics is None else torch.from_numpy(np.ascontiguousarray(intrinsics)), p, depth)
    return out.numpy()
