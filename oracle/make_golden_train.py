"""Generates tests/golden/train_*.npz: reference gradients of the TRAINING step for the GPU backward tests.

The reference's training forward (train.py:140-160) is its model in .train() mode (BatchNorm on batch statistics)
followed by geodesic_loss on lietorch SE3 objects.  lietorch is absent here (parity unpinned, see geom_oracle.py), so
the chain is assembled from the two pinned/self-checked pieces:
  * oracle/torch_port.forward(train=True): the PyTorch-CPU port of ViTEss.forward (pinned to the reference's golden
    vectors in eval mode; the only train-mode difference is F.batch_norm(training=True)), differentiated by autograd
    in float64;
  * oracle/geom_oracle.geodesic_loss_grad: d(10 tr + 10 rot)/d(pose) in lietorch's tangent-space convention.
Stored: poses, the pose gradient, every parameter gradient (full for small tensors, a strided sample + norm for large
ones) and the updated BatchNorm running statistics.  Run in the build container: python oracle/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import geom_oracle as G  # noqa: E402
import torch_port  # noqa: E402
from rel_pose_b200 import synthetic as S  # noqa: E402

SAMPLE_STRIDE = 97
FULL_BELOW = 20000


def target_poses(seed, B):
    """Ground-truth SE3 poses [B,2,7]: identity for view 0, a moderate random motion for view 1."""
    xi = S.hash_normal(seed, "gt_pose", B * 6).reshape(B, 6) * np.array([0.3, 0.3, 0.3, 0.2, 0.2, 0.2])
    P = np.zeros((B, 2, 7)); P[..., 6] = 1
    P[:, 1] = G.se3_exp(xi)
    return P


def run_case(name, seed, profile, B, H, W):
    sd = S.make_state_dict(seed, profile)
    p = {k: v.double().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and "num_batches" not in k)
         if v.dtype.is_floating_point else v.clone() for k, v in sd.items()}
    for k in p:
        if "running" in k:
            p[k] = sd[k].double().clone()
    images = torch.from_numpy(S.make_images_numpy(seed, B, H, W, True)).double()
    intr = torch.from_numpy(S.make_intrinsics_numpy(B)).double()
    Gs = torch.zeros(B, 2, 7, dtype=torch.float64); Gs[..., 6] = 1
    out = torch_port.forward(images, Gs, intr, p, train=True)
    poses = out.detach().numpy()
    Ps = target_poses(seed, B)
    g_pose = G.geodesic_loss_grad(Ps, poses, 10.0, 10.0)          # [B,2,7], lietorch convention
    out.backward(torch.from_numpy(g_pose))
    rec = {"meta": np.array([seed, B, H, W]), "profile": np.array(profile), "poses": poses, "target": Ps, "g_pose": g_pose}
    for k, v in p.items():
        if "running" in k:
            rec["buf/" + k] = v.numpy()
        if not (torch.is_tensor(v) and v.requires_grad):
            continue
        if v.grad is None:
            rec["none/" + k] = np.zeros(0)
            continue
        g = v.grad.numpy().ravel()
        rec["norm/" + k] = np.array([np.linalg.norm(g), np.abs(g).max()])
        rec["grad/" + k] = g if g.size <= FULL_BELOW else g[::SAMPLE_STRIDE].copy()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **rec)
    n_none = sum(1 for k in rec if k.startswith("none/"))
    print(name, "poses", poses[0, 1], "|g_pose|", np.abs(g_pose).max(), "params with grad", sum(1 for k in rec if k.startswith("grad/")),
          "without", n_none, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    run_case("train_b2_64x80", 7, "stress", 2, 64, 80)
    run_case("train_b1_48x48_init", 8, "init", 1, 48, 48)
