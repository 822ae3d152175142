"""Generates tests/golden/train_*.npz: reference gradients of the TRAINING step for the GPU backward tests.

The reference's training step (train.py:140-160) is its model in .train() mode (BatchNorm on batch statistics, running
statistics updated) followed by geodesic_loss on lietorch SE3 objects and loss.backward().  The goldens come from the
UNMODIFIED reference model (oracle/ref_loader.py, /root/reference) run in .train() and float64:
  * forward: `ViTEss(args).train().double()(images, SE3(Gs), intrinsics)` -> poses_est[0].data  [B,2,7];
  * the loss gradient is injected at that tensor, `poses_est[0].data.backward(g_pose)` (the point where lietorch hands
    its tangent-space gradient back to ordinary autograd, src/geom/losses.py:8-14), and torch's autograd differentiates
    the reference's own modules from there: every parameter gradient below is the reference's;
  * g_pose = d(10 tr + 10 rot)/d(pose) from oracle/geom_oracle.geodesic_loss_grad in lietorch's left-perturbation
    convention.  lietorch itself is absent (parity of that 7-vector unpinned, see geom_oracle.py); everything upstream
    of it -- 19.3 M parameters' gradients and the BatchNorm buffers -- is pinned to the real reference.
The PyTorch port (oracle/torch_port.forward(train=True)) is run beside it and must agree to 1e-9: the port stays pinned
in train mode too.  Stored: poses, the pose gradient, every parameter gradient (full for small tensors, a strided sample
+ norm for large ones) and the updated BatchNorm running statistics.
Run in the build container: python oracle/make_golden_train.py
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
import ref_loader  # noqa: E402
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


def port_step(sd, images, intr, Gs, Ps):
    """Same step through the port (float64): returns (poses, {name: grad}, {name: buffer})."""
    p = {k: (v.double().clone().requires_grad_("running" not in k) if v.dtype.is_floating_point else v.clone())
         for k, v in sd.items()}
    out = torch_port.forward(images, Gs, intr, p, train=True)
    g_pose = G.geodesic_loss_grad(Ps, out.detach().numpy(), 10.0, 10.0)
    out.backward(torch.from_numpy(g_pose))
    grads = {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    return out.detach().numpy(), grads, {k: v.detach() for k, v in p.items() if "running" in k}


def run_case(name, seed, profile, B, H, W):
    sd = S.make_state_dict(seed, profile)
    images = torch.from_numpy(S.make_images_numpy(seed, B, H, W, True)).double()
    intr = torch.from_numpy(S.make_intrinsics_numpy(B)).double()
    Gs = torch.zeros(B, 2, 7, dtype=torch.float64); Gs[..., 6] = 1
    # ---- the unmodified reference, train mode, float64
    model, SE3 = ref_loader.load_reference_model()
    model.load_state_dict(sd)
    model = model.double().train()
    for q in list(model.resnet.layer3.parameters()) + list(model.resnet.layer4.parameters()):
        q.requires_grad = False                                   # train.py:60-64
    poses_est = model(images.clone(), SE3(Gs), intrinsics=intr.clone())
    out = poses_est[0].data
    poses = out.detach().numpy()
    Ps = target_poses(seed, B)
    g_pose = G.geodesic_loss_grad(Ps, poses, 10.0, 10.0)          # [B,2,7], lietorch convention
    out.backward(torch.from_numpy(g_pose))
    params = dict(model.named_parameters())
    bufs = {k: v for k, v in model.state_dict().items() if "running" in k}
    # ---- the port must agree (keeps oracle/torch_port.py pinned in train mode)
    poses_p, grads_p, bufs_p = port_step(sd, images, intr, Gs, Ps)
    worst = float(np.abs(poses_p - poses).max())
    gscale = float(np.median([float(q.grad.abs().max()) for q in params.values() if q.grad is not None]))
    for k, gp in grads_p.items():
        gr = params[k].grad
        assert gr is not None, k
        # biases in front of a train-mode BatchNorm have an exactly-zero gradient (1e-14 of rounding on both sides):
        # measured against the step's typical gradient scale
        worst = max(worst, float((gp - gr).abs().max() / max(float(gr.abs().max()), 1e-6 * gscale)))
    for k, bp in bufs_p.items():
        if "extractor_final_conv.downsample.1" in k:
            continue       # the reference's downsample[1] IS norm3 (extractor.py:44-48); the port updates it under norm3.*
        worst = max(worst, float((bp - bufs[k]).abs().max() / (bufs[k].abs().max() + 1e-30)))
    # (the reference evaluates its positional table in float32 even in a float64 model: agreement is ~4e-8, not 1e-15)
    assert worst < 1e-6, f"port disagrees with the reference in train mode: {worst:.3e}"
    rec = {"meta": np.array([seed, B, H, W]), "profile": np.array(profile), "poses": poses, "target": Ps, "g_pose": g_pose,
           "source": np.array("unmodified reference (train mode, float64); port agreement %.2e" % worst)}
    for k, v in bufs.items():
        rec["buf/" + k] = v.detach().numpy()
    for k, v in params.items():
        if not v.requires_grad:
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
          "without", n_none, "port-vs-reference", f"{worst:.2e}", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    run_case("train_b2_64x80", 7, "stress", 2, 64, 80)
    run_case("train_b1_48x48_init", 8, "init", 1, 48, 48)
