"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference from /root/reference.

Used solely by oracle/make_golden.py (golden-vector generation in the build container)
and by bench.py's `--impl reference` / cpu_baseline leg when the reference tree is
present.  Nothing in rel_pose_b200/ may import this file.

The reference (crockwell/rel_pose @35d1352) cannot be imported as-is offline:
  * `lietorch` (third-party C++/CUDA ext, environment.yml:20) is not installed
    -> a minimal stand-in module exposing `SE3` is registered (src/model.py:9).
  * `models.resnet18(pretrained=True)` (src/model.py:31) needs the network
    -> wrapped so that `pretrained` is ignored (random init; weights are always
       overwritten by load_state_dict afterwards).
  * `get_positional_encodings(...).cuda()` (src/modules/vision_transformer.py:211)
    is unconditional -> `.cuda()` becomes identity when no CUDA device exists.
None of this changes the reference's arithmetic.
"""
import os
import sys
import types
import argparse

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def find_reference_root():
    """$RELPOSE_REFERENCE_ROOT, else /root/reference (build container), else baseline/_ref (the byte-for-byte copy
    oracle/install_reference.py makes; it travels to the GPU box).  None when no tree is present."""
    cands = [os.environ.get("RELPOSE_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "model.py")):
            return c
    return None


REFERENCE_ROOT = find_reference_root() or "/root/reference"


class _StubSE3:
    """Container-only stand-in for lietorch.SE3: what ViTEss.forward touches
    (src/model.py:146-152,163) plus the group ops geodesic_loss needs
    (src/geom/losses.py:8-10), written with plain torch in float64-safe form."""

    def __init__(self, data):
        self.data = data

    def __getitem__(self, idx):
        return _StubSE3(self.data[idx])

    @staticmethod
    def IdentityLike(other):
        d = torch.zeros_like(other.data)
        d[..., 6] = 1.0
        return _StubSE3(d)

    def detach(self):
        return _StubSE3(self.data.detach())

    # --- group maths (forward only; gradients here are plain autograd, NOT lietorch's) ---
    @staticmethod
    def _qmul(a, b):
        ax, ay, az, aw = a.unbind(-1)
        bx, by, bz, bw = b.unbind(-1)
        return torch.stack([
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz], -1)

    @staticmethod
    def _qrot(q, v):
        qv, w = q[..., :3], q[..., 3:]
        uv = 2.0 * torch.cross(qv, v, dim=-1)
        return v + w * uv + torch.cross(qv, uv, dim=-1)

    def __mul__(self, other):
        t1, q1 = self.data[..., :3], self.data[..., 3:]
        t2, q2 = other.data[..., :3], other.data[..., 3:]
        return _StubSE3(torch.cat([t1 + self._qrot(q1, t2), self._qmul(q1, q2)], -1))

    def inv(self):
        t, q = self.data[..., :3], self.data[..., 3:]
        qi = q * torch.tensor([-1.0, -1.0, -1.0, 1.0], dtype=q.dtype, device=q.device)
        return _StubSE3(torch.cat([-self._qrot(qi, t), qi], -1))


def install_shims():
    if "lietorch" not in sys.modules:
        m = types.ModuleType("lietorch")
        m.SE3 = _StubSE3
        sys.modules["lietorch"] = m
    import torchvision.models as tvm
    if not getattr(tvm.resnet18, "_relpose_offline", False):
        _orig = tvm.resnet18

        def resnet18(pretrained=False, **kw):
            kw.pop("weights", None)              # callers that already use the new API (rel_pose_b200.model) pass weights=None
            return _orig(weights=None, **kw)

        resnet18._relpose_offline = True
        tvm.resnet18 = resnet18
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self


class cpu_only:
    """Context manager: run the reference on the CPU of a box that HAS a CUDA device.  The reference moves its positional
    table with an unconditional `.cuda()` (vision_transformer.py:211); for a CPU run of the model that call must be the
    identity (exactly what install_shims() does permanently on a GPU-less box)."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig
        return False


def default_args(**over):
    """Flags every reference script uses (scripts/train_matterport.sh:6-9, demo.py:36-46)."""
    d = dict(noess=False, pool_size=60, fc_hidden_size=512, fusion_transformer=True,
             transformer_depth=6, cross_features=False, use_single_softmax=False,
             no_pos_encoding=False, l1_pos_encoding=False)
    d.update(over)
    return argparse.Namespace(**d)


def load_reference_model(**over):
    """Returns (ViTEss instance on CPU in eval mode, SE3 stub class).  The model may be moved to a CUDA device by the
    caller (bench.py's gpu_eager_baseline leg): the reference then runs exactly as its scripts run it."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # `src` must resolve to the reference package, not anything of ours
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        f = getattr(sys.modules[k], "__file__", "") or ""
        if not f.startswith(REFERENCE_ROOT):
            del sys.modules[k]
    from src.model import ViTEss  # noqa: E402  (the reference's own class)
    model = ViTEss(default_args(**over)).eval()
    return model, sys.modules["lietorch"].SE3
