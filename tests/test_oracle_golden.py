"""Pins the numpy oracle (oracle/relpose_oracle.py) to golden vectors produced by the real
reference (oracle/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

import relpose_oracle as O
from rel_pose_b200 import synthetic as S
from conftest import GOLDEN

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "posenc" not in p and not os.path.basename(p).startswith("train_"))
TOK = (slice(None), slice(None, None, 9), slice(None, None, 4))


def _load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    seed, B, H, W, integer = (int(v) for v in g["meta"])
    profile = str(g["profile"])
    ikind = str(g["intrinsics_kind"])
    flags = tuple(f for f in (str(g["flags"]).split(",") if "flags" in g.files else []) if f)   # ablation branches (8 f-4)
    return g, seed, B, H, W, bool(integer), profile, (None if ikind == "none" else ikind), flags


def test_have_cases():
    assert len(CASES) >= 5


def test_linspace_matches_torch():
    import torch
    np.testing.assert_allclose(O.linspace_pm1(24), torch.linspace(-1, 1, 24).numpy(), rtol=0, atol=6e-8)


def test_nearest_index_matches_torch():
    import torch
    import torch.nn.functional as F
    for n in (384, 512, 96, 128, 64, 80, 48, 480, 640, 224, 100):
        src = torch.arange(n, dtype=torch.float32).reshape(1, 1, 1, n)
        got = F.interpolate(src, size=(1, 224)).reshape(-1).numpy().astype(np.int64)
        assert np.array_equal(O.nearest_src_index(224, n), got), n


def test_posenc_golden_bit_exact():
    g = np.load(os.path.join(GOLDEN, "posenc.npz"))
    np.testing.assert_allclose(O.positional_encodings(2, None), g["none"], rtol=0, atol=1.2e-7)
    for nm in ("matterport_24", "square", "odd"):
        got = O.positional_encodings(1, g[nm + "_k"])
        # reference goes through torch.inverse (LU) and a 3x3 matvec: allow 1 ulp-level noise
        np.testing.assert_allclose(got, g[nm], rtol=0, atol=2.5e-7)


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    g, seed, B, H, W, integer, profile, ikind, flags = _load(name)
    p = S.make_state_dict_numpy(seed, profile, noess="noess" in flags, cnn_only="cnn_only" in flags)
    images = S.make_images_numpy(seed, B, H, W, integer)
    intr = None if ikind is None else S.make_intrinsics_numpy(B, ikind, seed)
    Gs = np.zeros((B, 2, 7), np.float32)
    Gs[..., 6] = 1
    st = {}
    poses, intr_after = O.vitess_forward(images, Gs, intr, p, np.float32, stages=st, flags=flags)
    # bit-exact index work
    assert np.array_equal(st["preprocessed"][:, :, ::7, ::5], g["stage_preprocessed"])
    if intr is not None:
        assert np.array_equal(intr_after, g["intrinsics_after"])
    tol = dict(rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(st["tokens"][TOK], g["stage_tokens"], **tol)
    for i in range(0 if "cnn_only" in flags else 5):
        np.testing.assert_allclose(st[f"block{i}"][TOK], g[f"stage_block{i}"], rtol=5e-4, atol=5e-4)
    # absolute floor relative to the magnitude of the forms: with --use_single_softmax they are ~576x larger
    for kk in (() if ("noess" in flags or "cnn_only" in flags) else ("bilinear1", "bilinear2")):
        np.testing.assert_allclose(st[kk], g["stage_" + kk], rtol=1e-3, atol=2e-5 + 2e-7 * np.abs(g["stage_" + kk]).max())
    if "noess" in flags:
        np.testing.assert_allclose(st["cross"][TOK], g["stage_cross"], rtol=5e-4, atol=5e-4)
    np.testing.assert_allclose(st["features"][:, ::3], g["stage_features"], rtol=1e-3, atol=1e-3)
    rot = O.rotation_error_rad(poses[:, 1, 3:], g["poses"][:, 1, 3:])
    tr = O.translation_rel_error(poses[:, 1, :3], g["poses"][:, 1, :3])
    assert rot.max() < 1e-4, rot
    assert tr.max() < 1e-4, tr
    assert np.array_equal(poses[:, 0], Gs[:, 0])


@pytest.mark.parametrize("name", CASES)
def test_torch_port_matches_reference(name):
    """The CPU-baseline port (oracle/torch_port.py) is pinned to the same golden vectors."""
    import torch_port
    g, seed, B, H, W, integer, profile, ikind, flags = _load(name)
    if flags:
        pytest.skip("the CPU-baseline port covers the default configuration only")
    p = S.make_state_dict_numpy(seed, profile)
    images = S.make_images_numpy(seed, B, H, W, integer)
    intr = None if ikind is None else S.make_intrinsics_numpy(B, ikind, seed)
    Gs = np.zeros((B, 2, 7), np.float32)
    Gs[..., 6] = 1
    poses = torch_port.forward_numpy(images, Gs, intr, p)
    rot = O.rotation_error_rad(poses[:, 1, 3:], g["poses"][:, 1, 3:])
    tr = O.translation_rel_error(poses[:, 1, :3], g["poses"][:, 1, :3])
    assert rot.max() < 1e-4 and tr.max() < 1e-4, (rot, tr)
    assert np.array_equal(poses[:, 0], Gs[:, 0])
