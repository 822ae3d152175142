"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host mirror of
the reference interface has the reference's state-dict layout, the SE3 shim's container behaviour,
and the product path refuses to run without CUDA (no silent fallback)."""
import argparse
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from rel_pose_b200 import _lib, ops, synthetic as S
from rel_pose_b200.lietorch import SE3, install_as_lietorch
from conftest import ROOT


def _args(**kw):
    d = dict(noess=False, pool_size=60, fc_hidden_size=512, fusion_transformer=True, transformer_depth=6,
             cross_features=False, use_single_softmax=False, no_pos_encoding=False, l1_pos_encoding=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "relpose_b200.h")).read()
    declared = set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.exported_symbols())
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    L = _lib.lib()
    assert L.rp_version() >= 100
    # pure host queries (no GPU needed)
    assert L.rp_linear_workspace_bytes(64, 512, 26880) > 0
    assert L.rp_linear_workspace_bytes(73728, 576, 192) == 0
    assert L.rp_essential_workspace_bytes(2) > 0


def test_state_dict_layout_matches_reference_spec():
    from rel_pose_b200 import ViTEss
    m = ViTEss(_args())
    sd = m.state_dict()
    spec = S.state_dict_spec()
    assert len(sd) == 227
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    # norm3 and downsample.1 are the same tensors (extractor.py:43-46)
    assert m.extractor_final_conv.norm3 is m.extractor_final_conv.downsample[1]
    res = m.load_state_dict(S.make_state_dict(0, "stress"), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == 29751950
    for p in list(m.resnet.layer3.parameters()) + list(m.resnet.layer4.parameters()):
        p.requires_grad = False                          # train.py:60-64
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 19258510


def test_noess_state_dict_layout():
    """--noess (model.py:71-88): proj instead of proj_fundamental, the pool_attn head, a 24768-wide regressor.  The spec
    is the one oracle/make_golden.py loads strictly into the real reference."""
    from rel_pose_b200 import ViTEss
    m = ViTEss(_args(noess=True))
    sd = m.state_dict()
    spec = S.state_dict_spec(noess=True)
    assert sorted(sd.keys()) == sorted(k for k, _, _ in spec)
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert "fusion_transformer.blocks.5.cross_attn.proj.weight" in sd and m.H == 24768
    assert tuple(sd["pool_attn.0.weight"].shape) == (96, 384, 1, 1) and tuple(sd["pool_attn.3.weight"].shape) == (43, 96, 1, 1)
    res = m.load_state_dict(S.make_state_dict(0, "stress", noess=True), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert ViTEss(_args()).noess is False and not hasattr(ViTEss(_args()), "pool_attn")


@pytest.mark.parametrize("variant", ["noess", "cnn_only"])
def test_pool_head_parameter_preparation(variant):
    """Host logic of the two pooling heads (model.py:62-88,179-187): BatchNorm folded into the 1x1 convolutions, the CNN-only
    head's weight widened over the token matrix, pose_regressor.0's columns permuted to pixel-major order.  Checked against
    the torch modules of the parameter containers themselves (eval mode), i.e. the reference's formula."""
    from rel_pose_b200 import ViTEss
    cnn_only = variant == "cnn_only"
    m = ViTEss(_args(fusion_transformer=False) if cnn_only else _args(noess=True)).eval()
    m.load_state_dict(S.make_state_dict(5, "stress", noess=not cnn_only, cnn_only=cnn_only))
    B = 2
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2 * B, 576, 192, generator=g, dtype=torch.float64)      # tokens (CNN-only) / final-norm output (noess)
    m = m.double()
    with torch.no_grad():
        # the reference: model.py:138-139,179-181 (CNN-only) / :183-187 (noess)
        if cnn_only:
            feats = x[:, :, :96].reshape(-1, 24, 24, 192).permute(0, 3, 1, 2)
            pooled = m.pool_transformer_output(feats)
        else:
            feats = x.reshape(B, 24, 24, -1).permute(0, 3, 1, 2)
            pooled = m.pool_attn(feats)
        ref = m.pose_regressor[0](pooled.reshape(B, -1))
        # the product's parameter preparation, evaluated with plain matmuls in the order the kernels run them
        w1, b1, w2, b2, w0 = m._pool_head_params()
        f = x.reshape(B * 576, 384)
        h = torch.relu(f @ w1.T + b1)
        y = (h @ w2.T + b2).reshape(B, m.H)
        got = y @ w0.T + m.pose_regressor[0].bias
    assert tuple(w1.shape) == (96, 384) and tuple(w0.shape) == (512, m.H)
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-10), float((got - ref).abs().max())
    assert m._pool_head_params()[4] is w0                  # cached until a parameter changes
    with torch.no_grad():
        m.pose_regressor[0].weight.mul_(2.0)
    assert m._pool_head_params()[4] is not w0


def test_ablation_flags():
    """--no_pos_encoding (and --noess without a transformer) cannot run in the reference either and are rejected loudly;
    --noess, the CNN-only model and the Essential-Matrix-Module variants are accepted."""
    from rel_pose_b200 import ViTEss, ops
    with pytest.raises(NotImplementedError):
        ViTEss(_args(no_pos_encoding=True))
    with pytest.raises(NotImplementedError):
        ViTEss(_args(fusion_transformer=False, noess=True))
    m = ViTEss(_args(fusion_transformer=False))
    spec = S.state_dict_spec(cnn_only=True)
    assert sorted(m.state_dict().keys()) == sorted(k for k, _, _ in spec) and m.fusion_transformer is None and m.H == 34560
    res = m.load_state_dict(S.make_state_dict(0, "init", cnn_only=True), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m = ViTEss(_args(cross_features=True, use_single_softmax=True, l1_pos_encoding=True))
    assert m.em_flags == (ops.EM_SINGLE_SOFTMAX | ops.EM_CROSS_FEATURES) and m.l1_pos_encoding
    assert ViTEss(_args()).em_flags == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-CUDA behaviour")
def test_no_cpu_fallback():
    from rel_pose_b200 import ViTEss
    m = ViTEss(_args()).eval()
    img = torch.zeros(1, 2, 3, 32, 32)
    with pytest.raises(_lib.RelposeLibraryError):
        m(img, SE3(torch.zeros(1, 2, 7)))
    with pytest.raises(_lib.RelposeLibraryError):
        ops.layernorm(torch.zeros(4, 192), torch.ones(192), torch.zeros(192))
    with pytest.raises(_lib.RelposeLibraryError):
        SE3(torch.zeros(3, 7)).inv()


def test_se3_container_behaviour():
    d = torch.arange(2 * 2 * 7, dtype=torch.float32).reshape(2, 2, 7)
    G = SE3(d)
    assert G.shape == (2, 2) and G[0][1].data.shape == (7,)
    assert torch.equal(G[:, :1].data, d[:, :1])
    jj = torch.tensor([1, 0])
    assert torch.equal(G[:, jj].data, d[:, jj])
    I = SE3.IdentityLike(G)
    assert torch.equal(I.data[..., 6], torch.ones(2, 2)) and I.data[..., :6].abs().sum() == 0
    G2 = SE3(torch.clone(G.data))
    G2.data[:, :, 3:] = 0.5            # slice-assign like model.py:151
    assert G2.data[0, 0, 3] == 0.5 and G.data[0, 0, 3] == 3.0
    assert isinstance(G.detach(), SE3)
    with pytest.raises(ValueError):
        SE3(torch.zeros(3, 6))
    mod = install_as_lietorch(force=True)
    import lietorch
    assert lietorch.SE3 is SE3 and mod.SE3 is SE3


def test_synthetic_generators_are_stable():
    u = S.hash_uniform(0, "abc", 5)
    np.testing.assert_array_equal(u, S.hash_uniform(0, "abc", 5))
    assert u.min() >= 0 and u.max() < 1
    a = S.make_images_numpy(3, 1, 8, 8)
    assert a.shape == (1, 2, 3, 8, 8) and a.dtype == np.float32 and a.max() <= 255 and np.all(a == np.floor(a))
    assert float(a.sum()) == float(S.make_images_numpy(3, 1, 8, 8).sum())


def test_one_cycle_lr_closed_form_matches_torch():
    """rel_pose_b200.optim.one_cycle_lr == torch OneCycleLR as train.py:71-73 configures it (div_factor 25, cos)."""
    import torch
    from rel_pose_b200.optim import one_cycle_lr
    w = torch.nn.Parameter(torch.zeros(1))
    for total, warm, lr in [(120000, 10000, 5e-4), (60, 7, 1e-3), (1000, 300, 2e-4)]:
        opt = torch.optim.Adam([w], lr=lr)
        sch = torch.optim.lr_scheduler.OneCycleLR(opt, lr, total, pct_start=warm / total, div_factor=25, cycle_momentum=False)
        for k in range(min(total, 1500)):
            assert abs(opt.param_groups[0]["lr"] - one_cycle_lr(k, lr, total, warm / total)) <= 1e-12 * lr
            opt.step(); sch.step()


def test_one_cycle_lr_refuses_to_run_past_the_schedule():
    """torch's OneCycleLR raises once the step count exceeds total_steps; so does the closed form (no silent climb back up)."""
    from rel_pose_b200.optim import one_cycle_lr
    one_cycle_lr(10, 1e-3, 10, 0.3)
    with pytest.raises(ValueError):
        one_cycle_lr(11, 1e-3, 10, 0.3)


def test_fused_optimizer_refuses_cpu_parameters():
    import torch
    from rel_pose_b200 import _lib
    from rel_pose_b200.optim import FusedAdamOneCycle
    with pytest.raises(_lib.RelposeLibraryError):
        FusedAdamOneCycle([torch.nn.Parameter(torch.zeros(4))], 1e-3, 10, 0.3)
