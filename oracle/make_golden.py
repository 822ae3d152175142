"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_loader.py) on deterministic synthetic weights and inputs.

Run in the build container only (`python oracle/make_golden.py`); the GPU box has no
/root/reference, it only reads the committed fixtures.  Fixtures hold outputs and strided
samples of intermediate activations; weights/inputs are regenerated from (seed, profile) by
rel_pose_b200.synthetic at test time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from rel_pose_b200 import synthetic as S  # noqa: E402

CASES = [
    # name, seed, profile, B, H, W, intrinsics kind, integer pixels
    ("stress_b2_96x128", 0, "stress", 2, 96, 128, "matterport", True),
    ("init_b1_384x384", 1, "init", 1, 384, 384, "square", True),
    ("stress_b3_64x80_varied", 2, "stress", 3, 64, 80, "varied", False),
    ("stress_b1_nointr", 3, "stress", 1, 48, 48, None, True),
    ("init_b1_384x512_demo_shape", 4, "init", 1, 384, 512, "matterport", True),
    # the benchmark's shape at a batch the CPU reference still finishes quickly (round 2)
    ("init_b8_384x384", 5, "init", 8, 384, 384, "matterport", True),
]

# ablation branches of the Essential Matrix Module: same generator, reference built with the flag(s)
ABLATION_CASES = [
    ("ablate_single_softmax_b2_64x80", 11, "stress", 2, 64, 80, "matterport", True, ("use_single_softmax",)),
    ("ablate_cross_features_b2_64x80", 12, "stress", 2, 64, 80, "varied", True, ("cross_features",)),
    ("ablate_l1_pos_b2_64x80", 13, "stress", 2, 64, 80, "matterport", True, ("l1_pos_encoding",)),
    ("ablate_all_three_b1_96x128", 14, "stress", 1, 96, 128, "matterport", True, ("use_single_softmax", "cross_features", "l1_pos_encoding")),
    # --noess: plain cross attention + pool_attn head instead of the module (model.py:71-88,183-187)
    ("ablate_noess_b2_64x80", 15, "stress", 2, 64, 80, "matterport", True, ("noess",)),
    ("ablate_noess_b3_96x128_init", 16, "init", 3, 96, 128, None, False, ("noess",)),
    # no --fusion_transformer: CNN front end + pool_transformer_output head only (model.py:62-69,179-181)
    ("ablate_cnn_only_b3_64x80", 17, "stress", 3, 64, 80, "matterport", True, ("cnn_only",)),
]

TOK_SAMPLE = (slice(None), slice(None, None, 9), slice(None, None, 4))


def sample(name, t):
    a = t.detach().cpu().numpy()
    if name in ("tokens",) or name.startswith("block") or (name == "cross" and a.shape[1] == 576):
        return np.ascontiguousarray(a[TOK_SAMPLE])
    if name == "preprocessed":
        return np.ascontiguousarray(a[:, :, ::7, ::5])
    if name == "features":
        return np.ascontiguousarray(a[:, ::3])
    return a


def run_case(name, seed, profile, B, H, W, ikind, integer, flags=()):
    cnn_only = "cnn_only" in flags
    model, SE3 = ref_loader.load_reference_model(**{("fusion_transformer" if f == "cnn_only" else f): f != "cnn_only" for f in flags})
    noess = "noess" in flags
    model.load_state_dict(S.make_state_dict(seed, profile, noess=noess, cnn_only=cnn_only))
    model.eval()
    images = torch.from_numpy(S.make_images_numpy(seed, B, H, W, integer))
    intr = None if ikind is None else torch.from_numpy(S.make_intrinsics_numpy(B, ikind, seed))
    Gs_data = torch.zeros(B, 2, 7)
    Gs_data[..., 6] = 1.0
    stages = {}
    vt = model.fusion_transformer
    hooks = []

    def put(key, fn):
        def h(m, i, o):
            stages[key] = fn(i, o)   # returns None: a non-None hook result would replace the output
        return h

    hooks.append(model.extractor_final_conv.register_forward_hook(
        put("tokens", lambda i, o: o.reshape(o.shape[0], 192, 576).permute(0, 2, 1))))
    hooks.append(model.resnet.conv1.register_forward_hook(put("preprocessed", lambda i, o: i[0])))
    for li in range(0 if cnn_only else 5):
        hooks.append(vt.blocks[li].register_forward_hook(put(f"block{li}", lambda i, o: o)))
    if not cnn_only:
        hooks.append(vt.blocks[5].register_forward_hook(put("cross", lambda i, o: o)))
    zs = []

    def _z_hook(m, i):
        zs.append(i[0])

    if not noess and not cnn_only:
        hooks.append(vt.blocks[5].cross_attn.proj_fundamental.register_forward_pre_hook(_z_hook))

    def _reg_hook(m, i, o):
        stages["features"] = i[0]
        stages["raw_pose"] = o

    hooks.append(model.pose_regressor.register_forward_hook(_reg_hook))
    with torch.no_grad():
        out = model(images, SE3(Gs_data), intrinsics=intr)
    for h in hooks:
        h.remove()
    # proj_fundamental is applied to z2 first, then z1 (vision_transformer.py:233-234);
    # z[b,c,h*70+a] = F[b,h,a,c]  ->  recover F[b,h,a,c]
    if not noess and not cnn_only:
        z2, z1 = zs
        unz = lambda z: z.reshape(B, 70, 3, 70).permute(0, 2, 3, 1)
        stages["bilinear1"] = unz(z1)
        stages["bilinear2"] = unz(z2)
    rec = {"poses": out[0].data.numpy()}
    if intr is not None:
        rec["intrinsics_after"] = intr.numpy()          # mutated in place by the reference
    for k, v in stages.items():
        rec["stage_" + k] = sample(k, v)
    rec["meta"] = np.array([seed, B, H, W, int(integer)], np.int64)
    rec["profile"] = np.array(profile)
    rec["intrinsics_kind"] = np.array("none" if ikind is None else ikind)
    rec["flags"] = np.array(",".join(flags))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **rec)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "pose1[0] =", rec["poses"][0, 1])


def posenc_golden():
    """The reference's own get_positional_encodings (host double loop) for a few intrinsics."""
    ref_loader.load_reference_model()  # makes `src` importable
    from src.modules.vision_transformer import get_positional_encodings
    rec = {}
    rec["none"] = get_positional_encodings(2, 576, None).numpy()
    for nm, k in (("matterport_24", [32.373125, 32.373125, 20.0, 15.0]),
                  ("square", [8.0, 8.0, 8.0, 8.0]), ("odd", [17.31, 23.9, 11.2, 13.7])):
        intr = torch.tensor([[k, k]], dtype=torch.float32)
        rec[nm] = get_positional_encodings(1, 576, intr).numpy()
        rec[nm + "_k"] = intr.numpy()
    path = os.path.join(ROOT, "tests", "golden", "posenc.npz")
    np.savez_compressed(path, **rec)
    print("posenc ->", path)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only in [c[0] for c in CASES]:              # a single default-configuration case by name
        run_case(*[c for c in CASES if c[0] == only][0])
        sys.exit(0)
    if only not in ("ablations", "noess", "cnn_only"):
        posenc_golden()
        for c in CASES:
            run_case(*c)
    for c in ABLATION_CASES:
        if only not in ("noess", "cnn_only") or only in c[-1]:
            run_case(*c)
