"""Where does the bf16x3 engine lose accuracy on the CNN-only golden?  Runs the front end and the head/regressor in
each engine separately (4 combinations) and prints the pose error against the reference golden."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import relpose_oracle as O
from rel_pose_b200 import ViTEss, ops, synthetic as S

g = np.load(os.path.join(ROOT, "tests/golden/ablate_cnn_only_b3_64x80.npz"))
seed, B, H, W, integer = (int(v) for v in g["meta"])
a = argparse.Namespace(noess=False, pool_size=60, fc_hidden_size=512, fusion_transformer=False, transformer_depth=6,
                       cross_features=False, use_single_softmax=False, no_pos_encoding=False, l1_pos_encoding=False)
m = ViTEss(a); m.load_state_dict(S.make_state_dict(seed, "stress", cnn_only=True)); m = m.cuda().eval()
images = torch.from_numpy(S.make_images_numpy(seed, B, H, W, bool(integer))).cuda()
# float64 truth from the oracle
p64 = {k: v.astype(np.float64) if v.dtype == np.float32 else v for k, v in S.make_state_dict_numpy(seed, "stress", cnn_only=True).items()}
Gs = np.zeros((B, 2, 7)); Gs[..., 6] = 1
st64 = {}
truth, _ = O.vitess_forward(images.cpu().numpy(), Gs, None, p64, np.float64, stages=st64, flags=("cnn_only",))

def err(tag, poses):
    for nm, ref in (("golden", g["poses"]), ("f64", truth)):
        rot = O.rotation_error_rad(poses[:, 1, 3:], ref[:, 1, 3:]); tr = O.translation_rel_error(poses[:, 1, :3], ref[:, 1, :3])
        print(f"{tag:46s} vs {nm:6s}: rot {rot.max():.3e} rad  trans {tr.max():.3e}")

rot = O.rotation_error_rad(g["poses"][:, 1, 3:], truth[:, 1, 3:])
print("golden (fp32 reference) vs f64 truth: rot", rot.max())
toks = {}
with torch.no_grad():
    for prec in ("fp32", "bf16x3"):
        m.precision = prec
        x = ops.preprocess_nhwc4(images.float()) if prec == "fp32" else ops.preprocess_stem_windows(images.float(), 2)
        toks[prec] = m._cnn_front_end(x)
        t = toks[prec].cpu().numpy().astype(np.float64)
        e = np.abs(t - st64["tokens"]); print(f"tokens[{prec}] vs f64: max abs {e.max():.3e} rms {np.sqrt((e**2).mean()):.3e} ref rms {np.sqrt((st64['tokens']**2).mean()):.3e}")
    for fe in ("fp32", "bf16x3"):
        for he in ("fp32", "bf16x3"):
            for tcreg in ((True, False) if he == "bf16x3" else (False,)):
                m.precision = he; m.tc_regressor = tcreg
                feat, w0 = m._pool_attn_head(toks[fe], B)
                reg = m.pose_regressor
                if he == "bf16x3" and tcreg:
                    h = ops.linear_tc_splitk(ops.split_planes(feat.contiguous(), 2), m._planes(w0, 2), reg[0].bias, act=ops.ACT_RELU)
                else:
                    h = ops.linear(feat, w0, reg[0].bias, act=ops.ACT_RELU)
                raw = ops.regressor_tail(h, m._transposed(reg[2].weight), reg[2].bias, reg[4].weight.detach().contiguous(), reg[4].bias).reshape(B, 2, 7)
                f = feat.reshape(B, 576, 60).permute(0, 2, 1).reshape(B, -1).cpu().numpy().astype(np.float64)
                e = np.abs(f - st64["features"])
                poses = O.normalize_preds(Gs, raw.cpu().numpy().astype(np.float64))
                print(f"front={fe} head={he} tc_regressor={tcreg}: features max abs err {e.max():.3e} (ref rms {np.sqrt((st64['features']**2).mean()):.3e})  raw q {raw[0,1,3:].cpu().numpy()}")
                err(f"  front={fe} head={he} tcreg={tcreg}", poses)
