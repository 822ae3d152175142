"""A caller written against the REFERENCE's public surface only (the call sequence of its demo script:
argparse Namespace -> ViTEss(args) -> load_state_dict of a `module.`-prefixed checkpoint -> .cuda().eval() ->
model(images, SE3(poses), intrinsics=...) -> poses_est[0][0][1].data).  It imports `src.model` and `lietorch`
by those names; tests run it through `python -m rel_pose_b200.run` to prove the drop-in boundary."""
import argparse
from collections import OrderedDict

import cv2
import numpy as np
import torch
import torch.nn.functional as F
import lietorch                      # noqa: F401  (the reference imports the module too)
from lietorch import SE3
from src.model import ViTEss

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--img1"); ap.add_argument("--img2"); ap.add_argument("--ckpt")
    for flag in ("no_pos_encoding", "noess", "cross_features", "use_single_softmax", "l1_pos_encoding"):
        ap.add_argument("--" + flag, action="store_true")
    ap.add_argument("--fc_hidden_size", type=int, default=512)
    ap.add_argument("--pool_size", type=int, default=60)
    ap.add_argument("--transformer_depth", type=int, default=6)
    args = ap.parse_args()
    args.fusion_transformer = True
    k = [517.97, 517.97, 320, 240] if "matterport" in args.ckpt else [128, 128, 128, 128]
    intrinsics = torch.from_numpy(np.array([[k, k]], np.float32)).cuda()
    model = ViTEss(args)
    sd = OrderedDict((n.replace("module.", ""), v) for n, v in torch.load(args.ckpt)["model"].items())
    model.load_state_dict(sd)
    model = model.cuda().eval()
    images = torch.from_numpy(np.stack([cv2.imread(args.img1), cv2.imread(args.img2)]).astype(np.float32)).permute(0, 3, 1, 2)
    if "matterport" in args.ckpt:
        images = F.interpolate(images, size=[384, 512])
    images = images.unsqueeze(0).cuda()
    poses = torch.from_numpy(np.tile(np.array([0, 0, 0, 0, 0, 0, 1], np.float32), (1, 2, 1))).cuda()
    with torch.no_grad():
        poses_est = model(images, SE3(poses), intrinsics=intrinsics)
    preds = poses_est[0][0][1].data.cpu().numpy()
    print("POSE " + " ".join(f"{v:.9e}" for v in preds))
    print("INTRINSICS " + " ".join(f"{v:.9e}" for v in intrinsics.cpu().numpy().ravel()))
