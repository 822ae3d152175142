"""Throughput of the model variants of SURVEY.md 8 f-4 on one B200: 64 synthetic 384x384 pairs per step, images
resident on the device, CUDA events around the timed steps (same protocol as bench.py's `value`).  One JSON line.

    python tools/bench_ablations.py [--steps 10] [--warmup 3] [--batch 64]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rel_pose_b200 import ViTEss, ops, synthetic as S  # noqa: E402
from rel_pose_b200.lietorch import SE3  # noqa: E402
from bench import ClockSampler  # noqa: E402  (nvidia-smi clocks / throttle reasons sampled while the GPU is loaded)

VARIANTS = [
    ("default", {}),
    ("l1_pos_encoding", {"l1_pos_encoding": True}),
    ("use_single_softmax", {"use_single_softmax": True}),
    ("cross_features", {"cross_features": True}),
    ("noess", {"noess": True}),
    ("cnn_only (no --fusion_transformer)", {"fusion_transformer": False}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=384)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B = a.batch
    g = torch.Generator(device=dev).manual_seed(1234)
    images = (torch.rand(B, 2, 3, a.size, a.size, generator=g, device=dev) * 255).floor()
    intr0 = torch.from_numpy(S.make_intrinsics_numpy(B)).to(dev)
    Gs = SE3.Identity(B, 2, device=dev)
    out = {"workload": f"{B} synthetic {a.size}x{a.size} pairs per step, device-resident float32 images, "
                       f"{a.steps} timed steps after {a.warmup} warm-up steps, CUDA events", "variants": {}}
    sampler = ClockSampler(0)
    sampler.start()
    for name, over in VARIANTS:
        margs = argparse.Namespace(noess=False, pool_size=60, fc_hidden_size=512, fusion_transformer=True, transformer_depth=6,
                                   cross_features=False, use_single_softmax=False, no_pos_encoding=False, l1_pos_encoding=False)
        for k, v in over.items():
            setattr(margs, k, v)
        model = ViTEss(margs)
        model.load_state_dict(S.make_state_dict(0, "init", noess=bool(over.get("noess")),
                                                cnn_only=(over.get("fusion_transformer") is False)))
        model = model.to(dev).eval()
        model.precision = "bf16x3"
        with torch.no_grad():
            for _ in range(max(1, a.warmup)):
                model(images, Gs, intrinsics=intr0.clone())
            torch.cuda.synchronize()
            l0 = ops.launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                model(images, Gs, intrinsics=intr0.clone())
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out["variants"][name] = {"ms_per_step": round(ms, 3), "pairs_per_s": round(B / ms * 1e3, 1),
                                 "launches_per_step": (ops.launches() - l0) // a.steps,
                                 "engine": "fp32 SIMT" if model._tc_planes() == 0 else
                                           ("bf16x3 tcgen05" + (" (module flags on rp_essential_ex_tc)" if model.em_flags else ""))}
        out["variants"][name]["clocks"] = sampler.mark()
        del model
        torch.cuda.empty_cache()
    out["clocks"] = sampler.stop()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
