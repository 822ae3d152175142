"""Multi-GPU training exchange step (BASELINE.json config 5): DistributedDataParallel over NCCL must hand every rank
the mean of the per-rank gradients of the CUDA backward.  Needs >= 2 GPUs (skipped on the single-GPU box)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r"""
import argparse, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RP_ROOT"])
from rel_pose_b200 import ViTEss, SE3, synthetic as S
from rel_pose_b200.losses import geodesic_loss
from rel_pose_b200.train_synthetic import make_batch
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
margs = argparse.Namespace(noess=False, pool_size=60, fc_hidden_size=512, fusion_transformer=True, transformer_depth=6,
                           cross_features=False, use_single_softmax=False, no_pos_encoding=False, l1_pos_encoding=False)
def build():
    m = ViTEss(margs); m.load_state_dict(S.make_state_dict(3, "init")); m.to(dev).train()
    for p in list(m.resnet.layer3.parameters()) + list(m.resnet.layer4.parameters()): p.requires_grad = False
    return m
def grads(net, r):
    images, poses, intr = make_batch(0, r, 2, 64, 80, dev)
    net.zero_grad()
    Ps = SE3(poses)
    out = net(images, SE3.IdentityLike(Ps), intrinsics=intr)
    ltr, lrot, _ = geodesic_loss(Ps, out)
    (10 * ltr + 10 * lrot).backward()
    mod = net.module if hasattr(net, "module") else net
    return {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}
ddp = torch.nn.parallel.DistributedDataParallel(build(), device_ids=[local], find_unused_parameters=False)
g_ddp = grads(ddp, rank)
ref = None
for r in range(world):                      # every rank recomputes all shards locally, without DDP
    g = grads(build(), r)
    ref = g if ref is None else {k: ref[k] + g[k] for k in g}
worst = max(float((g_ddp[k] - ref[k] / world).abs().max() / (ref[k].abs().max() / world + 1e-12)) for k in ref)
print(f"rank {rank}: {len(ref)} gradients, worst relative deviation from the mean of per-rank gradients {worst:.3e}", flush=True)
assert len(g_ddp) == len(ref) == 123 and worst < 1e-4
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_allreduce_gives_mean_of_rank_gradients(tmp_path):
    script = tmp_path / "ddp_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RP_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=900)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.count("worst relative deviation") == 2
