"""Config 5 driver: the training loop of the reference's train.py (lines 38-73, 140-165) on SYNTHETIC pairs.

    python -m rel_pose_b200.train_synthetic --steps 50 --batch 6                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m rel_pose_b200.train_synthetic ...

Everything the reference does per step is reproduced with the same library calls it makes: DistributedDataParallel
(NCCL gradient all-reduce, bucketed and overlapped with the backward), Adam(lr, weight_decay), OneCycleLR(div_factor 25),
clip_grad_norm_(2.5), loss = w_tr * tr + w_rot * rot from geodesic_loss on SE3 objects.  The datasets are not available
offline, so pairs are synthesised: view 2 is view 1 translated by (dx, dy) pixels and the target pose encodes that
shift, which makes the loss learnable.  Prints one JSON line (steps/s, pairs/s, loss curve head/tail).
"""
import argparse
import json
import os

import torch
import torch.distributed as dist

from . import SE3, synthetic as S, train_path
from .losses import geodesic_loss


def make_batch(step, rank, B, H, W, device):
    g = torch.Generator(device="cpu").manual_seed(1000003 * rank + step)
    base = torch.rand(B, 3, H + 16, W + 16, generator=g) * 255
    base = (base + base.roll(1, 2) + base.roll(1, 3) + base.roll((1, 1), (2, 3))) * 0.25      # smooth a little
    dx = torch.randint(-8, 9, (B,), generator=g); dy = torch.randint(-8, 9, (B,), generator=g)
    v1 = base[:, :, 8:8 + H, 8:8 + W]
    v2 = torch.stack([base[b, :, 8 + int(dy[b]):8 + int(dy[b]) + H, 8 + int(dx[b]):8 + int(dx[b]) + W] for b in range(B)])
    images = torch.stack([v1, v2], 1).floor().contiguous()
    poses = torch.zeros(B, 2, 7); poses[..., 6] = 1
    poses[:, 1, 0] = dx.float() / 8.0; poses[:, 1, 1] = dy.float() / 8.0                       # translation encodes the shift
    ang = 0.05 * dx.float() / 8.0                                                              # and a small yaw
    poses[:, 1, 5] = torch.sin(ang / 2); poses[:, 1, 6] = torch.cos(ang / 2)
    intr = torch.from_numpy(S.make_intrinsics_numpy(B))
    return images.to(device), poses.to(device), intr.to(device)


def default_options(**over):
    """The reference's Matterport training flags (scripts/train_matterport.sh:6-9, train.py:200-233)."""
    d = dict(steps=30, warmup_steps=5, batch=6, size=[384, 512], lr=5e-4, weight_decay=1e-5, clip=2.5, w_tr=10.0, w_rot=10.0,
             total_steps=120000, warmup=10000, optimizer="fused", pool=8, measure_allreduce=True, graph=True, overlap_exchange=True)
    d.update(over)
    return argparse.Namespace(**d)


def train_loop(a, dev, rank=0, world=1, local=0):
    """Runs a.warmup_steps + a.steps optimizer steps of the reference's loop (train.py:140-165) and returns the result dict
    (rank 0; None elsewhere).  The process group must already be initialised when world > 1.  Timing: CUDA events on
    the launching stream, barrier + synchronize on both sides, max over ranks."""
    from . import ViTEss
    torch.manual_seed(0)
    margs = argparse.Namespace(noess=False, pool_size=60, fc_hidden_size=512, fusion_transformer=True, transformer_depth=6,
                               cross_features=False, use_single_softmax=False, no_pos_encoding=False, l1_pos_encoding=False)
    model = ViTEss(margs)
    model.load_state_dict(S.make_state_dict(0, "init"))
    model.to(dev).train()
    for p in list(model.resnet.layer4.parameters()) + list(model.resnet.layer3.parameters()):
        p.requires_grad = False
    net = model
    use_graph = bool(getattr(a, "graph", False)) and a.optimizer == "fused"
    if world > 1 and not use_graph:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False)
    if a.optimizer == "fused":
        from .optim import FusedAdamOneCycle
        opt = FusedAdamOneCycle(list(net.parameters()), a.lr, a.total_steps, a.warmup / a.total_steps, div_factor=25,
                                weight_decay=a.weight_decay, clip=a.clip)
        sched = None
    else:
        opt = torch.optim.Adam(net.parameters(), lr=a.lr, weight_decay=a.weight_decay)
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, a.lr, a.total_steps, pct_start=a.warmup / a.total_steps, div_factor=25,
                                                    cycle_momentum=False)
    H, W = a.size
    # the data loader is outside the path: batches are synthesised up front on the device and cycled
    pool = [make_batch(i, rank, a.batch, H, W, dev) for i in range(max(1, a.pool))]

    def one_step(step, evs=None, sync_grads=True):
        images, poses, intr = pool[step % len(pool)]
        intr = intr.clone()                                   # forward rescales the intrinsics in place
        opt.zero_grad()
        if evs:
            evs[0].record()
        Ps = SE3(poses)
        Gs = SE3.IdentityLike(Ps)
        poses_est = net(images, Gs, intrinsics=intr)
        ltr, lrot, _ = geodesic_loss(SE3(Ps.data.clone()), poses_est, sync_metrics=False) if a.optimizer == "fused" \
            else geodesic_loss(SE3(Ps.data.clone()), poses_est)
        loss = a.w_tr * ltr + a.w_rot * lrot
        if evs:
            evs[1].record()
        if sync_grads or world == 1 or net is model:
            loss.backward()
            if world > 1 and net is model and sync_grads:     # graph path that fell back to eager: same flat all-reduce
                dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)
        else:
            with net.no_sync():                               # same backward without the gradient exchange
                loss.backward()
        if evs:
            evs[2].record()
        if a.optimizer == "fused":
            gn = opt.step()
        else:
            gn = torch.nn.utils.clip_grad_norm_(net.parameters(), a.clip)
            opt.step()
            sched.step()
        if evs:
            evs[3].record()
        return loss.detach(), gn.reshape(())

    # ---- the step as CUDA graphs: a step is ~700 kernel launches of 5-30 us each, so issued one by one from Python it is
    # bound by the host (17 ms per step for 11 ms of kernels).  Two graphs are recorded once after a few eager steps and
    # replayed: G1 = zero_grad + forward + loss + backward, G2 = gradient norm + clipped Adam.  Between them, on more than
    # one rank, ONE NCCL all-reduce (average) of the flat gradient buffer -- every parameter's .grad is a view of it.
    # It is the whole exchange step of train.py:66-67 in one call: 77 MB take 0.3 ms over NVLink = 5 % of the 5.9 ms step at
    # 8 GPUs.  --overlap_exchange 1 (default) hides most of it: the gradients come in two buckets inside G1 -- everything
    # behind the CNN (transformer, Essential Matrix Module, regressor: 92 % of the bytes) is all-reduced on a side stream
    # from the moment the backward pass reaches the CNN (a hook on the token tensor's gradient, train_path.forward_train),
    # the CNN's own gradients when the backward ends.  Same sums, same order of the two-operand adds inside NCCL per
    # element: the losses are those of the single-call exchange.  (DistributedDataParallel's own reducer cannot be
    # captured: the attempt invalidated the capture, profiles/r02_train_2gpu_ddp_capture_attempt.json.)
    # The batch is copied into static buffers and the two per-step optimizer scalars into a 2-float device buffer in
    # front of each replay.
    graph = None
    graph_note = "off"
    if use_graph:
        opt.enable_device_scalars()
        trainable = [p for p in model.parameters() if p.requires_grad]
        offs, o = [], 0
        for p in trainable:
            offs.append(o); o += (p.numel() + 3) // 4 * 4
        flat_grad = torch.zeros(o, dtype=torch.float32, device=dev)
        flat_views = [flat_grad[off:off + p.numel()].view_as(p) for p, off in zip(trainable, offs)]
        for p, v in zip(trainable, flat_views):
            p.grad = v
        static = [t.clone() for t in pool[0]]
        g_out = {}
        # first parameter behind the CNN (model.parameters() follows the module order: resnet, extractor_final_conv, then the rest)
        cnn_ids = {id(p) for m in (model.resnet, model.extractor_final_conv) for p in m.parameters()}
        k0 = next((i for i, p in enumerate(trainable) if id(p) not in cnn_ids), len(trainable))
        tail_ok = world > 1 and a.overlap_exchange and 0 < k0 < len(trainable) and all(id(p) not in cnn_ids for p in trainable[k0:])
        ov = {"on": bool(tail_ok), "fired": False}
        xstream = torch.cuda.Stream(device=dev) if tail_ok else None

        def tokens_grad_hook(_g):
            # backward has reached the CNN: pack and all-reduce the gradients behind it while the CNN's backward runs
            tail = trainable[k0:]
            if not ov["on"] or ov["fired"] or any(p.grad is None for p in tail):
                return None
            ov["fired"] = True
            with torch.no_grad():
                torch._foreach_copy_(flat_views[k0:], [p.grad for p in tail])
            cur = torch.cuda.current_stream(dev)
            xstream.wait_stream(cur)
            with torch.cuda.stream(xstream):
                dist.all_reduce(flat_grad[offs[k0]:], op=dist.ReduceOp.AVG)
            return None

        def fwd_bwd():
            images, poses, intr = static
            intr_w = intr.clone()
            # no zero-fill + 123 accumulate kernels: autograd ASSIGNS fresh gradient tensors (p.grad is None), one
            # multi-tensor copy packs them into the flat buffer the exchange step and the optimizer read
            for p in trainable:
                p.grad = None
            Ps = SE3(poses)
            ov["fired"] = False
            train_path.TOKENS_GRAD_HOOK = tokens_grad_hook if ov["on"] else None      # attached to the token tensor by the forward
            try:
                poses_est = net(images, SE3.IdentityLike(Ps), intrinsics=intr_w)
            finally:
                train_path.TOKENS_GRAD_HOOK = None
            ltr, lrot, _ = geodesic_loss(SE3(Ps.data.clone()), poses_est, sync_metrics=False)
            loss = a.w_tr * ltr + a.w_rot * lrot
            loss.backward()
            lo = k0 if ov["fired"] else len(trainable)            # parameters not packed by the hook
            got = [(v, p.grad) for p, v in zip(trainable[:lo], flat_views[:lo]) if p.grad is not None]
            with torch.no_grad():
                torch._foreach_copy_([v for v, _ in got], [g for _, g in got])
            for p, v in zip(trainable, flat_views):
                p.grad = v
            if ov["on"]:
                # second bucket (or, if the hook could not fire, everything) on the main stream, then join the side stream
                dist.all_reduce(flat_grad[:offs[k0]] if ov["fired"] else flat_grad, op=dist.ReduceOp.AVG)
                torch.cuda.current_stream(dev).wait_stream(xstream)
            g_out["loss"] = loss.detach()

        def exchange_grads():
            if world > 1 and not ov["on"]:
                dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)

        def opt_part():
            g_out["gn"] = opt.step(_captured=True).reshape(())

        # the first n_pre steps of the run are issued eagerly on a side stream (the allocator warm-up a capture needs),
        # the graphs are recorded after them and replayed for every later step: same batches, same number of
        # optimizer steps as the eager loop
        n_pre = max(1, min(a.warmup_steps, 2))
        pre_log = []
        try:
            if ov["on"]:                                     # NCCL communicator of the side stream created outside the capture
                xstream.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(xstream):
                    dist.all_reduce(torch.zeros(8, device=dev))
                torch.cuda.current_stream(dev).wait_stream(xstream)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for i in range(n_pre):
                    for dst, src in zip(static, pool[i % len(pool)]):
                        dst.copy_(src)
                    opt.upload_step_scalars()
                    fwd_bwd(); exchange_grads(); opt_part()
                    opt.advance()
                    pre_log.append((g_out["loss"].clone(), g_out["gn"].clone()))
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()

            def capture():
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                opt.upload_step_scalars()
                with torch.cuda.graph(g1):
                    fwd_bwd()
                with torch.cuda.graph(g2, pool=g1.pool()):
                    opt_part()
                return g1, g2

            try:
                graph, graph2 = capture()
            except Exception:
                if not ov["on"]:
                    raise
                ov["on"] = False                             # the exchange would not capture: one all-reduce between the graphs
                torch.cuda.synchronize()
                graph, graph2 = capture()
            assert all(p.grad.data_ptr() == flat_grad.data_ptr() + 4 * off for p, off in zip(trainable, offs)), \
                "a gradient left the flat buffer"
            if ov["on"]:
                graph_note = ("G1 = forward + loss + backward WITH the gradient exchange inside: NCCL all-reduce (AVG) of the "
                              f"gradients behind the CNN ({(flat_grad.numel() - offs[k0]) * 4} bytes) on a side stream under the "
                              f"CNN's backward, of the CNN's ({offs[k0] * 4} bytes) at the end; G2 = clip + Adam: CUDA graphs "
                              f"(first {n_pre} steps eager)" + ("" if ov["fired"] else "; hook did not fire: single all-reduce"))
            else:
                graph_note = ("G1 = forward + loss + backward, " + ("one NCCL all-reduce (AVG) of the flat gradient buffer, "
                              if world > 1 else "") + f"G2 = clip + Adam: CUDA graphs (first {n_pre} steps eager)")
        except Exception as ex:
            graph = None
            graph_note = f"capture failed, eager: {type(ex).__name__}: {str(ex)[:160]}"
            torch.cuda.synchronize()

    def graphed_step(step, evs=None):
        for dst, src in zip(static, pool[step % len(pool)]):
            dst.copy_(src, non_blocking=True)
        opt.upload_step_scalars()
        if evs:
            evs[0].record()
        graph.replay()
        if evs:
            evs[1].record()
        exchange_grads()
        if evs:
            evs[2].record()
        graph2.replay()
        if evs:
            evs[3].record()
        opt.advance()
        return g_out["loss"], g_out["gn"]

    nsteps = a.warmup_steps + a.steps
    loss_log = torch.zeros(nsteps, 2, device=dev)            # read back once at the end: no per-step host sync
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(nsteps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first = 0
    if graph is not None:                                     # steps already taken while warming up for the capture
        first = n_pre
        for i, (l_, g_) in enumerate(pre_log):
            loss_log[i, 0].copy_(l_); loss_log[i, 1].copy_(g_)
    for step in range(first, nsteps):
        if step == max(a.warmup_steps, first):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
        if graph is not None:
            l_, g_ = graphed_step(step, ev[step])
            loss_log[step, 0].copy_(l_); loss_log[step, 1].copy_(g_)
        else:
            loss_log[step, 0], loss_log[step, 1] = one_step(step, ev[step])
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    losses = [(float(x), float(y)) for x, y in loss_log.cpu().tolist()]
    timed = range(max(a.warmup_steps, first), nsteps)
    phase = [sum(ev[i][k].elapsed_time(ev[i][k + 1]) for i in timed) / len(timed) for k in range(3)]

    # ---- the exchange step by itself (world > 1): backward with and without the gradient all-reduce, and the same
    # payload (one flat float32 buffer of all trainable gradients) all-reduced alone on an idle GPU
    exchange = None
    if world > 1 and getattr(a, "measure_allreduce", True) and graph is None and net is not model:
        n_extra = max(3, min(8, a.steps))
        evn = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_extra)]
        for i in range(n_extra):
            one_step(nsteps + i, evn[i], sync_grads=False)
        numel = sum(p.numel() for p in model.parameters() if p.requires_grad)
        flat = torch.zeros(numel, dtype=torch.float32, device=dev)
        for _ in range(2):
            dist.all_reduce(flat)
        dist.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(flat)
        a1.record()
        torch.cuda.synchronize()
        bw_nosync = sum(evn[i][1].elapsed_time(evn[i][2]) for i in range(1, n_extra)) / (n_extra - 1)
        ar_ms = a0.elapsed_time(a1) / 5
        tt = torch.tensor([bw_nosync, ar_ms, phase[1]], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        bw_nosync, ar_ms, bw_sync = (float(x) for x in tt.tolist())
        exposed = max(0.0, bw_sync - bw_nosync)
        exchange = {"payload_bytes": numel * 4, "allreduce_alone_ms": round(ar_ms, 3),
                    "bus_bandwidth_GBps": round(2 * (world - 1) / world * numel * 4 / (ar_ms * 1e-3) / 1e9, 1),
                    "backward_ms_with_allreduce": round(bw_sync, 3), "backward_ms_no_sync": round(bw_nosync, 3),
                    "exposed_allreduce_ms": round(exposed, 3),
                    "overlap_fraction": round(1.0 - min(1.0, exposed / ar_ms), 3) if ar_ms > 0 else None,
                    "how": "DistributedDataParallel buckets (25 MB) reduced by NCCL while the backward kernels still run; "
                           "exposed = backward with the exchange minus the same backward under no_sync(), max over ranks"}
    if rank != 0:
        return None
    ls = [l for l, _ in losses]
    return {"metric": "training steps/sec (config 5: train.py loop on synthetic pairs; fp32-class split-bf16 products on tcgen05 in forward and backward, flash-style attention / Essential-Matrix-Module gradients, implicit-GEMM weight gradients)",
            "n_gpus": world, "steps": a.steps, "warmup_steps": a.warmup_steps, "pairs_per_gpu": a.batch, "image_size": [H, W],
            "ms_per_step": ms / len(timed), "steps_per_s": len(timed) / (ms * 1e-3),
            "pairs_per_s": world * a.batch * len(timed) / (ms * 1e-3),
            "loss_first5": [round(x, 4) for x in ls[:5]], "loss_last5": [round(x, 4) for x in ls[-5:]],
            "grad_norm_first": round(losses[0][1], 3), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 2),
            "ddp": world > 1 and graph is None, "optimizer": a.optimizer, "cuda_graph": graph_note,
            "phase_ms": ({"forward+loss+backward (G1)": round(phase[0], 3), "gradient all-reduce": round(phase[1], 3),
                          "clip+adam+lr (G2)": round(phase[2], 3)} if graph is not None else
                         {"forward+loss": round(phase[0], 3), "backward(+allreduce)": round(phase[1], 3),
                          "clip+adam+lr": round(phase[2], 3)}),
            "exchange": exchange if exchange is not None else (
                "single rank: no gradient exchange" if world == 1 else
                {"payload_bytes": int(flat_grad.numel()) * 4, "how": (
                    "two NCCL all-reduces (average) inside G1: the gradients behind the CNN on a side stream while the CNN's backward "
                    "runs, the CNN's at the end; phase_ms['gradient all-reduce'] is then ~0 and the exposed part is inside G1"
                    if (graph is not None and ov["on"]) else
                    "one NCCL all-reduce (average) of the flat gradient buffer between the two graphs; its time is "
                    "phase_ms['gradient all-reduce'] (CUDA events, max over ranks not taken)")})}


def main():
    ap = argparse.ArgumentParser()
    d = default_options()
    ap.add_argument("--steps", type=int, default=d.steps)
    ap.add_argument("--warmup_steps", type=int, default=d.warmup_steps, help="untimed steps before the timed region")
    ap.add_argument("--batch", type=int, default=d.batch)                # scripts/train_matterport.sh: --batch=6 per GPU
    ap.add_argument("--size", type=int, nargs=2, default=d.size)         # the Matterport pipeline's image size
    ap.add_argument("--lr", type=float, default=d.lr)
    ap.add_argument("--weight_decay", type=float, default=d.weight_decay)
    ap.add_argument("--clip", type=float, default=d.clip)
    ap.add_argument("--w_tr", type=float, default=d.w_tr)
    ap.add_argument("--w_rot", type=float, default=d.w_rot)
    ap.add_argument("--total_steps", type=int, default=d.total_steps)
    ap.add_argument("--warmup", type=int, default=d.warmup)
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: rel_pose_b200.optim.FusedAdamOneCycle (clip + Adam + OneCycle in 3 launches, no host sync); "
                         "torch: the reference's own calls (clip_grad_norm_, Adam.step, OneCycleLR.step)")
    ap.add_argument("--pool", type=int, default=d.pool, help="synthetic batches generated up front on the device and cycled")
    ap.add_argument("--overlap_exchange", type=int, default=1,
                    help="graphs + several ranks: all-reduce the gradients behind the CNN under the CNN's backward (0: one call between the graphs)")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the whole step as one CUDA graph (fused optimizer only); 0: eager")
    a = ap.parse_args()
    a.measure_allreduce = True
    a.graph = bool(a.graph)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = train_loop(a, dev, rank, world, local)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
