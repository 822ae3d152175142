#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the rel_pose hot path (ViTEss.forward) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one forward pass of the hot path over one batch of synthetic 384x384 image pairs per GPU
(BASELINE.json configs[1]: batch=64 pairs, fp32 arithmetic, random-init weights).  Pairs are
independent, so N GPUs run N shards with no data-path collective (weak scaling).  Rank 0 prints ONE
JSON line.  `--impl reference` times the reference's CPU implementation of the same path on the
host cores (the reference tree when present, else its pinned PyTorch-CPU port in oracle/torch_port.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-pairs/sec @384x384"
UNIT = "pairs/s"
FLOP_PER_PAIR = 16.79e9          # BASELINE.md section 2 (torch.utils.flop_counter on the reference)


def model_args():
    return argparse.Namespace(noess=False, pool_size=60, fc_hidden_size=512, fusion_transformer=True,
                              transformer_depth=6, cross_features=False, use_single_softmax=False,
                              no_pos_encoding=False, l1_pos_encoding=False)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_forward_factory(batch, size):
    """Returns (callable running one CPU forward over `batch` pairs, kind, description)."""
    import numpy as np
    import torch
    from rel_pose_b200 import synthetic as S
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    images = (torch.rand(batch, 2, 3, size, size, generator=g) * 255).floor()
    intr = torch.from_numpy(S.make_intrinsics_numpy(batch))
    Gs = torch.zeros(batch, 2, 7); Gs[..., 6] = 1
    sd = S.make_state_dict(0, "init")
    ref_root = os.environ.get("RELPOSE_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(ref_root, "src")):
        import ref_loader
        model, SE3 = ref_loader.load_reference_model()
        model.load_state_dict(sd)
        model.eval()

        def run():
            with torch.no_grad():
                return model(images, SE3(Gs), intrinsics=intr.clone())[0].data
        return run, "reference", f"unmodified reference ViTEss from {ref_root}, PyTorch CPU fp32"
    import torch_port
    p = {k: v for k, v in sd.items() if v.dtype != torch.int64}

    def run():
        with torch.no_grad():
            return torch_port.forward(images, Gs, intr, p)
    return run, "port", "oracle/torch_port.py (PyTorch-CPU port pinned to the reference's golden vectors), fp32"


def time_cpu(run, batch, min_seconds, max_iters):
    run()                                   # warm-up (oneDNN primitive creation, page faults)
    t0 = time.perf_counter(); n = 0
    while True:
        run(); n += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or n >= max_iters:
            break
    return batch * n / dt, n, dt


_JSON_FD = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." to stdout
    when NCCL_DEBUG=VERSION is set in the environment): everything written to fd 1 during the run goes to stderr, the
    JSON line goes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = a.ref_batch
    run, kind, desc = cpu_reference_forward_factory(batch, a.size)
    for _ in range(max(1, a.warmup if a.warmup < 2 else 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    dt = time.perf_counter() - t0
    val = batch * a.steps / dt
    cores = os.cpu_count() or 1
    sample = f"{a.steps} steps x {batch} pairs of {a.size}x{a.size} on {cores} host threads; {desc}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batch={batch} synthetic {a.size}x{a.size} pair inference on host CPU "
                                   "(bounded sample of configs[1])", "precision": "fp32"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="pairs per GPU per step (configs[1]: 64)")
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--micro-batch", type=int, default=0,
                    help="process the per-GPU batch in chunks of this many pairs (config 4: --batch 4096/G --micro-batch 256)")
    ap.add_argument("--u8", action="store_true", help="device-resident images as uint8 (config 4 at 4096 pairs: 3.6 GB instead of 14.5 GB)")
    ap.add_argument("--ref-batch", type=int, default=8, help="pairs per step of the CPU reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"],
                    help="arithmetic of the transformer GEMMs (fp32 SIMT | tcgen05 split-bf16 | tcgen05 bf16)")
    a = ap.parse_args()
    guard_stdout()
    if a.impl == "reference":
        return run_reference_arm(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    from rel_pose_b200 import ViTEss, SE3, ops, synthetic as S, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.lib().rp_device_arch(local) >= 100
    W = max(3, a.warmup)

    model = ViTEss(model_args())
    model.load_state_dict(S.make_state_dict(0, "init"))
    model = model.to(dev).eval()
    model.precision = a.precision
    B, size = a.batch, a.size
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    MB = a.micro_batch if 0 < a.micro_batch < B else B
    if a.u8:
        images = torch.empty((B, 2, 3, size, size), dtype=torch.uint8, device=dev)
        for lo in range(0, B, MB):                              # generated on the device, chunk by chunk
            images[lo:lo + MB] = (torch.rand(min(MB, B - lo), 2, 3, size, size, generator=g, device=dev) * 255).to(torch.uint8)
    else:
        images = (torch.rand(B, 2, 3, size, size, generator=g, device=dev) * 255).floor()     # 226 MB at B=64
    intr0 = torch.from_numpy(S.make_intrinsics_numpy(B)).to(dev)
    Gs = SE3.Identity(B, 2, device=dev)

    def step():
        if MB == B:
            return model(images, Gs, intrinsics=intr0.clone())      # fresh clone: forward rescales in place
        outs = [model(images[lo:lo + MB], Gs[lo:lo + MB], intrinsics=intr0[lo:lo + MB].clone())[0].data for lo in range(0, B, MB)]
        return [SE3(torch.cat(outs, 0))]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # clocks / throttle reasons are sampled (nvidia-smi every 100 ms) from the warm-up to the end of the end-to-end
        # legs: the 10-step timed region alone is ~50 ms, every sample of the window is taken under the same load
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        for _ in range(W):
            step()
        # ---- device-resident throughput ("value") ----
        barrier()
        l0 = ops.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            out = step()
        e1.record()
        barrier()
        launches = ops.launches() - l0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

        # ---- per-kernel timing with CUDA events on the launching stream (roofline) ----
        timer = ops.StageTimer()
        ops.set_timer(timer)
        for _ in range(min(3, a.steps)):
            step()
        torch.cuda.synchronize()
        ops.set_timer(None)
        stages = timer.summary()

        # ---- end to end through the public API: pinned host -> H2D, forward, D2H of the poses ----
        # rel_pose_b200.parallel.StreamedInference is the user-facing call for batched inference from host
        # memory: the H2D copy of step k+1 runs on a side stream while step k computes.  Every step copies its
        # own inputs from pinned host memory and reads its result back; all of it is inside the timed region.
        from rel_pose_b200.parallel import StreamedInference
        host_images = torch.empty((B, 2, 3, size, size), dtype=torch.float32, pin_memory=True)
        host_images.copy_(images)
        host_u8 = torch.empty((B, 2, 3, size, size), dtype=torch.uint8, pin_memory=True)
        host_u8.copy_(images.to(torch.uint8))
        host_intr = torch.from_numpy(S.make_intrinsics_numpy(B)).pin_memory()
        host_Gs = SE3.Identity(B, 2).data.pin_memory()
        runner = StreamedInference(model, dev)

        def e2e_run(src, n):
            last = None
            for last in runner.run((src, host_Gs, host_intr) for _ in range(n)):
                pass
            return last

        def e2e_time(src):
            if a.no_e2e:
                return 1e-9, 1e-9
            e2e_run(src, 2)
            barrier()
            t0 = time.perf_counter()
            e0.record()
            e2e_run(src, a.steps)
            e1.record()
            barrier()
            wall_ = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([max(e0.elapsed_time(e1), 1e-9), wall_], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0].item()), float(t[1].item())

        ms_e2e, wall = e2e_time(host_images)
        img_h2d = runner.last_image_h2d_bytes or host_images.numel() * 4
        ms_e2e_u8, wall_u8 = e2e_time(host_u8)
        img_h2d_u8 = runner.last_image_h2d_bytes or host_u8.numel()
        # host wall clock is the honest end-to-end figure (it includes the final D2H wait); events agree within noise
        ms_e2e, ms_e2e_u8 = max(ms_e2e, wall), max(ms_e2e_u8, wall_u8)
        clocks = sampler.stop() if sampler else None
        if clocks is not None:
            clocks["window"] = "warm-up + timed steps + per-kernel timing + end-to-end legs (GPU under the bench load throughout)"
        # bytes that actually cross PCIe per step: StreamedInference copies only the 224 of `size` image rows the
        # nearest resize reads (rp_copy_rows_h2d), plus intrinsics and Gs
        h2d = img_h2d + host_intr.numel() * 4 + host_Gs.numel() * 4
        h2d_u8 = img_h2d_u8 + host_intr.numel() * 4 + host_Gs.numel() * 4
        d2h = B * 2 * 7 * 4

    if rank == 0:
        pk = peaks()
        total_pairs = world * B * a.steps
        value = total_pairs / (ms * 1e-3)
        # dominant kernel of the step by measured time
        tot_ms = sum(d["ms"] for d in stages.values()) or 1.0
        dom = max(stages, key=lambda k: stages[k]["ms"])
        d = stages[dom]
        tfl = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        traffic = None
        try:          # per-launch DRAM bytes of the same kernel from the committed ncu capture (same batch / precision only)
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("batch") == B and tj.get("precision") == a.precision:
                traffic = tj["traffic_bytes"].get(dom)
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "tensor", "achieved": tfl, "peak": pk["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": tfl / pk["bf16_tflops_sustained"], "traffic": traffic,
                    "peak_source": pk["source"] + " (sustained bf16 dense; kernel timed inside a long step)",
                    "avg_launch_ms": d["ms"] / d["calls"], "share_of_step": d["ms"] / tot_ms,
                    "note": ("algorithmic FLOPs (2MNK, one product per MAC) over the CUDA-event duration of the kernel with the "
                             "largest share of the step; bf16x3 issues 3 tcgen05.mma per algorithmic product, so its "
                             "tensor-pipe occupancy is ~3x this fraction" if a.precision == "bf16x3" else
                             "algorithmic FLOPs over the CUDA-event duration of the kernel with the largest share of the step")}
        stage_table = {k: {"calls": v["calls"], "ms": round(v["ms"], 4), "share": round(v["ms"] / tot_ms, 4),
                           "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3) if v["ms"] > 0 else 0.0,
                           "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0}
                       for k, v in sorted(stages.items(), key=lambda kv: -kv[1]["ms"])}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x3": "bf16x3+f32", "bf16": "bf16+f32"}[a.precision], "data": "synthetic",
                "config": {"workload": f"batch={B} synthetic {size}x{size} pair inference per GPU, full CNN+ViT+EM "
                                       "module, fp32 (BASELINE.json configs[1]); random-init weights",
                           "pairs_per_gpu_per_step": B, "micro_batch": MB, "device_image_dtype": "u8" if a.u8 else "f32", "precision": {"fp32": "fp32 operands, fp32 accumulate (SIMT)",
                                         "bf16x3": "transformer GEMMs on tcgen05 with split-bf16 operands (a0b0+a0b1+a1b0), fp32 accumulate; rest fp32",
                                         "bf16": "transformer GEMMs on tcgen05 in bf16, fp32 accumulate; rest fp32"}[a.precision],
                           "l2_policy": f"inputs larger than L2 ({images.numel() * 4 / 1e6:.0f} MB of images per step vs 126 MB L2)",
                           "parallelism": f"{world} shard(s), no collective",
                           "cnn": "own implicit-GEMM convolutions (NHWC, BN folded), no cuDNN"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": total_pairs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / a.steps,
                        "api": "rel_pose_b200.parallel.StreamedInference (pinned float32 host images [B,2,3,H,W], the reference's input dtype; "
                               "row-selective H2D: only the rows the 224x224 nearest resize reads are copied)",
                        "host_bytes_per_step": host_images.numel() * 4},
                "e2e_u8": {"value": total_pairs / (ms_e2e_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_u8,
                           "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_u8 / a.steps,
                           "api": "same call with uint8 host images (cv2.imread's dtype, demo.py:65): identical poses"},
                "roofline": roofline, "stages": stage_table,
                "achieved_tflops_whole_step": value * FLOP_PER_PAIR / 1e12}
        if world == 1 and not a.no_cpu_baseline:
            run, kind, desc = cpu_reference_forward_factory(a.ref_batch, size)
            v, n, dt = time_cpu(run, a.ref_batch, a.cpu_seconds, 50)
            cores = os.cpu_count() or 1
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{n} forwards x {a.ref_batch} pairs of {size}x{size} in {dt:.1f} s on "
                                              f"{cores} host threads; {desc}"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
