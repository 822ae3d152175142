#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the rel_pose hot path (ViTEss.forward) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs main,parity,eager,config4,config5,geometry]

A step = one forward pass of the hot path over one batch of synthetic 384x384 image pairs per GPU
(BASELINE.json configs[1]: batch=64 pairs, fp32-class arithmetic, random-init weights).  Pairs are
independent, so N GPUs run N shards with no data-path collective (weak scaling).  Rank 0 prints ONE
JSON line.  Besides the headline (`value`, `e2e`, `roofline`, `cpu_baseline`) the same line carries, each
with its own clock record:
  parity              rotation / translation error of the TIMED batch against the reference (CPU, fp32)
  gpu_eager_baseline  the reference's own GPU path (torch eager: cuBLAS / cuDNN) on the same B200, fp32 with
                      TF32 off, TF32 on, and bf16 autocast -- the bar the kernels have to beat (N = 1 only)
  config4             BASELINE.json configs[3]: 4096 pairs in total, bf16 operands, uint8 images, strong-scaled
                      4096/N per rank in micro-batches of 256, device-resident and end to end
  config5             BASELINE.json configs[4]: the train.py loop on synthetic pairs, 6 pairs per GPU,
                      DistributedDataParallel (NCCL) when N > 1, all-reduce time and overlap
  geometry            BASELINE.json configs[2]: 3x3 SVD / E->(R,t) / SE3 microbench on 2^20 elements (N = 1 only)
`--impl reference` times the reference's CPU implementation of the same path on the host cores: the unmodified
reference tree when present (/root/reference, or its copy in baseline/_ref), else the pinned port oracle/torch_port.py.
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
# attention-type batched GEMMs per pair: self-attention QK^T + PV (5 blocks) 2.55 G, Essential Matrix
# Module scores 0.255 G, A [v|pos] 0.279 G, [v|pos]^T T 0.034 G
ATTN_FLOP_PER_PAIR = 3.12e9


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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md).  One process for
    the whole run; `mark()` returns the summary of the samples taken since the previous mark, so every leg of the
    line carries the clock record of its own window."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.pos = index, None, [], 0

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

    @staticmethod
    def _summarise(lines):
        sm, mx, reasons, power = [], None, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1]); power = max(power, float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": mx,
                "power_w_max": power, "reasons": sorted(reasons), "samples": len(sm)}

    def mark(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)                          # let the last 100 ms sample of the window arrive
        n = len(self.lines)
        out = self._summarise(self.lines[self.pos:n])
        self.pos = n
        return out

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self._summarise(self.lines)


# ---------------------------------------------------------------------------------------------- reference (CPU / GPU eager)
def reference_inputs(batch, size, seed=0):
    import torch
    from rel_pose_b200 import synthetic as S
    g = torch.Generator().manual_seed(seed)
    images = (torch.rand(batch, 2, 3, size, size, generator=g) * 255).floor()
    intr = torch.from_numpy(S.make_intrinsics_numpy(batch))
    Gs = torch.zeros(batch, 2, 7); Gs[..., 6] = 1
    return images, Gs, intr


def reference_forward_factory(device="cpu", sd=None, dtype=None):
    """Returns (fn(images, Gs, intr) -> [B,2,7] tensor on `device`, kind, description): the unmodified reference ViTEss when
    its tree is present (kind "reference"), else the pinned port (kind "port").  Test / baseline infrastructure."""
    import torch
    from rel_pose_b200 import synthetic as S
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    sd = sd if sd is not None else S.make_state_dict(0, "init")
    root = ref_loader.find_reference_root()
    if root is not None:
        model, SE3 = ref_loader.load_reference_model()
        model.load_state_dict(sd)
        model = model.to(device).eval()
        if dtype is not None:
            model = model.to(dtype)

        def run(images, Gs, intr):
            with torch.no_grad():
                if str(device) == "cpu":
                    with ref_loader.cpu_only():       # its unconditional .cuda() (vision_transformer.py:211) stays on the CPU
                        return model(images, SE3(Gs), intrinsics=intr.clone())[0].data
                return model(images, SE3(Gs), intrinsics=intr.clone())[0].data
        return run, "reference", f"unmodified reference ViTEss from {root}"
    import torch_port
    p = {k: (v.to(device) if dtype is None else v.to(device, dtype)) for k, v in sd.items() if v.dtype != torch.int64}

    def run(images, Gs, intr):
        with torch.no_grad():
            return torch_port.forward(images, Gs, intr, p)
    return run, "port", "oracle/torch_port.py (PyTorch port pinned to the reference's golden vectors)"


def cpu_reference_forward_factory(batch, size):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    images, Gs, intr = reference_inputs(batch, size)
    fn, kind, desc = reference_forward_factory("cpu")
    return (lambda: fn(images, Gs, intr)), kind, desc + ", PyTorch CPU fp32"


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
                                   "(bounded sample of configs[1]; CPU pairs/s is flat in the batch size, BASELINE.md section 2)",
                       "precision": "fp32"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def pose_errors(got, ref):
    """rotation 2 acos|<q^,q>| (test_matterport.py:41) and |t^-t|/|t| of pose 1 -> dict of median / max."""
    import numpy as np
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    qg, qr = got[:, 1, 3:], ref[:, 1, 3:]
    dot = np.abs((qg * qr).sum(-1)) / (np.linalg.norm(qg, axis=-1) * np.linalg.norm(qr, axis=-1))
    rot = 2.0 * np.arccos(np.clip(dot, 0.0, 1.0))
    tr = np.linalg.norm(got[:, 1, :3] - ref[:, 1, :3], axis=-1) / np.linalg.norm(ref[:, 1, :3], axis=-1)
    return {"rot_median_rad": float(np.median(rot)), "rot_max_rad": float(rot.max()),
            "trans_median_rel": float(np.median(tr)), "trans_max_rel": float(tr.max()), "pairs": int(got.shape[0])}


# ---------------------------------------------------------------------------------------------- legs
def leg_gpu_eager(dev, B, size, steps):
    """The reference's GPU path -- torch eager on this B200 (cuBLAS GEMMs, cuDNN convolutions, separate softmax / LN /
    GELU kernels, 576x576 attention matrices in HBM) -- CUDA-event timed on the same kind of batch.  SURVEY.md 8(d)."""
    import torch
    out = {"batch": B, "steps": steps, "timing": "CUDA events, 2 warm-up forwards, device-resident inputs"}
    images, Gs, intr = reference_inputs(B, size, seed=1)
    images, Gs, intr = images.to(dev), Gs.to(dev), intr.to(dev)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port
    from rel_pose_b200 import synthetic as S
    sd = S.make_state_dict(0, "init")
    p = {k: v.to(dev) for k, v in sd.items() if v.dtype != torch.int64}

    def port(images, Gs, intr):
        with torch.no_grad():
            return torch_port.forward(images, Gs, intr, p)
    impls = [("port", port, "oracle/torch_port.py on cuda (closed-form positional table: FASTER than the reference)")]
    import ref_loader
    if ref_loader.find_reference_root() is not None:
        fn, kind, desc = reference_forward_factory(dev, sd)
        impls.append(("reference", fn, desc + " on cuda, incl. its host-side positional-encoding loop (vision_transformer.py:139-151)"))

    def timed(fn, n):
        for _ in range(2):
            fn(images, Gs, intr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn(images, Gs, intr)
        e1.record()
        torch.cuda.synchronize()
        return B * n / (max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0))
    tf32_state = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, fn, desc in impls:
            r = {"what": desc}
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            r["fp32"] = timed(fn, steps)
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
            r["tf32"] = timed(fn, steps)
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            with torch.autocast("cuda", dtype=torch.bfloat16):
                r["bf16_autocast"] = timed(fn, steps)
            out[name] = r
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32_state
    best = impls[-1][0]
    out["fp32"], out["bf16"], out["tf32"] = out[best]["fp32"], out[best]["bf16_autocast"], out[best]["tf32"]
    out["unit"] = UNIT
    out["kind"] = best
    out["note"] = ("fp32 = TF32 disabled (the precision the 1e-4 parity bar is stated in); bf16 = torch.autocast(bfloat16); "
                   "top-level fp32/bf16/tf32 are the unmodified reference's when its tree is present, else the port's")
    del images, p
    torch.cuda.empty_cache()
    return out


def leg_config4(model, dev, rank, world, size, total_pairs, micro, passes):
    """BASELINE.json configs[3]: `total_pairs` pairs in total, bf16 operands (fp32 accumulate), uint8 source images,
    strong-scaled over the ranks, micro-batches of `micro` pairs; device-resident and end to end (pinned uint8 host
    micro-batches through StreamedInference, H2D of every micro-batch inside the timed region)."""
    import torch
    import torch.distributed as dist
    from rel_pose_b200 import SE3, synthetic as S
    from rel_pose_b200.parallel import StreamedInference, shard_range, micro_batches
    lo, hi = shard_range(total_pairs, world, rank)
    n = hi - lo
    g = torch.Generator(device=dev).manual_seed(4096 + rank)
    images = torch.empty((n, 2, 3, size, size), dtype=torch.uint8, device=dev)
    mbs = micro_batches(0, n, micro)
    for a0, a1 in mbs:                                       # generated on the device, chunk by chunk
        images[a0:a1] = (torch.rand(a1 - a0, 2, 3, size, size, generator=g, device=dev) * 255).to(torch.uint8)
    intr0 = torch.from_numpy(S.make_intrinsics_numpy(micro)).to(dev)
    Gs = SE3.Identity(micro, 2, device=dev)
    prev = model.precision
    model.precision = "bf16"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass():
        for a0, a1 in mbs:
            model(images[a0:a1], Gs[:a1 - a0], intrinsics=intr0[:a1 - a0].clone())
    try:
        with torch.no_grad():
            for a0, a1 in mbs[:2]:
                model(images[a0:a1], Gs[:a1 - a0], intrinsics=intr0[:a1 - a0].clone())
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(passes):
                one_pass()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            # end to end: a ring of three pinned host micro-batches, every micro-batch copied H2D inside the timed region
            ring = [torch.empty((micro, 2, 3, size, size), dtype=torch.uint8, pin_memory=True) for _ in range(3)]
            for i, h in enumerate(ring):
                h.copy_(images[(i * micro) % max(1, n - micro + 1):][:micro])
            host_intr = torch.from_numpy(S.make_intrinsics_numpy(micro)).pin_memory()
            host_Gs = SE3.Identity(micro, 2).data.pin_memory()
            runner = StreamedInference(model, dev)

            def feed(k):
                for i in range(k):
                    a0, a1 = mbs[i % len(mbs)]
                    yield ring[i % 3][:a1 - a0], host_Gs[:a1 - a0], host_intr[:a1 - a0]
            for _ in runner.run(feed(2)):
                pass
            barrier()
            t0 = time.perf_counter()
            for _ in runner.run(feed(len(mbs) * passes)):
                pass
            barrier()
            ms_e2e = (time.perf_counter() - t0) * 1e3
            h2d = runner.last_image_h2d_bytes
            t = torch.tensor([ms, ms_e2e], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_e2e = float(t[0].item()), float(t[1].item())
    finally:
        model.precision = prev
        del images
        torch.cuda.empty_cache()
    pairs = total_pairs * passes
    val = pairs / (ms * 1e-3)
    pk = peaks()
    return {"workload": f"{total_pairs} synthetic {size}x{size} pairs in total, bf16 operands / fp32 accumulate, uint8 images generated on "
                        f"the device, {world} rank(s) x {n} pairs (strong scaling), micro-batches of {micro}, {passes} pass(es)",
            "value": val, "unit": UNIT, "scaling": "strong", "n_gpus": world, "pairs_per_rank": n, "micro_batch": micro,
            "ms_per_pass": ms / passes, "dtype": "bf16+f32",
            "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_pass": ms_e2e / passes,
                    "h2d_bytes_per_micro_batch": h2d, "d2h_bytes_per_micro_batch": micro * 2 * 7 * 4,
                    "api": "rel_pose_b200.parallel.StreamedInference over pinned uint8 host micro-batches (row-selective H2D)"},
            "achieved_tflops_whole_step": val * FLOP_PER_PAIR / 1e12,
            "frac_of_sustained_bf16_peak_all_gpus": val * FLOP_PER_PAIR / 1e12 / (pk["bf16_tflops_sustained"] * world)}


def reference_train_eager(dev, a, steps, warm):
    """The GPU bar of config 5: the UNMODIFIED reference model in .train() on this B200 under torch eager (cuDNN / cuBLAS
    forward, torch autograd backward), with the optimizer stack train.py builds (Adam + OneCycleLR + clip_grad_norm_, train.py:69-73,
    161-165).  lietorch is absent, so the loss backward -- a few hundred FLOPs on [B,2,7] -- is replaced by a fixed upstream pose
    gradient injected at poses_est[0].data (what make_golden_train.py does); everything that costs time is the reference's own.
    Returns ms per step for fp32 (TF32 off) / TF32 / bf16 autocast, or None when the reference tree is not present."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    if ref_loader.find_reference_root() is None:
        return None
    from rel_pose_b200 import synthetic as S, train_synthetic as T
    out = {"what": "unmodified reference ViTEss.train() + torch autograd + Adam/OneCycleLR/clip_grad_norm_, torch eager on cuda",
           "pairs_per_gpu": a.batch, "image_size": list(a.size), "steps": steps, "warmup_steps": warm,
           "timing": "CUDA events around the timed steps, device-resident batches"}
    batches = [T.make_batch(i, 0, a.batch, a.size[0], a.size[1], dev) for i in range(2)]
    g = torch.Generator().manual_seed(5)
    gpose = (torch.randn(a.batch, 2, 7, generator=g) * 0.1).to(dev)
    tf32_state = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)

    def run(mode):
        model, SE3 = ref_loader.load_reference_model()
        model.load_state_dict(S.make_state_dict(0, "init"))
        model = model.to(dev).train()
        for q in list(model.resnet.layer3.parameters()) + list(model.resnet.layer4.parameters()):
            q.requires_grad = False                                                  # train.py:60-64
        params = [q for q in model.parameters() if q.requires_grad]
        opt = torch.optim.Adam(params, lr=a.lr, weight_decay=a.weight_decay)
        sch = torch.optim.lr_scheduler.OneCycleLR(opt, a.lr, a.total_steps, pct_start=a.warmup / a.total_steps, div_factor=25,
                                                  anneal_strategy="cos")
        tf = mode == "tf32"
        torch.backends.cuda.matmul.allow_tf32 = tf
        torch.backends.cudnn.allow_tf32 = tf

        def step(i):
            images, poses, intr = batches[i % 2]
            Gs = torch.zeros(a.batch, 2, 7, device=dev); Gs[..., 6] = 1
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                est = model(images, SE3(Gs), intrinsics=intr.clone())[0].data
            est.backward(gpose.to(est.dtype))
            torch.nn.utils.clip_grad_norm_(params, a.clip)
            opt.step()
            sch.step()
        for i in range(warm):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / steps
        del model, opt, sch
        torch.cuda.empty_cache()
        return ms
    try:
        for mode in ("fp32", "tf32", "bf16_autocast"):
            out["ms_per_step_" + mode] = run(mode)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32_state
    return out


def leg_config5(dev, rank, world, local, steps, warm):
    from rel_pose_b200 import train_synthetic as T
    a = T.default_options(steps=steps, warmup_steps=warm, pool=4)
    res = T.train_loop(a, dev, rank, world, local)
    if rank == 0 and res is not None and world == 1:
        try:
            res["gpu_eager_baseline"] = reference_train_eager(dev, a, min(steps, 10), 3)
        except Exception as e:                                                       # baseline leg: never takes the line down
            res["gpu_eager_baseline"] = {"error": repr(e)[:300]}
    return res


def leg_geometry(dev):
    """BASELINE.json configs[2]: N = 2^20 elements; 8 back-to-back launches per CUDA graph (inputs rotate over 8 copies,
    larger than L2), CUDA events around 13 replays.  HBM roofline: N x algorithmic bytes / time vs the measured copy peak."""
    import torch
    from rel_pose_b200 import ops
    N = 1 << 20
    pk = peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    E = torch.randn(N, 3, 3, generator=g, device=dev)
    xi = torch.randn(N, 6, generator=g, device=dev) * torch.tensor([1, 1, 1, .5, .5, .5], device=dev)
    X = ops.se3_exp_fwd(xi)
    Y = ops.se3_exp_fwd(xi.flip(0).contiguous())
    Es = [E.clone() for _ in range(8)]; Xs = [X.clone() for _ in range(8)]; xis = [xi.clone() for _ in range(8)]
    cnt = [0]

    def rot(lst):
        cnt[0] += 1
        return lst[cnt[0] % 8]

    def timeit(fn, iters=104, warm=8):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(st):
            for _ in range(8):
                fn()
            st.synchronize()
            with torch.cuda.graph(gr, stream=st):
                keep = [fn() for _ in range(8)]
            reps = iters // 8
            gr.replay()
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                gr.replay()
            e1.record(st)
            st.synchronize()
        del keep
        return e0.elapsed_time(e1) / (reps * 8) * 1e-3
    rows = {}
    for name, fn, nbytes in [
        ("svd3", lambda: ops.svd3(rot(Es)), 36 + 84),
        ("essential_to_rt", lambda: ops.essential_to_rt(rot(Es)), 36 + 84),
        ("se3_log", lambda: ops.se3_log_fwd(rot(Xs)), 28 + 24),
        ("se3_exp", lambda: ops.se3_exp_fwd(rot(xis)), 24 + 28),
        ("se3_mul", lambda: ops.se3_mul_fwd(rot(Xs), Y), 56 + 28),
        ("se3_inv", lambda: ops.se3_inv_fwd(rot(Xs)), 28 + 28)]:
        t = timeit(fn)
        gbs = N * nbytes / t / 1e9
        rows[name] = {"us": round(t * 1e6, 2), "GB/s": round(gbs, 1), "frac_of_measured_hbm": round(gbs / pk["hbm_gbs"], 3),
                      "bytes_per_element": nbytes}
    del Es, Xs, xis
    torch.cuda.empty_cache()
    return {"workload": "N = 2^20 elements, fp32, inputs rotated over 8 copies (larger than L2), 8 launches per CUDA graph, "
                        "13 replays, CUDA events", "hbm_peak_gbs": pk["hbm_gbs"], "peak_source": pk["source"], "kernels": rows}


class Watchdog:
    """The extra legs run after the headline has been measured.  If one of them hangs (a rank that died inside a
    collective leaves the others waiting), rank 0 still prints the line it has and every rank exits."""

    def __init__(self, seconds, on_fire):
        self.t = threading.Timer(seconds, self._fire)
        self.t.daemon = True
        self.on_fire = on_fire

    def _fire(self):
        try:
            self.on_fire()
        finally:
            os._exit(0)

    def start(self):
        self.t.start()

    def cancel(self):
        self.t.cancel()


def guarded(name, fn, sampler, line):
    """A leg never takes the headline down with it: failures are recorded in the leg's slot."""
    try:
        res = fn()
    except Exception as e:                     # noqa: BLE001
        import traceback
        sys.stderr.write(f"[bench] leg {name} failed:\n{traceback.format_exc()}\n")
        res = {"error": f"{type(e).__name__}: {e}"[:500]}
    if sampler is not None and isinstance(res, dict):
        res["clocks"] = sampler.mark()
    if line is not None and res is not None:
        line[name] = res
    return res


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
                    help="process the per-GPU batch in chunks of this many pairs")
    ap.add_argument("--u8", action="store_true", help="device-resident images as uint8")
    ap.add_argument("--ref-batch", type=int, default=8, help="pairs per step of the CPU reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end legs (e2e = null)")
    ap.add_argument("--legs", default="main,parity,eager,config4,config5,geometry",
                    help="comma list of the extra legs to run (main is always run)")
    ap.add_argument("--legs-timeout", type=float, default=420.0, help="watchdog for the extra legs (seconds)")
    ap.add_argument("--config4-pairs", type=int, default=4096)
    ap.add_argument("--config4-micro", type=int, default=256)
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"],
                    help="arithmetic of the GEMMs (fp32 SIMT | tcgen05 split-bf16 | tcgen05 bf16)")
    a = ap.parse_args()
    guard_stdout()
    if a.impl == "reference":
        return run_reference_arm(a)
    legs = set(x.strip() for x in a.legs.split(",") if x.strip())

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
        # one process per GPU; spread the ranks over the host cores (every rank drives its own copy engine from pinned memory)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except Exception:
            pass
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
        # clocks / throttle reasons are sampled (nvidia-smi every 100 ms) during every leg; `clocks` of the headline covers
        # warm-up + timed steps + per-kernel timing + the end-to-end legs, each later leg carries its own window
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
        timed_out = out[0].data.float().cpu().numpy()          # poses of the TIMED batch (parity leg)

        # ---- per-kernel timing with CUDA events on the launching stream (roofline) ----
        timer = ops.StageTimer()
        ops.set_timer(timer)
        for _ in range(min(3, a.steps)):
            step()
        torch.cuda.synchronize()
        ops.set_timer(None)
        stages = timer.summary()
        n_timer_steps = min(3, a.steps)

        # ---- end to end through the public API: pinned host -> H2D, forward, D2H of the poses ----
        # rel_pose_b200.parallel.StreamedInference is the user-facing call for batched inference from host
        # memory: the H2D copy of step k+1 runs on a side stream while step k computes.  Every step copies its
        # own inputs from pinned host memory and reads its result back; all of it is inside the timed region.
        from rel_pose_b200.parallel import StreamedInference
        host_intr = torch.from_numpy(S.make_intrinsics_numpy(B)).pin_memory()
        host_Gs = SE3.Identity(B, 2).data.pin_memory()
        runner = StreamedInference(model, dev)

        def e2e_run(src, n):
            last = None
            for last in runner.run((src, host_Gs, host_intr) for _ in range(n)):
                pass
            return last

        def e2e_time(src):
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
            # host wall clock is the honest end-to-end figure (it includes the final D2H wait); events agree within noise
            return max(float(t[0].item()), float(t[1].item()))

        e2e = e2e_f32 = None
        if not a.no_e2e:
            # uint8 is what the images are at the source (cv2.imread, demo.py:65); the reference converts them to float32 on
            # the host before .cuda() (demo.py:71-76) -- both entry dtypes are accepted and give identical poses
            host_u8 = torch.empty((B, 2, 3, size, size), dtype=torch.uint8, pin_memory=True)
            host_u8.copy_(images.to(torch.uint8))
            ms_u8 = e2e_time(host_u8)
            h2d_u8 = (runner.last_image_h2d_bytes or host_u8.numel()) + host_intr.numel() * 4 + host_Gs.numel() * 4
            host_images = torch.empty((B, 2, 3, size, size), dtype=torch.float32, pin_memory=True)
            host_images.copy_(images)
            ms_f32 = e2e_time(host_images)
            h2d_f32 = (runner.last_image_h2d_bytes or host_images.numel() * 4) + host_intr.numel() * 4 + host_Gs.numel() * 4
            d2h = B * 2 * 7 * 4
            tp = world * B * a.steps
            e2e = {"value": tp / (ms_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_u8 / a.steps, "host_bytes_per_step": host_u8.numel(),
                   "api": "rel_pose_b200.parallel.StreamedInference, pinned uint8 host images [B,2,3,H,W] (cv2.imread's dtype, "
                          "demo.py:65); row-selective H2D: only the rows the 224x224 nearest resize reads are copied; poses "
                          "identical to the float32 entry"}
            e2e_f32 = {"value": tp / (ms_f32 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_f32, "d2h_bytes_per_step": d2h,
                       "ms_per_step": ms_f32 / a.steps, "host_bytes_per_step": host_images.numel() * 4,
                       "api": "same call with pinned float32 host images (what the reference's callers hand to .cuda())"}
            del host_images, host_u8
        clocks = sampler.mark() if sampler else None
        if clocks is not None:
            clocks["window"] = "warm-up + timed steps + per-kernel timing + end-to-end legs (GPU under the bench load throughout)"

    line = None
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
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
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
        # attention-type batched GEMMs (north_star: "achieved fraction of the attention-GEMM roofline"): the 3.12 GFLOP/pair
        # attention subset over the time of the kernels that execute it (self-attention x5 + Essential Matrix Module)
        att_ms = sum(v["ms"] for k, v in stages.items() if k.startswith("self_attention") or k.startswith("essential")) / n_timer_steps
        att_tfl = B * ATTN_FLOP_PER_PAIR / (att_ms * 1e-3) / 1e12 if att_ms > 0 else 0.0
        stage_table = {k: {"calls": v["calls"], "ms": round(v["ms"], 4), "share": round(v["ms"] / tot_ms, 4),
                           "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 3) if v["ms"] > 0 else 0.0,
                           "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0}
                       for k, v in sorted(stages.items(), key=lambda kv: -kv[1]["ms"])}
        prec_name = {"fp32": "fp32 operands, fp32 accumulate (SIMT)",
                     "bf16x3": "fp32-class: every GEMM / convolution / attention product on tcgen05 with split-bf16 operands "
                               "(a0b0+a0b1+a1b0), fp32 accumulate; elementwise fp32",
                     "bf16": "tcgen05 in bf16, fp32 accumulate; elementwise fp32"}[a.precision]
        cfg_name = "BASELINE.json configs[1]" if (B == 64 and a.precision != "bf16") else "custom"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x3": "bf16x3+f32", "bf16": "bf16+f32"}[a.precision], "data": "synthetic",
                "config": {"workload": f"batch={B} synthetic {size}x{size} pair inference per GPU, full CNN+ViT+EM module, "
                                       f"precision {a.precision} ({cfg_name}); random-init weights",
                           "pairs_per_gpu_per_step": B, "micro_batch": MB, "device_image_dtype": "u8" if a.u8 else "f32",
                           "precision": prec_name,
                           "l2_policy": f"inputs larger than L2 ({images.numel() * images.element_size() / 1e6:.0f} MB of images per step vs 126 MB L2)",
                           "parallelism": f"{world} shard(s), no collective",
                           "cnn": "own implicit-GEMM convolutions (NHWC, BN folded), no cuDNN"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": e2e, "e2e_f32": e2e_f32,
                "roofline": roofline,
                "attention_gemm_roofline_frac": att_tfl / pk["bf16_tflops_sustained"],
                "attention_gemm": {"achieved": att_tfl, "unit": "TFLOP/s", "peak": pk["bf16_tflops_sustained"],
                                   "frac": att_tfl / pk["bf16_tflops_sustained"], "ms_per_step": att_ms,
                                   "flop_per_pair": ATTN_FLOP_PER_PAIR,
                                   "kernels": "self_attention_tc x5 + essential_tc (stats + accumulate)"},
                "stages": stage_table,
                "achieved_tflops_whole_step": value * FLOP_PER_PAIR / 1e12}

    # ---- parity of the timed batch against the reference (CPU fp32): rank 0's shard ----
    if rank == 0 and "parity" in legs:
        def parity():
            torch.set_num_threads(os.cpu_count() or 1)
            fn, kind, desc = reference_forward_factory("cpu")
            t0 = time.perf_counter()
            ref = fn(images.float().cpu(), Gs.data.cpu(), intr0.cpu()).numpy()
            r = pose_errors(timed_out, ref)
            # the same reference in float64 = the exact result of the model: how far the float32 reference itself is from it
            # (its own rounding, the noise floor of the 1e-4 bar) and how far this implementation is
            fn64, _, _ = reference_forward_factory("cpu", dtype=torch.float64)
            ref64 = fn64(images.double().cpu(), Gs.data.double().cpu(), intr0.double().cpu()).numpy()
            r["vs_float64_reference"] = {"ours": pose_errors(timed_out, ref64), "float32_reference": pose_errors(ref, ref64)}
            r.update({"against": desc + ", CPU fp32, same inputs and weights as the timed batch", "kind": kind,
                      "bar": "1e-4 rad / 1e-4 relative translation (north_star)", "cpu_seconds": round(time.perf_counter() - t0, 2),
                      "pass": bool(r["rot_max_rad"] < 1e-4 and r["trans_max_rel"] < 1e-4) if a.precision != "bf16" else None})
            return r
        guarded("parity", parity, None, line)
    if world > 1:
        dist.barrier()

    def on_timeout():
        if rank == 0 and line is not None:
            line["legs_timeout"] = f"an extra leg did not finish within {a.legs_timeout:.0f} s; legs present in this line did"
            emit(line)
    dog = Watchdog(a.legs_timeout, on_timeout)
    dog.start()
    if "eager" in legs and world == 1:
        guarded("gpu_eager_baseline", lambda: leg_gpu_eager(dev, B, size, max(3, min(a.steps, 10))), sampler, line)
    del images
    torch.cuda.empty_cache()
    if "config4" in legs:
        r = guarded("config4", lambda: leg_config4(model, dev, rank, world, size, a.config4_pairs, a.config4_micro, 2),
                    sampler, line if rank == 0 else None)
    if "config5" in legs:
        guarded("config5", lambda: leg_config5(dev, rank, world, local, max(10, a.steps), 4), sampler, line if rank == 0 else None)
    if "geometry" in legs and world == 1:
        guarded("geometry", lambda: leg_geometry(dev), sampler, line)
    if sampler:
        sampler.stop()
    dog.cancel()

    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            def cpu_base():
                run, kind, desc = cpu_reference_forward_factory(a.ref_batch, size)
                v, n, dt = time_cpu(run, a.ref_batch, a.cpu_seconds, 50)
                cores = os.cpu_count() or 1
                return {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{n} forwards x {a.ref_batch} pairs of {size}x{size} in {dt:.1f} s on {cores} host threads; {desc}"}
            guarded("cpu_baseline", cpu_base, None, line)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
