"""Fused training-step tail -- what train.py:69-73,161-165 does with

    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    scheduler = OneCycleLR(optimizer, args.lr, args.steps, pct_start=args.warmup/args.steps, div_factor=25, cycle_momentum=False)
    ...
    torch.nn.utils.clip_grad_norm_(model.parameters(), args.clip); optimizer.step(); scheduler.step()

as ONE object whose `step()` is three launches of the CUDA library (csrc/optim.cu: gradient norm, its reduction, the
clipped Adam update) and no device->host synchronisation.  Same arithmetic as torch's `_single_tensor_adam`, same
learning-rate curve as `OneCycleLR(anneal_strategy="cos", three_phase=False)` (closed form below, checked against
torch in tests/test_host_cpu.py).  No CPU fallback: parameters must live on a CUDA device.
"""
import math

import numpy as np
import torch

from . import _lib, ops

_DESC = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8")])     # struct TensorDesc, 40 bytes


def one_cycle_lr(step_num, max_lr, total_steps, pct_start=0.3, div_factor=25.0, final_div_factor=1e4):
    """Learning rate OneCycleLR has set for optimizer step number `step_num` (0-based) -- torch/optim/lr_scheduler.py."""
    if step_num > total_steps:          # torch: "Tried to step {step_num} times. The specified number of total steps is ..."
        raise ValueError(f"Tried to step {step_num} times. The specified number of total steps is {total_steps}")
    initial_lr = max_lr / div_factor
    min_lr = initial_lr / final_div_factor
    phases = [(float(pct_start * total_steps) - 1, initial_lr, max_lr), (total_steps - 1, max_lr, min_lr)]
    start_step = 0.0
    lr = initial_lr
    for i, (end_step, lo, hi) in enumerate(phases):
        if step_num <= end_step or i == len(phases) - 1:
            pct = (step_num - start_step) / (end_step - start_step)
            lr = hi + (lo - hi) / 2.0 * (math.cos(math.pi * pct) + 1)
            break
        start_step = end_step
    return lr


class FusedAdamOneCycle:
    """Adam(lr, betas, eps, weight_decay as L2-in-gradient) + clip_grad_norm_(clip) + OneCycleLR in one `step()`."""

    def __init__(self, params, max_lr, total_steps, pct_start, div_factor=25.0, final_div_factor=1e4, betas=(0.9, 0.999),
                 eps=1e-8, weight_decay=0.0, clip=None, chunk=65536):
        self.params = [p for p in params]
        assert self.params, "no parameters"
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.RelposeLibraryError("FusedAdamOneCycle: parameters must live on a CUDA device (no CPU fallback)")
        for p in self.params:
            assert p.device == dev and p.dtype == torch.float32 and p.is_contiguous()
        self.device = dev
        self.max_lr, self.total_steps, self.pct_start = float(max_lr), int(total_steps), float(pct_start)
        self.div_factor, self.final_div_factor = float(div_factor), float(final_div_factor)
        self.betas, self.eps, self.weight_decay = (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.clip = None if clip is None else float(clip)
        self.chunk = int(chunk)
        self.step_num = 0                                     # optimizer steps taken so far
        # moments of ALL parameters in two flat buffers (views per parameter): stable addresses, one allocation
        total = sum(p.numel() for p in self.params)
        pad = lambda n: (n + 3) // 4 * 4                      # keep every view 16-byte aligned
        offs, o = [], 0
        for p in self.params:
            offs.append(o); o += pad(p.numel())
        self._m = torch.zeros(o, dtype=torch.float32, device=dev)
        self._v = torch.zeros(o, dtype=torch.float32, device=dev)
        self.exp_avg = [self._m[a:a + p.numel()].view_as(p) for a, p in zip(offs, self.params)]
        self.exp_avg_sq = [self._v[a:a + p.numel()].view_as(p) for a, p in zip(offs, self.params)]
        self.total_numel = total
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)      # written by every step(), never read on the host
        self._key = None
        self._tables = None
        self._upload_done = None
        self._ever_updated = set()

    # --------------------------------------------------------------------------------------------
    def lr_at(self, step_num):
        return one_cycle_lr(step_num, self.max_lr, self.total_steps, self.pct_start, self.div_factor, self.final_div_factor)

    @property
    def lr(self):
        """Learning rate the NEXT step() will use (what `optimizer.param_groups[0]['lr']` shows in the reference)."""
        return self.lr_at(self.step_num)

    def zero_grad(self, set_to_none=False):
        """Gradients are zeroed in place by default (stable addresses -> the descriptor table is uploaded once)."""
        grads = [p.grad for p in self.params if p.grad is not None]
        if set_to_none:
            for p in self.params:
                p.grad = None
        elif grads:
            torch._foreach_zero_(grads)

    def _build_tables(self, active):
        """Device tables for the multi-tensor kernels; rebuilt only when a pointer or shape changed."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) for _, p in active)
        if key == self._key:
            return self._tables
        if self._upload_done is not None:
            self._upload_done.synchronize()                   # the pinned staging buffers are about to be rewritten
        desc = np.zeros(len(active), _DESC)
        blk_t, blk_o = [], []
        for j, (i, p) in enumerate(active):
            g = p.grad
            assert g.dtype == torch.float32 and g.is_contiguous() and g.device == self.device
            desc[j] = (p.data_ptr(), g.data_ptr(), self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr(), p.numel())
            for off in range(0, p.numel(), self.chunk):
                blk_t.append(j); blk_o.append(off)
        h_desc = torch.from_numpy(desc.view(np.uint8).copy()).pin_memory()
        h_t = torch.tensor(blk_t, dtype=torch.int32).pin_memory()
        h_o = torch.tensor(blk_o, dtype=torch.int64).pin_memory()
        d_desc, d_t, d_o = (h.to(self.device, non_blocking=True) for h in (h_desc, h_t, h_o))
        partial = torch.empty(len(blk_t), dtype=torch.float32, device=self.device)
        self._upload_done = torch.cuda.Event()
        self._upload_done.record()
        self._host_keepalive = (h_desc, h_t, h_o)
        self._key, self._tables = key, (d_desc, d_t, d_o, partial, len(blk_t))
        return self._tables

    # ---- CUDA-graph support: the per-step scalars live in device memory --------------------------------------
    def step_scalars(self, step_num=None):
        """(lr_t / (1 - beta1^t), sqrt(1 - beta2^t)) for optimizer step `step_num` (0-based; default: the next one)."""
        k = self.step_num if step_num is None else step_num
        t = k + 1
        return (self.lr_at(k) / (1.0 - self.betas[0] ** t), math.sqrt(1.0 - self.betas[1] ** t))

    def enable_device_scalars(self):
        """After this call step() reads the learning rate / bias corrections from a 2-float device buffer, so a step
        captured in a CUDA graph can be replayed: call `upload_step_scalars()` (an 8-byte async copy) before each replay
        and `advance()` after it.  Eager `step()` keeps working (it uploads and advances by itself)."""
        if getattr(self, "_hyper_dev", None) is None:
            # a RING of pinned slots: the host runs many steps ahead of the device, and a slot must not be rewritten
            # before its asynchronous copy has executed (the launch queue is far shorter than the ring)
            self._hyper_host = torch.zeros((self._HYPER_RING, 2), dtype=torch.float32).pin_memory()
            self._hyper_dev = torch.zeros(2, dtype=torch.float32, device=self.device)
        return self

    _HYPER_RING = 4096

    def upload_step_scalars(self):
        a, b = self.step_scalars()
        slot = self._hyper_host[self.step_num % self._HYPER_RING]
        slot[0] = a; slot[1] = b
        self._hyper_dev.copy_(slot, non_blocking=True)

    def advance(self):
        """Book-keeping of one replayed step (the kernels ran inside the graph)."""
        ops.bump_param_generation()
        self.step_num += 1

    @torch.no_grad()
    def step(self, _captured=False):
        """clip -> Adam -> advance the schedule.  Returns the device tensor holding the pre-clip gradient norm.
        _captured: called while a CUDA graph is being recorded -- launches only, no host-side state changes."""
        active = [(i, p) for i, p in enumerate(self.params) if p.grad is not None]
        if not active:
            self.step_num += 1
            return self.grad_norm
        self._ever_updated.update(i for i, _ in active)
        d_desc, d_t, d_o, partial, nblk = self._build_tables(active)
        L = _lib.lib()
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        st = torch.cuda.current_stream(dev).cuda_stream
        import ctypes
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        if self.clip is not None:
            _lib.check(L.rp_grad_norm_multi(P(d_desc), P(d_t), P(d_o), nblk, self.chunk, P(partial), P(self.grad_norm), dev,
                                            ctypes.c_void_p(st)), "rp_grad_norm_multi")
        if getattr(self, "_hyper_dev", None) is not None:
            if not _captured:
                self.upload_step_scalars()
            _lib.check(L.rp_adam_clip_step_multi_dev(P(d_desc), P(d_t), P(d_o), nblk, self.chunk, P(self.grad_norm),
                                                     self.clip if self.clip is not None else 0.0, P(self._hyper_dev),
                                                     self.betas[0], self.betas[1], self.eps, self.weight_decay, dev,
                                                     ctypes.c_void_p(st)), "rp_adam_clip_step_multi_dev")
            if not _captured:
                self.advance()
            return self.grad_norm
        lr = self.lr_at(self.step_num)
        _lib.check(L.rp_adam_clip_step_multi(P(d_desc), P(d_t), P(d_o), nblk, self.chunk, P(self.grad_norm),
                                             self.clip if self.clip is not None else 0.0, lr, self.betas[0], self.betas[1],
                                             self.eps, self.weight_decay, self.step_num + 1, dev, ctypes.c_void_p(st)),
                   "rp_adam_clip_step_multi")
        ops.bump_param_generation()          # parameters were written through raw pointers: derived-data caches are stale
        self.step_num += 1
        return self.grad_norm

    # ---- checkpoints: the layouts train.py:191-194 stores under "optimizer" and "scheduler" ---------------------
    def _updated(self):
        """Indices of the parameters torch.optim.Adam would hold state for (those that have received a gradient)."""
        return sorted(self._ever_updated)

    def state_dict(self):
        """torch.optim.Adam.state_dict() layout: {"state": {index: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]},
        so `torch.optim.Adam(model.parameters(), ...).load_state_dict(sd)` in the reference's train.py (:96) accepts it and
        `load_state_dict` here accepts what the reference saved (:191-194).  The param group carries the keys OneCycleLR adds
        (initial_lr / max_lr / min_lr)."""
        state = {}
        if self.step_num > 0:
            for i in self._updated():
                state[i] = {"step": torch.tensor(float(self.step_num)), "exp_avg": self.exp_avg[i].clone(),
                            "exp_avg_sq": self.exp_avg_sq[i].clone()}
        initial_lr = self.max_lr / self.div_factor
        group = {"lr": self.lr_at(min(self.step_num, self.total_steps)), "betas": self.betas, "eps": self.eps,
                 "weight_decay": self.weight_decay, "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "decoupled_weight_decay": False, "initial_lr": initial_lr,
                 "max_lr": self.max_lr, "min_lr": initial_lr / self.final_div_factor, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group],
                "relpose": {"clip": self.clip, "pct_start": self.pct_start, "total_steps": self.total_steps,
                            "div_factor": self.div_factor, "final_div_factor": self.final_div_factor}}

    def scheduler_state_dict(self):
        """What `OneCycleLR(...).state_dict()` holds after `step_num` scheduler steps (train.py:193), produced by torch's own
        class on a dummy optimizer so that the key set matches the installed torch version."""
        dummy = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=self.max_lr)
        sch = torch.optim.lr_scheduler.OneCycleLR(dummy, self.max_lr, self.total_steps, pct_start=self.pct_start,
                                                  div_factor=self.div_factor, final_div_factor=self.final_div_factor,
                                                  cycle_momentum=False)
        sd = sch.state_dict()
        sd["last_epoch"] = self.step_num
        sd["_step_count"] = self.step_num + 1
        sd["_last_lr"] = [self.lr_at(min(self.step_num, self.total_steps))]
        return sd

    def load_scheduler_state_dict(self, sd):
        self.step_num = int(sd["last_epoch"])
        if "total_steps" in sd and int(sd["total_steps"]) != self.total_steps:
            raise ValueError(f"scheduler total_steps {sd['total_steps']} != {self.total_steps}")

    def load_state_dict(self, sd):
        """Accepts torch.optim.Adam's layout (a checkpoint written by the reference or by state_dict() above) and the
        private layout of earlier versions of this class ({"step", "exp_avg", "exp_avg_sq", "hyper"}).  Lengths, shapes
        and the group's hyper-parameters are validated: a mismatched parameter list is an error, not a silent partial load."""
        if "param_groups" in sd:
            groups = sd["param_groups"]
            order = [i for g in groups for i in g["params"]]
            if len(order) != len(self.params):
                raise ValueError(f"optimizer state has {len(order)} parameters, this optimizer has {len(self.params)}")
            g0 = groups[0]
            for key, mine in (("betas", self.betas), ("eps", self.eps), ("weight_decay", self.weight_decay)):
                if key in g0 and tuple(np.atleast_1d(g0[key]).tolist()) != tuple(np.atleast_1d(mine).tolist()):
                    raise ValueError(f"optimizer state was written with {key}={g0[key]}, this optimizer uses {mine}")
            steps = set()
            pos = {pid: k for k, pid in enumerate(order)}
            for pid, st in sd["state"].items():
                k = pos[pid]
                if tuple(st["exp_avg"].shape) != tuple(self.params[k].shape):
                    raise ValueError(f"state of parameter {pid}: shape {tuple(st['exp_avg'].shape)} != {tuple(self.params[k].shape)}")
                self.exp_avg[k].copy_(st["exp_avg"])
                self.exp_avg_sq[k].copy_(st["exp_avg_sq"])
                self._ever_updated.add(k)
                steps.add(int(float(st["step"])))
            if len(steps) > 1:
                raise ValueError(f"parameters disagree on the step count: {sorted(steps)}")
            if steps:
                self.step_num = steps.pop()
            return
        if len(sd["exp_avg"]) != len(self.params) or len(sd["exp_avg_sq"]) != len(self.params):
            raise ValueError(f"optimizer state has {len(sd['exp_avg'])} parameters, this optimizer has {len(self.params)}")
        hyper = sd.get("hyper", {})
        for key in ("max_lr", "total_steps", "pct_start", "weight_decay", "eps"):
            if key in hyper and float(hyper[key]) != float(getattr(self, key)):
                raise ValueError(f"optimizer state was written with {key}={hyper[key]}, this optimizer uses {getattr(self, key)}")
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            if tuple(dst.shape) != tuple(src.shape):
                raise ValueError(f"moment shape {tuple(src.shape)} != {tuple(dst.shape)}")
        self.step_num = int(sd["step"])
        if self.step_num > 0:
            self._ever_updated.update(i for i, p in enumerate(self.params) if p.requires_grad)
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            dst.copy_(src)
