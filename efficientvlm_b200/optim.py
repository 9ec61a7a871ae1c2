"""Optimizer + data-parallel gradient exchange over flat fp32 arenas.

Replaces, for the hot path, the reference's
  * `optim.py:23-69` create_optimizer (4 name-based groups: decay / no-decay x lr / lr*lr_mult, HF AdamW eps 1e-8,
    betas (0.9, 0.98)) and `optim.py:4-21` create_L0_optimizer (gate optimizer + NEGATIVE-lr Lagrangian optimizer);
  * `accelerators/apex_ddp_accelerator.py:74-101`: parameter broadcast, `Apex_DDP(delay_allreduce=True)` mean-allreduce of all
    gradients after backward, and global-norm clipping.

Design: parameters, gradients and both Adam moments of every group live in ONE contiguous fp32 arena each (p.data / p.grad
become views), so the gradient allreduce is a single NCCL call on the arena (NVLink/NVSwitch, no per-tensor buckets), the
global grad norm is one reduction launch, and the AdamW update is one launch per group — instead of the hundreds of
per-parameter kernels HF AdamW issues.  Parameter NAMES still decide the grouping exactly like the reference.
"""
import os

import torch
import torch.distributed as dist

from . import kernels as K
from . import ops

ARENA_ALIGN = 64  # floats (256 bytes)
NO_DECAY = {"bias", "LayerNorm.bias", "LayerNorm.weight", "norm.bias", "norm.weight", "norm1.bias", "norm1.weight", "norm2.bias",
            "norm2.weight"}


def group_parameters(model, lr, weight_decay, lr_mult=1):
    """optim.py:23-65: the four (weight_decay, lr) groups chosen by substring of the parameter name / init_params."""
    groups = [{"params": [], "names": [], "weight_decay": weight_decay, "lr": lr}, {"params": [], "names": [], "weight_decay": 0.0, "lr": lr},
              {"params": [], "names": [], "weight_decay": weight_decay, "lr": lr * lr_mult},
              {"params": [], "names": [], "weight_decay": 0.0, "lr": lr * lr_mult}]
    large_lr = model.init_params if hasattr(model, "init_params") else {}
    seen = set()
    for n, p in model.named_parameters():
        if not p.requires_grad or id(p) in seen:
            continue
        seen.add(id(p))
        nd = any(s in n for s in NO_DECAY)
        gi = (3 if n in large_lr else 1) if nd else (2 if n in large_lr else 0)
        groups[gi]["params"].append(p)
        groups[gi]["names"].append(n)
    return groups


OVERLAP_STAGES = int(os.environ.get("EVLM_OVERLAP_STAGES", "2"))     # profiling knob: 1 = only the exchange at the vision tower's output


class FlatAdamW:
    """HF-AdamW semantics (decoupled decay AFTER the Adam update, bias correction) on flat arenas, with the data-parallel
    gradient mean-allreduce and global-norm clip folded into `step()`."""

    def __init__(self, param_groups, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, process_group=None, clip_grad_norm=0.0):
        self.betas, self.eps = betas, eps
        self.clip_grad_norm = clip_grad_norm
        self.process_group = process_group
        self.param_groups = []
        self.state_step = 0
        self._orphans = set()          # ids of parameters a later FlatAdamW took over
        for g in param_groups:
            params = [p for p in g["params"]]
            if not params:
                continue
            dev = params[0].device
            # every parameter starts on a 256-byte boundary (the kernels use 16-byte vector / TMA accesses on parameters);
            # the zero padding in between has zero gradients and stays zero under AdamW
            offsets, n = [], 0
            for p in params:
                offsets.append(n)
                n += (p.numel() + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN
            arena_p = torch.zeros(n, dtype=torch.float32, device=dev)
            arena_g = torch.zeros(n, dtype=torch.float32, device=dev)
            for p, off in zip(params, offsets):
                k = p.numel()
                arena_p[off:off + k].copy_(p.data.reshape(-1))
                p.data = arena_p[off:off + k].view(p.shape)
                p.grad = arena_g[off:off + k].view(p.shape)
                p._evlm_main_grad = p.grad          # ops.py accumulates weight / bias gradients straight into this view
                # ONE owner per parameter.  The reference's main AdamW also lists the l0_module gates (they are sub-module
                # parameters, optim.py:49-63) and create_L0_optimizer lists them again (quirk Q11); here the optimizer built LAST
                # owns the parameter (the L0 optimizers, as the drivers build them after the main one) and the earlier one
                # leaves its orphaned arena slot alone: zero gradient, never re-bound.
                prev = getattr(p, "_evlm_owner", None)
                if prev is not None and prev is not self:
                    prev._orphans.add(id(p))
                p._evlm_owner = self
            self.param_groups.append({"params": params, "offsets": offsets, "names": list(g.get("names", [])), "size": n,
                                      "lr": g.get("lr", lr), "initial_lr": g.get("lr", lr),
                                      "weight_decay": g.get("weight_decay", 0.0), "p": arena_p, "g": arena_g,
                                      "m": torch.zeros_like(arena_p), "v": torch.zeros_like(arena_p)})
        dev = self.param_groups[0]["p"].device
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._coef = torch.ones(1, dtype=torch.float32, device=dev)
        # per-step scalars (step_size, lr*wd per group) live in device memory so that a captured step can be replayed with
        # a moving learning rate / bias correction (graph.py): the host refreshes them with one tiny launch per step
        self._hyper = torch.zeros(2 * len(self.param_groups), dtype=torch.float32, device=dev)
        ops.invalidate_weight_cache()

    # -- torch.optim.Optimizer-like surface used by the drivers
    def zero_grad(self, set_to_none=False):
        for g in self.param_groups:
            g["g"].zero_()
            # autograd may have replaced .grad (e.g. first backward after set_to_none): re-attach the arena views
            for p, off in zip(g["params"], g["offsets"]):
                k = p.numel()
                if id(p) in self._orphans:
                    continue
                if p.grad is None or p.grad.data_ptr() != g["g"].data_ptr() + 4 * off:
                    p.grad = g["g"][off:off + k].view(p.shape)
                    p._evlm_main_grad = p.grad

    def _gather_stray_grads(self):
        """If autograd re-bound p.grad to a fresh tensor, fold it back into the arena (keeps `loss.backward()` drop-in)."""
        for g in self.param_groups:
            for p, off in zip(g["params"], g["offsets"]):
                k = p.numel()
                view = g["g"][off:off + k].view(p.shape)
                if id(p) in self._orphans:
                    continue
                if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                    p.grad = view

    def broadcast_parameters(self, src=0):
        """apex_ddp_accelerator.py:74-77 as one broadcast per arena."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            for g in self.param_groups:
                dist.broadcast(g["p"], src, group=self.process_group)
            ops.invalidate_weight_cache()

    def _reduce(self, t):
        if t.numel() == 0:
            return
        if t.is_cuda:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.process_group)      # NCCL
        else:                                                                        # gloo (CPU tests): no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.process_group)
            t.div_(dist.get_world_size(self.process_group))

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1

    def enable_overlap(self, model, first_prefix="vision_encoder."):
        """Overlap the gradient exchange with the backward (the reference's `Eff_*` drivers get this from torch DDP's 25 MB buckets,
        Eff_VQA.py:325-328; apex DDP with delay_allreduce=True, apex_ddp_accelerator.py:79, does not).  Every arena lists the
        parameters in `named_parameters()` order, so the vision tower (`first_prefix`) is a PREFIX of each arena — and the last part
        of the model whose gradients complete.  When autograd reaches the output of `model.vision_encoder` (eff_vit.py) the SUFFIX
        of every arena is final: its all-reduce starts on a side stream and runs under the vision tower's backward.  A second stage
        does the same inside the tower: its upper half (layers depth/2 .. and the final LayerNorm) is a contiguous run at the end of
        the prefix and leaves when autograd has passed layer depth/2 (hook in eff_vit.CLIPEncoder); `step()` then only exchanges what
        is left — the lower half of the vision tower.  Same values as the single blocking exchange (tests/test_gpu_distributed.py)."""
        vision = getattr(model, first_prefix.rstrip("."))
        layers = getattr(getattr(vision, "encoder", None), "layers", None)
        mid = len(layers) // 2 if layers is not None and len(layers) >= 2 else None
        mid_prefixes = tuple("%sencoder.layers.%d." % (first_prefix, i) for i in range(mid, len(layers))) + (first_prefix + "post_layernorm.",) \
            if mid is not None else ()
        for g in self.param_groups:
            k = 0
            names = g["names"]
            while k < len(names) and names[k].startswith(first_prefix):
                k += 1
            if any(n.startswith(first_prefix) for n in names[k:]) or len(names) != len(g["params"]):
                g["split"] = g["split2"] = g["size"]    # not a clean prefix (or unnamed parameters): nothing leaves early
                continue
            g["split"] = g["offsets"][k] if k < len(names) else g["size"]
            # second stage: the vision tower's upper half (layers mid.. and the final LayerNorm) is a contiguous run at the END of the
            # vision prefix; it is final as soon as autograd has passed layer `mid`
            j = k
            while j > 0 and names[j - 1].startswith(mid_prefixes):
                j -= 1
            clean = mid is not None and not any(n.startswith(mid_prefixes) for n in names[:j])
            g["split2"] = (g["offsets"][j] if j < len(names) else g["size"]) if clean else g["split"]
        self._early = {"stream": None, "done": False, "done2": False}
        hooks = vision.__dict__.setdefault("_evlm_grad_ready", [])
        if self._early_allreduce not in hooks:
            hooks.append(self._early_allreduce)
        if OVERLAP_STAGES >= 2 and mid is not None and any(g["split2"] < g["split"] for g in self.param_groups):
            slot = vision.encoder.__dict__.setdefault("_evlm_grad_mid", [mid, []])
            if self._mid_allreduce not in slot[1]:
                slot[1].append(self._mid_allreduce)

    def _side_reduce(self, ranges):
        e = self._early
        if self.param_groups[0]["g"].is_cuda:
            if e["stream"] is None:
                e["stream"] = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            e["stream"].wait_stream(cur)                # every gradient kernel issued so far precedes the exchange
            with torch.cuda.stream(e["stream"]):
                for t in ranges:
                    self._reduce(t)
        else:
            for t in ranges:
                self._reduce(t)

    def _early_allreduce(self):
        e = getattr(self, "_early", None)
        if e is None or e["done"] or not self._distributed() or not ops.accumulating_into_main_grads():
            return
        self._side_reduce([g["g"][g["split"]:] for g in self.param_groups if g["split"] < g["size"]])
        e["done"] = True

    def _mid_allreduce(self):
        e = getattr(self, "_early", None)
        if e is None or e["done2"] or not e["done"] or not self._distributed() or not ops.accumulating_into_main_grads():
            return
        self._side_reduce([g["g"][g["split2"]:g["split"]] for g in self.param_groups if g["split2"] < g["split"]])
        e["done2"] = True

    def allreduce_gradients(self):
        """Mean-allreduce of every gradient: ONE NCCL call per arena (4 per step) over NVLink/NVSwitch — or, after
        `enable_overlap()`, only of the part that was not exchanged under the backward."""
        if not self._distributed():
            return
        e = getattr(self, "_early", None)
        if e is not None and e["done"]:
            if e["stream"] is not None:
                torch.cuda.current_stream().wait_stream(e["stream"])
            for g in self.param_groups:
                end = g["split2"] if e["done2"] else g["split"]
                if end > 0:
                    self._reduce(g["g"][:end])
            e["done"] = e["done2"] = False
            return
        for g in self.param_groups:
            self._reduce(g["g"])

    def invalidate_shadows(self):
        for g in self.param_groups:
            ops.invalidate_weight_cache(g["params"])

    def _refresh_shadows(self):
        """bf16 shadows of this optimizer's weights, rebuilt by one multi-tensor cast right after the update."""
        if self.param_groups[0]["p"].is_cuda:
            if getattr(self, "_all_params", None) is None:
                self._all_params = [p for g in self.param_groups for p in g["params"]]
            ops.refresh_weight_shadows(self._all_params)

    def begin_step(self):
        """Host half of step(): advance t and push this step's (step_size, lr*wd) per group to the device.  `step()` calls it
        itself in eager mode; a captured step is replayed as `begin_step(); graph.replay()` (graph.GraphedTrainStep)."""
        self.state_step += 1
        b1, b2 = self.betas
        t = self.state_step
        corr = (1.0 - b2 ** t) ** 0.5 / (1.0 - b1 ** t)
        vals = []
        for g in self.param_groups:
            vals += [float(g["lr"]) * corr, float(g["lr"]) * float(g["weight_decay"])]
        K.store_f32(self._hyper, vals)

    def step(self, allreduce=True):
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        self._gather_stray_grads()
        if allreduce:
            self.allreduce_gradients()
        if not capturing:
            self.begin_step()
        scale = None
        if self.clip_grad_norm and self.clip_grad_norm > 0:
            self._sumsq.zero_()
            for g in self.param_groups:
                K.sumsq(g["g"], self._sumsq)
            K.clip_coef(self._sumsq, float(self.clip_grad_norm), self._coef)
            scale = self._coef
        K.adamw_step([dict(p=g["p"], g=g["g"], m=g["m"], v=g["v"], p_bf16=None, lr=float(g["lr"]), beta1=self.betas[0],
                           beta2=self.betas[1], eps=self.eps, weight_decay=float(g["weight_decay"]), step=max(1, self.state_step))
                      for g in self.param_groups], scale, self._hyper)
        self.invalidate_shadows()
        self._refresh_shadows()

    def grad_norm(self):
        """sqrt of the last global sum of squares (device tensor; no host sync)."""
        return self._sumsq.sqrt()

    # -- checkpointing: what the drivers put into `save_obj` / `training_states` (GeneralDistill.py:422,430; Eff_VQA.py:398,405;
    #    Eff_Retrieval.py:537) and read back on resume (GeneralDistill.py:517)
    def state_dict(self):
        """`torch.optim.Optimizer.state_dict()` layout: {"state": {index: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups":
        [{..., "params": [indices]}]} with parameters numbered in group order, so a checkpoint written by the reference's HF AdamW
        over the same groups loads here and the other way round.  The moment tensors are VIEWS of the arenas (no copy, like torch)."""
        state, groups, idx = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p, off in zip(g["params"], g["offsets"]):
                k = p.numel()
                if self.state_step > 0:
                    state[idx] = {"step": self.state_step, "exp_avg": g["m"][off:off + k].view(p.shape),
                                  "exp_avg_sq": g["v"][off:off + k].view(p.shape)}
                ids.append(idx)
                idx += 1
            groups.append({"lr": g["lr"], "initial_lr": g["initial_lr"], "weight_decay": g["weight_decay"], "betas": tuple(self.betas),
                           "eps": self.eps, "correct_bias": True, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        """Copies the moments INTO the existing arenas (views held by captured graphs and by `p.grad` stay valid).  One global step
        count: the largest per-parameter `step` of the checkpoint (HF keeps one per parameter; they only differ for parameters that
        never received a gradient — see the class note in DESIGN.md 4, optimizer deviations)."""
        saved = sd["param_groups"]
        if len(saved) != len(self.param_groups):
            raise ValueError("optimizer state has %d parameter groups, this optimizer %d" % (len(saved), len(self.param_groups)))
        step = 0
        for g, sg in zip(self.param_groups, saved):
            if len(sg["params"]) != len(g["params"]):
                raise ValueError("parameter group size mismatch: %d vs %d" % (len(sg["params"]), len(g["params"])))
            g["lr"] = sg.get("lr", g["lr"])
            g["initial_lr"] = sg.get("initial_lr", g["initial_lr"])
            g["weight_decay"] = sg.get("weight_decay", g["weight_decay"])
            for p, off, i in zip(g["params"], g["offsets"], sg["params"]):
                st = sd["state"].get(i, sd["state"].get(str(i)))
                k = p.numel()
                if st is None:
                    g["m"][off:off + k].zero_()
                    g["v"][off:off + k].zero_()
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError("optimizer state %d has shape %s, parameter %s" % (i, tuple(st["exp_avg"].shape), tuple(p.shape)))
                g["m"][off:off + k].copy_(st["exp_avg"].reshape(-1))
                g["v"][off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                step = max(step, int(st["step"]))
        if saved and "betas" in saved[0]:
            self.betas, self.eps = tuple(saved[0]["betas"]), saved[0].get("eps", self.eps)
        self.state_step = step


def create_optimizer(args, model, clip_grad_norm=0.0, process_group=None):
    """Drop-in for optim.py:create_optimizer (args has .lr, .weight_decay, optional .lr_mult)."""
    get = (lambda k, d=None: args.get(k, d)) if isinstance(args, dict) else (lambda k, d=None: getattr(args, k, d))
    groups = group_parameters(model, get("lr"), get("weight_decay"), get("lr_mult", 1) or 1)
    return FlatAdamW(groups, lr=get("lr"), betas=(0.9, 0.98), eps=1e-8, clip_grad_norm=clip_grad_norm, process_group=process_group)


def create_L0_optimizer(args, l0_module):
    """optim.py:4-21: gate optimizer (lr = reg_learning_rate) and Lagrangian optimizer (lr = -reg_learning_rate: ascent)."""
    rl = args["reg_learning_rate"] if isinstance(args, dict) else args.reg_learning_rate
    l0 = FlatAdamW([{"params": [p for n, p in l0_module.named_parameters() if "lambda" not in n], "weight_decay": 0.0, "lr": rl}],
                   eps=1e-8, betas=(0.9, 0.98))
    lag = FlatAdamW([{"params": [p for n, p in l0_module.named_parameters() if "lambda" in n], "weight_decay": 0.0, "lr": -rl}],
                    eps=1e-8, betas=(0.9, 0.98))
    return l0, lag


class LinearWarmupDecay:
    """scheduler.py:17-24 (`LambdaLR` with linear warm-up then linear decay) for an optimizer that is not a `torch.optim.Optimizer`.
    Keeps the parts of the `LambdaLR` surface the reference's drivers touch: `.optimizer`, `.last_epoch`, `.base_lrs`,
    `get_last_lr()`, `state_dict()` / `load_state_dict()` with `LambdaLR`'s keys, and re-initialisation through
    `scheduler.__init__(optimizer, lr_lambda, last_epoch=-1)` (Captioning_pretrain.py:32-50, NLVR_pretrain.py:175)."""

    def __init__(self, optimizer, num_training_steps, num_warmup_steps=None, last_epoch=-1):
        self.optimizer = self.opt = optimizer
        if callable(num_training_steps):                 # LambdaLR-style: (optimizer, lr_lambda, last_epoch=-1)
            self.lr_lambda = num_training_steps
            self.total = self.warm = None
        else:
            if isinstance(num_warmup_steps, float):
                assert 0 <= num_warmup_steps < 1
                num_warmup_steps = int(num_training_steps * num_warmup_steps)
            self.total, self.warm = num_training_steps, num_warmup_steps
            self.lr_lambda = None
        self.base_lrs = [g["initial_lr"] for g in optimizer.param_groups]
        self.last_epoch = last_epoch
        self._step_count = 0
        self.step()

    @property
    def last_step(self):
        return self.last_epoch

    def factor(self, s):
        if self.lr_lambda is not None:
            return self.lr_lambda(s)
        if s < self.warm:
            return float(s) / float(max(1, self.warm))
        return max(0.0, float(self.total - s) / float(max(1, self.total - self.warm)))

    def step(self):
        self.last_epoch += 1
        self._step_count += 1
        f = self.factor(self.last_epoch)
        for g, base in zip(self.optimizer.param_groups, self.base_lrs):
            g["lr"] = base * f
        self._last_lr = [g["lr"] for g in self.optimizer.param_groups]

    def get_last_lr(self):
        return self._last_lr

    def state_dict(self):
        """What the drivers checkpoint as `lr_scheduler` (GeneralDistill.py:423): `LambdaLR.state_dict()`'s keys (the lambda itself is
        not saved there either) plus the two step counts of the closed-form schedule."""
        return {"last_epoch": self.last_epoch, "_step_count": self._step_count, "base_lrs": list(self.base_lrs),
                "_last_lr": list(self._last_lr), "lr_lambdas": [None], "total": self.total, "warm": self.warm}

    def load_state_dict(self, state):
        if state.get("total") is not None:
            self.total, self.warm, self.lr_lambda = state["total"], state["warm"], None
        if "base_lrs" in state:
            self.base_lrs = list(state["base_lrs"])
        last = state["last_epoch"] if "last_epoch" in state else state["last_step"]       # "last_step": round-1 checkpoints
        self._step_count = state.get("_step_count", last + 1) - 1
        self.last_epoch = last - 1
        self.step()
