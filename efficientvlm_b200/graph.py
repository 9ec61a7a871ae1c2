"""Whole-step CUDA graph for the fixed-shape training step of the hot path.

The reference's train loop (Train.py:60-110 -> accelerators/apex_ddp_accelerator.py:84-101) re-issues every kernel of
forward, backward and the optimizer from Python each iteration.  The B200 step is ~1300 launches of 5-300 us each, so
the Python/ctypes enqueue cost (~45 ms per step) sits right under the GPU time; one captured graph per (model, batch
shape) removes it and the inter-launch gaps.  Nothing about the math changes: the graph replays exactly the launches
the eager step issues.

What a replay must be able to change lives in device memory:
  * the batch            -> static input tensors, refreshed with `copy_` (H2D from pinned memory) before the replay;
  * dropout seeds        -> by-value seeds are baked into the graph; its first node advances the device seed offset every
                            dropout site adds (evlm_rng_bind / evlm_rng_advance, include/evlm.h);
  * lr / bias correction -> FlatAdamW.begin_step() pushes (step_size, lr*wd) per group with one eager launch before the
                            replay and the captured AdamW launches read them (evlm_adamw_step_dev).
Host-side Python that runs inside `step_fn` (schedulers, counters, logging) is NOT replayed: keep it outside, or pass
it as `host_fn`, which runs after every replay.
"""
import torch

from . import kernels as K
from . import ops

_rng_state = {}

RNG_STRIDE = 0x9E3779B97F4A7C15      # odd: seed + k*stride never repeats within 2^64 replays


def rng_state(device):
    """The bound device seed-offset word (created and bound on first use; one per process/device)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _rng_state.get(key)
    if st is None:
        st = torch.zeros(1, dtype=torch.int64, device=device)
        K.rng_bind(st)
        _rng_state[key] = st
    return st


def reset_rng_offset():
    """Back to offset 0 (explicit-seed parity tests replay the masks of a given seed on the host)."""
    for st in _rng_state.values():
        K.rng_advance(st, 0, set_value=True)


class GraphedTrainStep:
    """Capture `step_fn(*static_inputs)` — forward, loss, backward, `optimizer.step()`, `optimizer.zero_grad()` — once and
    replay it.  `step_fn` returns a tensor or a tuple of tensors (losses); replays return the same static tensors.

    optimizers: the FlatAdamW instances stepped inside `step_fn` (their `begin_step()` runs eagerly before every replay).
    warmup: eager steps run on a side stream before capture (allocator, bf16 shadows, lazy initialisation settle).
    """

    def __init__(self, step_fn, example_inputs, optimizers=(), warmup=2, host_fn=None):
        self.step_fn, self.host_fn = step_fn, host_fn
        self.optimizers = list(optimizers)
        dev = example_inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (no CPU path)")
        self.static_inputs = [t.clone() for t in example_inputs]
        self.rng = rng_state(dev)
        # the parameters' AccumulateGrad nodes were created on the default stream; warm-up and capture run on side streams
        # (torch orders them with events, which is exactly what the capture needs)
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                out = step_fn(*self.static_inputs)
                if host_fn is not None:
                    host_fn()
            del out
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = K.launch_count()
        with torch.cuda.graph(self.graph):
            K.rng_advance(self.rng, RNG_STRIDE)
            self.static_outputs = step_fn(*self.static_inputs)
        # the capture pass itself executed nothing: optimizer state / parameters are exactly as after the warm-up steps,
        # but bf16 shadows the capture pass "refreshed" were not actually written
        ops.invalidate_weight_cache()
        self.captured_launches = K.launch_count() - n0     # libevlm kernel nodes one replay executes
        self.replays = 0

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        for opt in self.optimizers:
            opt.begin_step()
        self.graph.replay()
        self.replays += 1
        for opt in self.optimizers:          # parameters moved behind the eager shadow cache's back
            opt.invalidate_shadows()
        if self.host_fn is not None:
            self.host_fn()
        return self.static_outputs
