"""CapturedStep — one whole training (or inference) step of the hot path as a CUDA graph.

A Conformer step is ~1300-1700 kernel launches; at small per-GPU batches (BASELINE configs[2], configs[3]: 16 and 2
utterances per GPU) the host cannot enqueue them as fast as the GPU retires them, and the step time IS the Python /
launch time.  Capturing the step (mel front-end → encoder forward → loss → backward → gradient all-reduce →
B200FusedAdamW) once and replaying it removes the host from the loop: one `cudaGraphLaunch` per step.

What makes the step capturable (all of it device-side state):
  * dropout / dither keys are `seed + *counter * golden` (csrc/common.cuh: seed_eff): the by-value seeds frozen in the graph
    are offset by a device counter that the first node of the graph advances, so every replay draws fresh masks while the
    forward and backward of one replay agree;
  * B200FusedAdamW keeps its step count on the device and evaluates the Noam rate there;
  * BatchNorm's `num_batches_tracked` is advanced by the bn_finalize kernel;
  * the bf16 operand copies of the weights are refilled inside the graph (one multi-tensor kernel);
  * activations come from the graph's private memory pool (the bump arena allocates inside the capture).

Usage:
    step = CapturedStep(lambda: train_step(static_wav, static_len))   # closure over STATIC input tensors
    static_wav.copy_(next_batch, non_blocking=True); loss = step()      # replay; `loss` is the closure's (static) result
"""
from __future__ import annotations

import torch

from . import _lib as L

_COUNTER = {}


def seed_counter(device=None) -> torch.Tensor:
    """The device-resident RNG counter of `device` (uint64 as int64 tensor), registered with the library on first use."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _COUNTER:
        with torch.cuda.device(key[1]):
            t = torch.zeros(1, dtype=torch.int64, device=dev)
            L.check(L.lib.b2c_set_seed_counter(t.data_ptr()), "b2c_set_seed_counter")
        _COUNTER[key] = t
    return _COUNTER[key]


class CapturedStep:
    """Captures `fn()` into a CUDA graph after `warmup` eager runs on a side stream (lazy allocations, cudaFuncSetAttribute
    calls, caches and NCCL channel setup must all have happened before the capture)."""

    def __init__(self, fn, warmup: int = 3, pool=None):
        if not torch.cuda.is_available():
            raise RuntimeError("CapturedStep needs a CUDA device")
        self.fn = fn
        self.counter = seed_counter()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._advance()
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.lib.b2c_launch_count()
        with torch.cuda.graph(self.graph, pool=pool):
            self._advance()
            self.result = fn()
        self.launches_per_replay = L.lib.b2c_launch_count() - n0  # kernels of this library inside one replay

    def _advance(self):
        L.check(L.lib.b2c_counter_add(self.counter.data_ptr(), 1, L.stream()), "b2c_counter_add")

    def __call__(self):
        self.graph.replay()
        return self.result
