"""Data-parallel plumbing for the encoder hot path: the utterance batch is sharded across ranks (one process per GPU,
weights replicated) and the ONLY collective is the gradient sum all-reduce — the reference's
Lightning `DDPStrategy` reducer (examples/asr/conf/conformer/conformer_ctc_bpe.yaml:201) re-thought for this engine:

* the backward kernels write every parameter gradient straight into ONE flat fp32 buffer (`p.grad` are views of it),
  so there is no bucket packing pass at all;
* as soon as a layer's backward kernels are enqueued, its contiguous slice of the flat buffer is all-reduced with
  `async_op=True` — NCCL (NVLS on the NVSwitch domain) runs on its own stream while the remaining layers' backward
  kernels keep the SMs busy;
* the reference's per-forward `all_reduce(MAX)` of the sequence length (conformer_encoder.py:770-776) and SyncBatchNorm
  (`trainer.sync_batchnorm`) are deliberately NOT reproduced (north_star: gradient all-reduce is the sole collective).
  Multi-GPU results therefore equal a reference DDP run with `sync_batchnorm=false`.

Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def grad_layout(named_params):
    """Order of the parameters inside the flat gradient buffer: `named_parameters()` order, except that inside every
    self-attention block linear_q / linear_k / linear_v weights (and biases) are adjacent — the fused q|k|v projection's
    weight gradient [3d, d] (and bias gradient [3d]) is then ONE GEMM output written in place, with no scatter-adds.
    Returns [(name, param)]; every layer's parameters stay contiguous (per-layer all-reduce spans)."""
    items = list(named_params)
    out, i = [], 0
    while i < len(items):
        n = items[i][0]
        if n.endswith("self_attn.linear_q.weight"):
            pfx = n[: -len("linear_q.weight")]
            want = [pfx + "linear_q.weight", pfx + "linear_q.bias", pfx + "linear_k.weight", pfx + "linear_k.bias",
                    pfx + "linear_v.weight", pfx + "linear_v.bias"]
            got = [x[0] for x in items[i:i + 6]]
            if got == want:
                blk = dict(items[i:i + 6])
                for k in (0, 2, 4, 1, 3, 5):
                    out.append((want[k], blk[want[k]]))
                i += 6
                continue
        out.append(items[i])
        i += 1
    return out


class FlatGradients:
    """Owns the flat gradient buffer of a module and the layer-wise overlapped all-reduce."""

    def __init__(self, module: torch.nn.Module, process_group=None, grad_dtype=torch.float32, defer: bool = False,
                 bucket_bytes: int = 96 << 20):
        self.module = module
        self.group = process_group
        self.defer = defer
        # finished layer slices are coalesced into buckets of at least this many bytes before their all-reduce is launched:
        # a handful of large NCCL calls per step instead of one per layer (each call's channel CTAs share the SMs with the
        # backward GEMMs that are still running; fewer, larger calls cost less of their throughput — profiles/r01_scaling.md)
        self.bucket_bytes = int(bucket_bytes)
        self._pending = None  # [lo, hi) of finished-but-not-yet-reduced gradients (contiguous: backward walks the layers in order)
        params = grad_layout((n, p) for n, p in module.named_parameters() if p.requires_grad)
        self.names = [n for n, _ in params]
        self.sizes = [p.numel() for _, p in params]
        dev = params[0][1].device
        self.flat = torch.zeros(sum(self.sizes), dtype=grad_dtype, device=dev)
        self.offsets = {}
        off = 0
        for (n, p), sz in zip(params, self.sizes):
            self.offsets[n] = (off, off + sz)
            p.grad = self.flat[off:off + sz].view_as(p)
            off += sz
        self._handles: List = []
        self._reduced = False  # an all-reduce has been launched on this buffer since the last zero()
        self._params = dict(params)
        module.__dict__["_flat_grads"] = self  # encoder backward writes here and calls `ready(lo, hi)`

    def bind(self):
        """Called at the start of every backward: (re-)attach `p.grad` to the flat views — `optimizer.zero_grad()` with
        the default set_to_none=True (what Lightning / NeMo call) detaches them — and refuse a second backward into a
        buffer that has already been all-reduced (that would reduce the first micro-batch twice).  Use `zero()` instead
        of `zero_grad()`; for gradient accumulation construct with `defer=True` (one all-reduce in `finish()`)."""
        self._zeroed = False
        if self._reduced and self._world() > 1:
            raise RuntimeError("FlatGradients: backward into an already all-reduced buffer; call zero() between steps "
                               "(gradient accumulation: FlatGradients(..., defer=True))")
        for n, p in self._params.items():
            lo, hi = self.offsets[n]
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + lo * self.flat.element_size():
                view = self.flat[lo:hi].view_as(p)
                if g is not None:
                    view.add_(g.to(view.dtype))  # keep what the caller accumulated outside the buffer
                p.grad = view

    # ---- called by the backward orchestration ------------------------------------------------------------------
    def view(self, name: str) -> torch.Tensor:
        lo, hi = self.offsets[name]
        return self.flat[lo:hi]

    def block(self, first: str, count: int):
        """flat view covering `count` consecutive parameters starting at `first` (None unless they are adjacent)"""
        i = self.names.index(first) if first in self.offsets else -1
        if i < 0 or i + count > len(self.names):
            return None
        lo, hi = self.offsets[self.names[i]][0], self.offsets[self.names[i + count - 1]][1]
        return self.flat[lo:hi] if hi - lo == sum(self.sizes[i:i + count]) else None

    def span(self, prefix: str):
        los = [v[0] for k, v in self.offsets.items() if k.startswith(prefix)]
        his = [v[1] for k, v in self.offsets.items() if k.startswith(prefix)]
        return (min(los), max(his)) if los else None

    def ready_prefix(self, prefix: str):
        """All gradients of the parameters under `prefix` are final (a fully frozen sub-module has no span: no-op)."""
        sp = self.span(prefix)
        if sp is not None:
            self.ready(*sp)

    def ready(self, lo: int, hi: int):
        """Gradients in flat[lo:hi] are final (their kernels are enqueued on the current stream)."""
        if self.defer or self._world() <= 1 or hi <= lo:
            return
        if self._pending is not None and (lo == self._pending[1] or hi == self._pending[0]):
            self._pending = [min(lo, self._pending[0]), max(hi, self._pending[1])]
        else:
            self._flush()
            self._pending = [lo, hi]
        if (self._pending[1] - self._pending[0]) * self.flat.element_size() >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        if self._pending is not None:
            lo, hi = self._pending
            self._pending = None
            self._reduced = True
            self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # ---- called by the training loop ---------------------------------------------------------------------------
    def _world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def zero(self):
        self.flat.zero_()
        self.mark_zeroed()

    def mark_zeroed(self):
        """the flat buffer is all zeros again (after zero(), or after an optimizer step that zeroed what it consumed)"""
        self._reduced = False
        self._pending = None
        self._zeroed = True

    def is_zeroed(self) -> bool:
        return bool(self.__dict__.get("_zeroed", False))

    def finish(self):
        """Wait for the outstanding all-reduces (the current stream waits; no host sync for NCCL)."""
        if self.defer and self._world() > 1:
            self._reduced = True
            self._handles.append(dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self._flush()
        for h in self._handles:
            h.wait()
        self._handles.clear()

    def loss_scale(self) -> float:
        """Scale the local loss by this so the summed gradients equal the global-batch mean."""
        return 1.0 / self._world()


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous, near-even shard [lo, hi) of the utterance batch for `rank` (no data-path collective)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
