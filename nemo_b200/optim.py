"""B200FusedAdamW — the optimizer step of the Conformer recipe as ONE multi-tensor sm_100a kernel over the
flat gradient buffer the backward kernels (and the gradient all-reduce) already wrote:

    optim: {name: adamw, lr: 2.0, betas: [0.9, 0.98], weight_decay: 1e-3,
            sched: {name: NoamAnnealing, d_model: ${model.encoder.d_model}, warmup_steps: 10000, min_lr: 1e-6}}
    (examples/asr/conf/conformer/conformer_ctc_bpe.yaml:175-192; built by nemo/core/classes/modelPT.py:650-823 as
     torch.optim.AdamW + nemo/core/optim/lr_scheduler.py:518-575)

Arithmetic = torch.optim.AdamW (decoupled weight decay, bias corrections) with the NoamAnnealing rate of the matching
scheduler step, evaluated on the device from a device-resident step counter: no host scalar enters the kernel, so the
step is capturable in a CUDA graph (nemo_b200.graph.CapturedStep).  exp_avg / exp_avg_sq are flat fp32 buffers with the
layout of `FlatGradients.flat`.  There is no CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .parallel import FlatGradients


class B200FusedAdamW:
    def __init__(self, flat_grads: FlatGradients, lr: float = 2.0, betas=(0.9, 0.98), eps: float = 1e-8,
                 weight_decay: float = 1e-3, noam_d_model: Optional[int] = None, noam_warmup_steps: int = 10000,
                 noam_min_lr: float = 1e-6, grad_scale: float = 1.0, zero_grads_in_step: bool = True):
        fg = self.fg = flat_grads
        dev = fg.flat.device
        if dev.type != "cuda":
            raise RuntimeError("B200FusedAdamW runs on CUDA tensors only (no CPU fallback)")
        if fg.flat.dtype != torch.float32:
            raise NotImplementedError("B200FusedAdamW reads fp32 flat gradients")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.noam = (int(noam_d_model), int(noam_warmup_steps), float(noam_min_lr)) if noam_d_model else None
        self.grad_scale = float(grad_scale)
        # the kernel zeroes each gradient element right after consuming it: `FlatGradients.zero()` / `zero_grad()` after a
        # step are then free (no separate fill pass over the flat buffer: 2.5 GB at Conformer-XL)
        self.zero_grads_in_step = bool(zero_grads_in_step)
        self.exp_avg = torch.zeros_like(fg.flat)
        self.exp_avg_sq = torch.zeros_like(fg.flat)
        self.state = torch.zeros(1, dtype=torch.float64, device=dev)   # optimizer step count (device-resident)
        self.scalars = torch.zeros(4, dtype=torch.float32, device=dev)  # lr, lr/bc1, 1/sqrt(bc2), 1 - lr*wd of the last step
        self._table = None

    # ---- parameter table (rebuilt when a parameter's storage moved, e.g. after .to()) ---------------------------
    def _ensure_table(self):
        ptrs = tuple(p.data_ptr() for p in self.fg._params.values())
        if self._table is not None and self._table[0] == ptrs:
            return self._table[1:]
        entries, numels = [], []
        for n, p in self.fg._params.items():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise NotImplementedError(f"B200FusedAdamW: parameter {n} must be contiguous fp32")
            lo, hi = self.fg.offsets[n]
            entries.append(L.TensorEntry(p.data_ptr(), lo, hi - lo, 0, 0))
            numels.append(hi - lo)
        tab, blocks, nb = L.upload_table(entries, numels, self.fg.flat.device)
        self._table = (ptrs, tab, blocks, nb)
        return tab, blocks, nb

    def step(self):
        tab, blocks, nb = self._ensure_table()
        h = L.AdamWHyper(self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                         self.noam[0] if self.noam else 0, self.noam[1] if self.noam else 0, self.noam[2] if self.noam else 0.0)
        L.check(L.lib.b2c_adamw_step(tab.data_ptr(), blocks.data_ptr(), nb, self.fg.flat.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), self.state.data_ptr(), self.scalars.data_ptr(), C.byref(h),
                                     self.grad_scale, 1 if self.zero_grads_in_step else 0, L.stream()), "b2c_adamw_step")
        if self.zero_grads_in_step:
            self.fg.mark_zeroed()
        # the weights changed through raw pointers: tensor version counters did not move
        enc = self.fg.module
        if hasattr(enc, "invalidate_weight_cache"):
            enc.invalidate_weight_cache()

    def zero_grad(self, set_to_none: bool = False):
        if not (self.zero_grads_in_step and self.fg.is_zeroed()):
            self.fg.zero()

    @property
    def last_lr(self) -> float:
        """learning rate of the most recent step (host sync: for logging only)"""
        return float(self.scalars[0].item())

    def state_dict(self):
        return {"step": self.state.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "names": list(self.fg.names), "hyper": dict(lr=self.lr, betas=self.betas, eps=self.eps,
                                                            weight_decay=self.weight_decay, noam=self.noam)}

    def load_state_dict(self, sd):
        if list(sd["names"]) != list(self.fg.names):
            raise ValueError("B200FusedAdamW.load_state_dict: parameter layout differs")
        self.state.copy_(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])


def noam_lr(base_lr: float, d_model: int, warmup_steps: int, min_lr: float, opt_step: int) -> float:
    """Host restatement of the rate B200FusedAdamW applies at optimizer step `opt_step` (1-based): NoamAnnealing at
    scheduler step max(1, opt_step - 1) (lr_scheduler.py:555-575).  Used by the tests and for logging."""
    step = max(1, opt_step - 1)
    norm = d_model ** -0.5
    mult = norm * (min(step ** -0.5, step * warmup_steps ** -1.5) if warmup_steps > 0 else step ** -0.5)
    out = base_lr * mult
    if step > warmup_steps:
        out = max(out, min_lr)
    return out
