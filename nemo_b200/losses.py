"""B200CTCLoss — drop-in for `nemo.collections.asr.losses.ctc.CTCLoss` (losses/ctc.py:25-82): constructor
`(num_classes, zero_infinity=False, reduction='mean_batch')`, `forward(log_probs=[B,T,V+1], targets=[B,U], input_lengths=[B],
target_lengths=[B])`, blank = num_classes, reductions none / sum / mean / mean_batch / mean_volume.  The alpha / beta
recursions run in csrc/ctc.cu (one CTA per utterance) directly on the decoder's [B,T,V+1] layout; no CPU fallback."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L

F32 = torch.float32


class _CTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank, zero_infinity):
        if not log_probs.is_cuda:
            raise RuntimeError("B200CTCLoss runs on CUDA tensors only (no CPU fallback)")
        B, T, V1 = log_probs.shape
        dev = log_probs.device
        lp = log_probs.detach().to(F32).contiguous()
        tg = targets.to(device=dev, dtype=torch.int64).contiguous()
        if tg.dim() != 2:
            raise NotImplementedError("B200CTCLoss takes padded 2-D targets [B, U] (what EncDecCTCModel passes)")
        Umax = max(int(tg.shape[1]), 1)
        if tg.shape[1] == 0:
            tg = torch.zeros(B, 1, dtype=torch.int64, device=dev)
        il = input_lengths.to(device=dev, dtype=torch.int64).contiguous()
        tl = target_lengths.to(device=dev, dtype=torch.int64).contiguous()
        ws = torch.empty(L.lib.b2c_ctc_workspace(B, T, Umax) // 4, dtype=F32, device=dev)
        nll = torch.empty(B, dtype=F32, device=dev)
        L.call("ctc_loss_fwd", lp, tg, il, tl, ws, nll, B, T, V1, Umax, int(blank))
        ctx.save_for_backward(lp, tg, il, tl, ws, nll)
        ctx.meta = (B, T, V1, Umax, int(blank), bool(zero_infinity))
        if zero_infinity:
            return torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
        return nll

    @staticmethod
    def backward(ctx, grad_out):
        lp, tg, il, tl, ws, nll = ctx.saved_tensors
        B, T, V1, Umax, blank, zi = ctx.meta
        d = torch.empty(B, T, V1, dtype=F32, device=lp.device)
        L.call("ctc_loss_bwd", lp, tg, il, tl, ws, nll, grad_out.to(F32).contiguous(), d, B, T, V1, Umax, blank, 1 if zi else 0)
        return d, None, None, None, None, None


class B200CTCLoss(nn.Module):
    def __init__(self, num_classes, zero_infinity=False, reduction="mean_batch"):
        super().__init__()
        if reduction not in ["none", "mean", "sum", "mean_batch", "mean_volume"]:
            raise ValueError("`reduction` must be one of [mean, sum, mean_batch, mean_volume]")  # ctc.py:47-48
        self._blank, self.zero_infinity, self.config_reduction = num_classes, zero_infinity, reduction

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        target_lengths = target_lengths.long()
        losses = _CTCFn.apply(log_probs, targets, input_lengths.long(), target_lengths, self._blank, self.zero_infinity)
        r = self.config_reduction
        if r == "mean_batch":
            return losses.mean()                                   # ctc.py:58
        if r == "mean_volume":
            return losses.sum() / target_lengths.to(losses.device).sum()  # ctc.py:60
        if r == "sum":
            return losses.sum()
        if r == "mean":  # torch.nn.CTCLoss 'mean': per-utterance loss / target length, then batch mean
            return (losses / target_lengths.to(losses.device).clamp_min(1).to(losses.dtype)).mean()
        return losses
