"""SpecAugment for the training step between the mel front-end and the encoder (SURVEY.md §8f "next"):
`nemo/collections/asr/parts/submodules/spectr_augment.py:25-215` (class SpecAugment) as one CUDA pass.

The random span boundaries are drawn exactly like the reference does — `torch.rand` on the input's device in the
vectorised mode (time masks first, then frequency masks: width draw, then start draw), Python's `random.Random.randint`
per utterance in the legacy mode — so that, given the same RNG state, the masks are identical to the reference's; the
masking itself (the reference: mask tensors of shape [B, n, axis] plus two full `masked_fill` passes) is one kernel."""
from __future__ import annotations

import random

import torch
from torch import nn

from . import _lib as L


class B200SpecAugment(nn.Module):
    """Constructor kwargs of SpecAugment (spectr_augment.py:63-92)."""

    FREQ_AXIS, TIME_AXIS = 1, 2

    def __init__(self, freq_masks=0, time_masks=0, freq_width=10, time_width=10, rng=None, mask_value=0.0,
                 use_vectorized_code=True):
        super().__init__()
        self._rng = random.Random() if rng is None else rng
        self.freq_masks, self.time_masks = freq_masks, time_masks
        self.freq_width, self.time_width = freq_width, time_width
        self.mask_value, self.use_vectorized_code = mask_value, use_vectorized_code
        if isinstance(time_width, int):
            self.adaptive_temporal_width = False
        else:
            if time_width > 1.0 or time_width < 0.0:
                raise ValueError("If `time_width` is a float value, must be in range [0, 1]")
            self.adaptive_temporal_width = True

    @classmethod
    def from_reference(cls, ref):
        return cls(freq_masks=ref.freq_masks, time_masks=ref.time_masks, freq_width=ref.freq_width,
                   time_width=ref.time_width, rng=ref._rng, mask_value=ref.mask_value,
                   use_vectorized_code=ref.use_vectorized_code)

    # ---- span boundaries ------------------------------------------------------------------------------------------
    def _bounds_vectorized(self, B, axis_length, num_masks, length, width, axis, device):
        """spectr_augment.py:176-195 — identical draws and arithmetic (float32 products truncated by .long())"""
        if axis == self.TIME_AXIS and isinstance(width, float):
            width = torch.clamp(width * length, max=axis_length).unsqueeze(1)
        mask_width = (torch.rand((B, num_masks), device=device, dtype=torch.float32) * width).long()
        mask_start = torch.rand((B, num_masks), device=device, dtype=torch.float32)
        if axis == self.TIME_AXIS:
            mask_start = mask_start * (length.unsqueeze(1) - mask_width)
        else:
            mask_start = mask_start * (axis_length - mask_width)
        mask_start = mask_start.long()
        return mask_start.to(torch.int32).contiguous(), (mask_start + mask_width).to(torch.int32).contiguous()

    def _bounds_legacy(self, B, F, lengths_cpu):
        """spectr_augment.py:102-128 — the same host draws in the same order; slices clip at the tensor edge"""
        fs, fe, ts, te = [], [], [], []
        f_hi = F - self.freq_width
        for idx in range(B):
            f0, f1, t0, t1 = [], [], [], []
            for _ in range(self.freq_masks):
                start = self._rng.randint(0, f_hi)
                width = self._rng.randint(0, self.freq_width)
                f0.append(start); f1.append(start + width)
            if self.adaptive_temporal_width:
                tmax = max(1, int(lengths_cpu[idx] * self.time_width))
            else:
                tmax = self.time_width
            t_hi = max(1, int(lengths_cpu[idx]) - tmax)
            for _ in range(self.time_masks):
                start = self._rng.randint(0, t_hi)
                width = self._rng.randint(0, tmax)
                t0.append(start); t1.append(start + width)
            fs.append(f0); fe.append(f1); ts.append(t0); te.append(t1)
        return fs, fe, ts, te

    @torch.no_grad()
    def forward(self, input_spec, length):
        if not input_spec.is_cuda:
            raise RuntimeError("B200SpecAugment runs on CUDA tensors only (no CPU fallback)")
        x = input_spec.contiguous()
        if x.dtype != torch.float32:
            raise TypeError("B200SpecAugment expects the fp32 features of the mel front-end")
        B, F, T = x.shape
        dev = x.device
        if self.use_vectorized_code:
            # the reference applies the time masks first, then the frequency masks (spectr_augment.py:135-154)
            t0, t1 = self._bounds_vectorized(B, T, self.time_masks, length, self.time_width, self.TIME_AXIS, dev)
            f0, f1 = self._bounds_vectorized(B, F, self.freq_masks, length, self.freq_width, self.FREQ_AXIS, dev)
        else:
            fs, fe, ts, te = self._bounds_legacy(B, F, length.cpu().numpy())
            mk = lambda v, n: torch.tensor(v, dtype=torch.int32, device=dev).reshape(B, n)
            f0, f1 = mk(fs, self.freq_masks), mk(fe, self.freq_masks)
            t0, t1 = mk(ts, self.time_masks), mk(te, self.time_masks)
        out = torch.empty_like(x)
        L.call("spec_augment", x, out, t0, t1, f0, f1, B, F, T, self.time_masks, self.freq_masks, float(self.mask_value))
        return out
