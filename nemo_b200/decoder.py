"""B200ConvASRDecoder — drop-in for the CTC head `ConvASRDecoder` (nemo/collections/asr/modules/conv_asr.py:407-468):
1x1 Conv1d (feat_in → num_classes+1) + log_softmax, the op that turns encoder activations into the "encoder logits"
the benchmark metric names.  Same state_dict keys (`decoder_layers.0.{weight,bias}`), same
`forward(encoder_output=[B, D, T]) -> log_probs [B, T, V+1]`.  The conv is the tcgen05 GEMM, the softmax a warp-per-row
kernel; backward = log-softmax backward + the usual dgrad / wgrad GEMMs."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L

BF16, F32 = torch.bfloat16, torch.float32


class _CTCHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        B, D, T = x.shape
        M, V = B * T, weight.shape[0]
        dev = x.device
        xt = x.detach().transpose(1, 2).contiguous().to(F32)  # [B, T, D] (a no-op view for B200ConformerEncoder outputs)
        x16 = torch.empty(M, D, dtype=BF16, device=dev)
        L.call("cast_bf16", xt, x16, M * D, 1.0)
        w16 = torch.empty(V, D, dtype=BF16, device=dev)
        L.call("cast_bf16", weight.detach().contiguous(), w16, V * D, 1.0)
        ld = (V + 15) // 16 * 16
        logits = torch.empty(M, ld, dtype=BF16, device=dev)
        L.gemm(x16, w16, M, V, D, D, D, 0, 0, L.EPI_BF16, out0=logits, bias=bias.detach() if bias is not None else None, ldo=ld)
        logp = torch.empty(M, V, dtype=F32, device=dev)
        L.call("log_softmax_fwd", logits, ld, logp, M, V)
        ctx.save_for_backward(x16, w16, logp)
        ctx.meta = (B, D, T, V, ld, bias is not None)
        return logp.view(B, T, V)

    @staticmethod
    def backward(ctx, dlogp):
        x16, w16, logp = ctx.saved_tensors
        B, D, T, V, ld, has_bias = ctx.meta
        M, dev = B * T, x16.device
        dlogits = torch.empty(M, ld, dtype=BF16, device=dev)
        L.call("log_softmax_bwd", dlogp.contiguous().to(F32), logp, dlogits, ld, M, V)
        gw = torch.zeros(V, D, dtype=F32, device=dev)
        L.gemm(dlogits, x16, V, D, M, ld, D, 1, 1, L.EPI_ACCUM_F32, out0=gw, ldo=D)
        gb = None
        if has_bias:
            gb = torch.zeros(V, dtype=F32, device=dev)
            L.call("colsum_bf16", dlogits, gb, M, V, ld)
        dx = torch.empty(M, D, dtype=F32, device=dev)
        L.gemm(dlogits, w16, M, D, V, ld, D, 0, 1, L.EPI_RESID_F32, out0=dx, ldo=D)
        return dx.view(B, T, D).transpose(1, 2), gw.view(V, D, 1), gb


class B200ConvASRDecoder(nn.Module):
    """Constructor kwargs of ConvASRDecoder (conv_asr.py:407-455); `num_classes` excludes the CTC blank."""

    def __init__(self, feat_in, num_classes, init_mode="xavier_uniform", vocabulary=None, add_blank=True):
        super().__init__()
        if vocabulary is not None and num_classes <= 0:
            num_classes = len(vocabulary)
        self._feat_in, self.__vocabulary = feat_in, vocabulary
        self._num_classes = num_classes + 1 if add_blank else num_classes
        self.decoder_layers = nn.Sequential(nn.Conv1d(feat_in, self._num_classes, kernel_size=1, bias=True))
        if init_mode == "xavier_uniform":
            nn.init.xavier_uniform_(self.decoder_layers[0].weight, gain=1.0)
        self.temperature = 1.0
        self.precision = "bf16"  # "fp32": forward-only parity mode (nemo_b200/parity.py), fp32-equivalent logits

    @classmethod
    def from_reference(cls, ref):
        new = cls(feat_in=ref._feat_in, num_classes=ref._num_classes, add_blank=False)
        new.load_state_dict(ref.state_dict())
        return new.to(next(ref.parameters()).device)

    @property
    def vocabulary(self):
        return self.__vocabulary

    @property
    def num_classes_with_blank(self):
        return self._num_classes

    def forward(self, encoder_output):
        if not encoder_output.is_cuda:
            raise RuntimeError("B200ConvASRDecoder runs on CUDA tensors only (no CPU fallback)")
        if self.temperature != 1.0:
            raise NotImplementedError("temperature != 1.0")
        conv = self.decoder_layers[0]
        if self.precision == "fp32":
            from .parity import _linear

            with torch.no_grad():
                B, D, T = encoder_output.shape
                V = conv.weight.shape[0]
                x = encoder_output.detach().transpose(1, 2).contiguous().to(F32).view(B * T, D)
                logits = _linear(x, conv.weight, conv.bias, B * T, V, D)
                logp = torch.empty(B * T, V, dtype=F32, device=x.device)
                L.call("p_log_softmax", logits, V, logp, B * T, V)
                self._last_logits = logits.view(B, T, V)  # pre-softmax logits (north_star's metric names them)
                return logp.view(B, T, V)
        return _CTCHeadFn.apply(encoder_output, conv.weight, conv.bias)
