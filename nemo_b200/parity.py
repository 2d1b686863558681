"""fp32-accurate PARITY MODE of the encoder forward (SURVEY.md §8c tolerance policy (i)).

`B200ConformerEncoder(..., precision="fp32")` (or `enc.precision = "fp32"`) routes `forward` here.  Same algorithm as the
bf16 speed path and the same tensor-core GEMM core (`b2c_gemm`, tcgen05 / TMEM), but every tensor between kernels is fp32 and
every contraction is evaluated with fp32-equivalent operands: x = hi + lo in bf16 (`b2c_p_split2`), A·B = hi·hi + hi·lo + lo·hi
accumulated in fp32 (`B2C_EPI_ACCUM_F32` into a bias-prefilled fp32 output).  The elementwise stages (LayerNorm, SiLU, GLU,
masks, depthwise convs, BatchNorm, the rel-shifted masked softmax) are the plain fp32 kernels of csrc/parity.cu.

Purpose: north_star's "encoder logits within 1e-3 relative of the reference" can only be met by an fp32-accurate path — the
reference's OWN bf16-autocast run is 1.3e-2 away from its fp32 run at 18 layers (tests/test_full_depth_gpu.py).
This mode is the like-for-like check of the algorithm (masks, rel_shift, layouts, BatchNorm, flatten orders) at the BASELINE
shapes against the fp32 oracle; it is forward-only (inference / logits), dropout must be off, and it is not a fast path
(~15x slower than speed mode: 3 tensor-core passes per contraction, per-(batch, head) attention GEMMs, materialised scores)."""
from __future__ import annotations

import math

import torch

from . import _lib as L

BF16, F32 = torch.bfloat16, torch.float32
P_SILU, P_AXPY, P_GLU_MASK, P_RELU_MASK, P_ADD_ROWVEC = range(5)


class _Split:
    """fp32 tensor → bf16 (hi, lo) with x ≈ hi + lo to 2^-17 relative; same shape / strides as the (contiguous) source"""

    __slots__ = ("hi", "lo")

    def __init__(self, x: torch.Tensor):
        x = x.detach()
        if x.dtype != F32 or not x.is_contiguous():
            x = x.to(F32).contiguous()
        self.hi = torch.empty(x.shape, dtype=BF16, device=x.device)
        self.lo = torch.empty(x.shape, dtype=BF16, device=x.device)
        L.call("p_split2", x, self.hi, self.lo, x.numel())

    def view(self, fn):
        s = object.__new__(_Split)
        s.hi, s.lo = fn(self.hi), fn(self.lo)
        return s


def _mm(A: _Split, B: _Split, M, N, K, out, *, lda, ldb, ldo, a_mn=0, b_mn=0, bias=None, prefill=True):
    """out[M,N] (fp32, pitch ldo) = bias + A·Bᵀ with fp32-equivalent operands (three tcgen05 passes)"""
    if prefill:
        L.call("p_fill_rows", out, bias, M, N, ldo)
    for a, b in ((A.hi, B.lo), (A.lo, B.hi), (A.hi, B.hi)):  # small terms first
        L.gemm(a, b, M, N, K, lda, ldb, a_mn, b_mn, L.EPI_ACCUM_F32, out0=out, ldo=ldo)


def _pw(mode, a, b, out, M, N, lda=None, alpha=1.0, lengths=None, rows_per_batch=1, rows_inner=1):
    L.call("p_pointwise", mode, a, b, out, M, N, N if lda is None else lda, float(alpha), lengths, rows_per_batch, rows_inner)


def _linear(x: torch.Tensor, w: torch.Tensor, bias, M, N, K, out=None, ldo=None):
    """fp32 [M,K] · W[N,K]ᵀ + b"""
    dev = x.device
    if out is None:
        out = torch.empty(M, N, dtype=F32, device=dev)
    _mm(_Split(x), _Split(w.reshape(N, K)), M, N, K, out, lda=K, ldb=K, ldo=N if ldo is None else ldo,
        bias=None if bias is None else bias.detach())
    return out


@torch.no_grad()
def forward_fp32(enc, mel: torch.Tensor, length: torch.Tensor):
    if enc.training and (enc.dropout > 0 or enc.dropout_pre_encoder > 0 or enc.dropout_att > 0 or enc.dropout_emb > 0):
        raise NotImplementedError("precision='fp32' (parity mode) needs eval() or all dropout probabilities = 0")
    dev = mel.device
    training = enc.training
    B, Fm, Tm = mel.shape
    d, H, dff, k = enc.d_model, enc.n_heads, enc.d_ff, enc.conv_kernel_size
    dk = d // H
    pe = enc.pre_encode
    C, F2, n_conv = pe._conv_channels, pe._f_out, pe._sampling_num
    mel = mel.detach().to(F32).contiguous()
    length = length.detach().to(device=dev, dtype=torch.int64).contiguous()
    E = lambda *s, dtype=F32: torch.empty(*s, dtype=dtype, device=dev)  # noqa: E731

    len_in32 = E(B, dtype=torch.int32)
    len_stage = E(n_conv * B, dtype=torch.int32)
    out_len = E(B, dtype=torch.int64)
    L.call("subsample_lengths", length, len_in32, len_stage, out_len, B, n_conv)
    lens = [len_stage[i * B:(i + 1) * B] for i in range(n_conv)]
    Ts, Fs = [Tm], [Fm]
    for _ in range(n_conv):
        Ts.append((Ts[-1] - 1) // 2 + 1)
        Fs.append((Fs[-1] - 1) // 2 + 1)
    T = Ts[-1]
    M = B * T
    enc_len32 = lens[-1]

    # ---- subsampling (subsampling.py:142-253,385-436,725-759) -------------------------------------------------------
    rows1 = B * Ts[1] * Fs[1]
    patches = E(rows1, 16)
    L.call("p_im2col_mel", mel, len_in32, patches, B, Fm, Tm)
    w16 = torch.zeros(C, 16, dtype=F32, device=dev)
    w16[:, :9] = pe.conv[0].weight.detach().reshape(C, 9)  # layout shuffle of a weight (no arithmetic)
    pre1 = _linear(patches, w16, pe.conv[0].bias, rows1, C, 16)
    x1 = E(B, Ts[1], Fs[1], C)
    _pw(P_RELU_MASK, pre1, None, x1, rows1, C, lengths=lens[0], rows_per_batch=Ts[1] * Fs[1], rows_inner=Fs[1])
    del pre1, patches
    if enc.subsampling == "striding":
        rows2 = B * Ts[2] * Fs[2]
        p2 = E(rows2, 9 * C)
        L.call("p_im2col_nhwc", x1, p2, B, Ts[1], Fs[1], C)
        w2 = E(C, 9, C)
        L.call("permute3_f32", pe.conv[2].weight.detach().contiguous(), w2, C, C, 9, 1, 0)  # [co, ci, 9] → [co, tap, ci]
        pre2 = _linear(p2, w2, pe.conv[2].bias, rows2, C, 9 * C)
        del p2
        xl = E(B, Ts[2], Fs[2], C)
        _pw(P_RELU_MASK, pre2, None, xl, rows2, C, lengths=lens[1], rows_per_batch=Ts[2] * Fs[2], rows_inner=Fs[2])
        del pre2
    else:
        xl = x1
        for s_i, ci in enumerate((2, 5)):
            rows = B * Ts[s_i + 2] * Fs[s_i + 2]
            dwo = E(B, Ts[s_i + 2], Fs[s_i + 2], C)
            L.call("p_dwconv2d", xl, pe.conv[ci].weight.detach().contiguous(), pe.conv[ci].bias.detach(), lens[s_i + 1], dwo,
                   B, Ts[s_i + 1], Fs[s_i + 1], C)
            pre = _linear(dwo, pe.conv[ci + 1].weight.detach(), pe.conv[ci + 1].bias, rows, C, C)
            xl = E(B, Ts[s_i + 2], Fs[s_i + 2], C)
            _pw(P_RELU_MASK, pre, None, xl, rows, C, lengths=lens[s_i + 1], rows_per_batch=Ts[s_i + 2] * Fs[s_i + 2],
                rows_inner=Fs[s_i + 2])
    wo = E(d, F2, C)
    L.call("permute3_f32", pe.out.weight.detach().contiguous(), wo, d, C, F2, 1, 0)  # [n, c, f] → [n, f, c] (our flatten order)
    xo = _linear(xl.view(M, F2 * C), wo, pe.out.bias, M, d, F2 * C)
    x = E(M, d)
    _pw(P_AXPY, None, xo, x, M, d, alpha=enc.xscale or 1.0)  # xscaling, multi_head_attention.py:1086-1087
    del xo, xl, x1

    # ---- relative positions (multi_head_attention.py:1056-1100) ------------------------------------------------------
    p16 = E(2 * T - 1, d, dtype=BF16)
    p32 = E(2 * T - 1, d)
    L.call("relpos_table", p16, p32, T, d)
    pos_split = _Split(p32)
    Tp, Rp = (T + 7) & ~7, (2 * T - 1 + 7) & ~7
    scale = 1.0 / math.sqrt(dk)
    taps = enc.__dict__.get("_debug_taps")

    def ln(xin, norm):
        y = E(M, d)
        L.call("layernorm_fwd", xin, norm.weight.detach(), norm.bias.detach(), None, y, None, None, M, d)
        return y

    def ffn(xin, norm, ff):
        y = ln(xin, norm)
        z = _linear(y, ff.linear1.weight, ff.linear1.bias, M, dff, d)
        h = E(M, dff)
        _pw(P_SILU, z, None, h, M, dff)
        o = _linear(h, ff.linear2.weight, ff.linear2.bias, M, d, dff)
        xn = E(M, d)
        _pw(P_AXPY, xin, o, xn, M, d, alpha=0.5)
        return xn

    for lyr in enc.layers:
        xa = ffn(x, lyr.norm_feed_forward1, lyr.feed_forward1)
        # ---- rel-pos self-attention (multi_head_attention.py:272-354) ------------------------------------------------
        at = lyr.self_attn
        y2 = ln(xa, lyr.norm_self_att)
        y2s = _Split(y2)
        qkv = E(M, 3 * d)
        for j, lin in enumerate((at.linear_q, at.linear_k, at.linear_v)):
            _mm(y2s, _Split(lin.weight), M, d, d, qkv[:, j * d:], lda=d, ldb=d, ldo=3 * d, bias=lin.bias.detach())
        pp = _linear(p32, at.linear_pos.weight, None, 2 * T - 1, d, d)
        qu, qv = E(M, d), E(M, d)
        _pw(P_ADD_ROWVEC, qkv, at.pos_bias_u.detach().reshape(-1).contiguous(), qu, M, d, lda=3 * d)
        _pw(P_ADD_ROWVEC, qkv, at.pos_bias_v.detach().reshape(-1).contiguous(), qv, M, d, lda=3 * d)
        qus, qvs, kvs, pps = _Split(qu), _Split(qv), _Split(qkv), _Split(pp)
        AC, R = E(B * H, T, Tp), E(B * H, T, Rp)
        for b in range(B):
            for h in range(H):
                bh = b * H + h
                a_u = qus.view(lambda t: t[b * T:(b + 1) * T, h * dk:(h + 1) * dk])
                a_v = qvs.view(lambda t: t[b * T:(b + 1) * T, h * dk:(h + 1) * dk])
                k_bh = kvs.view(lambda t: t[b * T:(b + 1) * T, d + h * dk:d + (h + 1) * dk])
                p_h = pps.view(lambda t: t[:, h * dk:(h + 1) * dk])
                _mm(a_u, k_bh, T, T, dk, AC[bh], lda=d, ldb=3 * d, ldo=Tp)
                _mm(a_v, p_h, T, 2 * T - 1, dk, R[bh], lda=d, ldb=d, ldo=Rp)
        P = E(B * H, T, Tp)
        L.call("p_attn_softmax", AC, R, enc_len32, P, B, H, T, Tp, Rp, scale)
        Ps = _Split(P)
        ctx = E(M, d)
        for b in range(B):
            for h in range(H):
                bh = b * H + h
                v_bh = kvs.view(lambda t: t[b * T:(b + 1) * T, 2 * d + h * dk:2 * d + (h + 1) * dk])
                _mm(Ps.view(lambda t: t[bh]), v_bh, T, dk, T, ctx[b * T:(b + 1) * T, h * dk:], lda=Tp, ldb=3 * d, ldo=d, b_mn=1)
        o = _linear(ctx, at.linear_out.weight, at.linear_out.bias, M, d, d)
        xb = E(M, d)
        _pw(P_AXPY, xa, o, xb, M, d, alpha=1.0)
        del AC, R, P, Ps
        # ---- convolution module (conformer_modules.py:320-350) -------------------------------------------------------
        cv = lyr.conv
        y3 = ln(xb, lyr.norm_conv)
        g = _linear(y3, cv.pointwise_conv1.weight, cv.pointwise_conv1.bias, M, 2 * d, d)
        u = E(M, d)
        _pw(P_GLU_MASK, g, None, u, M, d, lda=2 * d, lengths=enc_len32, rows_per_batch=T, rows_inner=1)
        c = E(M, d)
        L.call("p_dwconv1d", u, cv.depthwise_conv.weight.detach().contiguous(), cv.depthwise_conv.bias.detach(), c, B, T, d, k)
        bn = cv.batch_norm
        stats = torch.zeros(2, d, dtype=F32, device=dev)
        if training:
            L.call("p_colstats", c, stats[0], stats[1], M, d)
        bnp = E(4, d)
        L.call("bn_finalize", stats[0], stats[1], bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
               bnp[0], bnp[1], bnp[2], bnp[3], M, d, 1 if training else 0, bn.momentum, bn.eps,
               bn.num_batches_tracked if training else None)
        s_ = E(M, d)
        L.call("p_bn_silu", c, bnp[0], bnp[1], s_, M, d)
        o = _linear(s_, cv.pointwise_conv2.weight, cv.pointwise_conv2.bias, M, d, d)
        xc = E(M, d)
        _pw(P_AXPY, xb, o, xc, M, d, alpha=1.0)
        xd = ffn(xc, lyr.norm_feed_forward2, lyr.feed_forward2)
        x = ln(xd, lyr.norm_out)
        if taps is not None:
            taps.append(x.view(B, T, d).clone())
    return x.view(B, T, d).transpose(1, 2), out_len
