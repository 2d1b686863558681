"""fp32-accurate PARITY MODE of the encoder (SURVEY.md §8c tolerance policy (i)), forward AND backward.

`enc.precision = "fp32"` routes `B200ConformerEncoder.forward` here.  Same algorithm as the bf16 speed path and the same
tensor-core GEMM core (`b2c_gemm`, tcgen05 / TMEM), but every tensor between kernels is fp32 and every contraction is evaluated
with fp32-equivalent operands: x = hi + lo in bf16 (`b2c_p_split2`), A·B = hi·hi + hi·lo + lo·hi accumulated in fp32
(`B2C_EPI_ACCUM_F32` into a bias-prefilled fp32 output).  The elementwise stages (LayerNorm, SiLU, GLU, masks, depthwise convs,
BatchNorm, the rel-shifted masked softmax) and their backward are the plain fp32 kernels of csrc/parity.cu.

Purpose: north_star's "encoder logits within 1e-3 relative of the reference" can only be met by an fp32-accurate path — the
reference's OWN bf16-autocast run is 1.3e-2 away from its fp32 run at 18 layers (tests/test_full_depth_gpu.py).
This mode is the like-for-like check of the algorithm (masks, rel_shift, layouts, BatchNorm, flatten orders, every gradient) at
the BASELINE shapes against the fp32 oracle.  Dropout must be off; it is not a fast path (three tensor-core passes per
contraction, per-(batch, head) attention GEMMs, materialised [T,T] scores): use it on a few utterances."""
from __future__ import annotations

import math

import torch

from . import _lib as L

BF16, F32 = torch.bfloat16, torch.float32
P_SILU, P_AXPY, P_GLU_MASK, P_RELU_MASK, P_ADD_ROWVEC, P_DSILU, P_GLU_BWD, P_RELU_BWD = range(8)


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
    """out[M,N] (fp32, pitch ldo) (= bias +, or +=) A·Bᵀ with fp32-equivalent operands (three tcgen05 passes)"""
    if prefill:
        L.call("p_fill_rows", out, bias, M, N, ldo)
    for a, b in ((A.hi, B.lo), (A.lo, B.hi), (A.hi, B.hi)):  # small terms first
        L.gemm(a, b, M, N, K, lda, ldb, a_mn, b_mn, L.EPI_ACCUM_F32, out0=out, ldo=ldo)


def _pw(mode, a, b, out, M, N, lda=None, alpha=1.0, lengths=None, rows_per_batch=1, rows_inner=1):
    L.call("p_pointwise", mode, a, b, out, M, N, N if lda is None else lda, float(alpha), lengths, rows_per_batch, rows_inner)


def _linear(x, w, bias, M, N, K, out=None, ldo=None, xs=None, ws=None):
    """fp32 [M,K] · W[N,K]ᵀ + b  (xs / ws: already-split operands)"""
    if out is None:
        out = torch.empty(M, N, dtype=F32, device=(x if x is not None else xs.hi).device)
    _mm(xs or _Split(x), ws or _Split(w.reshape(N, K)), M, N, K, out, lda=K, ldb=K, ldo=N if ldo is None else ldo,
        bias=None if bias is None else bias.detach())
    return out


def _linear_bwd(dy, xs: _Split, ws: _Split, M, N, K, gw, gb, ldy=None, need_dx=True, dx=None, dx_prefill=True):
    """y = x·Wᵀ + b:  gw [N,K] += dyᵀ·x, gb [N] += colsum(dy), returns dx [M,K] = dy·W   (dy fp32 [M,N], pitch ldy)"""
    ldy = N if ldy is None else ldy
    dys = _Split(dy) if dy.is_contiguous() and ldy == N else _Split(dy.contiguous())
    ldy_s = N
    if gw is not None:
        _mm(dys, xs, N, K, M, gw, lda=ldy_s, ldb=K, ldo=K, a_mn=1, b_mn=1, prefill=False)
    if gb is not None:
        L.call("p_colsum", dy, ldy, gb, M, N, 1)
    if not need_dx:
        return None
    if dx is None:
        dx = torch.empty(M, K, dtype=F32, device=dy.device)
    _mm(dys, ws, M, K, N, dx, lda=ldy_s, ldb=K, ldo=K, b_mn=1, prefill=dx_prefill)
    return dx


class _Pass:
    """One forward (+ optional backward) of the encoder in fp32-equivalent arithmetic."""

    def __init__(self, enc, mel, length, save):
        self.enc, self.save = enc, save
        if enc.training and (enc.dropout > 0 or enc.dropout_pre_encoder > 0 or enc.dropout_att > 0 or enc.dropout_emb > 0):
            raise NotImplementedError("precision='fp32' (parity mode) needs eval() or all dropout probabilities = 0")
        self.dev = mel.device
        self.mel = mel.detach().to(F32).contiguous()
        self.length = length.detach().to(device=self.dev, dtype=torch.int64).contiguous()

    def E(self, *s, dtype=F32):
        return torch.empty(*s, dtype=dtype, device=self.dev)

    def Z(self, *s):
        return torch.zeros(*s, dtype=F32, device=self.dev)

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self):
        enc, E, dev = self.enc, self.E, self.dev
        mel, training = self.mel, enc.training
        B, Fm, Tm = mel.shape
        d, H, dff, k = enc.d_model, enc.n_heads, enc.d_ff, enc.conv_kernel_size
        dk = d // H
        pe = enc.pre_encode
        C, F2, n_conv = pe._conv_channels, pe._f_out, pe._sampling_num
        len_in32 = E(B, dtype=torch.int32)
        len_stage = E(n_conv * B, dtype=torch.int32)
        out_len = E(B, dtype=torch.int64)
        L.call("subsample_lengths", self.length, len_in32, len_stage, out_len, B, n_conv)
        lens = [len_stage[i * B:(i + 1) * B] for i in range(n_conv)]
        Ts, Fs = [Tm], [Fm]
        for _ in range(n_conv):
            Ts.append((Ts[-1] - 1) // 2 + 1)
            Fs.append((Fs[-1] - 1) // 2 + 1)
        T = Ts[-1]
        M = B * T
        enc_len32 = lens[-1]
        self.dims = dict(B=B, T=T, M=M, Ts=Ts, Fs=Fs, d=d, H=H, dff=dff, k=k, dk=dk, C=C, F2=F2, lens=lens, len_in32=len_in32)
        S = self.S = {}

        # ---- subsampling (subsampling.py:142-253,385-436,725-759) ---------------------------------------------------
        rows1 = B * Ts[1] * Fs[1]
        patches = E(rows1, 16)
        L.call("p_im2col_mel", mel, len_in32, patches, B, Fm, Tm)
        w16 = torch.zeros(C, 16, dtype=F32, device=dev)
        w16[:, :9] = pe.conv[0].weight.detach().reshape(C, 9)  # layout shuffle of a weight (no arithmetic)
        ps = _Split(patches)
        pre1 = _linear(None, w16, pe.conv[0].bias, rows1, C, 16, xs=ps)
        x1 = E(B, Ts[1], Fs[1], C)
        _pw(P_RELU_MASK, pre1, None, x1, rows1, C, lengths=lens[0], rows_per_batch=Ts[1] * Fs[1], rows_inner=Fs[1])
        del pre1, patches
        S["patches_s"], S["x1"] = (ps if self.save else None), x1
        if enc.subsampling == "striding":
            rows2 = B * Ts[2] * Fs[2]
            p2 = E(rows2, 9 * C)
            L.call("p_im2col_nhwc", x1, p2, B, Ts[1], Fs[1], C)
            w2 = E(C, 9, C)
            L.call("permute3_f32", pe.conv[2].weight.detach().contiguous(), w2, C, C, 9, 1, 0)  # [co, ci, 9] → [co, tap, ci]
            w2s = _Split(w2)
            pre2 = _linear(p2, None, pe.conv[2].bias, rows2, C, 9 * C, ws=w2s)
            del p2
            xl = E(B, Ts[2], Fs[2], C)
            _pw(P_RELU_MASK, pre2, None, xl, rows2, C, lengths=lens[1], rows_per_batch=Ts[2] * Fs[2], rows_inner=Fs[2])
            del pre2
            S["w2s"] = w2s
        else:
            xl = x1
            S["dw"] = []
            for s_i, ci in enumerate((2, 5)):
                rows = B * Ts[s_i + 2] * Fs[s_i + 2]
                dwo = E(B, Ts[s_i + 2], Fs[s_i + 2], C)
                L.call("p_dwconv2d", xl, pe.conv[ci].weight.detach().contiguous(), pe.conv[ci].bias.detach(), lens[s_i + 1], dwo,
                       B, Ts[s_i + 1], Fs[s_i + 1], C)
                dws, wps = _Split(dwo), _Split(pe.conv[ci + 1].weight.detach().reshape(C, C))
                pre = _linear(None, None, pe.conv[ci + 1].bias, rows, C, C, xs=dws, ws=wps)
                xin = xl
                xl = E(B, Ts[s_i + 2], Fs[s_i + 2], C)
                _pw(P_RELU_MASK, pre, None, xl, rows, C, lengths=lens[s_i + 1], rows_per_batch=Ts[s_i + 2] * Fs[s_i + 2],
                    rows_inner=Fs[s_i + 2])
                S["dw"].append((xin, dws, wps, xl))
        wo = E(d, F2, C)
        L.call("permute3_f32", pe.out.weight.detach().contiguous(), wo, d, C, F2, 1, 0)  # [n, c, f] → [n, f, c] (our flatten order)
        xls, wos = _Split(xl.view(M, F2 * C)), _Split(wo.view(d, F2 * C))
        xo = _linear(None, None, pe.out.bias, M, d, F2 * C, xs=xls, ws=wos)
        x = E(M, d)
        _pw(P_AXPY, None, xo, x, M, d, alpha=enc.xscale or 1.0)  # xscaling, multi_head_attention.py:1086-1087
        S["xl"], S["xls"], S["wos"] = xl, xls, wos
        del xo

        # ---- relative positions (multi_head_attention.py:1056-1100) --------------------------------------------------
        p16 = E(2 * T - 1, d, dtype=BF16)
        p32 = E(2 * T - 1, d)
        L.call("relpos_table", p16, p32, T, d)
        self.p32s = _Split(p32)
        Tp, Rp = (T + 7) & ~7, (2 * T - 1 + 7) & ~7
        self.dims.update(Tp=Tp, Rp=Rp)
        scale = 1.0 / math.sqrt(dk)
        taps = enc.__dict__.get("_debug_taps")
        self.layers_saved = []

        def ln(xin, norm):
            y = E(M, d)
            mean, rstd = E(M), E(M)
            L.call("layernorm_fwd", xin, norm.weight.detach(), norm.bias.detach(), None, y, mean, rstd, M, d)
            return y, mean, rstd

        def ffn(xin, norm, ff, sv, key):
            y, mean, rstd = ln(xin, norm)
            ys, w1s, w2s_ = _Split(y), _Split(ff.linear1.weight), _Split(ff.linear2.weight)
            z = _linear(None, None, ff.linear1.bias, M, dff, d, xs=ys, ws=w1s)
            h = E(M, dff)
            _pw(P_SILU, z, None, h, M, dff)
            hs = _Split(h)
            o = _linear(None, None, ff.linear2.bias, M, d, dff, xs=hs, ws=w2s_)
            xn = E(M, d)
            _pw(P_AXPY, xin, o, xn, M, d, alpha=0.5)
            if self.save:
                sv[key] = (xin, mean, rstd, ys, z, hs, w1s, w2s_)
            return xn

        for lyr in enc.layers:
            sv = {}
            xa = ffn(x, lyr.norm_feed_forward1, lyr.feed_forward1, sv, "ff1")
            # ---- rel-pos self-attention (multi_head_attention.py:272-354) --------------------------------------------
            at = lyr.self_attn
            y2, m2, r2 = ln(xa, lyr.norm_self_att)
            y2s = _Split(y2)
            qkv = E(M, 3 * d)
            wqkv = [_Split(lin.weight) for lin in (at.linear_q, at.linear_k, at.linear_v)]
            for j, lin in enumerate((at.linear_q, at.linear_k, at.linear_v)):
                _mm(y2s, wqkv[j], M, d, d, qkv[:, j * d:], lda=d, ldb=d, ldo=3 * d, bias=lin.bias.detach())
            wps = _Split(at.linear_pos.weight)
            pp = _linear(None, None, None, 2 * T - 1, d, d, xs=self.p32s, ws=wps)
            qu, qv = E(M, d), E(M, d)
            _pw(P_ADD_ROWVEC, qkv, at.pos_bias_u.detach().reshape(-1).contiguous(), qu, M, d, lda=3 * d)
            _pw(P_ADD_ROWVEC, qkv, at.pos_bias_v.detach().reshape(-1).contiguous(), qv, M, d, lda=3 * d)
            qus, qvs, kvs, pps = _Split(qu), _Split(qv), _Split(qkv), _Split(pp)
            AC, R = E(B * H, T, Tp), E(B * H, T, Rp)
            for b in range(B):
                for h in range(H):
                    bh = b * H + h
                    rows, cq = slice(b * T, (b + 1) * T), slice(h * dk, (h + 1) * dk)
                    _mm(qus.view(lambda t: t[rows, cq]), kvs.view(lambda t: t[rows, d + h * dk:d + (h + 1) * dk]), T, T, dk, AC[bh],
                        lda=d, ldb=3 * d, ldo=Tp)
                    _mm(qvs.view(lambda t: t[rows, cq]), pps.view(lambda t: t[:, cq]), T, 2 * T - 1, dk, R[bh], lda=d, ldb=d, ldo=Rp)
            P = E(B * H, T, Tp)
            L.call("p_attn_softmax", AC, R, enc_len32, P, B, H, T, Tp, Rp, scale)
            del AC, R
            Ps = _Split(P)
            ctx = E(M, d)
            for b in range(B):
                for h in range(H):
                    bh = b * H + h
                    rows = slice(b * T, (b + 1) * T)
                    _mm(Ps.view(lambda t: t[bh]), kvs.view(lambda t: t[rows, 2 * d + h * dk:2 * d + (h + 1) * dk]), T, dk, T,
                        ctx[rows, h * dk:], lda=Tp, ldb=3 * d, ldo=d, b_mn=1)
            ctxs, wouts = _Split(ctx), _Split(at.linear_out.weight)
            o = _linear(None, None, at.linear_out.bias, M, d, d, xs=ctxs, ws=wouts)
            xb = E(M, d)
            _pw(P_AXPY, xa, o, xb, M, d, alpha=1.0)
            if self.save:
                sv["att"] = (xa, m2, r2, y2s, wqkv, wps, qus, qvs, kvs, pps, P, Ps, ctxs, wouts)
            else:
                del P, Ps
            # ---- convolution module (conformer_modules.py:320-350) ---------------------------------------------------
            cv = lyr.conv
            y3, m3, r3 = ln(xb, lyr.norm_conv)
            y3s, wp1s = _Split(y3), _Split(cv.pointwise_conv1.weight.detach().reshape(2 * d, d))
            g = _linear(None, None, cv.pointwise_conv1.bias, M, 2 * d, d, xs=y3s, ws=wp1s)
            u = E(M, d)
            _pw(P_GLU_MASK, g, None, u, M, d, lda=2 * d, lengths=enc_len32, rows_per_batch=T, rows_inner=1)
            c = E(M, d)
            L.call("p_dwconv1d", u, cv.depthwise_conv.weight.detach().contiguous(), cv.depthwise_conv.bias.detach(), c, B, T, d, k)
            bn = cv.batch_norm
            stats = self.Z(2, d)
            if training:
                L.call("p_colstats", c, stats[0], stats[1], M, d)
            bnp = E(4, d)
            L.call("bn_finalize", stats[0], stats[1], bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                   bnp[0], bnp[1], bnp[2], bnp[3], M, d, 1 if training else 0, bn.momentum, bn.eps,
                   bn.num_batches_tracked if training else None)
            s_ = E(M, d)
            L.call("p_bn_silu", c, bnp[0], bnp[1], s_, M, d)
            ss, wp2s = _Split(s_), _Split(cv.pointwise_conv2.weight.detach().reshape(d, d))
            o = _linear(None, None, cv.pointwise_conv2.bias, M, d, d, xs=ss, ws=wp2s)
            xc = E(M, d)
            _pw(P_AXPY, xb, o, xc, M, d, alpha=1.0)
            if self.save:
                sv["conv"] = (xb, m3, r3, y3s, wp1s, g, u, c, bnp, ss, wp2s)
            xd = ffn(xc, lyr.norm_feed_forward2, lyr.feed_forward2, sv, "ff2")
            x, m5, r5 = ln(xd, lyr.norm_out)
            if self.save:
                sv["out"] = (xd, m5, r5)
                self.layers_saved.append(sv)
            if taps is not None:
                taps.append(x.view(B, T, d).clone())
        return x.view(B, T, d).transpose(1, 2), out_len

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, dout):
        enc, E, Z, S, dev = self.enc, self.E, self.Z, self.S, self.dev
        D = self.dims
        B, T, M, Ts, Fs, d, H, dff, k, dk, C, F2 = (D[x] for x in ("B", "T", "M", "Ts", "Fs", "d", "H", "dff", "k", "dk", "C", "F2"))
        Tp, Rp, lens = D["Tp"], D["Rp"], D["lens"]
        training = enc.training
        scale = 1.0 / math.sqrt(dk)
        G = {n: torch.zeros_like(p, dtype=F32) for n, p in enc.named_parameters()}
        dr = dout.detach().transpose(1, 2).contiguous().to(F32).view(M, d).clone()

        def ln_bwd(dy, xin, name, norm, mean, rstd, accumulate):
            L.call("layernorm_bwd", dy, 0, xin, norm.weight.detach(), mean, rstd, dr, accumulate, G[name + ".weight"], G[name + ".bias"],
                   None, 1.0, 0.0, 0, 0, M, d)

        def ffn_bwd(pfx, name, ff, norm_name, norm, saved):
            xin, mean, rstd, ys, z, hs, w1s, w2s_ = saved
            t = E(M, d)
            _pw(P_AXPY, None, dr, t, M, d, alpha=0.5)
            dh = _linear_bwd(t, hs, w2s_, M, d, dff, G[f"{pfx}{name}.linear2.weight"], G[f"{pfx}{name}.linear2.bias"])
            dz = E(M, dff)
            _pw(P_DSILU, dh, z, dz, M, dff)
            dy = _linear_bwd(dz, ys, w1s, M, dff, d, G[f"{pfx}{name}.linear1.weight"], G[f"{pfx}{name}.linear1.bias"])
            ln_bwd(dy, xin, pfx + norm_name, norm, mean, rstd, 1)

        for li in range(enc.n_layers - 1, -1, -1):
            lyr, sv, pfx = enc.layers[li], self.layers_saved[li], f"layers.{li}."
            xd, m5, r5 = sv["out"]
            ln_bwd(dr, xd, pfx + "norm_out", lyr.norm_out, m5, r5, 0)
            ffn_bwd(pfx, "feed_forward2", lyr.feed_forward2, "norm_feed_forward2", lyr.norm_feed_forward2, sv["ff2"])
            # ---- convolution module
            xb, m3, r3, y3s, wp1s, g, u, c, bnp, ss, wp2s = sv["conv"]
            cv = lyr.conv
            ds = _linear_bwd(dr, ss, wp2s, M, d, d, G[pfx + "conv.pointwise_conv2.weight"].view(d, d), G[pfx + "conv.pointwise_conv2.bias"])
            dz_s, dzx_s, sums, dc = E(M, d), E(M, d), E(2 * d), E(M, d)
            L.call("p_bn_silu_bwd", ds, c, cv.batch_norm.weight.detach(), bnp[0], bnp[1], bnp[2], bnp[3], dz_s, dzx_s, sums, dc, M, d,
                   1 if training else 0)
            G[pfx + "conv.batch_norm.bias"].add_(sums[:d])    # plumbing: two [d] adds per layer (parity mode only)
            G[pfx + "conv.batch_norm.weight"].add_(sums[d:])
            du = E(M, d)
            L.call("p_dwconv1d_bwd", dc, u, cv.depthwise_conv.weight.detach().contiguous(), lens[-1], du,
                   G[pfx + "conv.depthwise_conv.weight"], B, T, d, k)
            L.call("p_colsum", dc, d, G[pfx + "conv.depthwise_conv.bias"], M, d, 1)
            dg = E(M, 2 * d)
            _pw(P_GLU_BWD, du, g, dg, M, d, lda=d, lengths=lens[-1], rows_per_batch=T, rows_inner=1)
            dy3 = _linear_bwd(dg, y3s, wp1s, M, 2 * d, d, G[pfx + "conv.pointwise_conv1.weight"].view(2 * d, d),
                              G[pfx + "conv.pointwise_conv1.bias"])
            ln_bwd(dy3, xb, pfx + "norm_conv", lyr.norm_conv, m3, r3, 1)
            # ---- self-attention
            xa, m2, r2, y2s, wqkv, wps, qus, qvs, kvs, pps, P, Ps, ctxs, wouts = sv["att"]
            at = lyr.self_attn
            dctx = _linear_bwd(dr, ctxs, wouts, M, d, d, G[pfx + "self_attn.linear_out.weight"], G[pfx + "self_attn.linear_out.bias"])
            dctxs = _Split(dctx)
            dP = E(B * H, T, Tp)
            dqkv = Z(M, 3 * d)
            for b in range(B):
                for h in range(H):
                    bh = b * H + h
                    rows, cq = slice(b * T, (b + 1) * T), slice(h * dk, (h + 1) * dk)
                    vsl = lambda t: t[rows, 2 * d + h * dk:2 * d + (h + 1) * dk]  # noqa: E731
                    # dP = dctx · Vᵀ ;  dV = Pᵀ · dctx
                    _mm(dctxs.view(lambda t: t[rows, cq]), kvs.view(vsl), T, T, dk, dP[bh], lda=d, ldb=3 * d, ldo=Tp)
                    _mm(Ps.view(lambda t: t[bh]), dctxs.view(lambda t: t[rows, cq]), T, dk, T, dqkv[rows, 2 * d + h * dk:], lda=Tp, ldb=d,
                        ldo=3 * d, a_mn=1, b_mn=1, prefill=False)
            dS, dR = E(B * H, T, Tp), E(B * H, T, Rp)
            L.call("p_attn_softmax_bwd", P, dP, dS, dR, B, H, T, Tp, Rp, scale)
            del dP
            dSs, dRs = _Split(dS), _Split(dR)
            dqu, dqv, dpp = Z(M, d), Z(M, d), Z(2 * T - 1, d)
            for b in range(B):
                for h in range(H):
                    bh = b * H + h
                    rows, cq = slice(b * T, (b + 1) * T), slice(h * dk, (h + 1) * dk)
                    ksl = lambda t: t[rows, d + h * dk:d + (h + 1) * dk]  # noqa: E731
                    a_ds, a_dr = dSs.view(lambda t: t[bh]), dRs.view(lambda t: t[bh])
                    # dQu = dS · K ; dK = dSᵀ · Qu ; dQv = dR · p ; dp += dRᵀ · Qv
                    _mm(a_ds, kvs.view(ksl), T, dk, T, dqu[rows, cq], lda=Tp, ldb=3 * d, ldo=d, b_mn=1, prefill=False)
                    _mm(a_ds, qus.view(lambda t: t[rows, cq]), T, dk, T, dqkv[rows, d + h * dk:], lda=Tp, ldb=d, ldo=3 * d, a_mn=1, b_mn=1,
                        prefill=False)
                    _mm(a_dr, pps.view(lambda t: t[:, cq]), T, dk, 2 * T - 1, dqv[rows, cq], lda=Rp, ldb=d, ldo=d, b_mn=1, prefill=False)
                    _mm(a_dr, qvs.view(lambda t: t[rows, cq]), 2 * T - 1, dk, T, dpp[:, cq], lda=Rp, ldb=d, ldo=d, a_mn=1, b_mn=1,
                        prefill=False)
            del dS, dR, dSs, dRs
            L.call("p_colsum", dqu, d, G[pfx + "self_attn.pos_bias_u"], M, d, 1)
            L.call("p_colsum", dqv, d, G[pfx + "self_attn.pos_bias_v"], M, d, 1)
            dq = E(M, d)
            _pw(P_AXPY, dqu, dqv, dq, M, d, alpha=1.0)
            dqkv[:, :d].copy_(dq)  # plumbing: place dq next to dk | dv
            _linear_bwd(dpp, self.p32s, wps, 2 * T - 1, d, d, G[pfx + "self_attn.linear_pos.weight"], None, need_dx=False)
            dy2 = Z(M, d)
            for j, nm in enumerate(("linear_q", "linear_k", "linear_v")):
                blk = dqkv[:, j * d:(j + 1) * d].contiguous()
                _linear_bwd(blk, y2s, wqkv[j], M, d, d, G[f"{pfx}self_attn.{nm}.weight"], G[f"{pfx}self_attn.{nm}.bias"], dx=dy2,
                            dx_prefill=False)
            ln_bwd(dy2, xa, pfx + "norm_self_att", lyr.norm_self_att, m2, r2, 1)
            ffn_bwd(pfx, "feed_forward1", lyr.feed_forward1, "norm_feed_forward1", lyr.norm_feed_forward1, sv["ff1"])
            self.layers_saved[li] = None

        # ---- pre_encode: xscale, out Linear, convolutions
        pe = enc.pre_encode
        dxo = E(M, d)
        _pw(P_AXPY, None, dr, dxo, M, d, alpha=enc.xscale or 1.0)
        gwo = Z(d, F2 * C)
        dxl = _linear_bwd(dxo, S["xls"], S["wos"], M, d, F2 * C, gwo, G["pre_encode.out.bias"])
        dbg = enc.__dict__.get("_debug_parity")  # tests / tools/debug: intermediates of the subsampling stack
        if dbg is not None:
            dbg["dr"], dbg["dxl"] = dr.clone(), dxl.clone()
            # the ReLU outputs in forward order ([B,T,F,C]): their signs are the gates the backward used
            dbg["relu_out"] = [S["x1"]] + ([st[3] for st in S["dw"]] if "dw" in S else [S["xl"]])
        L.call("permute3_f32", gwo, G["pre_encode.out.weight"], d, F2, C, 1, 1)  # [n, f, c] → [n, c*F2 + f]

        def conv1_bwd(dx1):
            rows1 = B * Ts[1] * Fs[1]
            dpre1 = E(rows1, C)
            _pw(P_RELU_BWD, dx1, S["x1"].view(rows1, C), dpre1, rows1, C)
            g16 = Z(C, 16)
            _linear_bwd(dpre1, S["patches_s"], None, rows1, C, 16, g16, G["pre_encode.conv.0.bias"], need_dx=False)
            G["pre_encode.conv.0.weight"].view(C, 9).add_(g16[:, :9])  # plumbing: un-pad the [C,16] stem gradient

        if enc.subsampling == "striding":
            rows2 = B * Ts[2] * Fs[2]
            dpre2 = E(rows2, C)
            _pw(P_RELU_BWD, dxl, S["xl"].view(rows2, C), dpre2, rows2, C)
            p2 = E(rows2, 9 * C)
            L.call("p_im2col_nhwc", S["x1"], p2, B, Ts[1], Fs[1], C)
            gw2 = Z(C, 9 * C)
            dp2 = _linear_bwd(dpre2, _Split(p2), S["w2s"], rows2, C, 9 * C, gw2, G["pre_encode.conv.2.bias"])
            del p2
            L.call("permute3_f32", gw2, G["pre_encode.conv.2.weight"], C, 9, C, 1, 1)  # [co, tap, ci] → [co, ci, tap]
            dx1 = E(B, Ts[1], Fs[1], C)
            L.call("p_col2im_nhwc", dp2, dx1, B, Ts[1], Fs[1], C)
            conv1_bwd(dx1.view(-1, C))
        else:
            dcur = dxl  # gradient w.r.t. the last stage's ReLU+mask output, [rows, C]
            for s_i, ci in reversed(list(enumerate((2, 5)))):
                xin, dws, wps, xout = S["dw"][s_i]
                rows = B * Ts[s_i + 2] * Fs[s_i + 2]
                dpre = E(rows, C)
                _pw(P_RELU_BWD, dcur.view(rows, C), xout.view(rows, C), dpre, rows, C)
                ddw = _linear_bwd(dpre, dws, wps, rows, C, C, G[f"pre_encode.conv.{ci + 1}.weight"].view(C, C),
                                  G[f"pre_encode.conv.{ci + 1}.bias"])
                dxin = E(B, Ts[s_i + 1], Fs[s_i + 1], C)
                L.call("p_dwconv2d_bwd", ddw, xin, pe.conv[ci].weight.detach().contiguous(), lens[s_i + 1], dxin,
                       G[f"pre_encode.conv.{ci}.weight"], G[f"pre_encode.conv.{ci}.bias"], B, Ts[s_i + 1], Fs[s_i + 1], C)
                dcur = dxin
                if dbg is not None:
                    dbg[f"dpre{s_i}"], dbg[f"ddw{s_i}"], dbg[f"dxin{s_i}"] = dpre.clone(), ddw.clone(), dxin.clone()
            conv1_bwd(dcur.view(-1, C))
        return G


class _EncoderFn32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, mel, length, *params):
        p = _Pass(enc, mel, length, save=True)
        out, out_len = p.forward()
        ctx.p, ctx.names = p, [n for n, _ in enc.named_parameters()]
        ctx.req = [q.requires_grad for q in params]
        ctx.mark_non_differentiable(out_len)
        return out, out_len

    @staticmethod
    def backward(ctx, dout, _dlen):
        if ctx.p is None:
            raise RuntimeError("B200ConformerEncoder (fp32 parity mode): backward through the same graph a second time")
        G = ctx.p.backward(dout)
        ctx.p = None
        return (None, None, None, *[G[n] if r else None for n, r in zip(ctx.names, ctx.req)])


def forward_fp32(enc, mel: torch.Tensor, length: torch.Tensor):
    params = [p for _, p in enc.named_parameters()]
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _EncoderFn32.apply(enc, mel, length, *params)
    with torch.no_grad():
        return _Pass(enc, mel, length, save=False).forward()
