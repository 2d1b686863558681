"""GPU unit tests: every CUDA kernel of libb200conformer against a plain PyTorch fp32 reference of the same op
(called through the C ABI).  Run on a B200:  python -m pytest tests -m gpu -x -q"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from nemo_b200 import _lib as L
    DEV = torch.device("cuda:0")


def rnd(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def report(name, got, ref, tol):
    e = relerr(got, ref)
    nan = int(torch.isnan(got.float()).sum().item())
    print(f"[{name}] max-abs-err/max|ref| = {e:.3e} (tol {tol:.1e}) nan={nan} shape={tuple(ref.shape)}")
    assert nan == 0, f"{name}: {nan} NaNs"
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol:.1e}"


# ------------------------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [(128, 128, 64), (300, 256, 512), (1000, 2048, 512), (777, 1536, 512), (520, 512, 2048), (130, 384, 200)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_kmajor_bias_bf16(M, N, K):
    A, B = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=0.05, dtype=torch.bfloat16)
    bias = rnd(N)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_BF16, out0=out, bias=bias)
    ref = A.float() @ B.float().T + bias
    report(f"gemm_tn_{M}x{N}x{K}", out, ref, 1e-2)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (512, 2048, 1000), (640, 128, 333), (2048, 512, 4000)])
def test_gemm_mn_major(a_mn, b_mn, M, N, K):
    A = rnd(K, M, dtype=torch.bfloat16) if a_mn else rnd(M, K, dtype=torch.bfloat16)
    B = rnd(K, N, scale=0.05, dtype=torch.bfloat16) if b_mn else rnd(N, K, scale=0.05, dtype=torch.bfloat16)
    Af = A.float().T if a_mn else A.float()
    Bf = B.float().T if b_mn else B.float()
    ref = Af @ Bf.T
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    L.gemm(A, B, M, N, K, A.shape[1], B.shape[1], a_mn, b_mn, L.EPI_BF16, out0=out)
    report(f"gemm_mn{a_mn}{b_mn}_{M}x{N}x{K}", out, ref, 1e-2)
    # fp32 accumulate (weight-gradient epilogue, split-K atomics)
    acc = rnd(M, N)
    acc0 = acc.clone()
    L.gemm(A, B, M, N, K, A.shape[1], B.shape[1], a_mn, b_mn, L.EPI_ACCUM_F32, out0=acc, alpha=0.5)
    report(f"gemm_accum_mn{a_mn}{b_mn}_{M}x{N}x{K}", acc, acc0 + 0.5 * ref, 2e-3)


def test_gemm_epilogue_silu_and_dsilu():
    M, N, K = 500, 1024, 256
    A, B, bias = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=0.08, dtype=torch.bfloat16), rnd(N)
    z = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    h = torch.empty_like(z)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_SILU, out0=z, out1=h, bias=bias)
    zr = (A.float() @ B.float().T + bias)
    report("silu_z", z, zr, 1e-2)
    report("silu_h", h, torch.nn.functional.silu(z.float()), 1e-2)
    # dsilu: out = bf16(acc) * silu'(z)
    G, W = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=0.08, dtype=torch.bfloat16)
    dz = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    L.gemm(G, W, M, N, K, K, K, 0, 0, L.EPI_DSILU, out0=dz, aux=z)
    zf = z.float()
    s = torch.sigmoid(zf)
    ref = (G.float() @ W.float().T) * (s * (1 + zf * (1 - s)))
    report("dsilu", dz, ref, 1.5e-2)


def test_gemm_epilogue_resid_glu_relu():
    M, N, K = 3 * 100, 512, 256
    A, B, bias = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=0.08, dtype=torch.bfloat16), rnd(N)
    x = rnd(M, N)
    out = torch.empty_like(x)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_RESID_F32, out0=out, aux=x, bias=bias, alpha=0.5)
    lin = (A.float() @ B.float().T + bias)
    report("resid", out, x + 0.5 * lin, 5e-3)
    out2 = torch.empty_like(x)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_RESID_F32, out0=out2, bias=bias, alpha=3.0)
    report("resid_noaux", out2, 3.0 * lin, 1e-2)
    # GLU: interleaved (value, gate) columns, rows = (b, t) with T=100, lengths mask
    lengths = torch.tensor([100, 37, 0], dtype=torch.int32, device=DEV)
    g = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    u = torch.empty(M, N // 2, dtype=torch.bfloat16, device=DEV)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_GLU_MASK, out0=g, out1=u, bias=bias, lengths=lengths, rows_per_batch=100)
    report("glu_g", g, lin, 1e-2)
    gf = g.float()
    t = torch.arange(M, device=DEV) % 100
    valid = (t < lengths.long().repeat_interleave(100)).float()[:, None]
    report("glu_u", u, gf[:, 0::2] * torch.sigmoid(gf[:, 1::2]) * valid, 1e-2)
    # ReLU + time mask with rows = (b, t, f): rows_per_batch = T*F, rows_inner = F
    F_ = 4
    lengths2 = torch.tensor([25, 10, 3], dtype=torch.int32, device=DEV)
    o = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_RELU_MASK, out0=o, bias=bias, lengths=lengths2, rows_per_batch=100,
           rows_inner=F_)
    tt = (torch.arange(M, device=DEV) % 100) // F_
    valid2 = (tt < lengths2.long().repeat_interleave(100)).float()[:, None]
    report("relu_mask", o, torch.relu(lin) * valid2, 1e-2)


def test_gemm_dropout_masks_are_consistent_and_unbiased():
    M, N, K = 512, 512, 64
    A, B = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=0.1, dtype=torch.bfloat16)
    o1 = torch.empty(M, N, device=DEV)
    o2 = torch.empty(M, N, device=DEV)
    o0 = torch.empty(M, N, device=DEV)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_RESID_F32, out0=o0)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_RESID_F32, out0=o1, seed=7, drop_site=3, drop_p=0.1)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_RESID_F32, out0=o2, seed=7, drop_site=3, drop_p=0.1)
    assert torch.equal(o1, o2)
    kept = (o1 != 0) | (o0 == 0)
    frac = kept.float().mean().item()
    print("dropout keep fraction", frac)
    assert abs(frac - 0.9) < 0.01
    assert torch.allclose(o1[kept], o0[kept] / 0.9, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------------- mel
@pytest.mark.parametrize("B,secs,ragged", [(3, 1.0, True), (2, 4.3, True), (4, 30.0, False)])
def test_mel_frontend_vs_oracle(B, secs, ragged):
    from nemo_b200.preprocessor import B200MelPreprocessor
    from oracle import conformer_oracle as O

    x, length = O.synthetic_waveforms(B, secs, ragged=ragged)
    if ragged:
        length[-1] = min(799, x.shape[1])
    ref, ref_len, ref_raw = O.mel_frontend(x, length, return_prenorm=True)
    _, _, truth_raw = O.mel_frontend(x, length, return_prenorm=True, dtype=torch.float64)  # same formulas in double
    pre = B200MelPreprocessor(sample_rate=16000, window_size=0.025, window_stride=0.01, features=80, n_fft=512,
                              dither=1e-5, pad_to=0).to(DEV).eval()
    mel, mlen = pre(input_signal=x.to(DEV), length=length.to(DEV))
    assert torch.equal(mlen.cpu(), ref_len)  # bit-exact frame counts
    assert mel.shape == ref.shape
    raw, _ = pre.featurizer.forward_raw(x.to(DEV), length.to(DEV))
    valid = (torch.arange(ref.shape[-1])[None, :] < ref_len[:, None])[:, None, :]
    err_raw = ((raw.cpu() - ref_raw) * valid).abs().max().item()
    err_truth = ((raw.cpu().double() - truth_raw) * valid).abs().max().item()
    ref_truth = ((ref_raw.double() - truth_raw) * valid).abs().max().item()
    err = (mel.cpu() - ref).abs().max().item()
    print(f"[mel B={B} {secs}s] log-mel max-abs err: cuda vs torch.stft(fp32) {err_raw:.2e}; vs float64 truth: cuda {err_truth:.2e} | "
          f"torch.stft(fp32) {ref_truth:.2e}; normalised abs err {err:.2e}")
    # SURVEY §8c policy: <= 1e-4 on the pre-normalisation log-mel.  Both fp32 pipelines sit ~1e-4 from the float64 truth on the
    # worst bin (log of a sum of ~1e-6 powers amplifies the FFT's 1e-7 rounding), so the gate is on the distance to the TRUTH:
    # the CUDA front-end may not be further from it than the reference's own fp32 torch.stft path (x1.5), nor than 1e-4 when
    # the reference is that accurate itself.
    assert err_truth <= max(1e-4, 1.5 * ref_truth)
    assert err_raw <= max(2e-4, 2.5 * ref_truth)
    assert err < 5e-3
    assert torch.equal(mel.cpu()[~valid.expand_as(ref)], torch.zeros_like(ref)[~valid.expand_as(ref)])


# ------------------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("M,d", [(1000, 512), (333, 1024), (77, 176)])
def test_layernorm_fwd_bwd(M, d):
    x = rnd(M, d, scale=2.0).requires_grad_(True)
    gamma, beta = (1 + 0.1 * rnd(d)).requires_grad_(True), (0.1 * rnd(d, seed=1)).requires_grad_(True)
    y16 = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    y32 = torch.empty(M, d, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    L.call("layernorm_fwd", x.detach(), gamma.detach(), beta.detach(), y16, y32, mean, rstd, M, d)
    ref = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5)
    report("ln_fwd32", y32, ref.detach(), 1e-5)
    report("ln_fwd16", y16, ref.detach(), 5e-3)
    dy = rnd(M, d, seed=5)
    ref.backward(dy)
    for bf in (0, 1):
        dr = rnd(M, d, seed=9)
        dr0 = dr.clone()
        dg, db = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
        nxt = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
        dyy = dy.to(torch.bfloat16) if bf else dy
        L.call("layernorm_bwd", dyy, bf, x.detach(), gamma.detach(), mean, rstd, dr, 1, dg, db, nxt, 0.5, 0.0, 0, 0,
               M, d)
        tol = 1e-2 if bf else 1e-4
        report(f"ln_bwd_dx(bf16={bf})", dr - dr0, x.grad, tol)
        report(f"ln_bwd_dgamma(bf16={bf})", dg, gamma.grad, tol)
        report(f"ln_bwd_dbeta(bf16={bf})", db, beta.grad, tol)
        report(f"ln_bwd_next(bf16={bf})", nxt, 0.5 * dr, 5e-3)


def test_colsum_cast_permute():
    M, N = 3000, 1000
    a = rnd(M, N, dtype=torch.bfloat16)
    out = torch.ones(N, device=DEV)
    L.call("colsum_bf16", a, out, M, N, N)
    report("colsum", out, 1 + a.float().sum(0), 1e-4)
    w = rnd(6, 10, 14)
    o = torch.empty(10, 6, 14, dtype=torch.bfloat16, device=DEV)
    L.call("permute3_bf16", w, o, 6, 10, 14, 0)
    assert torch.equal(o, w.permute(1, 0, 2).contiguous().to(torch.bfloat16))
    o2 = torch.empty(6, 14, 10, dtype=torch.bfloat16, device=DEV)
    L.call("permute3_bf16", w, o2, 6, 10, 14, 1)
    assert torch.equal(o2, w.permute(0, 2, 1).contiguous().to(torch.bfloat16))
    acc = torch.ones(6, 14, 10, device=DEV)
    L.call("permute3_f32", w, acc, 6, 10, 14, 1, 1)
    assert torch.allclose(acc, 1 + w.permute(0, 2, 1))
    c = torch.empty(6 * 10 * 14, dtype=torch.bfloat16, device=DEV)
    L.call("cast_bf16", w, c, 6 * 10 * 14, 2.0)
    assert torch.equal(c, (w * 2).to(torch.bfloat16).flatten())


def test_relpos_table_and_lengths():
    from oracle import conformer_oracle as O

    T, d = 131, 512
    p16 = torch.empty(2 * T - 1, d, dtype=torch.bfloat16, device=DEV)
    p32 = torch.empty(2 * T - 1, d, device=DEV)
    L.call("relpos_table", p16, p32, T, d)
    ref = O.rel_pos_table(T, d)
    assert (p32.cpu() - ref).abs().max() < 2e-4
    lens = torch.tensor([0, 1, 2, 3, 100, 3000, 3001, 479999], dtype=torch.int64, device=DEV)
    n = lens.numel()
    l32 = torch.empty(n, dtype=torch.int32, device=DEV)
    st = torch.empty(3 * n, dtype=torch.int32, device=DEV)
    o64 = torch.empty(n, dtype=torch.int64, device=DEV)
    L.call("subsample_lengths", lens, l32, st, o64, n, 3)
    for k in (1, 2, 3):
        assert torch.equal(st[(k - 1) * n:k * n].cpu().long(), O.conv_out_len(lens.cpu(), k).long())
    assert torch.equal(o64.cpu(), O.conv_out_len(lens.cpu(), 3).long())


# ------------------------------------------------------------------------------------------------------- conv module
@pytest.mark.parametrize("B,T,d,k", [(3, 131, 512, 31), (2, 75, 256, 9), (2, 40, 1024, 5), (2, 50, 176, 31)])
def test_conv_module_core(B, T, d, k):
    Fn = torch.nn.functional
    M = B * T
    lengths = torch.tensor([T, max(T // 2, 1), 3][:B], dtype=torch.int32, device=DEV)
    valid = (torch.arange(T, device=DEV)[None, :] < lengths[:, None].long()).float()
    u16 = (rnd(B, T, d) * valid[:, :, None]).to(torch.bfloat16)
    w = rnd(d, k, scale=k ** -0.5)
    bias = rnd(d, scale=0.1)
    gamma, beta = 1 + 0.1 * rnd(d, seed=3), 0.1 * rnd(d, seed=4)
    # ---- forward
    c = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    ssum, ssq = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    L.call("dwconv_fwd", u16, w, bias, c, ssum, ssq, B, T, d, k)
    uf = u16.float().requires_grad_(True)
    wf, bf_ = w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    gf, btf = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    cref = Fn.conv1d(Fn.pad(uf.transpose(1, 2), ((k - 1) // 2, (k - 1) // 2)), wf[:, None, :], bf_, groups=d)  # [B,d,T]
    report("dwconv_fwd", c.view(B, T, d), cref.transpose(1, 2).detach(), 1e-2)
    cq = c.float().view(B, T, d)
    report("bn_sum", ssum, cq.sum((0, 1)), 1e-3)
    report("bn_sumsq", ssq, (cq * cq).sum((0, 1)), 1e-3)
    rm, rv = 0.1 * rnd(d, seed=6), 1 + 0.1 * rnd(d, seed=7).abs()
    rm0, rv0 = rm.clone(), rv.clone()
    scale, shift, mean, rstd = (torch.empty(d, device=DEV) for _ in range(4))
    L.call("bn_finalize", ssum, ssq, gamma, beta, rm, rv, scale, shift, mean, rstd, M, d, 1, 0.1, 1e-5, None)
    bm, bv = cq.mean((0, 1)), cq.var((0, 1), unbiased=False)
    report("bn_mean", mean, bm, 1e-3)
    report("bn_rstd", rstd, torch.rsqrt(bv + 1e-5), 1e-3)
    report("bn_running_mean", rm, 0.9 * rm0 + 0.1 * bm, 1e-3)
    report("bn_running_var", rv, 0.9 * rv0 + 0.1 * cq.var((0, 1), unbiased=True), 1e-3)
    s = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    L.call("bn_silu_fwd", c, scale, shift, s, M, d)
    # reference on the same (bf16-rounded) conv output so only this op's error is measured
    cl = c.float().view(B, T, d).requires_grad_(True)
    bn = (cl - cl.mean((0, 1))) * torch.rsqrt(cl.var((0, 1), unbiased=False) + 1e-5) * gf + btf
    sref = Fn.silu(bn)
    report("bn_silu_fwd", s.view(B, T, d), sref.detach(), 1.5e-2)
    # ---- backward
    ds16 = rnd(B, T, d, seed=11).to(torch.bfloat16)
    sref.backward(ds16.float())
    sdz, sdzx = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    L.call("bn_silu_bwd_reduce", ds16, c, scale, shift, mean, rstd, sdz, sdzx, M, d)
    dc = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    dgam, dbet = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    L.call("bn_silu_bwd_apply", ds16, c, gamma, scale, shift, mean, rstd, sdz, sdzx, dc, dgam, dbet, M, d, 1)
    report("bn_bwd_dc", dc.view(B, T, d), cl.grad, 2e-2)
    report("bn_bwd_dgamma", dgam, gf.grad, 2e-2)
    report("bn_bwd_dbeta", dbet, btf.grad, 2e-2)
    dcr = dc.float().view(B, T, d)
    cref.backward(dcr.transpose(1, 2))
    du = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    dw, dbias = torch.zeros(d, k, device=DEV), torch.zeros(d, device=DEV)
    L.call("dwconv_bwd", dc, u16, w, lengths, du, dw, dbias, B, T, d, k)
    report("dwconv_bwd_du", du.view(B, T, d), uf.grad * valid[:, :, None], 1e-2)
    report("dwconv_bwd_dw", dw, wf.grad, 1e-2)
    report("dwconv_bwd_dbias", dbias, bf_.grad, 1e-2)
    # GLU backward
    g16 = rnd(M, 2 * d, seed=13).to(torch.bfloat16)
    dg = torch.empty(M, 2 * d, dtype=torch.bfloat16, device=DEV)
    L.call("glu_bwd", du, g16, dg, M, d)
    a_, b_ = g16.float()[:, 0::2], g16.float()[:, 1::2]
    sg = torch.sigmoid(b_)
    ref = torch.stack([du.float() * sg, du.float() * a_ * sg * (1 - sg)], dim=-1).reshape(M, 2 * d)
    report("glu_bwd", dg, ref, 1e-2)


# ------------------------------------------------------------------------------------------------------- attention
def attn_reference(q, k, v, p, bu, bv, lengths):
    """fp32 restatement of multi_head_attention.py:300-349 + 125-146 on per-head tensors
    q,k,v [B,T,h,dk]; p [2T-1,h,dk]; returns ctx [B,T,h*dk]."""
    B, T, h, dk = q.shape
    qu = (q + bu).to(torch.bfloat16).float().transpose(1, 2)  # kernels round q+bias to bf16 like autocast does
    qv = (q + bv).to(torch.bfloat16).float().transpose(1, 2)
    kk, vv = k.transpose(1, 2), v.transpose(1, 2)
    ac = qu @ kk.transpose(-1, -2)
    bd_full = qv @ p.permute(1, 2, 0)[None]  # [B,h,T,2T-1]
    idx = (T - 1 - torch.arange(T, device=q.device)[:, None] + torch.arange(T, device=q.device)[None, :])
    bd = torch.gather(bd_full, 3, idx[None, None].expand(B, h, T, T))
    scores = (ac + bd) / math.sqrt(dk)
    valid = torch.arange(T, device=q.device)[None, :] < lengths[:, None].long()
    mask = ~(valid[:, None, :, None] & valid[:, None, None, :])
    scores = scores.masked_fill(mask, -10000.0)
    attn = torch.softmax(scores, -1).masked_fill(mask, 0.0)
    return (attn @ vv).transpose(1, 2).reshape(B, T, h * dk)


@pytest.mark.parametrize("B,T,h,dk", [(3, 26, 4, 64), (2, 64, 2, 64), (3, 131, 4, 64), (2, 200, 2, 128), (2, 751, 2, 64)])
def test_relpos_attention_fwd_bwd(B, T, h, dk):
    d = h * dk
    lengths = torch.tensor([T, (2 * T) // 3, 5][:B], dtype=torch.int32, device=DEV)
    qkv16 = rnd(B * T, 3 * d, scale=1.0).to(torch.bfloat16)
    pos16 = rnd(2 * T - 1, d, scale=1.0, seed=3).to(torch.bfloat16)
    bu, bv = rnd(h, dk, scale=0.3, seed=4), rnd(h, dk, scale=0.3, seed=5)
    ctx = torch.empty(B * T, d, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, h, T, device=DEV)
    ctx32 = torch.empty(B * T, d, device=DEV)
    L.call("relpos_attn_fwd", qkv16, pos16, bu, bv, lengths, ctx, ctx32, lse, B, T, h, dk, 0.0, 0, 0)
    qkvf = qkv16.float().view(B, T, 3, h, dk)
    q = qkvf[:, :, 0].clone().requires_grad_(True)
    k = qkvf[:, :, 1].clone().requires_grad_(True)
    v = qkvf[:, :, 2].clone().requires_grad_(True)
    p = pos16.float().view(2 * T - 1, h, dk).clone().requires_grad_(True)
    bur, bvr = bu.clone().requires_grad_(True), bv.clone().requires_grad_(True)
    ref = attn_reference(q, k, v, p, bur, bvr, lengths)
    report(f"attn_fwd_T{T}_dk{dk}", ctx.view(B, T, d), ref.detach(), 2e-2)
    # rows beyond the valid length are exactly zero (second masked_fill, multi_head_attention.py:137-138)
    for b in range(B):
        assert ctx.view(B, T, d)[b, int(lengths[b]):].abs().max().item() == 0 if int(lengths[b]) < T else True
    # ---- backward
    dctx16 = rnd(B * T, d, seed=8).to(torch.bfloat16)
    ref.backward(dctx16.float().view(B, T, d))
    dqkv = torch.zeros(B * T, 3 * d, dtype=torch.bfloat16, device=DEV)
    dpos = torch.zeros(2 * T - 1, d, device=DEV)
    dbu, dbv = torch.zeros(h, dk, device=DEV), torch.zeros(h, dk, device=DEV)
    nbytes = L.lib.b2c_relpos_attn_bwd_workspace(B, T, h)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    L.call("relpos_attn_bwd", qkv16, pos16, bu, bv, lengths, ctx, ctx32, dctx16, lse, dqkv, dpos, dbu, dbv, ws, nbytes, B, T,
           h, dk, 0.0, 0, 0)
    g = dqkv.float().view(B, T, 3, h, dk)
    report("attn_bwd_dq", g[:, :, 0], q.grad, 3e-2)
    report("attn_bwd_dk", g[:, :, 1], k.grad, 3e-2)
    report("attn_bwd_dv", g[:, :, 2], v.grad, 3e-2)
    report("attn_bwd_dpos", dpos.view(2 * T - 1, h, dk), p.grad, 3e-2)
    report("attn_bwd_dbias_u", dbu, bur.grad, 3e-2)
    report("attn_bwd_dbias_v", dbv, bvr.grad, 3e-2)


def test_relpos_attention_dropout_statistics():
    B, T, h, dk = 2, 192, 2, 64
    d = h * dk
    lengths = torch.tensor([T, T], dtype=torch.int32, device=DEV)
    qkv16 = rnd(B * T, 3 * d, scale=0.2).to(torch.bfloat16)
    pos16 = rnd(2 * T - 1, d, scale=0.2, seed=3).to(torch.bfloat16)
    bu, bv = rnd(h, dk, scale=0.1, seed=4), rnd(h, dk, scale=0.1, seed=5)
    outs = []
    for seed in (0, 1, 2, 3):
        ctx = torch.empty(B * T, d, dtype=torch.bfloat16, device=DEV)
        lse = torch.empty(B, h, T, device=DEV)
        L.call("relpos_attn_fwd", qkv16, pos16, bu, bv, lengths, ctx, None, lse, B, T, h, dk, 0.1 if seed else 0.0, seed, 5)
        outs.append(ctx.float())
    # dropout is unbiased: averaging a few seeds approaches the p=0 output; and differs per seed
    assert not torch.equal(outs[1], outs[2])
    mean_drop = (outs[1] + outs[2] + outs[3]) / 3
    e = ((mean_drop - outs[0]).abs().mean() / outs[0].abs().mean()).item()
    e1 = ((outs[1] - outs[0]).abs().mean() / outs[0].abs().mean()).item()
    print("attn dropout: mean-of-3 vs no-drop rel-L1", e, "single", e1)
    assert e < 0.7 * e1 and e1 < 0.5  # averaging seeds converges towards the p=0 output (unbiased mask + 1/(1-p))


# ------------------------------------------------------------------------------------------------------- subsampling
@pytest.mark.parametrize("B,F,T,C", [(2, 80, 131, 128), (3, 80, 301, 256), (2, 80, 64, 512), (2, 80, 203, 1024)])
def test_subsampling_convs(B, F, T, C):
    Fn = torch.nn.functional
    lengths = torch.tensor([T, max(T - 37, 3), 9][:B], dtype=torch.int64, device=DEV)
    n = 2
    l32 = torch.empty(B, dtype=torch.int32, device=DEV)
    st = torch.empty(n * B, dtype=torch.int32, device=DEV)
    o64 = torch.empty(B, dtype=torch.int64, device=DEV)
    L.call("subsample_lengths", lengths, l32, st, o64, B, n)
    len1, len2 = st[:B], st[B:]
    T1, F1 = (T - 1) // 2 + 1, (F - 1) // 2 + 1
    T2, F2 = (T1 - 1) // 2 + 1, (F1 - 1) // 2 + 1
    mel = rnd(B, F, T)
    tmask = lambda Tn, ln: (torch.arange(Tn, device=DEV)[None, :] < ln[:, None].long()).float()
    mel = mel * tmask(T, l32)[:, None, :]
    w1, b1 = rnd(C, 1, 3, 3, scale=0.3), rnd(C, scale=0.1)
    w2, b2 = rnd(C, C, 3, 3, scale=(9 * C) ** -0.5, seed=2), rnd(C, scale=0.1, seed=3)
    # ---- conv1 forward
    x1 = torch.empty(B, T1, F1, C, dtype=torch.bfloat16, device=DEV)
    L.call("subsample_conv1_fwd", mel, w1, b1, l32, len1, x1, B, F, T, C)
    xin = mel.transpose(1, 2)[:, None]  # [B,1,T,F]
    w1r = w1.to(torch.bfloat16).float().requires_grad_(True)
    b1r = b1.clone().requires_grad_(True)
    r1 = Fn.relu(Fn.conv2d(xin.to(torch.bfloat16).float(), w1r, b1r, stride=2, padding=1)) * tmask(T1, len1)[:, None, :, None]
    report("conv1_fwd", x1, r1.permute(0, 2, 3, 1).detach(), 1e-2)
    # ---- conv2 forward (implicit GEMM on tcgen05)
    w2p = torch.empty(C, 9, C, dtype=torch.bfloat16, device=DEV)
    L.call("permute3_bf16", w2, w2p, C, C, 9, 1)
    x2 = torch.empty(B, T2, F2, C, dtype=torch.bfloat16, device=DEV)
    L.call("conv3x3s2_fwd", x1, w2p, b2, len2, x2, B, T1, F1, C)
    x1f = x1.float().permute(0, 3, 1, 2).requires_grad_(True)  # [B,C,T1,F1]
    w2r = w2.to(torch.bfloat16).float().requires_grad_(True)
    b2r = b2.clone().requires_grad_(True)
    pre2 = Fn.conv2d(x1f, w2r, b2r, stride=2, padding=1)
    r2 = Fn.relu(pre2) * tmask(T2, len2)[:, None, :, None]
    report("conv2_fwd", x2, r2.permute(0, 2, 3, 1).detach(), 1e-2)
    # ---- conv2 backward: dy (already gated by relu/mask like the engine produces it)
    dy = (rnd(B, T2, F2, C, seed=7) * (x2.float() > 0)).to(torch.bfloat16)
    pre2.backward(dy.float().permute(0, 3, 1, 2))
    tmp = torch.empty(C, C, 9, device=DEV)
    L.call("permute3_f32", w2, tmp, C, C, 9, 0, 0)
    wd = torch.empty(C, 9, C, dtype=torch.bfloat16, device=DEV)
    L.call("permute3_bf16", tmp, wd, C, C, 9, 1)
    dx1 = torch.full((B, T1, F1, C), 7.0, dtype=torch.bfloat16, device=DEV)
    L.call("conv3x3s2_bwd_data", dy, wd, x1, dx1, B, T1, F1, C)
    ref_dx1 = x1f.grad.permute(0, 2, 3, 1) * (x1.float() > 0)
    report("conv2_bwd_data", dx1, ref_dx1, 1.5e-2)
    gw = torch.zeros(C, 9 * C, device=DEV)
    L.call("conv3x3s2_bwd_weight", dy, x1, gw, B, T1, F1, C)
    report("conv2_bwd_weight", gw.view(C, 9, C).permute(0, 2, 1).reshape(C, C, 3, 3), w2r.grad, 1e-2)
    # ---- conv1 backward
    r1.backward(dx1.float().permute(0, 3, 1, 2) * (x1.float().permute(0, 3, 1, 2) > 0))
    dw1, db1 = torch.zeros(C, 9, device=DEV), torch.zeros(C, device=DEV)
    L.call("subsample_conv1_bwd", mel, dx1, l32, dw1, db1, B, F, T, C)
    report("conv1_bwd_dw", dw1.view(C, 1, 3, 3), w1r.grad, 1e-2)
    report("conv1_bwd_db", db1, b1r.grad, 1e-2)


@pytest.mark.parametrize("B,F,T,C", [(2, 80, 131, 128), (3, 80, 301, 512)])
def test_stem_as_im2col_gemm_matches_the_stencil_kernels(B, F, T, C):
    """Conv2d(1→C) stem through the tensor-core route (im2col + K=16 GEMM) against the direct stencil kernels of the same
    library (which test_subsampling_convs pins to torch): forward bit-compatible up to bf16 rounding, weight / bias
    gradient from the GEMM's [C,16] output (column 9 = bias gradient)."""
    lengths = torch.tensor([T, max(T - 37, 3), 9][:B], dtype=torch.int64, device=DEV)
    l32 = torch.empty(B, dtype=torch.int32, device=DEV)
    st = torch.empty(2 * B, dtype=torch.int32, device=DEV)
    o64 = torch.empty(B, dtype=torch.int64, device=DEV)
    L.call("subsample_lengths", lengths, l32, st, o64, B, 2)
    len1 = st[:B]
    T1, F1 = (T - 1) // 2 + 1, (F - 1) // 2 + 1
    mel = rnd(B, F, T)  # NOT pre-masked: both routes apply the input length mask themselves
    w1, b1 = rnd(C, 1, 3, 3, scale=0.3), rnd(C, scale=0.1)
    x_ref = torch.empty(B, T1, F1, C, dtype=torch.bfloat16, device=DEV)
    L.call("subsample_conv1_fwd", mel, w1, b1, l32, len1, x_ref, B, F, T, C)
    rows = B * T1 * F1
    patches = torch.empty(rows, 16, dtype=torch.bfloat16, device=DEV)
    L.call("subsample_im2col", mel, l32, patches, B, F, T)
    assert (patches[:, 9].float() == 1).all() and (patches[:, 10:].float() == 0).all()
    w16 = torch.zeros(C, 16, device=DEV)
    w16[:, :9] = w1.reshape(C, 9)
    w16 = w16.to(torch.bfloat16)
    x = torch.empty(B, T1, F1, C, dtype=torch.bfloat16, device=DEV)
    L.gemm(patches, w16, rows, C, 16, 16, 16, 0, 0, L.EPI_RELU_MASK, out0=x, bias=b1, lengths=len1, rows_per_batch=T1 * F1,
           rows_inner=F1)
    report("stem_gemm_fwd", x, x_ref.float(), 1e-2)
    dy = (rnd(B, T1, F1, C, seed=11) * (x_ref.float() > 0)).to(torch.bfloat16)
    dw_ref, db_ref = torch.zeros(C, 9, device=DEV), torch.zeros(C, device=DEV)
    L.call("subsample_conv1_bwd", mel, dy, l32, dw_ref, db_ref, B, F, T, C)
    g16 = torch.zeros(C, 16, device=DEV)
    L.gemm(dy, patches, C, 16, rows, C, 16, 1, 1, L.EPI_ACCUM_F32, out0=g16, ldo=16)
    report("stem_gemm_dw", g16[:, :9], dw_ref, 1e-2)
    report("stem_gemm_db", g16[:, 9], db_ref, 1e-2)


@pytest.mark.parametrize("slope", [1.0, -1.0])
def test_relpos_attention_scores_drifting_across_key_tiles(slope):
    """The forward moves its softmax reference maximum lazily (from the PREVIOUS key tile's row maximum): logits that
    climb (or fall) by ~20 nats per 64-key tile exercise the rescale path and the 2^(v - m_ref) >> 1 range."""
    B, T, h, dk = 2, 600, 2, 64
    d = h * dk
    lengths = torch.tensor([T, 437], dtype=torch.int32, device=DEV)
    qkvf = rnd(B, T, 3, h, dk, scale=1.0)
    ramp = torch.linspace(0, 1, T, device=DEV) * 750.0 * slope  # (q0 ≈ 2) * ramp / sqrt(64): ~190 nats over 9.4 tiles
    qkvf[:, :, 0, :, 0] = 2.0
    qkvf[:, :, 1, :, 0] = ramp[None, :, None]
    qkv16 = qkvf.reshape(B * T, 3 * d).to(torch.bfloat16)
    pos16 = rnd(2 * T - 1, d, scale=1.0, seed=3).to(torch.bfloat16)
    bu, bv = rnd(h, dk, scale=0.3, seed=4), rnd(h, dk, scale=0.3, seed=5)
    ctx = torch.empty(B * T, d, dtype=torch.bfloat16, device=DEV)
    ctx32 = torch.empty(B * T, d, device=DEV)
    lse = torch.empty(B, h, T, device=DEV)
    L.call("relpos_attn_fwd", qkv16, pos16, bu, bv, lengths, ctx, ctx32, lse, B, T, h, dk, 0.0, 0, 0)
    q3 = qkv16.float().view(B, T, 3, h, dk)
    ref = attn_reference(q3[:, :, 0], q3[:, :, 1], q3[:, :, 2], pos16.float().view(2 * T - 1, h, dk), bu, bv, lengths)
    report(f"attn_fwd_drift_{slope:+.0f}", ctx32.view(B, T, d), ref, 2e-2)
    report(f"attn_fwd_drift_bf16_{slope:+.0f}", ctx.view(B, T, d), ref, 2e-2)


@pytest.mark.parametrize("time_width", [10, 0.05])
def test_spec_augment_module_matches_the_oracle(time_width):
    """B200SpecAugment (one CUDA pass) against the oracle restatement of the reference class under the same RNG state:
    bit-identical output in the vectorised (device RNG) and the legacy (host RNG) modes."""
    import random

    from nemo_b200.spec_augment import B200SpecAugment
    from oracle import conformer_oracle as O

    x = rnd(3, 80, 301)
    ln = torch.tensor([301, 250, 17], device=DEV)
    sa = B200SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=time_width, mask_value=0.0)
    torch.manual_seed(11)
    got = sa(input_spec=x, length=ln)
    torch.manual_seed(11)
    want = O.spec_augment_vectorized(x, ln, 2, 5, 27, time_width)
    assert torch.equal(got, want) and 0.05 < (got == 0).float().mean().item() < 0.9
    sa = B200SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=time_width, rng=random.Random(5),
                         mask_value=-1.5, use_vectorized_code=False)
    got = sa(input_spec=x, length=ln)
    want = O.spec_augment_legacy(x, ln, random.Random(5), 2, 5, 27, time_width, mask_value=-1.5)
    assert torch.equal(got, want)
    # no masks at all: identity
    assert torch.equal(B200SpecAugment()(input_spec=x, length=ln), x)


def test_gemm_epilogues_on_odd_tile_counts_narrow_n_and_shared_dropout_masks():
    """3 row tiles (the CTA pair's second tile of the last pair is a dummy), N = 384 (128-wide tiles, 3 of them), ragged
    K; the SiLU forward and the SiLU' backward epilogues must drop exactly the same elements for the same (seed, site)."""
    M, N, K, p = 333, 384, 200, 0.2
    A, B, bias = rnd(M, K, dtype=torch.bfloat16), rnd(N, K, scale=0.08, dtype=torch.bfloat16), rnd(N)
    z = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    h = torch.empty_like(z)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_SILU, out0=z, out1=h, bias=bias, seed=9, drop_site=21, drop_p=p)
    zr = A.float() @ B.float().T + bias
    report("silu_z_odd", z, zr, 1e-2)
    h_full = torch.nn.functional.silu(z.float()).to(torch.bfloat16).float()
    kept = h.float() != 0
    frac = (kept | (h_full == 0)).float().mean().item()
    print("keep fraction", frac)
    assert abs(frac - (1 - p)) < 0.01
    assert torch.allclose(h.float()[kept], (h_full / (1 - p))[kept], rtol=1e-2, atol=1e-3)
    G, W = rnd(M, K, seed=5, dtype=torch.bfloat16), rnd(N, K, scale=0.08, seed=6, dtype=torch.bfloat16)
    dz = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    L.gemm(G, W, M, N, K, K, K, 0, 0, L.EPI_DSILU, out0=dz, aux=z, seed=9, drop_site=21, drop_p=p)
    zf = z.float()
    s = torch.sigmoid(zf)
    full = (G.float() @ W.float().T).to(torch.bfloat16).float() * (s * (1 + zf * (1 - s))) / (1 - p)
    mask = kept | (h_full == 0)  # elements the forward kept (h_full == 0 cannot tell; treat as kept and compare loosely)
    report("dsilu_odd_kept", dz.float() * kept, full * kept, 1.5e-2)
    dropped = ~mask
    assert dz.float()[dropped].abs().max().item() == 0.0
    # plain bf16 and the fp32 residual epilogue on the same shape
    o = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_BF16, out0=o, bias=bias)
    report("bf16_odd", o, zr, 1e-2)
    x = rnd(M, N)
    r = torch.empty_like(x)
    L.gemm(A, B, M, N, K, K, K, 0, 0, L.EPI_RESID_F32, out0=r, aux=x, bias=bias, alpha=0.5)
    report("resid_odd", r, x + 0.5 * zr, 5e-3)


@pytest.mark.parametrize("B,T1,F1,C", [(3, 101, 20, 64), (2, 203, 40, 256), (2, 77, 10, 24)])
def test_depthwise_3x3_stride2_kernels_vs_torch(B, T1, F1, C):
    """dw_striding stages (subsampling.py:142-215): depthwise Conv2d(C, 3x3, s2, p1, groups=C) + bias with the output time mask,
    its data gradient (through the producer's ReLU gate) and weight / bias gradients against F.conv2d autograd on the same
    bf16-rounded operands.  C = 64 / 256 take the 8-channel kernels, C = 24 (C/8 = 3 does not divide the CTA) the 2-channel ones."""
    Fn = torch.nn.functional
    T2, F2 = (T1 - 1) // 2 + 1, (F1 - 1) // 2 + 1
    lin = torch.tensor([T1, max(T1 - 31, 3), 9][:B], device=DEV)
    lout = ((lin - 1) // 2 + 1).to(torch.int32)
    x = torch.relu(rnd(B, T1, F1, C)) * (torch.arange(T1, device=DEV)[None, :] < lin[:, None]).float()[:, :, None, None]
    x16 = x.to(torch.bfloat16)
    w, bias = rnd(C, 1, 3, 3, scale=0.3, seed=2), rnd(C, scale=0.1, seed=3)
    y = torch.empty(B, T2, F2, C, dtype=torch.bfloat16, device=DEV)
    L.call("dwconv3x3s2_fwd", x16, w.contiguous(), bias, lout, y, B, T1, F1, C)
    xr = x16.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    m_out = (torch.arange(T2, device=DEV)[None, :] < lout[:, None]).float()[:, None, :, None]
    yr = Fn.conv2d(xr, wr, br, stride=2, padding=1, groups=C) * m_out
    report(f"dw3x3_fwd_C{C}", y, yr.detach().permute(0, 2, 3, 1), 1e-2)
    dy = rnd(B, T2, F2, C, seed=7).to(torch.bfloat16)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.full((B, T1, F1, C), 7.0, dtype=torch.bfloat16, device=DEV)
    dw, db = torch.zeros(C, 9, device=DEV), torch.zeros(C, device=DEV)
    L.call("dwconv3x3s2_bwd", dy, x16, w.contiguous(), lout, dx, dw, db, B, T1, F1, C)
    report(f"dw3x3_bwd_dx_C{C}", dx, xr.grad.permute(0, 2, 3, 1) * (x16.float() > 0), 1.5e-2)
    report(f"dw3x3_bwd_dw_C{C}", dw.view(C, 1, 3, 3), wr.grad, 1e-2)
    report(f"dw3x3_bwd_db_C{C}", db, br.grad, 1e-2)
