"""Launch the hot kernels once or twice at the bench shapes (Conformer-CTC-Large, 32 x 30 s: M = 24000 tokens, T = 750)
through the C ABI, so one `ncu --set full` run captures all of them without paying for a whole training step.
  ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels -f python tools/prof_kernels.py
Also prints CUDA-event timings (not under ncu: python tools/prof_kernels.py --time)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemo_b200 import _lib as L  # noqa: E402

DEV = torch.device("cuda:0")
B, T, H, DK, D, FF = 32, 750, 8, 64, 512, 2048
M = B * T
TIME = "--time" in sys.argv
REPS = 20 if TIME else 1


def rnd(*s, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*s, device=DEV) * scale).to(dtype)


def timed(name, fn, flops=None, bytes_=None):
    fn()
    if not TIME:
        return
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    tot = 0.0
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / REPS * 1e3
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{name:34s} {us:9.1f} us{extra}", flush=True)


x16 = rnd(M, D)
w1, w2 = rnd(FF, D, scale=0.05), rnd(D, FF, scale=0.05)
b1, b2 = torch.zeros(FF, device=DEV), torch.zeros(D, device=DEV)
z, hcopy = torch.empty(M, FF, dtype=torch.bfloat16, device=DEV), torch.empty(M, FF, dtype=torch.bfloat16, device=DEV)
resid = torch.randn(M, D, device=DEV)
resid_out = torch.empty(M, D, device=DEV)
out16 = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=DEV)
wqkv = rnd(3 * D, D, scale=0.05)
lengths = torch.full((B,), T, dtype=torch.int32, device=DEV)
gw = torch.zeros(FF, D, device=DEV)
dz = rnd(M, FF)

timed("ffn1 fwd  silu  24000x2048x512 p=.1", lambda: L.gemm(x16, w1, M, FF, D, D, D, 0, 0, L.EPI_SILU, out0=z, out1=hcopy, bias=b1,
                                                              seed=1, drop_site=5, drop_p=0.1), 2.0 * M * FF * D)
timed("ffn2 fwd  resid 24000x512x2048", lambda: L.gemm(hcopy, w2, M, D, FF, FF, FF, 0, 0, L.EPI_RESID_F32, out0=resid_out,
                                                        aux=resid, bias=b2, alpha=0.5), 2.0 * M * FF * D)
timed("qkv  fwd  bf16  24000x1536x512", lambda: L.gemm(x16, wqkv, M, 3 * D, D, D, D, 0, 0, L.EPI_BF16, out0=out16), 2.0 * M * 3 * D * D)
dzz = torch.empty(M, FF, dtype=torch.bfloat16, device=DEV)
timed("ffn2 dgrad dsilu 24000x2048x512 p=.1", lambda: L.gemm(x16, w2, M, FF, D, D, FF, 0, 1, L.EPI_DSILU, out0=dzz, aux=z, seed=1,
                                                               drop_site=5, drop_p=0.1), 2.0 * M * FF * D)
timed("ffn2 dgrad bf16 24000x2048x512", lambda: L.gemm(x16, w2, M, FF, D, D, FF, 0, 1, L.EPI_BF16, out0=dzz), 2.0 * M * FF * D)
timed("ffn1 dgrad bf16 24000x512x2048", lambda: L.gemm(dz, w1, M, D, FF, FF, D, 0, 1, L.EPI_BF16, out0=x16.clone()), 2.0 * M * FF * D)
timed("ffn1 wgrad acc  2048x512x24000", lambda: L.gemm(dz, x16, FF, D, M, FF, D, 1, 1, L.EPI_ACCUM_F32, out0=gw), 2.0 * M * FF * D)
gw2 = torch.zeros(D, FF, device=DEV)
timed("ffn2 wgrad acc  512x2048x24000", lambda: L.gemm(x16, dz, D, FF, M, D, FF, 1, 1, L.EPI_ACCUM_F32, out0=gw2), 2.0 * M * FF * D)

# ---- attention
qkv = rnd(M, 3 * D)
pos = rnd(2 * T - 1, D)
bu, bv = torch.randn(H, DK, device=DEV) * 0.3, torch.randn(H, DK, device=DEV) * 0.3
ctx = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
ctx32 = torch.empty(M, D, device=DEV)
lse = torch.empty(B, H, T, device=DEV)
attn_flops_f = 2.0 * B * H * T * T * DK * 2 + 2.0 * B * H * T * (T + 128) * DK
timed("relpos_attn_fwd (p=0.1)", lambda: L.call("relpos_attn_fwd", qkv, pos, bu, bv, lengths, ctx, ctx32, lse, B, T, H, DK, 0.1, 1, 3),
      attn_flops_f)
dctx = rnd(M, D)
dqkv = torch.zeros(M, 3 * D, dtype=torch.bfloat16, device=DEV)
dpos = torch.zeros(2 * T - 1, D, device=DEV)
dbu, dbv = torch.zeros(H, DK, device=DEV), torch.zeros(H, DK, device=DEV)
nbytes = L.lib.b2c_relpos_attn_bwd_workspace(B, T, H)
ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
timed("relpos_attn_bwd (q+kv+pos, p=0.1)",
      lambda: L.call("relpos_attn_bwd", qkv, pos, bu, bv, lengths, ctx, ctx32, dctx, lse, dqkv, dpos, dbu, dbv, ws, nbytes, B, T, H, DK,
                     0.1, 1, 3), 2.5 * attn_flops_f)
# ---- HBM-bound kernels: LayerNorm backward, depthwise conv backward (k = 31), mel front-end (32 x 30 s)
xr = torch.randn(M, D, device=DEV)
gam = torch.ones(D, device=DEV)
mean, rstd = torch.zeros(M, device=DEV), torch.ones(M, device=DEV)
drr = torch.randn(M, D, device=DEV)
dgam, dbet = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
nxt = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
dy16 = rnd(M, D)
timed("layernorm_bwd 24000x512 (bf16 dy, accumulate, next)",
      lambda: L.call("layernorm_bwd", dy16, 1, xr, gam, mean, rstd, drr, 1, dgam, dbet, nxt, 1.0, 0.1, 1, 3, M, D),
      bytes_=M * D * (2 + 4 + 8 + 2))
K = 31
u16, dc16, du16 = rnd(M, D), rnd(M, D), torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
wdw, gdw, gdb = torch.randn(D, K, device=DEV) * 0.1, torch.zeros(D, K, device=DEV), torch.zeros(D, device=DEV)
timed("dwconv_bwd k=31 24000x512", lambda: L.call("dwconv_bwd", dc16, u16, wdw, lengths, du16, gdw, gdb, B, T, D, K), bytes_=M * D * 6)
c16 = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
bs, bq = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
timed("dwconv_fwd k=31 24000x512", lambda: L.call("dwconv_fwd", u16, wdw, gdb, c16, bs, bq, B, T, D, K), bytes_=M * D * 4)
from nemo_b200.preprocessor import B200MelPreprocessor  # noqa: E402

pre = B200MelPreprocessor(sample_rate=16000, window_size=0.025, window_stride=0.01, features=80, n_fft=512, dither=1e-5, pad_to=0).to(DEV).eval()
wav = (torch.rand(32, 480000, device=DEV) * 0.2 - 0.1)
wl = torch.full((32,), 480000, dtype=torch.int64, device=DEV)
timed("mel front-end 32x30 s (frames + normalize)", lambda: pre(input_signal=wav, length=wl), bytes_=32 * (4 * 480000 + 4 * 80 * 3001))
# ---- dw_striding depthwise 3x3 stride-2 kernels at the FastConformer shapes (16 x 30 s per GPU, C = 256)
Bq, Cq = 16, 256
for (T1, F1) in ((1501, 40), (751, 20)):
    T2, F2 = (T1 - 1) // 2 + 1, (F1 - 1) // 2 + 1
    xin = rnd(Bq, T1, F1, Cq)
    wq, bq_ = torch.randn(Cq, 9, device=DEV) * 0.2, torch.zeros(Cq, device=DEV)
    lo = torch.full((Bq,), T2, dtype=torch.int32, device=DEV)
    yq = torch.empty(Bq, T2, F2, Cq, dtype=torch.bfloat16, device=DEV)
    dxq = torch.empty(Bq, T1, F1, Cq, dtype=torch.bfloat16, device=DEV)
    gwq, gbq = torch.zeros(Cq, 9, device=DEV), torch.zeros(Cq, device=DEV)
    nb_in, nb_out = Bq * T1 * F1 * Cq * 2, Bq * T2 * F2 * Cq * 2
    timed(f"dwconv3x3s2_fwd  [{Bq},{T1},{F1},{Cq}]", lambda: L.call("dwconv3x3s2_fwd", xin, wq, bq_, lo, yq, Bq, T1, F1, Cq), bytes_=nb_in + nb_out)
    timed(f"dwconv3x3s2_bwd  [{Bq},{T1},{F1},{Cq}] (data+weight)",
          lambda: L.call("dwconv3x3s2_bwd", yq, xin, wq, lo, dxq, gwq, gbq, Bq, T1, F1, Cq), bytes_=3 * nb_in + 2 * nb_out)
torch.cuda.synchronize()
print("done")
