import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nemo_b200.encoder import B200ConformerEncoder
from nemo_b200.preprocessor import B200MelPreprocessor
from oracle import conformer_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = torch.device("cuda:0")
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()
g = torch.Generator().manual_seed(1)
nhwc = lambda t: t.permute(0, 2, 3, 1)
tm = O._time_mask

def gate_report(name, ours, ref):
    mism = (ours > 0) != (ref > 0)
    print(f"{name}: shape {tuple(ours.shape)} max|diff| {(ours - ref).abs().max().item():.3e} gate mismatches {int(mism.sum())}")
    idx = mism.nonzero()[:8]
    for i in idx:
        i = tuple(i.tolist())
        print("    at", i, "ours", ours[i].item(), "ref", ref[i].item())

# ---- (A) dw_striding n_layers=0
C = 64
cfg = dict(n_layers=0, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="dw_striding", subsampling_factor=8, subsampling_conv_channels=C)
sd = O.init_state_dict(seed=3, **cfg)
enc = B200ConformerEncoder(feat_in=80, n_layers=0, d_model=128, n_heads=2, conv_kernel_size=9, subsampling="dw_striding", subsampling_factor=8,
                           subsampling_conv_channels=C, dropout=0.0, dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
enc.load_state_dict(sd); enc = enc.to(DEV).train(); enc.precision = "fp32"
x = torch.randn(3, 80, 403, generator=g)
ln = torch.tensor([403, 300, 77])
for b in range(3):
    x[b, :, ln[b]:] = 0
x, ln = x.to(DEV), ln.to(DEV)
enc.__dict__["_debug_parity"] = dbg = {}
y, yl = enc(audio_signal=x, length=ln)
r = torch.randn(y.shape, generator=g).to(DEV)
(y * r).sum().backward()
P = {k: v.to(DEV) for k, v in sd.items()}
cur = ln.clone().float()
h = tm(x.transpose(1, 2).unsqueeze(1), cur.long())
h = F.conv2d(h, P["pre_encode.conv.0.weight"], P["pre_encode.conv.0.bias"], stride=2, padding=1)
cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1
h = F.relu(tm(h, cur.long())); h = tm(h, cur.long())
gate_report("dw x1", dbg["x1"], nhwc(h))
hd = F.conv2d(h, P["pre_encode.conv.2.weight"], P["pre_encode.conv.2.bias"], stride=2, padding=1, groups=C)
cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1
hd = tm(hd, cur.long())
hp = tm(F.conv2d(hd, P["pre_encode.conv.3.weight"], P["pre_encode.conv.3.bias"]), cur.long())
h = tm(F.relu(hp), cur.long())
gate_report("dw xout0", dbg["xout0"], nhwc(h))
print("bias conv.3 min/max", P["pre_encode.conv.3.bias"].min().item(), P["pre_encode.conv.3.bias"].max().item())

# ---- (B) striding at Conformer-L width, 1 layer, 2 x 30 s ragged (the failing shape)
cfg = dict(n_layers=1, d_model=512, n_heads=8, conv_kernel_size=31, subsampling_kind="striding", subsampling_factor=4, subsampling_conv_channels=512)
sd = O.init_state_dict(seed=11, **cfg)
enc = B200ConformerEncoder(feat_in=80, n_layers=1, d_model=512, n_heads=8, conv_kernel_size=31, dropout=0.0, dropout_pre_encoder=0.0,
                           dropout_emb=0.0, dropout_att=0.0)
enc.load_state_dict(sd); enc = enc.to(DEV).train(); enc.precision = "fp32"
pre = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0).to(DEV).eval()
wav, wlen = O.synthetic_waveforms(2, 30.0, ragged=True)
wlen[-1] = 30 * 16000 - 160 * 411
mel, mlen = pre(input_signal=wav.to(DEV), length=wlen.to(DEV))
enc.__dict__["_debug_parity"] = dbg = {}
yy, _ = enc(audio_signal=mel, length=mlen)
r = torch.randn(yy.shape, generator=g).to(DEV)
(yy * r).sum().backward()
params = {k: (v.to(DEV).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.to(DEV)) for k, v in sd.items()}
y_ref, _ = O.encoder_forward(params, mel, mlen, n_layers=1, d_model=512, n_heads=8, training=True)
(y_ref * r).sum().backward()
print("striding L-width fwd", rel(yy, y_ref))
for k, p in enc.named_parameters():
    if k.startswith("pre_encode") or "norm_feed_forward1" in k:
        print(f"  {k:48s} rel {rel(p.grad, params[k].grad):.3e}")
P = {k: v.to(DEV) for k, v in sd.items()}
cur = mlen.clone().float()
h = tm(mel.transpose(1, 2).unsqueeze(1), cur.long())
for idx in (0, 2):
    h = tm(h, cur.long())
    h = F.conv2d(h, P[f"pre_encode.conv.{idx}.weight"], P[f"pre_encode.conv.{idx}.bias"], stride=2, padding=1)
    cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1
    h = F.relu(tm(h, cur.long()))
    gate_report(f"striding conv.{idx} out", dbg["x1"] if idx == 0 else dbg["xl"], nhwc(tm(h, cur.long())))
