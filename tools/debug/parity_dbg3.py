import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nemo_b200.encoder import B200ConformerEncoder
from oracle import conformer_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = torch.device("cuda:0")
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()
g = torch.Generator().manual_seed(1)
# ---- (A) dw_striding, n_layers = 0: intermediates
C = 64
cfg = dict(n_layers=0, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="dw_striding", subsampling_factor=8, subsampling_conv_channels=C)
sd = O.init_state_dict(seed=3, **cfg)
enc = B200ConformerEncoder(feat_in=80, n_layers=0, d_model=128, n_heads=2, conv_kernel_size=9, subsampling="dw_striding", subsampling_factor=8,
                           subsampling_conv_channels=C, dropout=0.0, dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
enc.load_state_dict(sd)
enc = enc.to(DEV).train()
enc.precision = "fp32"
x = torch.randn(3, 80, 403, generator=g)
ln = torch.tensor([403, 300, 77])
for b in range(3):
    x[b, :, ln[b]:] = 0
x, ln = x.to(DEV), ln.to(DEV)
enc.__dict__["_debug_parity"] = dbg = {}
y, yl = enc(audio_signal=x, length=ln)
r = torch.randn(y.shape, generator=g).to(DEV)
(y * r).sum().backward()
# inline oracle with retained intermediates
P = {k: v.to(DEV).clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
tm = O._time_mask
cur = ln.clone().float()
h = x.transpose(1, 2).unsqueeze(1)
h = tm(h, cur.long())
h = F.conv2d(h, P["pre_encode.conv.0.weight"], P["pre_encode.conv.0.bias"], stride=2, padding=1)
cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1
h = F.relu(tm(h, cur.long()))
inter = {}
idx = 2
for s in range(2):
    h = tm(h, cur.long()); h.retain_grad(); inter[f"in{s}"] = h
    hd = F.conv2d(h, P[f"pre_encode.conv.{idx}.weight"], P[f"pre_encode.conv.{idx}.bias"], stride=2, padding=1, groups=C)
    cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1
    hd = tm(hd, cur.long()); hd.retain_grad(); inter[f"dwo{s}"] = hd
    hp = F.conv2d(hd, P[f"pre_encode.conv.{idx + 1}.weight"], P[f"pre_encode.conv.{idx + 1}.bias"])
    hp = tm(hp, cur.long()); hp.retain_grad(); inter[f"pre{s}"] = hp
    h = F.relu(hp)
    idx += 3
h = tm(h, cur.long()); h.retain_grad(); inter["xl"] = h
b_, c_, t_, f_ = h.shape
o = F.linear(h.transpose(1, 2).reshape(b_, t_, c_ * f_), P["pre_encode.out.weight"], P["pre_encode.out.bias"]) * math.sqrt(128)
yr = o.transpose(1, 2)
print("fwd", rel(y, yr))
(yr * r).sum().backward()
nhwc = lambda t: t.permute(0, 2, 3, 1)
print("dxl  ", rel(dbg["dxl"].view(nhwc(inter["xl"].grad).shape), nhwc(inter["xl"].grad)))
for s in (1, 0):
    print(f"stage {s}: dpre", rel(dbg[f"dpre{s}"].view(nhwc(inter[f'pre{s}'].grad).shape), nhwc(inter[f"pre{s}"].grad)),
          "ddw", rel(dbg[f"ddw{s}"].view(nhwc(inter[f'dwo{s}'].grad).shape), nhwc(inter[f"dwo{s}"].grad)),
          "dxin", rel(dbg[f"dxin{s}"] * (nhwc(inter[f'in{s}']) != 0), nhwc(inter[f"in{s}"].grad) * (nhwc(inter[f'in{s}']) != 0)),
          "dxin(raw)", rel(dbg[f"dxin{s}"], nhwc(inter[f"in{s}"].grad)))
# ---- (B) striding with layers: per-parameter errors
for nl in (1,):
    cfg = dict(n_layers=nl, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="striding", subsampling_factor=4)
    sd = O.init_state_dict(seed=3, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=nl, d_model=128, n_heads=2, conv_kernel_size=9, dropout=0.0, dropout_pre_encoder=0.0,
                               dropout_emb=0.0, dropout_att=0.0)
    enc.load_state_dict(sd)
    enc = enc.to(DEV).train()
    enc.precision = "fp32"
    enc.__dict__["_debug_parity"] = dbg = {}
    yy, _ = enc(audio_signal=x, length=ln)
    r = torch.randn(yy.shape, generator=g).to(DEV)
    (yy * r).sum().backward()
    params = {k: (v.to(DEV).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.to(DEV)) for k, v in sd.items()}
    y_ref, _, taps = O.encoder_forward(params, x, ln, n_layers=nl, d_model=128, n_heads=2, training=True, return_layers=True)
    (y_ref * r).sum().backward()
    print("layers", nl, "fwd", rel(yy, y_ref))
    for k, p in enc.named_parameters():
        print(f"  {k:48s} rel {rel(p.grad, params[k].grad):.3e}  |g| {params[k].grad.norm().item():.3e}")
