import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nemo_b200 import _lib as L
from nemo_b200.encoder import B200ConformerEncoder
from oracle import conformer_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = torch.device("cuda:0")
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()
# (1) unit: p_dwconv2d fwd / bwd vs F.conv2d autograd
B, C, T1, F1 = 3, 64, 101, 20
T2, F2 = (T1 - 1) // 2 + 1, (F1 - 1) // 2 + 1
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T1, F1, C, generator=g).to(DEV)
lin = torch.tensor([101, 75, 20], device=DEV)
lout = ((lin - 1) // 2 + 1).to(torch.int32)
x = x * (torch.arange(T1, device=DEV)[None, :] < lin[:, None]).float()[:, :, None, None]
w = torch.randn(C, 1, 3, 3, generator=g).to(DEV).requires_grad_(True)
bias = torch.randn(C, generator=g).to(DEV).requires_grad_(True)
xr = x.clone().requires_grad_(True)
y = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), w, bias, stride=2, padding=1, groups=C)
y = y * (torch.arange(T2, device=DEV)[None, :] < lout[:, None]).float()[:, None, :, None]
dy = torch.randn(y.shape, generator=g).to(DEV)
y.backward(dy)
out = torch.empty(B, T2, F2, C, device=DEV)
L.call("p_dwconv2d", x, w.detach().contiguous(), bias.detach(), lout, out, B, T1, F1, C)
print("p_dwconv2d fwd", rel(out, y.detach().permute(0, 2, 3, 1)))
dx = torch.empty(B, T1, F1, C, device=DEV)
dw = torch.zeros(C, 9, device=DEV); db = torch.zeros(C, device=DEV)
L.call("p_dwconv2d_bwd", dy.permute(0, 2, 3, 1).contiguous(), x, w.detach().contiguous(), lout, dx, dw, db, B, T1, F1, C)
print("p_dwconv2d bwd dx", rel(dx, xr.grad), "dw", rel(dw, w.grad.view(C, 9)), "db", rel(db, bias.grad))
# (2) striding encoder with 1 layer: per-parameter errors
for nl in (1, 2):
    cfg = dict(n_layers=nl, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="striding", subsampling_factor=4)
    sd = O.init_state_dict(seed=3, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=nl, d_model=128, n_heads=2, conv_kernel_size=9, dropout=0.0, dropout_pre_encoder=0.0,
                               dropout_emb=0.0, dropout_att=0.0)
    enc.load_state_dict(sd)
    enc = enc.to(DEV).train()
    enc.precision = "fp32"
    xx = torch.randn(3, 80, 403, generator=g)
    ln = torch.tensor([403, 300, 77])
    for b in range(3):
        xx[b, :, ln[b]:] = 0
    xx, ln = xx.to(DEV), ln.to(DEV)
    yy, _ = enc(audio_signal=xx, length=ln)
    r = torch.randn(yy.shape, generator=g).to(DEV)
    (yy * r).sum().backward()
    params = {k: v.to(DEV).clone().requires_grad_(True) for k, v in sd.items()}
    y_ref, _ = O.encoder_forward(params, xx, ln, n_layers=nl, d_model=128, n_heads=2, training=True)
    (y_ref * r).sum().backward()
    print("layers", nl, "fwd", rel(yy, y_ref))
    for k, p in enc.named_parameters():
        print(f"  {k:48s} rel {rel(p.grad, params[k].grad):.3e}  |g| {params[k].grad.norm().item():.3e}")
