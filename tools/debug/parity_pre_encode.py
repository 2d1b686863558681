"""debug: parity-mode gradients of the pre_encode stack alone (n_layers = 0) against torch autograd, with intermediates"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nemo_b200.encoder import B200ConformerEncoder
from nemo_b200 import parity as Pm
from oracle import conformer_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = torch.device("cuda:0")
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()
for kind, fac, C in (("striding", 4, 128), ("dw_striding", 8, 64)):
    cfg = dict(n_layers=0, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind=kind, subsampling_factor=fac, subsampling_conv_channels=C)
    sd = O.init_state_dict(seed=3, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=0, d_model=128, n_heads=2, conv_kernel_size=9, subsampling=kind, subsampling_factor=fac,
                               subsampling_conv_channels=C, dropout=0.0, dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
    enc.load_state_dict(sd)
    enc = enc.to(DEV).train()
    enc.precision = "fp32"
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 80, 403, generator=g)
    ln = torch.tensor([403, 300, 77])
    for b in range(3):
        x[b, :, ln[b]:] = 0
    x, ln = x.to(DEV), ln.to(DEV)
    y, yl = enc(audio_signal=x, length=ln)
    r = torch.randn(y.shape, generator=g).to(DEV)
    (y * r).sum().backward()
    params = {k: v.to(DEV).clone().requires_grad_(True) for k, v in sd.items()}
    y_ref, _ = O.encoder_forward(params, x, ln, n_layers=0, d_model=128, n_heads=2, subsampling_kind=kind, subsampling_factor=fac, training=True)
    (y_ref * r).sum().backward()
    print(kind, "fwd rel", rel(y, y_ref))
    for k, p in enc.named_parameters():
        print(f"  {k:40s} rel {rel(p.grad, params[k].grad):.3e}  |g| {params[k].grad.norm().item():.3e}")
    # double-precision reference of the same oracle on the CPU
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    y64, _ = O.encoder_forward(p64, x.cpu().double(), ln.cpu(), n_layers=0, d_model=128, n_heads=2, subsampling_kind=kind, subsampling_factor=fac, training=True)
    (y64 * r.cpu().double()).sum().backward()
    for k, p in enc.named_parameters():
        print(f"  [vs float64 CPU] {k:40s} ours {rel(p.grad.cpu().double(), p64[k].grad):.3e} | torch-gpu-fp32 {rel(params[k].grad.cpu().double(), p64[k].grad):.3e}")
