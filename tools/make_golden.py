"""Generates tests/golden/*.npz by running the UNMODIFIED reference (exec'd from /root/reference via
oracle/ref_loader.py).  Run here (the GPU box has no /root/reference):  python tools/make_golden.py
The fixtures pin oracle/conformer_oracle.py (tests/test_oracle_pin.py) and travel to the GPU box."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sd_np(sd):
    return {f"sd::{k}": v.detach().cpu().numpy() for k, v in sd.items()}


def gen_mel():
    ref = rl.load()
    torch.manual_seed(7)
    fb = ref.FilterbankFeatures(sample_rate=16000, n_window_size=400, n_window_stride=160, n_fft=512, nfilt=80,
                                dither=1e-5, pad_to=0, normalize="per_feature")
    fb.eval()
    x = torch.rand(3, 8000) * 0.2 - 0.1
    length = torch.tensor([8000, 5280, 799])
    mel, seq = fb(x.clone(), length)
    fbn = ref.FilterbankFeatures(sample_rate=16000, n_window_size=400, n_window_stride=160, n_fft=512, nfilt=80,
                                 dither=1e-5, pad_to=0, normalize=None)
    fbn.eval()
    mel_raw, _ = fbn(x.clone(), length)
    lens = torch.cat([torch.arange(0, 1200), torch.randint(0, 2_000_000, (800,))])
    np.savez_compressed(os.path.join(OUT, "mel.npz"), x=x.numpy(), length=length.numpy(), mel=mel.numpy(),
                        mel_raw=mel_raw.numpy(), seq_len=seq.numpy(), fb=fb.fb[0].numpy(), window=fb.window.numpy(),
                        lens=lens.numpy(), lens_out=fb.get_seq_len(lens).numpy())


def gen_encoder(name, **kw):
    cfg = dict(feat_in=80, n_layers=2, d_model=32, n_heads=4, conv_kernel_size=9, dropout=0.0,
               dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
    cfg.update(kw)
    enc = rl.make_encoder(seed=3, **cfg)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 80, 131, generator=g)
    length = torch.tensor([131, 97, 40])
    # zero padded frames like the preprocessor does
    for b in range(3):
        x[b, :, length[b]:] = 0
    enc.eval()
    with torch.no_grad():
        y_eval, ylen = enc(audio_signal=x, length=length)
    sd0 = {k: v.clone() for k, v in enc.state_dict().items()}
    enc.train()
    y_train, _ = enc(audio_signal=x, length=length)
    # NB: mean(y^2) is ~invariant under the final LayerNorm (zero gradient) — use a fixed random projection
    r = torch.randn(y_train.shape, generator=g)
    loss = (y_train.float() * r).sum() / y_train.shape[0]
    loss.backward()
    grads = {f"grad::{k}": p.grad.detach().numpy() for k, p in enc.named_parameters()}
    bn = {f"bn::{k}": v.detach().numpy() for k, v in enc.state_dict().items() if "running_" in k}
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), x=x.numpy(), length=length.numpy(),
                        y_eval=y_eval.numpy(), y_train=y_train.detach().numpy(), ylen=ylen.numpy(),
                        loss=loss.detach().numpy(), r=r.numpy(), **sd_np(sd0), **grads, **bn)


def gen_spec_augment():
    """SpecAugment of the reference in both RNG modes: the masks depend only on torch's CPU generator (vectorised mode,
    manual_seed 7) or on random.Random(3) (legacy mode) — both reproducible anywhere."""
    import random

    ref = rl.load()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 80, 301, generator=g)
    length = torch.tensor([301, 250, 17])
    out = {"x": x.numpy(), "length": length.numpy()}
    for tag, tw in (("int", 10), ("adaptive", 0.05)):
        sa = ref.SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=tw, use_vectorized_code=True)
        torch.manual_seed(7)
        out[f"vec_{tag}"] = (sa(input_spec=x, length=length) == 0).numpy()
        sa = ref.SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=tw, rng=random.Random(3),
                             use_vectorized_code=False)
        out[f"legacy_{tag}"] = (sa(input_spec=x, length=length) == 0).numpy()
    np.savez_compressed(os.path.join(OUT, "spec_augment.npz"), **out)
    print("spec_augment.npz", {k: (v.shape, float(v.mean())) for k, v in out.items() if k not in ("x", "length")})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_mel()
    gen_encoder("enc_striding", subsampling="striding", subsampling_factor=4)
    gen_encoder("enc_dw_striding", subsampling="dw_striding", subsampling_factor=8, subsampling_conv_channels=16,
                n_heads=2)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def gen_encoder_cuda_cfg():
    """A configuration the CUDA path supports (head dim 64): weights = oracle.init_state_dict(seed) loaded into the
    reference module, so the fixture only carries inputs, outputs and (sub-sampled) gradients."""
    from oracle import conformer_oracle as O
    cfgs = {
        "enc_cuda_striding": dict(n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="striding",
                                  subsampling_factor=4, subsampling_conv_channels=-1, seed=5),
    }
    for name, c in cfgs.items():
        sd = O.init_state_dict(**c)
        enc = rl.make_encoder(seed=0, feat_in=80, n_layers=c["n_layers"], d_model=c["d_model"], n_heads=c["n_heads"],
                              conv_kernel_size=c["conv_kernel_size"], subsampling=c["subsampling_kind"],
                              subsampling_factor=c["subsampling_factor"],
                              subsampling_conv_channels=c["subsampling_conv_channels"], dropout=0.0,
                              dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
        enc.load_state_dict(sd)
        g = torch.Generator().manual_seed(21)
        x = torch.randn(3, 80, 131, generator=g)
        length = torch.tensor([131, 97, 40])
        for b in range(3):
            x[b, :, length[b]:] = 0
        enc.eval()
        with torch.no_grad():
            y_eval, ylen = enc(audio_signal=x, length=length)
        enc.train()
        y_train, _ = enc(audio_signal=x, length=length)
        r = torch.randn(y_train.shape, generator=g)
        loss = (y_train.float() * r).sum() / y_train.shape[0]
        loss.backward()
        grads = {}
        for k, p in enc.named_parameters():
            gflat = p.grad.detach().flatten()
            grads[f"grad::{k}"] = (gflat if gflat.numel() <= 20000 else gflat[::97]).numpy()
        bn = {f"bn::{k}": v.detach().numpy() for k, v in enc.state_dict().items() if "running_" in k}
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), x=x.numpy(), length=length.numpy(), y_eval=y_eval.numpy(),
                            y_train=y_train.detach().numpy(), ylen=ylen.numpy(), loss=loss.detach().numpy(), r=r.numpy(),
                            cfg=np.array(repr(c)), **grads, **bn)


def gen_ctc_head():
    """ConvASRDecoder (conv_asr.py:407-468) + CTCLoss (losses/ctc.py:25-82; its arithmetic is torch.nn.CTCLoss) of the
    unmodified reference: log-probs, loss (mean_batch) and gradients for a small ragged batch."""
    ref = rl.load()
    torch.manual_seed(3)
    dec = ref.ConvASRDecoder(feat_in=64, num_classes=30)
    with torch.no_grad():
        dec.decoder_layers[0].bias.normal_(0, 0.1)
    x = torch.randn(3, 64, 41, requires_grad=True)
    lp = dec(encoder_output=x)
    in_len = torch.tensor([41, 33, 12])
    tgt_len = torch.tensor([9, 5, 4])
    tgt = torch.randint(0, 30, (3, 9))
    tgt[0, 3] = tgt[0, 2]  # a repeated label (needs the blank between)
    # CTCLoss.forward, losses/ctc.py:70-82, reduction 'mean_batch' (:57-60): torch CTC with blank = num_classes, then mean
    losses = torch.nn.functional.ctc_loss(lp.transpose(1, 0), tgt.long(), in_len.long(), tgt_len.long(), blank=30,
                                          reduction="none", zero_infinity=False)
    loss = losses.mean()
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "ctc_head.npz"), x=x.detach().numpy(), logp=lp.detach().numpy(),
                        in_len=in_len.numpy(), tgt_len=tgt_len.numpy(), tgt=tgt.numpy(), losses=losses.detach().numpy(),
                        loss=loss.detach().numpy(), dx=x.grad.numpy(),
                        dW=dec.decoder_layers[0].weight.grad.numpy(), db=dec.decoder_layers[0].bias.grad.numpy(),
                        **sd_np(dec.state_dict()))


if __name__ == "__main__":
    gen_ctc_head()
    gen_encoder_cuda_cfg()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    gen_spec_augment()
