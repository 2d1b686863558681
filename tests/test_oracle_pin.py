"""Pins oracle/conformer_oracle.py against (i) golden fixtures produced by the unmodified reference
(tools/make_golden.py) and (ii) the live reference exec'd from /root/reference when present."""
import os

import numpy as np
import pytest
import torch

from oracle import conformer_oracle as O
from oracle import ref_loader as rl

needs_ref = pytest.mark.skipif(not rl.available(), reason="/root/reference not present (GPU box)")


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    return z, sd


def test_golden_seq_len_bit_exact(golden_dir):
    z, _ = _load(golden_dir, "mel.npz")
    got = O.mel_seq_len(torch.from_numpy(z["lens"]))
    # reference get_seq_len has no len==0 special case (forward applies it): compare on len>0
    ref = torch.from_numpy(z["lens_out"])
    nz = torch.from_numpy(z["lens"]) > 0
    assert torch.equal(got[nz], ref[nz])
    assert got[~nz].eq(0).all()


def test_golden_mel(golden_dir):
    z, _ = _load(golden_dir, "mel.npz")
    x, length = torch.from_numpy(z["x"]), torch.from_numpy(z["length"])
    mel, seq, raw = O.mel_frontend(x, length, return_prenorm=True)
    assert torch.equal(seq, torch.from_numpy(z["seq_len"]))
    np.testing.assert_allclose(O.slaney_mel_filterbank(), z["fb"], atol=1e-7)
    np.testing.assert_allclose(torch.hann_window(400, periodic=False).numpy(), z["window"], atol=0)
    # same torch ops in the same order → expect agreement to float rounding
    valid = (torch.arange(raw.shape[-1])[None, :] < seq[:, None])[:, None, :]
    np.testing.assert_allclose(torch.where(valid, raw, 0.0).numpy(), z["mel_raw"], atol=2e-5)
    np.testing.assert_allclose(mel.numpy(), z["mel"], atol=2e-4)


@pytest.mark.parametrize("name,kind,factor,heads", [("enc_striding.npz", "striding", 4, 4),
                                                     ("enc_dw_striding.npz", "dw_striding", 8, 2)])
def test_golden_encoder(golden_dir, name, kind, factor, heads):
    z, sd = _load(golden_dir, name)
    x, length = torch.from_numpy(z["x"]), torch.from_numpy(z["length"])
    kw = dict(n_layers=2, d_model=32, n_heads=heads, subsampling_kind=kind, subsampling_factor=factor)
    y, ylen = O.encoder_forward(sd, x, length, training=False, **kw)
    assert torch.equal(ylen, torch.from_numpy(z["ylen"]))
    np.testing.assert_allclose(y.numpy(), z["y_eval"], atol=2e-5, rtol=1e-5)
    # train mode (BatchNorm batch stats) + gradients through autograd
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    bn = {}
    yt, _ = O.encoder_forward(params, x, length, training=True, bn_out=bn, **kw)
    np.testing.assert_allclose(yt.detach().numpy(), z["y_train"], atol=2e-5, rtol=1e-5)
    loss = (yt.float() * torch.from_numpy(z["r"])).sum() / yt.shape[0]
    np.testing.assert_allclose(loss.item(), float(z["loss"]), rtol=1e-5)
    loss.backward()
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad::"))
    for k in z.files:
        if k.startswith("grad::"):
            g = params[k[6:]].grad
            ref = z[k]
            # (depthwise_conv.bias has an analytically-zero gradient under train-mode BatchNorm: abs floor)
            assert np.abs(g.numpy() - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-5 * gmax, k
    # running-stat update: momentum 0.1, unbiased var
    for p, (mean, var_u) in bn.items():
        rm = 0.9 * sd[f"{p}batch_norm.running_mean"] + 0.1 * mean
        rv = 0.9 * sd[f"{p}batch_norm.running_var"] + 0.1 * var_u
        np.testing.assert_allclose(rm.numpy(), z[f"bn::{p}batch_norm.running_mean"], atol=1e-5)
        np.testing.assert_allclose(rv.numpy(), z[f"bn::{p}batch_norm.running_var"], atol=1e-5)


@needs_ref
def test_live_reference_mel_ragged():
    ref = rl.load()
    fb = ref.FilterbankFeatures(sample_rate=16000, n_window_size=400, n_window_stride=160, n_fft=512, nfilt=80,
                                dither=1e-5, pad_to=0)
    fb.eval()
    x, length = O.synthetic_waveforms(4, 2.0, ragged=True)
    length[3] = 16001
    m_ref, l_ref = fb(x.clone(), length)
    m, l = O.mel_frontend(x, length)
    assert torch.equal(l, l_ref)
    assert (m - m_ref).abs().max() < 2e-4


@needs_ref
def test_live_reference_small_encoder_cfg1():
    """BASELINE config 1: Conformer-CTC-Small (16L, d=176, 4 heads) fwd on 2x1 s."""
    enc = rl.make_encoder(seed=0, feat_in=80, n_layers=16, d_model=176, n_heads=4, conv_kernel_size=31, dropout=0.0,
                          dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
    enc.eval()
    x, length = O.synthetic_waveforms(2, 1.0)
    mel, mlen = O.mel_frontend(x, length)
    with torch.no_grad():
        y_ref, l_ref = enc(audio_signal=mel, length=mlen)
        y, l = O.encoder_forward(enc.state_dict(), mel, mlen, n_layers=16, d_model=176, n_heads=4)
    assert torch.equal(l, l_ref) and y.shape == y_ref.shape == (2, 176, 26)
    assert (y - y_ref).abs().max() < 5e-5


@needs_ref
def test_init_state_dict_layout_matches_reference():
    for kind, fac, ch in (("striding", 4, -1), ("dw_striding", 8, 16)):
        enc = rl.make_encoder(feat_in=80, n_layers=2, d_model=32, n_heads=4, conv_kernel_size=9, subsampling=kind,
                              subsampling_factor=fac, subsampling_conv_channels=ch)
        ref_sd = enc.state_dict()
        sd = O.init_state_dict(n_layers=2, d_model=32, n_heads=4, conv_kernel_size=9, subsampling_kind=kind,
                               subsampling_factor=fac, subsampling_conv_channels=ch)
        assert set(sd) == set(ref_sd)
        for k in sd:
            assert tuple(sd[k].shape) == tuple(ref_sd[k].shape), k


@needs_ref
@pytest.mark.parametrize("time_width", [10, 0.05])
def test_live_reference_spec_augment(time_width):
    """oracle restatement of SpecAugment (both RNG modes) against the exec'd reference class: identical masks."""
    import random

    ref = rl.load()
    x = torch.randn(3, 80, 301)
    ln = torch.tensor([301, 250, 17])
    sa = ref.SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=time_width, use_vectorized_code=True)
    torch.manual_seed(7)
    want = sa(input_spec=x, length=ln)
    torch.manual_seed(7)
    got = O.spec_augment_vectorized(x, ln, 2, 5, 27, time_width)
    assert torch.equal(want, got) and 0.05 < (got == 0).float().mean() < 0.9
    sa = ref.SpecAugment(freq_masks=2, time_masks=5, freq_width=27, time_width=time_width, rng=random.Random(3),
                         use_vectorized_code=False)
    assert torch.equal(sa(input_spec=x, length=ln), O.spec_augment_legacy(x, ln, random.Random(3), 2, 5, 27, time_width))


@pytest.mark.parametrize("tag,time_width", [("int", 10), ("adaptive", 0.05)])
def test_golden_spec_augment(golden_dir, tag, time_width):
    """masks produced by the reference's SpecAugment (tools/make_golden.py) vs the oracle restatement, both RNG modes"""
    import random

    z = np.load(os.path.join(golden_dir, "spec_augment.npz"))
    x, ln = torch.from_numpy(z["x"]), torch.from_numpy(z["length"])
    torch.manual_seed(7)
    got = O.spec_augment_vectorized(x, ln, 2, 5, 27, time_width)
    assert np.array_equal((got == 0).numpy(), z[f"vec_{tag}"])
    got = O.spec_augment_legacy(x, ln, random.Random(3), 2, 5, 27, time_width)
    assert np.array_equal((got == 0).numpy(), z[f"legacy_{tag}"])


@needs_ref
def test_live_reference_ctc_head():
    """O.ctc_head against the exec'd reference ConvASRDecoder (conv_asr.py:407-468): bit-identical log-probs."""
    ref = rl.load()
    torch.manual_seed(0)
    dec = ref.ConvASRDecoder(feat_in=48, num_classes=37)
    x = torch.randn(2, 48, 29)
    want = dec(encoder_output=x)
    got = O.ctc_head(dec.state_dict(), x)
    assert list(dec.state_dict().keys()) == ["decoder_layers.0.weight", "decoder_layers.0.bias"]
    assert torch.equal(want, got)


def test_golden_ctc_head_and_loss(golden_dir):
    """fixture from the unmodified reference (tools/make_golden.py::gen_ctc_head): head log-probs, CTC loss and gradients"""
    z, sd = _load(golden_dir, "ctc_head.npz")
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lp = O.ctc_head(sdr, x)
    assert torch.equal(lp.detach(), torch.from_numpy(z["logp"]))
    losses = O.ctc_loss(lp, torch.from_numpy(z["tgt"]), torch.from_numpy(z["in_len"]), torch.from_numpy(z["tgt_len"]), blank=30,
                        reduction="none")
    assert np.allclose(losses.detach().numpy(), z["losses"], rtol=1e-6, atol=1e-6)
    losses.mean().backward()
    assert np.allclose(x.grad.numpy(), z["dx"], rtol=1e-5, atol=1e-7)
    assert np.allclose(sdr["decoder_layers.0.weight"].grad.numpy(), z["dW"], rtol=1e-5, atol=1e-6)
