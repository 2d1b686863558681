"""Full-depth parity at the BASELINE.json configurations: the complete 18 / 17 / 24-layer encoders of
configs[1..3] at 30 s utterances, forward + backward, against the fp32 oracle AND the oracle under torch.autocast(bf16)
(= what the reference itself computes with `trainer.precision=bf16`), error printed per depth.

Gates (speed mode = bf16 operands): err(cuda, oracle_fp32) <= 1.5 * err(autocast_oracle, oracle_fp32) + 2e-3
for the output and every layer tap; gradients <= 2 * autocast error + 2e-2.  Lengths bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from nemo_b200.encoder import B200ConformerEncoder
    from nemo_b200.preprocessor import B200MelPreprocessor
    from oracle import conformer_oracle as O
    DEV = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


FULL = {
    # name: (model cfg, batch, ragged)
    "cfg2_conformer_ctc_large_18L": (dict(n_layers=18, d_model=512, n_heads=8, conv_kernel_size=31, subsampling_kind="striding",
                                          subsampling_factor=4, subsampling_conv_channels=512), 2),
    "cfg3_fastconformer_large_17L": (dict(n_layers=17, d_model=512, n_heads=8, conv_kernel_size=9, subsampling_kind="dw_striding",
                                          subsampling_factor=8, subsampling_conv_channels=256), 2),
    "cfg4_conformer_transducer_xl_24L": (dict(n_layers=24, d_model=1024, n_heads=8, conv_kernel_size=5, subsampling_kind="striding",
                                              subsampling_factor=4, subsampling_conv_channels=1024), 1),
}


def _oracle(sd, mel, mlen, cfg, autocast, r=None, taps=False):
    params = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v) for k, v in sd.items()}
    kw = dict(n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"], subsampling_kind=cfg["subsampling_kind"],
              subsampling_factor=cfg["subsampling_factor"], training=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        res = O.encoder_forward(params, mel, mlen, return_layers=taps, **kw)
    y, ylen = res[0], res[1]
    grads = None
    if r is not None:
        ((y.float() * r).sum() / y.shape[0]).backward()
        grads = {k: p.grad for k, p in params.items() if isinstance(p, torch.Tensor) and p.requires_grad and p.grad is not None}
    return y.detach().float(), ylen, grads, ([t.detach().float() for t in res[2]] if taps else None)


@pytest.mark.parametrize("name", list(FULL))
def test_full_depth_forward_backward(name):
    cfg, B = FULL[name]
    sd = O.init_state_dict(seed=11, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"],
                               conv_kernel_size=cfg["conv_kernel_size"], subsampling=cfg["subsampling_kind"],
                               subsampling_factor=cfg["subsampling_factor"], subsampling_conv_channels=cfg["subsampling_conv_channels"],
                               dropout=0.0, dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).train()
    sd = {k: v.to(DEV) for k, v in sd.items()}
    pre = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0).to(DEV).eval()
    wav, wlen = O.synthetic_waveforms(B, 30.0, ragged=True)
    if B > 1:
        wlen[-1] = 30 * 16000 - 160 * 411  # a clearly shorter utterance: masks matter in every layer
    mel, mlen = pre(input_signal=wav.to(DEV), length=wlen.to(DEV))
    g = torch.Generator().manual_seed(3)
    enc.__dict__["_debug_taps"] = taps = []
    y, ylen = enc(audio_signal=mel, length=mlen)
    enc.__dict__["_debug_taps"] = None
    r = torch.randn(y.shape, generator=g).to(DEV)
    ((y.float() * r).sum() / y.shape[0]).backward()
    y_ref, ylen_ref, g_ref, t_ref = _oracle(sd, mel, mlen, cfg, False, r, taps=True)
    y_ac, _, g_ac, t_ac = _oracle(sd, mel, mlen, cfg, True, r, taps=True)
    assert torch.equal(ylen, ylen_ref)
    for i in (0, 1, 3, 7, 11, 15, cfg["n_layers"] - 1):
        if i < cfg["n_layers"]:
            print(f"[{name}] depth {i + 1:2d}: rel-L2 vs fp32 oracle  cuda {rel_l2(taps[i], t_ref[i]):.3e} | autocast-oracle {rel_l2(t_ac[i], t_ref[i]):.3e}")
    for i in range(cfg["n_layers"]):
        assert rel_l2(taps[i], t_ref[i]) <= 1.5 * rel_l2(t_ac[i], t_ref[i]) + 2e-3, f"layer {i}"
    e, e_ac = rel_l2(y, y_ref), rel_l2(y_ac, y_ref)
    mx = ((y.float() - y_ref).abs().max() / y_ref.abs().max()).item()
    print(f"[{name}] output rel-L2 cuda {e:.3e} | autocast-oracle {e_ac:.3e}; max-abs/max|y| cuda {mx:.3e}")
    assert e <= 1.5 * e_ac + 2e-3
    bad = []
    worst = (0.0, "")
    for k, p in enc.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        gr = g_ref[k]
        if gr.norm() <= 1e-6 * gr.numel() ** 0.5:
            continue
        eg, eg_ac = rel_l2(p.grad, gr), rel_l2(g_ac[k], gr)
        if eg > worst[0]:
            worst = (eg, k)
        if eg > 2.0 * eg_ac + 2e-2:
            bad.append((k, eg, eg_ac))
    print(f"[{name}] worst parameter-gradient rel-L2 {worst[0]:.3e} ({worst[1]})")
    for k, eg, eg_ac in bad:
        print(f"[{name}] GRAD FAIL {k}: cuda {eg:.3e} autocast-oracle {eg_ac:.3e}")
    assert not bad


def test_dither_is_gaussian_seeded_and_train_only():
    """a4: dither (features.py:435-436, training only): x += dither * randn.  The kernel draws its own Philox / Box-Muller
    normals, so the check is statistical: with pre-emphasis off and a zero waveform the STFT input is pure dither."""
    from nemo_b200 import _lib as L

    B, S = 4, 16000 * 4
    pre = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, dither=1e-5, pad_to=0).to(DEV)
    x = (torch.rand(B, S, generator=torch.Generator().manual_seed(0)) * 0.2 - 0.1).to(DEV)
    ln = torch.full((B,), S, dtype=torch.int64, device=DEV)
    pre.eval()
    e1, _ = pre(input_signal=x, length=ln)
    e2, _ = pre(input_signal=x, length=ln)
    assert torch.equal(e1, e2)  # eval: no dither, deterministic
    pre.train()
    t1, _ = pre.featurizer.forward_raw(x, ln)
    t2, _ = pre.featurizer.forward_raw(x, ln)
    pre.eval()
    e_raw, _ = pre.featurizer.forward_raw(x, ln)
    assert not torch.equal(t1, t2)            # a new draw every call
    assert not torch.equal(t1, e_raw)         # train != eval
    # dither 1e-5 on a 0.1-amplitude signal moves log-mel by ~1e-4 relative: tiny but non-zero, and never large
    assert (t1 - e_raw).abs().max().item() < 5e-2
    # same seed => same output (the module's seed counter is the Philox key)
    pre.train()
    pre.featurizer._seed = 41
    a, _ = pre.featurizer.forward_raw(x, ln)
    pre.featurizer._seed = 41
    b, _ = pre.featurizer.forward_raw(x, ln)
    assert torch.equal(a, b)
    # distribution: silence + dither sigma → power spectrum of white noise.  E[|X_k|^2] = sigma^2 * sum(w^2) for every bin, so
    # mel[m] = sigma^2 * sum(w^2) * sum_k fb[m,k]; compare the time-averaged linear mel energies with that expectation.
    sig = 1e-2
    pre2 = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, dither=sig, pad_to=0, preemph=0.0).to(DEV).train()
    z = torch.zeros(4, 16000 * 60, device=DEV)
    zl = torch.full((4,), z.shape[1], dtype=torch.int64, device=DEV)
    raw, rl = pre2.featurizer.forward_raw(z, zl)
    lin = (raw[:, :, 2:-3].exp() - 2 ** -24).mean(dim=(0, 2))  # interior frames only (no edge padding)
    w = pre2.featurizer.window.double()
    expect = (sig ** 2) * (w ** 2).sum() * pre2.featurizer.fb[0].double().sum(dim=1)
    ratio = (lin.double() / expect).cpu()
    print("dither: mel-energy / expectation: min %.3f max %.3f mean %.4f" % (ratio.min(), ratio.max(), ratio.mean()))
    assert (ratio - 1).abs().max().item() < 0.05  # ~4000 frames x several bins per filter: < 5 % sampling error
    assert abs(ratio.mean().item() - 1) < 0.01


@pytest.mark.parametrize("name", ["cfg2_conformer_ctc_large_18L", "cfg3_fastconformer_large_17L", "cfg4_conformer_transducer_xl_24L"])
def test_parity_mode_logits_within_1e3_of_fp32_reference(name):
    """north_star: "encoder logits matching the reference to <= 1e-3 relative".  precision='fp32' (parity mode: tcgen05 GEMMs on
    bf16 hi/lo-split operands = fp32-equivalent products, fp32 everywhere else) on the FULL-DEPTH BASELINE encoders, ragged
    2 x 30 s, through the CTC head, against the fp32 oracle: rel-L2 and max-abs/max|y| <= 1e-3 for the encoder output, the
    logits and every layer tap; log-probs <= 1e-3 absolute... (speed mode's numbers are printed next to them)."""
    from nemo_b200.decoder import B200ConvASRDecoder

    cfg, B = FULL[name]
    sd = O.init_state_dict(seed=11, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"],
                               conv_kernel_size=cfg["conv_kernel_size"], subsampling=cfg["subsampling_kind"],
                               subsampling_factor=cfg["subsampling_factor"], subsampling_conv_channels=cfg["subsampling_conv_channels"])
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).eval()
    torch.manual_seed(1)
    head = B200ConvASRDecoder(feat_in=cfg["d_model"], num_classes=1024).to(DEV)
    sd = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad():  # BatchNorm running statistics as a trained checkpoint has them (not the 0 / 1 init)
        for i in range(cfg["n_layers"]):
            bn = enc.layers[i].conv.batch_norm
            bn.running_mean.normal_(0, 0.1)
            bn.running_var.uniform_(0.75, 1.25)
            sd[f"layers.{i}.conv.batch_norm.running_mean"] = bn.running_mean.clone()
            sd[f"layers.{i}.conv.batch_norm.running_var"] = bn.running_var.clone()
    pre = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0).to(DEV).eval()
    wav, wlen = O.synthetic_waveforms(B, 30.0, ragged=True)
    if B > 1:
        wlen[-1] = 30 * 16000 - 160 * 411
    mel, mlen = pre(input_signal=wav.to(DEV), length=wlen.to(DEV))
    kw = dict(n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"], subsampling_kind=cfg["subsampling_kind"],
              subsampling_factor=cfg["subsampling_factor"], training=False)
    with torch.no_grad():
        y_ref, ylen_ref, t_ref = O.encoder_forward(sd, mel, mlen, return_layers=True, **kw)
        hsd = {k: v.detach() for k, v in head.state_dict().items()}
        logits_ref = torch.nn.functional.conv1d(y_ref, hsd["decoder_layers.0.weight"], hsd["decoder_layers.0.bias"]).transpose(1, 2)
        lp_ref = O.ctc_head(hsd, y_ref)
        y16, _ = enc(audio_signal=mel, length=mlen)
        lp16 = head(encoder_output=y16)
    enc.precision = head.precision = "fp32"
    enc.__dict__["_debug_taps"] = taps = []
    y, ylen = enc(audio_signal=mel, length=mlen)
    enc.__dict__["_debug_taps"] = None
    lp = head(encoder_output=y)
    logits = head._last_logits
    assert torch.equal(ylen, ylen_ref)
    relmax = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()  # noqa: E731
    for i in (0, 3, 7, 11, cfg["n_layers"] - 1):
        print(f"[parity {name}] depth {i + 1:2d}: rel-L2 {rel_l2(taps[i], t_ref[i]):.3e}  max-abs/max|y| {relmax(taps[i], t_ref[i]):.3e}")
    print(f"[parity {name}] encoder out: fp32 mode rel-L2 {rel_l2(y, y_ref):.3e} max-rel {relmax(y, y_ref):.3e} | speed mode rel-L2 {rel_l2(y16, y_ref):.3e}")
    print(f"[parity {name}] logits     : fp32 mode rel-L2 {rel_l2(logits, logits_ref):.3e} max-rel {relmax(logits, logits_ref):.3e}")
    print(f"[parity {name}] log-probs  : fp32 mode max-abs {(lp - lp_ref).abs().max().item():.3e} | speed mode {(lp16 - lp_ref).abs().max().item():.3e}")
    for i in range(cfg["n_layers"]):
        assert rel_l2(taps[i], t_ref[i]) <= 1e-3 and relmax(taps[i], t_ref[i]) <= 1e-3, f"layer {i}"
    assert rel_l2(y, y_ref) <= 1e-3 and relmax(y, y_ref) <= 1e-3
    assert rel_l2(logits, logits_ref) <= 1e-3 and relmax(logits, logits_ref) <= 1e-3
    assert (lp - lp_ref).abs().max().item() <= 1e-3 * lp_ref.abs().max().item()


def test_parity_mode_train_statistics_and_guards():
    """parity mode with BatchNorm batch statistics (train(), dropout 0) on a ragged small batch; dropout > 0 is refused"""
    cfg = dict(n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, subsampling_kind="striding", subsampling_factor=4)
    sd = O.init_state_dict(seed=5, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=2, d_model=128, n_heads=2, conv_kernel_size=9, dropout=0.0, dropout_pre_encoder=0.0,
                               dropout_emb=0.0, dropout_att=0.0)
    enc.load_state_dict(sd)
    enc = enc.to(DEV).train()
    enc.precision = "fp32"
    x = torch.randn(3, 80, 203, generator=torch.Generator().manual_seed(3))
    ln = torch.tensor([203, 152, 40])
    for b in range(3):
        x[b, :, ln[b]:] = 0
    y, yl = enc(audio_signal=x.to(DEV), length=ln.to(DEV))
    y_ref, yl_ref = O.encoder_forward({k: v.to(DEV) for k, v in sd.items()}, x.to(DEV), ln.to(DEV), n_layers=2, d_model=128, n_heads=2,
                                      training=True)
    print(f"[parity train-BN] rel-L2 {rel_l2(y, y_ref):.3e}")
    assert torch.equal(yl, yl_ref) and rel_l2(y, y_ref) <= 1e-3
    enc.dropout = 0.1
    with pytest.raises(NotImplementedError):
        enc(audio_signal=x.to(DEV), length=ln.to(DEV))


@pytest.mark.parametrize("name,n_layers", [("cfg2_conformer_ctc_large_18L", 18), ("cfg3_fastconformer_large_17L", 4),
                                           ("cfg4_conformer_transducer_xl_24L", 3)])
def test_parity_mode_gradients_within_1e3_of_fp32_reference(name, n_layers):
    """precision='fp32' backward: EVERY parameter gradient of the BASELINE architectures (config 2 at its full 18 layers; the
    dw_striding and d=1024 / head-dim-128 variants at reduced depth) within 1e-3 rel-L2 of the fp32 oracle's autograd, BatchNorm
    on batch statistics (train(), dropout 0), ragged 30 s utterances."""
    cfg, B = FULL[name]
    cfg = dict(cfg, n_layers=n_layers)
    sd = O.init_state_dict(seed=11, **cfg)
    enc = B200ConformerEncoder(feat_in=80, n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"],
                               conv_kernel_size=cfg["conv_kernel_size"], subsampling=cfg["subsampling_kind"],
                               subsampling_factor=cfg["subsampling_factor"], subsampling_conv_channels=cfg["subsampling_conv_channels"],
                               dropout=0.0, dropout_pre_encoder=0.0, dropout_emb=0.0, dropout_att=0.0)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).train()
    enc.precision = "fp32"
    sd = {k: v.to(DEV) for k, v in sd.items()}
    pre = B200MelPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0).to(DEV).eval()
    wav, wlen = O.synthetic_waveforms(max(B, 2), 30.0, ragged=True)
    wlen[-1] = 30 * 16000 - 160 * 411
    mel, mlen = pre(input_signal=wav.to(DEV), length=wlen.to(DEV))
    enc.__dict__["_debug_parity"] = dbg = {}
    y, ylen = enc(audio_signal=mel, length=mlen)
    r = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
    ((y.float() * r).sum() / y.shape[0]).backward()
    # The oracle's subsampling ReLUs use OUR gates: two fp32 implementations of a convolution disagree on the sign of the few
    # pre-activations that are within rounding of zero (27 of 15.4 M at Conformer-L), and a flipped gate passes / blocks a
    # full-size gradient (2e-3 of the stem's gradient norm) — gradients through ReLU are only comparable in the same linear
    # region.  The flips are counted and must all sit at rounding level.
    gates = [(t > 0).permute(0, 3, 1, 2) for t in dbg["relu_out"]]
    report = []
    params = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k) else v) for k, v in sd.items()}
    y_ref, ylen_ref = O.encoder_forward(params, mel, mlen, n_layers=cfg["n_layers"], d_model=cfg["d_model"], n_heads=cfg["n_heads"],
                                        subsampling_kind=cfg["subsampling_kind"], subsampling_factor=cfg["subsampling_factor"],
                                        training=True, relu_gates=gates, gate_report=report)
    ((y_ref.float() * r).sum() / y_ref.shape[0]).backward()
    g_ref = {k: p.grad for k, p in params.items() if isinstance(p, torch.Tensor) and p.requires_grad and p.grad is not None}
    y_ref = y_ref.detach()
    for i, (nflip, pre_flip, pre_max) in enumerate(report):
        print(f"[parity-grad {name}] ReLU {i}: {nflip} gate flips vs the oracle's own fp32 pre-activations, |pre| of the flipped <= {pre_flip:.2e} "
              f"(max |pre| {pre_max:.2e})")
        assert pre_flip <= 1e-4 * pre_max and nflip <= 1e-4 * gates[i].numel()
    assert torch.equal(ylen, ylen_ref)
    print(f"[parity-grad {name} {n_layers}L] output rel-L2 {rel_l2(y, y_ref):.3e}")
    assert rel_l2(y, y_ref) <= 1e-3
    rms_all = max(g.float().pow(2).mean().sqrt().item() for g in g_ref.values())
    worst, bad = (0.0, ""), []
    for k, p in enc.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        gr = g_ref[k]
        if gr.float().pow(2).mean().sqrt().item() < 1e-5 * rms_all:  # analytically-zero gradients (rounding noise in both)
            assert p.grad.float().pow(2).mean().sqrt().item() < 1e-3 * rms_all, k
            continue
        e = rel_l2(p.grad, gr)
        if e > worst[0]:
            worst = (e, k)
        if e > 1e-3:
            bad.append((k, e))
    print(f"[parity-grad {name} {n_layers}L] worst parameter-gradient rel-L2 {worst[0]:.3e} ({worst[1]})")
    for k, e in bad:
        print(f"[parity-grad {name}] FAIL {k}: {e:.3e}")
    assert not bad
