"""TEST INFRASTRUCTURE — CPU/fp32 restatement of the reference Conformer-ASR encoder hot path.

This is the ORACLE, not the product.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product path
(`nemo_b200/`) never does and fails loudly when its CUDA extension is missing.

It restates, in plain functional PyTorch (the reference itself is PyTorch, all of whose hot-path
arithmetic executes in torch ops — SURVEY.md §2c), the algorithm of

  * FilterbankFeatures.forward            nemo/collections/asr/parts/preprocessing/features.py:423-502
  * normalize_batch("per_feature")        .../features.py:59-93
  * ConvSubsampling (striding, dw_striding) + MaskedConvSequential + calc_length
                                          nemo/collections/asr/parts/submodules/subsampling.py:142-253,385-436,576-586,725-759
  * RelPositionalEncoding                 .../submodules/multi_head_attention.py:1007-1020,1056-1100
  * RelPositionMultiHeadAttention         .../submodules/multi_head_attention.py:104-146,259-354
  * ConformerFeedForward / ConformerConvolution / ConformerLayer
                                          .../submodules/conformer_modules.py:160-233,320-350,366-387
  * ConformerEncoder.forward_internal / _create_masks
                                          nemo/collections/asr/modules/conformer_encoder.py:593-759,794-848

operating on a state_dict with the reference's key layout.

Parity pin: `tests/test_oracle_pin.py` checks every function here against (i) the unmodified reference
sources exec'd from /root/reference through `oracle/ref_loader.py` when that tree is present (this
container), and (ii) the committed golden fixtures `tests/golden/*.npz`, which were produced by running
the reference itself (`tools/make_golden.py`).  The reference's own tests hold no golden vectors for
this path, so (i)+(ii) are the pin.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

INF_VAL = 10000.0  # multi_head_attention.py:52
LOG_GUARD = 2.0 ** -24  # features.py:274
NORM_CONSTANT = 1e-5  # features.py:56


# ----------------------------------------------------------------------------------------------
# mel front-end
# ----------------------------------------------------------------------------------------------
def slaney_mel_filterbank(sr=16000, n_fft=512, n_mels=80, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(norm='slaney', htk=False) — third-party (librosa>=0.10.1), restated from its
    published algorithm; call site features.py:339-341.  f64 maths, f32 result [n_mels, n_fft//2+1]."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def h2m(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)

    def m2h(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = m2h(np.linspace(h2m(fmin), h2m(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - freqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def mel_seq_len(length: torch.Tensor, n_fft=512, hop=160) -> torch.Tensor:
    """features.py:413-417 (+ the len==0 fix at :424-427). Integer result — must be bit-exact."""
    pad_amount = n_fft // 2 * 2
    s = torch.floor_divide(length + pad_amount - n_fft, hop).to(torch.long)
    return torch.where(length == 0, torch.zeros_like(s), s)


def mel_frontend(
    x: torch.Tensor,
    length: torch.Tensor,
    *,
    sample_rate=16000,
    win_length=400,
    hop=160,
    n_fft=512,
    n_mels=80,
    preemph=0.97,
    normalize="per_feature",
    pad_to=0,
    fb: Optional[torch.Tensor] = None,
    window: Optional[torch.Tensor] = None,
    return_prenorm=False,
    dtype=torch.float32,
):
    """FilterbankFeatures.forward in eval mode (no dither), features.py:423-502.
    x [B,S] f32, length [B] int → (mel [B,n_mels,T] f32, seq_len [B] i64).
    dtype=torch.float64 evaluates the same formulas in double precision: the "truth" against which the fp32 rounding noise
    of torch.stft itself (and of the CUDA FFT) is measured in the tests."""
    x = x.to(dtype)
    seq_len = mel_seq_len(length, n_fft, hop)
    if window is None:
        window = torch.hann_window(win_length, periodic=False, dtype=torch.float32)  # features.py:326
    window = window.to(dtype)
    if fb is None:
        fb = torch.from_numpy(slaney_mel_filterbank(sample_rate, n_fft, n_mels))
    fb = fb.reshape(-1, fb.shape[-1]).to(x.device)
    # pre-emphasis + zeroing of samples >= length (features.py:439-442)
    timemask = torch.arange(x.shape[1], device=x.device).unsqueeze(0) < length.unsqueeze(1)
    x = torch.cat((x[:, :1], x[:, 1:] - preemph * x[:, :-1]), dim=1)
    x = x.masked_fill(~timemask, 0.0)
    # torch.stft center=True, constant pad (features.py:386-396)
    spec = torch.stft(
        x, n_fft=n_fft, hop_length=hop, win_length=win_length, center=True,
        window=window.to(x.device), return_complex=True, pad_mode="constant",
    )
    mag = torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1))  # features.py:450-452
    power = mag.pow(2.0)  # features.py:460-461
    mel = torch.matmul(fb.to(power.dtype), power)  # features.py:471
    mel = torch.log(mel + LOG_GUARD)  # features.py:474-475
    prenorm = mel
    if normalize == "per_feature":  # features.py:59-93
        T = mel.shape[2]
        valid = torch.arange(T, device=x.device).unsqueeze(0) < seq_len.unsqueeze(1)
        n = valid.sum(1)
        mean = torch.where(valid.unsqueeze(1), mel, 0.0).sum(2) / n.unsqueeze(1)
        var = torch.sum(torch.where(valid.unsqueeze(1), mel - mean.unsqueeze(2), 0.0) ** 2, 2) / (n.unsqueeze(1) - 1.0)
        std = torch.sqrt(var)
        std = std.masked_fill(std.isnan(), 0.0) + NORM_CONSTANT
        mel = (mel - mean.unsqueeze(2)) / std.unsqueeze(2)
    T = mel.shape[-1]
    mask = torch.arange(T, device=x.device).unsqueeze(0) >= seq_len.unsqueeze(1)
    mel = mel.masked_fill(mask.unsqueeze(1), 0.0)  # features.py:489-493
    if pad_to and pad_to > 0 and T % pad_to:
        mel = F.pad(mel, (0, pad_to - T % pad_to), value=0.0)
    if return_prenorm:
        return mel, seq_len, prenorm
    return mel, seq_len


# ----------------------------------------------------------------------------------------------
# encoder
# ----------------------------------------------------------------------------------------------
def conv_out_len(lengths: torch.Tensor, repeat: int) -> torch.Tensor:
    """calc_length(all_paddings=2, kernel=3, stride=2, ceil_mode=False) — subsampling.py:576-586.
    (Float arithmetic in the reference; integer-equal for every length < 2**24.)"""
    l = lengths.to(torch.float32)
    for _ in range(repeat):
        l = torch.floor(torch.div(l + (2 - 3), 2) + 1.0)
    return l.to(torch.int)


def _time_mask(x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """MaskedConvSequential._create_mask + apply_channel_mask, subsampling.py:707-717,755-759."""
    t = x.shape[2]
    m = (torch.arange(t, device=x.device).unsqueeze(0) < lengths.unsqueeze(1)).to(x.dtype)
    return x * m[:, None, :, None]


def subsampling(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, lengths: torch.Tensor, kind: str, factor: int,
                relu_gates=None, gate_report=None):
    """ConvSubsampling.forward — subsampling.py:385-436 with MaskedConvSequential.forward :725-753.
    x [B,T,feat] → ([B,T',d], lengths').

    relu_gates (test aid): list of bool tensors [B,C,T,F], one per ReLU, used INSTEAD of (x > 0).  Two fp32 implementations of
    the same convolution disagree on the sign of pre-activations that are within rounding of zero (a few dozen of 15 M at
    Conformer-L), and each flipped gate passes / blocks a full-size gradient — comparing gradients through ReLUs is only
    well-posed inside the same linear region.  gate_report (list) receives (n_flipped, max |pre| among flipped, max |pre|)."""
    n = int(math.log2(factor))
    gates = list(relu_gates) if relu_gates is not None else None

    def relu(t):
        if gates is None:
            return F.relu(t)
        gt = gates.pop(0).to(t.device)
        if gate_report is not None:
            flip = gt != (t > 0)
            gate_report.append((int(flip.sum()), float(t.detach().abs()[flip].max()) if flip.any() else 0.0, float(t.detach().abs().max())))
        return t * gt.to(t.dtype)

    cur = lengths.clone().float()
    x = x.unsqueeze(1)
    w = lambda i, s: sd[f"{prefix}conv.{i}.{s}"]
    if kind == "striding":
        # Sequential: [conv, relu] * n  → indices 0,2,4...
        idx = 0
        for _ in range(n):
            x = _time_mask(x, cur.long())
            x = F.conv2d(x, w(idx, "weight"), w(idx, "bias"), stride=2, padding=1)
            cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1  # calculate_conv_output_size :720-722
            x = _time_mask(x, cur.long())
            x = relu(x)
            idx += 2
        x = _time_mask(x, cur.long())
    elif kind == "dw_striding":
        # Sequential: conv0, relu1, [dw, pw, relu] * (n-1) → conv idx 0, (2,3), (5,6) ...
        x = _time_mask(x, cur.long())
        x = F.conv2d(x, w(0, "weight"), w(0, "bias"), stride=2, padding=1)
        cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1
        x = _time_mask(x, cur.long())
        x = relu(x)
        idx = 2
        for _ in range(n - 1):
            x = _time_mask(x, cur.long())
            C = x.shape[1]
            x = F.conv2d(x, w(idx, "weight"), w(idx, "bias"), stride=2, padding=1, groups=C)
            cur = torch.div(cur + 2 - 3, 2, rounding_mode="floor") + 1
            x = _time_mask(x, cur.long())
            x = F.conv2d(x, w(idx + 1, "weight"), w(idx + 1, "bias"))
            x = _time_mask(x, cur.long())
            x = relu(x)
            idx += 3
        x = _time_mask(x, cur.long())
    else:
        raise NotImplementedError(kind)
    b, c, t, f = x.shape
    x = F.linear(x.transpose(1, 2).reshape(b, t, c * f), sd[f"{prefix}out.weight"], sd[f"{prefix}out.bias"])
    return x, cur.long()


def rel_pos_table(T: int, d_model: int, device=None) -> torch.Tensor:
    """RelPositionalEncoding: rows = positions T-1 … -(T-1) (multi_head_attention.py:1007-1020,1056-1100).
    Equal to pe[:, max_len-T : max_len+T-1] of the reference's 2*max_len-1 table."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32, device=device).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32, device=device) * -(math.log(INF_VAL) / d_model))
    pe = torch.zeros(2 * T - 1, d_model, device=device)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def rel_shift(x: torch.Tensor) -> torch.Tensor:
    """multi_head_attention.py:259-270."""
    b, h, qlen, pos_len = x.shape
    x = F.pad(x, (1, 0))
    x = x.view(b, h, -1, qlen)[:, :, 1:].view(b, h, qlen, pos_len)
    return x


def rel_pos_mha(sd, p: str, x: torch.Tensor, pos_emb: torch.Tensor, att_mask: torch.Tensor, n_heads: int, use_sdpa: bool = False):
    """RelPositionMultiHeadAttention.forward — multi_head_attention.py:272-354,125-146: the default materialised branch
    (:344-349) and, with use_sdpa, the `use_pytorch_sdpa` branch (:316-343: matrix_bd as a float attn_mask).
    Dropout is not restated (all parity runs use p=0 / eval)."""
    B, T, d = x.shape
    dk = d // n_heads
    lin = lambda name, inp, bias=True: F.linear(inp, sd[f"{p}{name}.weight"], sd[f"{p}{name}.bias"] if bias else None)
    q = lin("linear_q", x).view(B, T, n_heads, dk)
    k = lin("linear_k", x).view(B, T, n_heads, dk).transpose(1, 2)
    v = lin("linear_v", x).view(B, T, n_heads, dk).transpose(1, 2)
    pe = lin("linear_pos", pos_emb, bias=False).view(1, -1, n_heads, dk).transpose(1, 2)
    q_u = (q + sd[f"{p}pos_bias_u"]).transpose(1, 2)
    q_v = (q + sd[f"{p}pos_bias_v"]).transpose(1, 2)
    bd = rel_shift(torch.matmul(q_v, pe.transpose(-2, -1)))
    if use_sdpa:
        m = att_mask.unsqueeze(1)
        bd = (bd[:, :, :, :T] * (1.0 / math.sqrt(dk))).masked_fill(m, -INF_VAL)
        out = F.scaled_dot_product_attention(q_u, k, v, attn_mask=bd, dropout_p=0.0)
        out = out.masked_fill(torch.all(m, dim=-1).unsqueeze(-1).expand(-1, out.size(1), -1, out.size(-1)), 0.0)
        return lin("linear_out", out.transpose(1, 2).reshape(B, T, d))
    ac = torch.matmul(q_u, k.transpose(-2, -1))
    scores = (ac + bd[:, :, :, :T]) / math.sqrt(dk)
    m = att_mask.unsqueeze(1)
    scores = scores.masked_fill(m, -INF_VAL)
    attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
    ctx = torch.matmul(attn, v).transpose(1, 2).reshape(B, T, d)
    return lin("linear_out", ctx)


def feed_forward(sd, p: str, x: torch.Tensor):
    """ConformerFeedForward.forward — conformer_modules.py:382-387."""
    h = F.silu(F.linear(x, sd[f"{p}linear1.weight"], sd[f"{p}linear1.bias"]))
    return F.linear(h, sd[f"{p}linear2.weight"], sd[f"{p}linear2.bias"])


def conv_module(sd, p: str, x: torch.Tensor, pad_mask: torch.Tensor, training: bool, bn_out: Optional[dict] = None):
    """ConformerConvolution.forward — conformer_modules.py:320-350; CausalConv1D symmetric pad
    (causal_convs.py:130-150); nn.BatchNorm1d (train: batch stats over (B,T) incl. padded frames)."""
    x = x.transpose(1, 2)
    x = F.conv1d(x, sd[f"{p}pointwise_conv1.weight"], sd[f"{p}pointwise_conv1.bias"])
    x = F.glu(x, dim=1)
    x = x.masked_fill(pad_mask.unsqueeze(1), 0.0)
    wd = sd[f"{p}depthwise_conv.weight"]
    k = wd.shape[-1]
    x = F.conv1d(F.pad(x, ((k - 1) // 2, (k - 1) // 2)), wd, sd[f"{p}depthwise_conv.bias"], groups=wd.shape[0])
    if training:
        mean = x.mean(dim=(0, 2))
        var_b = x.var(dim=(0, 2), unbiased=False)
        if bn_out is not None:
            n = x.shape[0] * x.shape[2]
            bn_out[p] = (mean.detach(), (var_b * n / max(n - 1, 1)).detach())  # unbiased var feeds running_var
        x = (x - mean[None, :, None]) / torch.sqrt(var_b[None, :, None] + 1e-5)
    else:
        x = (x - sd[f"{p}batch_norm.running_mean"][None, :, None]) / torch.sqrt(
            sd[f"{p}batch_norm.running_var"][None, :, None] + 1e-5
        )
    x = x * sd[f"{p}batch_norm.weight"][None, :, None] + sd[f"{p}batch_norm.bias"][None, :, None]
    x = F.silu(x)
    x = F.conv1d(x, sd[f"{p}pointwise_conv2.weight"], sd[f"{p}pointwise_conv2.bias"])
    return x.transpose(1, 2)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{p}.weight"], sd[f"{p}.bias"], 1e-5)


def conformer_layer(sd, p: str, x, pos_emb, att_mask, pad_mask, n_heads: int, training: bool, bn_out=None, use_sdpa=False):
    """ConformerLayer.forward — conformer_modules.py:160-233 (dropout omitted: parity runs use p=0)."""
    r = x + 0.5 * feed_forward(sd, f"{p}feed_forward1.", _ln(sd, f"{p}norm_feed_forward1", x))
    r = r + rel_pos_mha(sd, f"{p}self_attn.", _ln(sd, f"{p}norm_self_att", r), pos_emb, att_mask, n_heads, use_sdpa)
    r = r + conv_module(sd, f"{p}conv.", _ln(sd, f"{p}norm_conv", r), pad_mask, training, bn_out)
    r = r + 0.5 * feed_forward(sd, f"{p}feed_forward2.", _ln(sd, f"{p}norm_feed_forward2", r))
    return _ln(sd, f"{p}norm_out", r)


def encoder_forward(
    sd: Dict[str, torch.Tensor],
    audio_signal: torch.Tensor,
    length: torch.Tensor,
    *,
    n_layers: int,
    d_model: int,
    n_heads: int,
    subsampling_kind: str = "striding",
    subsampling_factor: int = 4,
    xscaling: bool = True,
    training: bool = False,
    bn_out: Optional[dict] = None,
    return_layers: bool = False,
    use_sdpa: bool = False,
    relu_gates=None,
    gate_report=None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """ConformerEncoder.forward_internal — conformer_encoder.py:593-759, full-context offline path.
    audio_signal [B,feat_in,T_mel], length [B] → (outputs [B,d_model,T'], encoded_lengths [B] i64)."""
    x = audio_signal.transpose(1, 2)
    x, length = subsampling(sd, "pre_encode.", x, length, subsampling_kind, subsampling_factor, relu_gates, gate_report)
    length = length.to(torch.int64)
    B, T, _ = x.shape
    if xscaling:
        x = x * math.sqrt(d_model)  # multi_head_attention.py:1086-1087
    pos_emb = rel_pos_table(T, d_model, x.device).unsqueeze(0)
    # _create_masks, conformer_encoder.py:794-848 (att_context_size=[-1,-1], style 'regular')
    valid = torch.arange(T, device=x.device).unsqueeze(0) < length.unsqueeze(1)
    att_mask = ~(valid.unsqueeze(1) & valid.unsqueeze(2))
    pad_mask = ~valid
    taps = []
    for i in range(n_layers):
        x = conformer_layer(sd, f"layers.{i}.", x, pos_emb, att_mask, pad_mask, n_heads, training, bn_out, use_sdpa)
        if return_layers:
            taps.append(x)
    out = x.transpose(1, 2)
    if return_layers:
        return out, length, taps
    return out, length


def ctc_head(sd, x: torch.Tensor, prefix="decoder_layers.0."):
    """ConvASRDecoder.forward — nemo/collections/asr/modules/conv_asr.py:456-468: 1x1 Conv1d + log_softmax.
    x [B,d,T] → log-probs [B,T,V+1]."""
    y = F.conv1d(x, sd[f"{prefix}weight"], sd[f"{prefix}bias"])
    return F.log_softmax(y.transpose(1, 2), dim=-1)


# ----------------------------------------------------------------------------------------------
# SpecAugment (training-time masking between the front-end and the encoder)
# ----------------------------------------------------------------------------------------------
def ctc_loss(log_probs: torch.Tensor, targets: torch.Tensor, input_lengths: torch.Tensor, target_lengths: torch.Tensor,
             blank: int, reduction: str = "mean_batch") -> torch.Tensor:
    """CTCLoss.forward — nemo/collections/asr/losses/ctc.py:44-82: log_probs [B,T,V+1] → transpose to [T,B,V+1] →
    torch.nn.CTCLoss(blank=num_classes, reduction='none', zero_infinity=False), then 'mean_batch' = losses.mean()
    (:57-60), 'mean_volume' = losses.sum() / target_lengths.sum().  The alpha recursion itself is torch's (third party)."""
    losses = F.ctc_loss(log_probs.transpose(1, 0), targets.long(), input_lengths.long(), target_lengths.long(), blank=blank,
                        reduction="none", zero_infinity=False)
    if reduction == "mean_batch":
        return losses.mean()
    if reduction == "mean_volume":
        return losses.sum() / target_lengths.sum()
    return losses


def spec_augment_vectorized(spec: torch.Tensor, length: torch.Tensor, freq_masks: int, time_masks: int, freq_width: int,
                            time_width, mask_value: float = 0.0) -> torch.Tensor:
    """SpecAugment._forward_vectorized / _apply_masks (spectr_augment.py:135-215): time masks first, then frequency
    masks; per axis one `torch.rand` draw for the widths and one for the starts (device RNG of `spec`)."""

    def apply(x, num_masks, width, axis):
        B, n = x.shape[0], x.shape[axis]
        if axis == 2 and isinstance(width, float):
            width = torch.clamp(width * length, max=n).unsqueeze(1)
        w = (torch.rand((B, num_masks), device=x.device, dtype=torch.float32) * width).long()
        st = torch.rand((B, num_masks), device=x.device, dtype=torch.float32)
        st = (st * ((length.unsqueeze(1) if axis == 2 else n) - w)).long()
        idx = torch.arange(n, device=x.device)
        m = ((idx >= st.unsqueeze(-1)) & (idx < (st + w).unsqueeze(-1))).any(dim=1)
        m = m[:, None, :] if axis == 2 else m[:, :, None]
        return x.masked_fill(m.expand_as(x), mask_value)

    return apply(apply(spec, time_masks, time_width, 2), freq_masks, freq_width, 1)


def spec_augment_legacy(spec: torch.Tensor, length: torch.Tensor, rng, freq_masks: int, time_masks: int, freq_width: int,
                        time_width, mask_value: float = 0.0) -> torch.Tensor:
    """SpecAugment._forward_legacy (spectr_augment.py:102-133): host `random.Random` draws per utterance."""
    B, F, _ = spec.shape
    lens = length.cpu().numpy()
    mask = torch.zeros(spec.shape, dtype=torch.bool)
    for b in range(B):
        for _ in range(freq_masks):
            st = rng.randint(0, F - freq_width)
            w = rng.randint(0, freq_width)
            mask[b, st:st + w, :] = True
        tmax = max(1, int(lens[b] * time_width)) if not isinstance(time_width, int) else time_width
        hi = max(1, lens[b] - tmax)
        for _ in range(time_masks):
            st = rng.randint(0, hi)
            w = rng.randint(0, tmax)
            mask[b, :, st:st + w] = True
    return spec.masked_fill(mask.to(spec.device), mask_value)


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synthetic_waveforms(batch: int, seconds: float, seed=1234, sample_rate=16000, ragged=False):
    g = torch.Generator().manual_seed(seed)
    S = int(round(seconds * sample_rate))
    x = torch.rand(batch, S, generator=g, dtype=torch.float32) * 0.2 - 0.1
    if ragged:
        length = torch.tensor([max(S - 160 * ((37 * b) % 400), 400) for b in range(batch)], dtype=torch.int64)
    else:
        length = torch.full((batch,), S, dtype=torch.int64)
    return x, length


def init_state_dict(
    *, feat_in=80, n_layers, d_model, n_heads, conv_kernel_size=31, ff_expansion_factor=4,
    subsampling_kind="striding", subsampling_factor=4, subsampling_conv_channels=-1, seed=0, bias_std=0.02,
) -> Dict[str, torch.Tensor]:
    """Random weights in the reference's state_dict layout with torch-default-like
    uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) init.  Used where /root/reference is unavailable
    (GPU box, bench); parity tests here also load weights straight from the reference module."""
    g = torch.Generator().manual_seed(seed)
    C = d_model if subsampling_conv_channels == -1 else subsampling_conv_channels
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def lin(name, out_f, in_f, bias=True):
        sd[f"{name}.weight"] = uni((out_f, in_f), in_f)
        if bias:
            sd[f"{name}.bias"] = uni((out_f,), in_f)

    n = int(math.log2(subsampling_factor))
    f_out = feat_in
    for _ in range(n):
        f_out = (f_out + 2 - 3) // 2 + 1
    if subsampling_kind == "striding":
        cin = 1
        for i in range(n):
            sd[f"pre_encode.conv.{2*i}.weight"] = uni((C, cin, 3, 3), cin * 9)
            sd[f"pre_encode.conv.{2*i}.bias"] = uni((C,), cin * 9)
            cin = C
    else:
        sd["pre_encode.conv.0.weight"] = uni((C, 1, 3, 3), 9)
        sd["pre_encode.conv.0.bias"] = uni((C,), 9)
        idx = 2
        for _ in range(n - 1):
            sd[f"pre_encode.conv.{idx}.weight"] = uni((C, 1, 3, 3), 9)
            sd[f"pre_encode.conv.{idx}.bias"] = uni((C,), 9)
            sd[f"pre_encode.conv.{idx+1}.weight"] = uni((C, C, 1, 1), C)
            sd[f"pre_encode.conv.{idx+1}.bias"] = uni((C,), C)
            idx += 3
    lin("pre_encode.out", d_model, C * f_out)
    d_ff = d_model * ff_expansion_factor
    dk = d_model // n_heads
    for i in range(n_layers):
        p = f"layers.{i}."
        for nm in ("norm_feed_forward1", "norm_self_att", "norm_conv", "norm_feed_forward2", "norm_out"):
            sd[f"{p}{nm}.weight"] = 1.0 + 0.1 * torch.randn(d_model, generator=g)
            sd[f"{p}{nm}.bias"] = 0.1 * torch.randn(d_model, generator=g)
        for ff in ("feed_forward1", "feed_forward2"):
            lin(f"{p}{ff}.linear1", d_ff, d_model)
            lin(f"{p}{ff}.linear2", d_model, d_ff)
        for nm in ("linear_q", "linear_k", "linear_v", "linear_out"):
            lin(f"{p}self_attn.{nm}", d_model, d_model)
        lin(f"{p}self_attn.linear_pos", d_model, d_model, bias=False)
        sd[f"{p}self_attn.pos_bias_u"] = torch.randn(n_heads, dk, generator=g) * bias_std
        sd[f"{p}self_attn.pos_bias_v"] = torch.randn(n_heads, dk, generator=g) * bias_std
        sd[f"{p}conv.pointwise_conv1.weight"] = uni((2 * d_model, d_model, 1), d_model)
        sd[f"{p}conv.pointwise_conv1.bias"] = uni((2 * d_model,), d_model)
        sd[f"{p}conv.depthwise_conv.weight"] = uni((d_model, 1, conv_kernel_size), conv_kernel_size)
        sd[f"{p}conv.depthwise_conv.bias"] = uni((d_model,), conv_kernel_size)
        sd[f"{p}conv.batch_norm.weight"] = 1.0 + 0.1 * torch.randn(d_model, generator=g)
        sd[f"{p}conv.batch_norm.bias"] = 0.1 * torch.randn(d_model, generator=g)
        sd[f"{p}conv.batch_norm.running_mean"] = 0.1 * torch.randn(d_model, generator=g)
        sd[f"{p}conv.batch_norm.running_var"] = torch.rand(d_model, generator=g) * 0.5 + 0.75
        sd[f"{p}conv.batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        sd[f"{p}conv.pointwise_conv2.weight"] = uni((d_model, d_model, 1), d_model)
        sd[f"{p}conv.pointwise_conv2.bias"] = uni((d_model,), d_model)
    return sd
