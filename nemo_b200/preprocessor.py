"""B200MelPreprocessor — drop-in for `AudioToMelSpectrogramPreprocessor`
(nemo/collections/asr/modules/audio_preprocessing.py:111-305) backed by the fused CUDA mel front-end
(nemo_b200/csrc/mel.cu).  Same constructor kwargs, same `forward(input_signal=, length=)` contract, same
state_dict (`featurizer.window`, `featurizer.fb`), always-fp32 arithmetic, no gradients (the reference runs the
featurizer under torch.no_grad, features.py:363-364).  No CPU / torch fallback: CPU tensors raise."""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from . import _lib as L


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """Constants only: the Slaney-scale, slaney-normalised triangular filterbank that the reference obtains from
    librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, norm='slaney') at features.py:339-341 (f64 maths → f32)."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    lin_step, knee_hz = 200.0 / 3.0, 1000.0
    knee_mel, log_step = knee_hz / lin_step, math.log(6.4) / 27.0
    to_mel = lambda f: np.where(f >= knee_hz, knee_mel + np.log(np.maximum(f, 1e-300) / knee_hz) / log_step, f / lin_step)
    to_hz = lambda m: np.where(m >= knee_mel, knee_hz * np.exp(log_step * (m - knee_mel)), lin_step * m)
    bins = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = to_hz(np.linspace(to_mel(np.float64(fmin)), to_mel(np.float64(fmax)), n_mels + 2))
    width = np.diff(edges)
    fb = np.zeros((n_mels, bins.size))
    for m in range(n_mels):
        rise = (bins - edges[m]) / width[m]
        fall = (edges[m + 2] - bins) / width[m + 1]
        fb[m] = np.clip(np.minimum(rise, fall), 0.0, None) * (2.0 / (edges[m + 2] - edges[m]))
    return fb.astype(np.float32)


class B200FilterbankFeatures(nn.Module):
    """Mirror of `FilterbankFeatures` (features.py:246-502): attributes poked by callers (`dither`, `pad_to`,
    `filter_banks`) and persistent buffers `window`, `fb` ([1, n_mels, n_fft/2+1]) keep the reference's names."""

    def __init__(self, sample_rate=16000, n_window_size=320, n_window_stride=160, window="hann", normalize="per_feature",
                 n_fft=None, preemph=0.97, nfilt=64, lowfreq=0, highfreq=None, log=True, log_zero_guard_type="add",
                 log_zero_guard_value=2 ** -24, dither=1e-5, pad_to=16, frame_splicing=1, exact_pad=False, pad_value=0,
                 mag_power=2.0, nb_augmentation_prob=0.0, mel_norm="slaney", **unused):
        super().__init__()
        unsupported = []
        if window != "hann": unsupported.append(f"window={window}")
        if not log or log_zero_guard_type != "add" or log_zero_guard_value != 2 ** -24: unsupported.append("log guard")
        if frame_splicing != 1: unsupported.append("frame_splicing")
        if exact_pad: unsupported.append("exact_pad")
        if mag_power != 2.0: unsupported.append("mag_power")
        if nb_augmentation_prob > 0: unsupported.append("nb_augmentation")
        if mel_norm != "slaney": unsupported.append("mel_norm")
        if normalize not in ("per_feature", None): unsupported.append(f"normalize={normalize}")
        if preemph is None: unsupported.append("preemph=None")
        self.n_fft = n_fft or 2 ** math.ceil(math.log2(n_window_size))
        if self.n_fft != 512: unsupported.append(f"n_fft={self.n_fft}")
        if unsupported:
            raise NotImplementedError("B200MelPreprocessor supports the Conformer recipe only; got " + ", ".join(unsupported))
        self.sample_rate, self.win_length, self.hop_length = sample_rate, n_window_size, n_window_stride
        self.normalize, self.preemph, self.nfilt = normalize, preemph, nfilt
        self.dither, self.pad_to, self.pad_value = dither, pad_to, pad_value
        self.register_buffer("window", torch.hann_window(self.win_length, periodic=False))  # features.py:326
        fb = slaney_mel_filterbank(sample_rate, self.n_fft, nfilt, lowfreq, highfreq or sample_rate / 2)
        self.register_buffer("fb", torch.from_numpy(fb).unsqueeze(0))
        self.register_buffer("_fb_ranges", torch.zeros(2 * nfilt, dtype=torch.int32), persistent=False)
        self._ranges_ready = None
        self._seed = 0

    @property
    def filter_banks(self):
        return self.fb

    def get_seq_len(self, seq_len: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(seq_len, dtype=torch.int64)
        L.call("mel_seq_len", seq_len.to(torch.int64).contiguous(), out, seq_len.numel(), self.n_fft, self.hop_length)
        return out

    def _run(self, x, seq_len, normalize):
        if not x.is_cuda:
            raise RuntimeError("B200MelPreprocessor runs on CUDA tensors only (no CPU fallback)")
        x = x.to(torch.float32).contiguous()
        length = seq_len.to(device=x.device, dtype=torch.int64).contiguous()
        B, S = x.shape
        T = S // self.hop_length + 1
        key = (self.fb.data_ptr(), self.fb._version)
        if self._ranges_ready != key:
            L.call("mel_fb_ranges", self.fb, self.nfilt, self.n_fft // 2 + 1, self._fb_ranges)
            self._ranges_ready = key
        mel = torch.empty(B, self.nfilt, T, dtype=torch.float32, device=x.device)
        out_len = torch.empty(B, dtype=torch.int64, device=x.device)
        dither = float(self.dither) if (self.training and self.dither > 0) else 0.0
        self._seed += 1
        L.call("mel_frontend", x, length, self.window, self.fb, self._fb_ranges, mel, out_len, B, S, T, self.nfilt,
               self.win_length, self.hop_length, self.n_fft, float(self.preemph), dither, self._seed,
               1 if normalize else 0, float(self.pad_value))
        return mel, out_len

    @torch.no_grad()
    def forward(self, x, seq_len, linear_spec=False):
        if linear_spec:
            raise NotImplementedError("linear_spec=True is outside the hot path")
        mel, out_len = self._run(x, seq_len, self.normalize == "per_feature")
        pad_to = self.pad_to
        if pad_to == "max":
            raise NotImplementedError("pad_to='max'")
        if pad_to and pad_to > 0 and mel.shape[-1] % pad_to:
            mel = nn.functional.pad(mel, (0, pad_to - mel.shape[-1] % pad_to), value=self.pad_value)
        return mel, out_len

    @torch.no_grad()
    def forward_raw(self, x, seq_len):
        """log-mel before normalisation (frames >= seq_len zeroed) — used by the parity tests."""
        return self._run(x, seq_len, False)


class B200MelPreprocessor(nn.Module):
    """Constructor kwargs of AudioToMelSpectrogramPreprocessor.__init__ (audio_preprocessing.py:214-246)."""

    def __init__(self, sample_rate=16000, window_size=0.02, window_stride=0.01, n_window_size=None, n_window_stride=None,
                 window="hann", normalize="per_feature", n_fft=None, preemph=0.97, features=64, lowfreq=0, highfreq=None,
                 log=True, log_zero_guard_type="add", log_zero_guard_value=2 ** -24, dither=1e-5, pad_to=16,
                 frame_splicing=1, exact_pad=False, pad_value=0, mag_power=2.0, rng=None, nb_augmentation_prob=0.0,
                 nb_max_freq=4000, use_torchaudio: bool = False, mel_norm="slaney", stft_exact_pad=False, stft_conv=False):
        super().__init__()
        self._sample_rate = sample_rate
        if window_size and n_window_size:
            raise ValueError(f"{self} received both window_size and n_window_size. Only one should be specified.")
        if window_stride and n_window_stride:
            raise ValueError(f"{self} received both window_stride and n_window_stride. Only one should be specified.")
        if window_size:
            n_window_size = int(window_size * self._sample_rate)
        if window_stride:
            n_window_stride = int(window_stride * self._sample_rate)
        if use_torchaudio:
            raise NotImplementedError("use_torchaudio=True")
        self.win_length, self.hop_length = n_window_size, n_window_stride
        self.register_buffer("dtype_sentinel_tensor", torch.tensor((), dtype=torch.float32), persistent=False)
        self.featurizer = B200FilterbankFeatures(
            sample_rate=sample_rate, n_window_size=n_window_size, n_window_stride=n_window_stride, window=window,
            normalize=normalize, n_fft=n_fft, preemph=preemph, nfilt=features, lowfreq=lowfreq, highfreq=highfreq, log=log,
            log_zero_guard_type=log_zero_guard_type, log_zero_guard_value=log_zero_guard_value, dither=dither,
            pad_to=pad_to, frame_splicing=frame_splicing, exact_pad=exact_pad, pad_value=pad_value, mag_power=mag_power,
            nb_augmentation_prob=nb_augmentation_prob, mel_norm=mel_norm)

    @classmethod
    def from_reference(cls, ref_module):
        """Build from a live reference AudioToMelSpectrogramPreprocessor (copies window / fb buffers)."""
        f = ref_module.featurizer
        new = cls(sample_rate=f.sample_rate, window_size=None, window_stride=None, n_window_size=f.win_length,
                  n_window_stride=f.hop_length, normalize=f.normalize, n_fft=f.n_fft, preemph=f.preemph, features=f.nfilt,
                  dither=f.dither, pad_to=f.pad_to, pad_value=f.pad_value)
        new.load_state_dict(ref_module.state_dict(), strict=False)
        return new.to(f.fb.device)

    @property
    def filter_banks(self):
        return self.featurizer.filter_banks

    def get_features(self, input_signal, length):
        return self.featurizer(input_signal, length)

    @torch.no_grad()
    def forward(self, input_signal, length):
        # audio_preprocessing.py:93-103: always fp32 inside; output cast to the sentinel dtype
        processed_signal, processed_length = self.get_features(input_signal.to(torch.float32), length)
        return processed_signal.to(self.dtype_sentinel_tensor.dtype), processed_length
