"""TEST INFRASTRUCTURE — loads the *unmodified* reference sources from /root/reference.

`import nemo` is impossible in this image (omegaconf / hydra / lightning / librosa are absent), so the
hot-path source files are exec'd file-by-file under their real dotted names after seeding `sys.modules`
with stub parents.  Nothing here restates reference arithmetic except
`_slaney_mel` (librosa.filters.mel is a third-party dependency that is not under /root/reference;
`librosa>=0.10.1`, call site nemo/collections/asr/parts/preprocessing/features.py:339-341).

Only tests/ and tools/make_golden.py import this module.  It cannot work on the GPU box
(/root/reference does not exist there): callers must check `available()` first.
"""
from __future__ import annotations

import contextlib
import dataclasses
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("B2C_REFERENCE_ROOT", "/root/reference")

_FILES = [
    "nemo/utils/cast_utils.py",
    "nemo/collections/asr/parts/submodules/causal_convs.py",
    "nemo/collections/asr/parts/submodules/batchnorm.py",
    "nemo/collections/asr/parts/utils/activations.py",
    "nemo/collections/asr/parts/utils/regularization_utils.py",
    "nemo/collections/asr/parts/mixins/streaming.py",
    "nemo/collections/asr/parts/submodules/multi_head_attention.py",
    "nemo/collections/asr/parts/submodules/subsampling.py",
    "nemo/collections/asr/parts/submodules/conformer_modules.py",
    "nemo/collections/asr/modules/conformer_encoder.py",
    "nemo/collections/asr/parts/preprocessing/features.py",
]


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_ROOT, f)) for f in _FILES)


def _slaney_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, norm="slaney", **_):
    """librosa.filters.mel (htk=False) restated from its published algorithm: Slaney mel scale
    (linear below 1 kHz with f_sp = 200/3 Hz/mel, log above with step ln(6.4)/27), triangular
    weights, slaney area normalisation 2/(f[i+2]-f[i]); computed in f64, returned as f32."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.linspace(0.0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, None]
    return weights.astype(np.float32)


class _Permissive(types.ModuleType):
    """Module whose unknown attributes resolve to empty shell classes (neural-type names etc.)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        shell = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, shell)
        return shell


def _mod(name, permissive=False, **attrs):
    m = (_Permissive if permissive else types.ModuleType)(name)
    m.__path__ = []  # behave like a package
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


class _Logging:
    def __getattr__(self, name):
        return lambda *a, **k: None


class _LoggingMode:
    ONCE = 1
    EACH = 0


def _typecheck(*dargs, **dkw):
    def deco(fn):
        return fn

    return deco


_typecheck.disable_checks = contextlib.nullcontext


class _AccessMixin:
    def is_access_enabled(self, *a, **k):
        return False

    @classmethod
    def set_access_enabled(cls, *a, **k):
        pass


class _AttentionAdapterModuleMixin:
    def is_adapter_available(self):
        return False


@dataclasses.dataclass
class _CacheAwareStreamingConfig:
    # nemo/collections/asr/models/configs/asr_models_config.py:120-141 (field names + defaults only)
    chunk_size: int = 0
    shift_size: int = 0
    cache_drop_size: int = 0
    last_channel_cache_size: int = 0
    valid_out_len: int = 0
    pre_encode_cache_size: int = 0
    drop_extra_pre_encoded: int = 0
    last_channel_num: int = 0
    last_time_num: int = 0


_loaded = None


def load():
    """Returns a namespace with the reference classes; raises if /root/reference is absent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_ROOT}")

    from torch import nn

    _mod("nemo")
    utils = _mod("nemo.utils", logging=_Logging(), logging_mode=_LoggingMode)
    _mod("nemo.collections")
    _mod("nemo.collections.asr")
    _mod("nemo.collections.asr.parts")
    _mod("nemo.collections.asr.parts.submodules")
    _mod("nemo.collections.asr.parts.submodules.adapters")
    _mod(
        "nemo.collections.asr.parts.submodules.adapters.attention_adapter_mixin",
        AttentionAdapterModuleMixin=_AttentionAdapterModuleMixin,
    )
    _mod("nemo.collections.asr.parts.utils", adapter_utils=_Permissive("adapter_utils"))
    _mod("nemo.collections.asr.parts.mixins")
    _mod("nemo.collections.asr.parts.preprocessing")
    _mod("nemo.collections.asr.parts.preprocessing.perturb", permissive=True)
    _mod("nemo.collections.asr.parts.preprocessing.segment", permissive=True)
    _mod("nemo.collections.asr.modules")
    _mod("nemo.collections.asr.models")
    _mod("nemo.collections.asr.models.configs", CacheAwareStreamingConfig=_CacheAwareStreamingConfig)
    _mod("nemo.collections.common")
    _mod("nemo.collections.common.parts")
    # nemo/collections/common/parts/utils.py:29-37
    _mod(
        "nemo.collections.common.parts.utils",
        activation_registry={
            "identity": nn.Identity,
            "hardtanh": nn.Hardtanh,
            "relu": nn.ReLU,
            "selu": nn.SELU,
            "swish": nn.SiLU,
            "silu": nn.SiLU,
            "gelu": nn.GELU,
        },
    )
    _mod("nemo.core")
    _mod("nemo.core.classes")
    _mod("nemo.core.classes.common", typecheck=_typecheck)
    _mod("nemo.core.classes.exportable", Exportable=type("Exportable", (), {}))
    _mod("nemo.core.classes.mixins", AccessMixin=_AccessMixin, adapter_mixins=_Permissive("adapter_mixins"))
    _mod("nemo.core.classes.module", NeuralModule=nn.Module)
    _mod("nemo.core.neural_types", permissive=True)
    if "omegaconf" not in sys.modules:
        _mod(
            "omegaconf",
            DictConfig=type("DictConfig", (dict,), {}),
            ListConfig=type("ListConfig", (list,), {}),
            open_dict=contextlib.nullcontext,
        )
    if "librosa" not in sys.modules:
        lib = _mod("librosa")
        _mod("librosa.filters", mel=_slaney_mel)
        lib.filters = sys.modules["librosa.filters"]

    ns = types.SimpleNamespace()
    for rel in _FILES:
        name = rel[:-3].replace("/", ".")
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], child, m)
        spec.loader.exec_module(m)
        if name == "nemo.utils.cast_utils":
            utils.avoid_float16_autocast_context = m.avoid_float16_autocast_context
            utils.cast_utils = m
    ns.ConformerEncoder = sys.modules["nemo.collections.asr.modules.conformer_encoder"].ConformerEncoder
    ns.ConformerLayer = sys.modules["nemo.collections.asr.parts.submodules.conformer_modules"].ConformerLayer
    ns.ConvSubsampling = sys.modules["nemo.collections.asr.parts.submodules.subsampling"].ConvSubsampling
    mha = sys.modules["nemo.collections.asr.parts.submodules.multi_head_attention"]
    ns.RelPositionMultiHeadAttention = mha.RelPositionMultiHeadAttention
    ns.RelPositionalEncoding = mha.RelPositionalEncoding
    ns.FilterbankFeatures = sys.modules["nemo.collections.asr.parts.preprocessing.features"].FilterbankFeatures
    ns.slaney_mel = _slaney_mel
    # SpecAugment: optional — only the pin test needs it
    rel = "nemo/collections/asr/parts/submodules/spectr_augment.py"
    if os.path.exists(os.path.join(REF_ROOT, rel)):
        core = sys.modules["nemo.core.classes"]
        core.Typing = type("Typing", (), {})
        core.typecheck = _typecheck
        name = rel[:-3].replace("/", ".")
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        ns.SpecAugment = m.SpecAugment
    # CTC head: ConvASRDecoder lives in conv_asr.py next to the Jasper encoder, so jasper.py and
    # tdnn_attention.py are exec'd too (unmodified); only the adapter mixin needs a stub class
    rels = ["nemo/collections/asr/parts/submodules/jasper.py", "nemo/collections/asr/parts/submodules/tdnn_attention.py",
            "nemo/collections/asr/modules/conv_asr.py"]
    if all(os.path.exists(os.path.join(REF_ROOT, r)) for r in rels):

        class _AdapterModuleMixin:
            def set_accepted_adapter_types(self, *_a, **_k):
                pass

            def is_adapter_available(self):
                return False

        am = _mod("nemo.core.classes.mixins.adapter_mixins", AdapterModuleMixin=_AdapterModuleMixin,
                  get_registered_adapter=lambda cls: True, register_adapter=lambda **k: None)
        sys.modules["nemo.core.classes.mixins"].adapter_mixins = am
        om = sys.modules["omegaconf"]
        for nm in ("MISSING", "OmegaConf"):
            if not hasattr(om, nm):
                setattr(om, nm, "???" if nm == "MISSING" else type("OmegaConf", (), {}))
        for rel in rels:
            name = rel[:-3].replace("/", ".")
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = m
            spec.loader.exec_module(m)
        ns.ConvASRDecoder = sys.modules["nemo.collections.asr.modules.conv_asr"].ConvASRDecoder
    _loaded = ns
    return ns


def make_encoder(seed=0, bias_std=0.02, **kw):
    """Reference ConformerEncoder with default init under manual_seed(seed); pos_bias_u/v ~ N(0, bias_std)
    (zeros-init hides rel-pos bugs); BatchNorm running stats randomised for eval parity."""
    ref = load()
    torch.manual_seed(seed)
    enc = ref.ConformerEncoder(**kw)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for layer in enc.layers:
            layer.self_attn.pos_bias_u.copy_(torch.randn(layer.self_attn.pos_bias_u.shape, generator=g) * bias_std)
            layer.self_attn.pos_bias_v.copy_(torch.randn(layer.self_attn.pos_bias_v.shape, generator=g) * bias_std)
            bn = layer.conv.batch_norm
            bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.1)
            bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) * 0.5 + 0.75)
    return enc
