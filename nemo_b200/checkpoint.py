""".nemo checkpoint import / export for the hot-path modules.

A `.nemo` file is an (uncompressed; gzip for pre-1.7 files) tar whose members are `./model_config.yaml` and
`./model_weights.ckpt` — `torch.save` of the MODEL-level state_dict, i.e. keys `preprocessor.featurizer.{window,fb}`,
`encoder.pre_encode...`, `encoder.layers.N...`, `decoder.decoder_layers.0.{weight,bias}` — plus tokenizer artifacts
(nemo/core/connectors/save_restore_connector.py:44-45 file names, :599-607 `_make_nemo_file_from_folder` = tar "w:" with
arcname ".", :700-745 `_unpack_nemo_file` = "r:" then "r:gz", :753-759 torch.save / torch.load of the state_dict).
The B200 modules keep the reference's state_dict key layout, so published checkpoints (e.g. stt_en_conformer_ctc_large)
load unchanged and what we save loads back into the reference.  Host-side only; no GPU needed."""
from __future__ import annotations

import io
import os
import re
import tarfile
from typing import Dict, Optional, Tuple

import torch

MODEL_CONFIG_YAML = "model_config.yaml"    # save_restore_connector.py:44
MODEL_WEIGHTS_CKPT = "model_weights.ckpt"  # save_restore_connector.py:45
_PREFIXES = ("preprocessor", "encoder", "decoder")


def _open_tar(path: str) -> tarfile.TarFile:
    try:
        return tarfile.open(path, "r:")  # versions >= 1.7.0 (save_restore_connector.py:736-744)
    except tarfile.ReadError:
        return tarfile.open(path, "r:gz")


def _safe(name: str) -> bool:
    # save_restore_connector.py:622-640: no absolute paths, no '..' components
    norm = os.path.normpath(name)
    return not (os.path.isabs(norm) or ".." in norm.split(os.sep))


def _member(tar: tarfile.TarFile, basename: str):
    for m in tar.getmembers():
        if _safe(m.name) and os.path.normpath(m.name) == basename and m.isfile():
            return m
    return None


def load_nemo(path: str) -> Tuple[object, Dict[str, torch.Tensor]]:
    """Returns (config, state_dict): config is the parsed YAML (dict) when PyYAML is importable, else the raw text; `model:`
    is unwrapped like the reference does (save_restore_connector.py:172-174)."""
    with _open_tar(path) as tar:
        mw = _member(tar, MODEL_WEIGHTS_CKPT)
        if mw is None:
            raise FileNotFoundError(f"{path}: no {MODEL_WEIGHTS_CKPT} member")
        sd = torch.load(io.BytesIO(tar.extractfile(mw).read()), map_location="cpu", weights_only=True)
        cfg = None
        mc = _member(tar, MODEL_CONFIG_YAML)
        if mc is not None:
            text = tar.extractfile(mc).read().decode("utf-8")
            try:
                import yaml

                cfg = yaml.safe_load(text)
                if isinstance(cfg, dict) and "model" in cfg:
                    cfg = cfg["model"]
            except Exception:  # noqa: BLE001
                cfg = text
    return cfg, sd


def split_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    """model-level keys → {'preprocessor': {...}, 'encoder': {...}, 'decoder': {...}, 'other': {...}}"""
    out = {p: {} for p in _PREFIXES}
    out["other"] = {}
    for k, v in sd.items():
        head, _, rest = k.partition(".")
        (out[head] if head in out and rest else out["other"])[rest if head in _PREFIXES else k] = v
    return out


def restore(path: str, preprocessor=None, encoder=None, decoder=None, strict: bool = True):
    """Load the sub-module weights of a .nemo file into B200 modules (any subset).  Returns the parsed config."""
    cfg, sd = load_nemo(path)
    parts = split_state_dict(sd)
    for name, mod in (("preprocessor", preprocessor), ("encoder", encoder), ("decoder", decoder)):
        if mod is None:
            continue
        if not parts[name] and strict:
            raise KeyError(f"{path}: no '{name}.*' weights")
        missing, unexpected = mod.load_state_dict(parts[name], strict=False)
        if strict and (missing or unexpected):
            raise KeyError(f"{path}: {name}: missing {list(missing)[:5]} unexpected {list(unexpected)[:5]}")
    return cfg


_INTERP = re.compile(r"^\$\{([\w\.]+)\}$")


def resolve_config(cfg, root=None, _depth=0):
    """Resolve OmegaConf-style absolute interpolations (`${model.preprocessor.features}`) in a plain-YAML config tree.  A `.nemo`
    file stores its config already resolved (save_restore_connector.py:590-591, `resolve=True`); the training YAMLs under
    examples/asr/conf/ do not.  Only whole-value `${a.b.c}` references are handled (all these configs use); anything else is
    left as it is."""
    root = cfg if root is None else root
    if _depth > 8:
        return cfg
    if isinstance(cfg, dict):
        return {k: resolve_config(v, root, _depth) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [resolve_config(v, root, _depth) for v in cfg]
    if isinstance(cfg, str):
        m = _INTERP.match(cfg)
        if m:
            node = root
            for part in m.group(1).split("."):
                if not isinstance(node, dict) or part not in node:
                    return cfg
                node = node[part]
            return resolve_config(node, root, _depth + 1)
    return cfg


def _kwargs(cfg) -> dict:
    return {k: v for k, v in dict(cfg).items() if k != "_target_"}


def encoder_from_config(cfg: dict):
    """Instantiate B200ConformerEncoder from the (resolved) `encoder:` section of a NeMo model config (the `_target_` key names
    the reference class and is dropped — common.py:60-100 would not let a YAML name ours anyway)."""
    from .encoder import B200ConformerEncoder

    return B200ConformerEncoder(**_kwargs(cfg))


def preprocessor_from_config(cfg: dict):
    from .preprocessor import B200MelPreprocessor

    return B200MelPreprocessor(**_kwargs(cfg))


def decoder_from_config(cfg: dict):
    from .decoder import B200ConvASRDecoder

    return B200ConvASRDecoder(**_kwargs(cfg))


def save_nemo(path: str, config, preprocessor=None, encoder=None, decoder=None, extra: Optional[Dict[str, torch.Tensor]] = None):
    """Write a .nemo file in the reference's layout: model-level state_dict (`<module>.<key>`), YAML config, tar 'w:' with
    members under './'.  `config` is a dict (dumped with PyYAML) or YAML text."""
    sd: Dict[str, torch.Tensor] = {}
    for name, mod in (("preprocessor", preprocessor), ("encoder", encoder), ("decoder", decoder)):
        if mod is not None:
            for k, v in mod.state_dict().items():
                sd[f"{name}.{k}"] = v.detach().cpu()
    if extra:
        sd.update({k: v.detach().cpu() for k, v in extra.items()})
    if isinstance(config, str):
        text = config
    else:
        import yaml

        text = yaml.safe_dump(config, sort_keys=False)
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    with tarfile.open(path, "w:") as tar:
        root = tarfile.TarInfo(".")
        root.type = tarfile.DIRTYPE
        root.mode = 0o755
        tar.addfile(root)
        for name, blob in ((MODEL_CONFIG_YAML, text.encode("utf-8")), (MODEL_WEIGHTS_CKPT, _dump(sd))):
            ti = tarfile.TarInfo("./" + name)
            ti.size = len(blob)
            ti.mode = 0o644
            tar.addfile(ti, io.BytesIO(blob))
    return path


def _dump(sd) -> bytes:
    buf = io.BytesIO()
    torch.save(sd, buf)
    return buf.getvalue()
