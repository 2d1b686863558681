"""nemo_b200 — B200-native drop-in for NeMo's Conformer-ASR encoder hot path
(AudioToMelSpectrogramPreprocessor → ConvSubsampling → ConformerEncoder).  See DESIGN.md."""
from . import _lib  # noqa: F401  (loads libb200conformer.so; raises if it is missing)

__all__ = ["_lib"]
