"""nanowakeword_b200 — B200-native (sm_100a) engine for the nanowakeword per-window hot path.

Public surface mirrors ``nanowakeword.interpreter`` (reference nanowakeword/__init__.py:1-5,
interpreter/__init__.py:1-22): ``NanoInterpreter``, ``DetectionResult``; plus the session
duck type ``B200Session`` and the low-level ``Engine``.
"""
__version__ = "0.1.0"

from .interpreter import DetectionResult, NanoInterpreter  # noqa: F401
from .session import B200Session, Engine, load_artifacts, save_model  # noqa: F401
from .streams import CascadeBank, StreamBank  # noqa: F401


class _OutOfScope:
    """Names the reference's package exports (nanowakeword/__init__.py:1-5) whose arithmetic lives in downloaded
    binaries (interpreter/models/_registry.py:34-47: Silero VAD, the mel / embedding networks of the embedding mode):
    importing them works, constructing them says why they are not built here instead of degrading silently."""
    _what = ""

    def __init__(self, *a, **k):
        raise NotImplementedError(
            f"{type(self).__name__}: {self._what} is outside the B200 hot path (SURVEY.md §8: its weights are downloaded "
            "ONNX binaries, parity unpinned); e2e models need neither")


class VAD(_OutOfScope):
    _what = "the Silero voice-activity detector (interpreter/vad.py)"


class AudioFeatures(_OutOfScope):
    _what = "the embedding-mode preprocessor (data/AudioFeatures.py: melspectrogram.onnx + embedding_model.onnx)"


__all__ = ["NanoInterpreter", "DetectionResult", "VAD", "AudioFeatures", "B200Session", "Engine", "StreamBank", "CascadeBank",
           "load_artifacts", "save_model"]
