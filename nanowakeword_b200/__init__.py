"""nanowakeword_b200 — B200-native (sm_100a) engine for the nanowakeword per-window hot path.

Public surface mirrors ``nanowakeword.interpreter`` (reference nanowakeword/__init__.py:1-5,
interpreter/__init__.py:1-22): ``NanoInterpreter``, ``DetectionResult``; plus the session
duck type ``B200Session`` and the low-level ``Engine``.
"""
__version__ = "0.1.0"

from .interpreter import DetectionResult, NanoInterpreter  # noqa: F401
from .session import B200Session, Engine, load_artifacts, save_model  # noqa: F401
from .streams import CascadeBank, StreamBank  # noqa: F401
