"""``NanoInterpreter`` / ``DetectionResult`` with the reference's API, driving B200 sessions.

Mirrors reference nanowakeword/interpreter/nanointerpreter.py for the e2e path:
``load_model`` (:310-542), ``predict`` (:606-717) -> ``_predict_e2e`` (:735-814),
``predict_clip`` (:816-833), ``reset`` (:719-733), ``detected`` (:274-295), the score
properties (:196-260) and ``_apply_post_processing`` (:1034-1064).  Same names, argument
meaning, return types and exception classes; the model behind ``self.models[name]`` is a
``B200Session`` instead of an onnxruntime session.

Not mirrored (out of scope): the embedding-mode preprocessor (its mel/embedding
networks are downloaded binaries), Silero VAD, the websocket remote verifier.  Asking for them
raises NotImplementedError instead of silently degrading.

Additions that do not change the reference surface: ``predict_batch`` for many independent
windows in one call.
"""
from __future__ import annotations

import logging
import os
import threading
import time
import wave
from collections import defaultdict, deque
from functools import partial
from typing import Callable, Dict, List, Optional, Union

import numpy as np

from .session import B200Session

_VALID_PIPELINES = {"verifier_only", "full"}


class DetectionResult:
    """Result of one ``predict()`` call (reference nanointerpreter.py:45-115)."""

    __slots__ = ("scores", "model_name", "gate_name", "threshold", "_detected")

    def __init__(self, scores: dict, model_name: str, gate_name: Optional[str], threshold: float = 0.0):
        self.scores = scores
        self.model_name = model_name
        self.gate_name = gate_name
        self.threshold = threshold
        self._detected = None

    @property
    def score(self) -> float:
        return self.scores.get(self.model_name, 0.0)

    @property
    def gate_score(self) -> float:
        return self.scores.get(self.gate_name, 0.0) if self.gate_name else 0.0

    @property
    def detected(self) -> bool:
        # always False unless a positive threshold was given (nanointerpreter.py:94-96)
        return self.score >= self.threshold if self.threshold > 0 else False

    def get(self, model_name: str, default: float = 0.0) -> float:
        return self.scores.get(model_name, default)

    def __getitem__(self, key: str) -> float:
        return self.scores[key]

    def __contains__(self, key: str) -> bool:
        return key in self.scores

    def __repr__(self) -> str:
        parts = [f"score={self.score:.4f}"]
        if self.gate_name:
            parts.append(f"gate={self.gate_score:.4f}")
        if self.threshold > 0:
            parts.append(f"detected={self.detected}")
        return f"DetectionResult({', '.join(parts)})"


class _PcmRing:
    """Last ``maxlen`` samples of a stream — the role of the reference's
    ``deque(maxlen=clip_samples)`` of Python floats (nanointerpreter.py:176-183, 751, 756).
    Kept as int16 (a window then goes to the device without a float round trip) for as long as
    every sample the caller gave is an integer in the int16 range — which ``predict`` documents;
    the first chunk that is not switches the ring to the reference's own representation,
    ``float32(x) / 32768``, for good (until ``clear``)."""

    def __init__(self, maxlen: int):
        self.maxlen = maxlen
        self.data = np.zeros(maxlen, dtype=np.int16)
        self.filled = 0

    def extend(self, x: np.ndarray):
        x = np.asarray(x).ravel()
        if x.dtype != np.int16:
            xf = x.astype(np.float32) / np.float32(32768.0)            # nanointerpreter.py:750, verbatim
            y = xf.astype(np.float64) * 32768.0
            if self.data.dtype == np.int16 and np.array_equal(y, np.rint(y)) and \
                    (y.size == 0 or (y.min() >= -32768.0 and y.max() <= 32767.0)):
                x = y.astype(np.int16)
            else:
                if self.data.dtype == np.int16:
                    self.data = self.data.astype(np.float32) / np.float32(32768.0)
                x = xf
        elif self.data.dtype != np.int16:
            x = x.astype(np.float32) / np.float32(32768.0)
        n = len(x)
        if n == 0:                 # deque.extend([]) is a no-op in the reference
            return
        if n >= self.maxlen:
            self.data[:] = x[-self.maxlen:]
            self.filled = self.maxlen
            return
        self.data[:-n] = self.data[n:]
        self.data[-n:] = x
        self.filled = min(self.maxlen, self.filled + n)

    def clear(self):
        self.data = np.zeros(self.maxlen, dtype=np.int16)
        self.filled = 0

    def __len__(self):
        return self.filled


class NanoInterpreter:
    """Inference engine front door.  Create with :meth:`load_model`."""

    session_factory = B200Session        # the seam: anything with get_inputs()/run() works

    def __init__(self, wakeword_models: List[str], **kwargs):
        self.models: Dict[str, object] = {}
        self.model_input_names: Dict[str, List[str]] = {}
        self.model_feature_length: Dict[str, int] = {}
        self.class_mapping: Dict[str, Dict[str, str]] = {}
        self.is_stateful: Dict[str, bool] = {}
        self.hidden_states: Dict[str, object] = {}
        self.raw_scores: Dict[str, float] = {}
        self.post_processed_scores: Dict[str, float] = {}
        self.is_e2e: Dict[str, bool] = {}
        self.e2e_clip_samples: Dict[str, int] = {}
        self.e2e_input_ndim: Dict[str, int] = {}
        self.e2e_buffer: Dict[str, _PcmRing] = {}
        self.e2e_buffer_samples: Dict[str, int] = {}

        device = kwargs.pop("device", 0)
        sessions = kwargs.pop("sessions", None) or {}
        for mdl_path in wakeword_models:
            name = os.path.splitext(os.path.basename(mdl_path))[0]
            if name in self.models:
                logging.warning(f"Model with name '{name}' is already loaded. Skipping.")
                continue
            session = sessions.get(name) or self.session_factory(mdl_path, device=device)
            self._register(name, session)
        self._setup_components(**kwargs)
        self.cascade_config: dict = {}
        self._listen_thread: Optional[threading.Thread] = None
        self._stop_event: Optional[threading.Event] = None

    def _register(self, name: str, session) -> None:
        self.models[name] = session
        inputs = session.get_inputs()
        self.model_input_names[name] = [i.name for i in inputs]
        self.model_feature_length[name] = inputs[0].shape[1]
        self.is_stateful[name] = "hidden_in" in self.model_input_names[name]
        if self.is_stateful[name]:
            self.hidden_states[name] = None
        self.class_mapping[name] = {"0": name}
        self.raw_scores[name] = 0.0
        self.post_processed_scores[name] = 0.0
        shape = inputs[0].shape
        # every model served by this engine takes raw PCM: e2e by construction
        # (the reference decides with metadata / a shape heuristic, nanointerpreter.py:968-992)
        self.is_e2e[name] = True
        clip = shape[-1]
        self.e2e_clip_samples[name] = clip
        self.e2e_input_ndim[name] = len(shape)
        self.e2e_buffer[name] = _PcmRing(clip)
        self.e2e_buffer_samples[name] = 0
        logging.info(f"[NanoInterpreter] E2E model '{name}' detected, clip_samples={clip}")

    def _setup_components(self, **kwargs):
        self.prediction_buffer = defaultdict(partial(deque, maxlen=30))
        enable_nr = kwargs.pop("enable_noise_reduction", False)
        self.noise_reducer_enabled = False
        if enable_nr:
            try:
                import noisereduce  # noqa: F401
                self.noise_reducer_enabled = True
            except ImportError:
                logging.warning("`enable_noise_reduction` is True, but `noisereduce` is not installed. Disabling feature.")
        self.vad_threshold = kwargs.pop("vad_threshold", 0)
        if self.vad_threshold > 0:
            raise NotImplementedError(
                "vad_threshold > 0 needs the Silero VAD model the reference downloads at first use "
                "(interpreter/models/_registry.py:34-47); it is outside the B200 hot path")
        self.preprocessor = None          # all models are e2e (nanointerpreter.py:1017-1019)

    # ------------------------------------------------------------------ properties (:196-260)
    @property
    def is_cascade(self) -> bool:
        return bool(self.cascade_config)

    @property
    def model_name(self) -> str:
        if self.is_cascade:
            return self.cascade_config["verifier"]
        return next(iter(self.models))

    @property
    def gate_name(self) -> Optional[str]:
        return self.cascade_config.get("gate")

    @property
    def gate_score(self) -> float:
        return self.post_processed_scores.get(self.gate_name, 0.0) if self.gate_name else 0.0

    @property
    def verifier_score(self) -> float:
        return self.post_processed_scores.get(self.model_name, 0.0)

    @property
    def score(self) -> float:
        return self.verifier_score

    @property
    def info(self) -> dict:
        return {
            "model_name": self.model_name,
            "is_cascade": self.is_cascade,
            "is_remote": False,
            "gate_name": self.gate_name,
            "gate_threshold": self.cascade_config.get("gate_threshold", None),
            "loaded_models": list(self.models.keys()),
            "score": self.score,
            "gate_score": self.gate_score,
            "raw_scores": dict(self.raw_scores),
        }

    def __repr__(self) -> str:
        if self.is_cascade:
            return (f"NanoInterpreter(model='{self.model_name}', gate='{self.gate_name}', "
                    f"gate_threshold={self.cascade_config.get('gate_threshold', 0.3)})")
        names = list(self.models.keys())
        return f"NanoInterpreter(model='{names[0]}')" if len(names) == 1 else f"NanoInterpreter(models={names})"

    def detected(self, threshold: float, model: Optional[str] = None) -> bool:
        return self.post_processed_scores.get(model or self.model_name, 0.0) >= threshold

    def stop(self) -> None:
        if self._stop_event is not None:
            self._stop_event.set()
        if self._listen_thread is not None and self._listen_thread.is_alive():
            self._listen_thread.join(timeout=2.0)
        self._listen_thread = None
        self._stop_event = None

    # ------------------------------------------------------------------ load_model (:310-542)
    @classmethod
    def load_model(cls, model: Union[str, List[str], None] = None, cascade: bool = False,
                   gate_model: Optional[str] = None, gate_threshold: float = 0.3,
                   remote_verifier: Optional[str] = None, remote_pipeline: str = "verifier_only",
                   remote_timeout: float = 2.0, remote_api_key: Optional[str] = None,
                   remote_token: Optional[str] = None, remote_ssl_certfile: Optional[str] = None,
                   remote_ssl_keyfile: Optional[str] = None, remote_ssl_ca_certs: Optional[str] = None, **kwargs):
        if remote_pipeline not in _VALID_PIPELINES:
            raise ValueError(f"Invalid remote_pipeline '{remote_pipeline}'. Choose from: {sorted(_VALID_PIPELINES)}")
        paths: List[str] = []
        if model is not None:
            if isinstance(model, str):
                paths = [model]
            elif isinstance(model, list):
                paths = model
            else:
                raise TypeError("`model` must be a string, list of strings, or None.")
            for p in paths:
                if not os.path.exists(p):
                    raise FileNotFoundError(f"Model file not found: {p}")
        if remote_verifier is not None:
            raise NotImplementedError("remote_verifier (websocket serving, remote_verifier.py) is outside the B200 hot path")
        if not paths:
            raise ValueError("`model` is required (no remote verifier to fall back on)")

        cascade_cfg: dict = {}
        if (cascade or gate_model is not None) and len(paths) == 1:
            main_path = paths[0]
            stem, ext = os.path.splitext(os.path.basename(main_path))
            gate_path = None
            if gate_model is not None:
                if not os.path.exists(gate_model):
                    raise FileNotFoundError(f"The specified gate model does not exist: {gate_model}")
                gate_path = gate_model
                gate_name = os.path.splitext(os.path.basename(gate_model))[0]
            else:
                gate_name = stem + "_lite"
                cand = os.path.join(os.path.dirname(os.path.abspath(main_path)), gate_name + ext)
                if os.path.exists(cand):
                    gate_path = cand
                else:
                    logging.warning(f"[NanoInterpreter] cascade=True but no lite model found at '{cand}'. "
                                    "Falling back to single-model mode.")
            if gate_path:
                paths = [gate_path, main_path]          # gate first: insertion order is evaluation order (:496)
                cascade_cfg = {"gate": gate_name, "verifier": stem, "gate_threshold": gate_threshold}

        inst = cls(wakeword_models=paths, **kwargs)
        inst.cascade_config = cascade_cfg
        return inst

    # ------------------------------------------------------------------ predict (:606-717, 735-814)
    def predict(self, x: np.ndarray, patience: dict = {}, threshold: dict = {}, debounce_time: float = 0.0) -> DetectionResult:
        if not isinstance(x, np.ndarray):
            raise ValueError("Input audio `x` must be a Numpy array.")
        if self.noise_reducer_enabled:
            x = self._reduce_noise(x)
        return self._predict_e2e(x, patience, threshold, debounce_time)

    def _predict_e2e(self, x, patience={}, threshold={}, debounce_time=0.0) -> DetectionResult:
        current: Dict[str, float] = {}
        for name, session in self.models.items():
            clip = self.e2e_clip_samples[name]
            ring = self.e2e_buffer[name]
            ring.extend(x)
            self.e2e_buffer_samples[name] += len(x)       # cumulative, cleared only by reset() (:752-753)
            if self.e2e_buffer_samples[name] >= clip:
                if self.cascade_config and name == self.cascade_config["verifier"]:
                    if current.get(self.cascade_config["gate"], 0.0) < self.cascade_config["gate_threshold"]:
                        current[name] = 0.0
                        self.raw_scores[name] = 0.0
                        continue
                window = ring.data.reshape(1, -1) if self.e2e_input_ndim.get(name, 2) != 3 else ring.data.reshape(1, 1, -1)
                score = float(session.run(None, {"input": window})[0].item())
            else:
                score = 0.0
            self.raw_scores[name] = score
            if len(self.prediction_buffer.get(name, [])) < 5:   # warm-up: first five reported as 0 (:789-790)
                score = 0.0
            current[name] = score

        final = dict(current)
        self._apply_post_processing(final, patience, threshold, debounce_time, len(x))
        for name, s in final.items():
            self.prediction_buffer[name].append(s)
            self.post_processed_scores[name] = s
        return DetectionResult(scores=dict(final), model_name=self.model_name, gate_name=self.gate_name)

    def predict_batch(self, pcm: np.ndarray, model: Optional[str] = None) -> np.ndarray:
        """Score B independent windows (B, clip_samples) int16 in one engine call; stateless
        (no ring buffer, warm-up or filters) — the batched counterpart of ``predict_clip``."""
        name = model or self.model_name
        out = self.models[name].run(None, {"input": np.asarray(pcm)})[0]
        return out.reshape(-1)

    def reset(self):
        self.prediction_buffer.clear()
        for name in self.hidden_states:
            self.hidden_states[name] = None
        for name in self.raw_scores:
            self.raw_scores[name] = 0.0
            self.post_processed_scores[name] = 0.0
        for name in self.e2e_buffer:
            self.e2e_buffer[name].clear()
            self.e2e_buffer_samples[name] = 0

    def predict_clip(self, clip: Union[str, np.ndarray], chunk_size: int = 1280, **kwargs) -> list:
        if isinstance(clip, str):
            with wave.open(clip, mode="rb") as f:
                if f.getframerate() != 16000 or f.getsampwidth() != 2 or f.getnchannels() != 1:
                    raise ValueError("Audio clip must be a 16kHz, 16-bit, single-channel WAV file.")
                data = np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16)
        elif isinstance(clip, np.ndarray):
            data = clip
        else:
            raise TypeError("`clip` must be a file path (string) or a numpy array.")
        return [self.predict(data, **kwargs)]              # e2e: one call on the whole clip (:828-830)

    def listen(self, on_detection: Optional[Callable[[str, float], None]] = None, threshold: float = 0.5,
               cooldown: float = 1.0, chunk_size: int = 1280, on_score: Optional[Callable[[float, float], None]] = None,
               on_audio: Optional[Callable[[np.ndarray], None]] = None, blocking: bool = True) -> None:
        """Microphone loop with the reference's signature and callback contract (nanointerpreter.py:835-945):
        every ``chunk_size`` frames ``on_audio(audio)``, ``predict(audio)``, ``on_score(verifier_score, gate_score)``;
        when the verifier score exceeds ``threshold`` and more than ``cooldown`` seconds passed since the last
        detection, ``on_detection(model_name, score)`` fires and the interpreter is ``reset()``.  Needs PyAudio,
        like the reference; ``blocking=False`` runs the loop on a daemon thread until ``stop()``."""
        try:
            import pyaudio
        except ImportError:
            raise ImportError("PyAudio is required for listen(). Install it with: pip install pyaudio")

        if on_detection is None:
            def on_detection(name: str, score: float) -> None:
                print(f"\nDetected '{name}'!  (score: {score:.5f})")

        def _loop():
            pa = pyaudio.PyAudio()
            stream = pa.open(format=pyaudio.paInt16, channels=1, rate=16000, input=True, frames_per_buffer=chunk_size)
            last_detection = 0.0
            stop_event = self._stop_event
            try:
                while not (stop_event and stop_event.is_set()):
                    audio = np.frombuffer(stream.read(chunk_size, exception_on_overflow=False), dtype=np.int16)
                    if on_audio is not None:
                        on_audio(audio)
                    self.predict(audio)
                    v_score = self.verifier_score
                    g_score = self.gate_score
                    if on_score is not None:
                        on_score(v_score, g_score)
                    now = time.monotonic()
                    if v_score > threshold and (now - last_detection) > cooldown:
                        on_detection(self.model_name, v_score)
                        last_detection = now
                        self.reset()
            except KeyboardInterrupt:
                pass
            finally:
                stream.stop_stream()
                stream.close()
                pa.terminate()

        if blocking:
            _loop()
        else:
            self._stop_event = threading.Event()
            self._listen_thread = threading.Thread(target=_loop, daemon=True)
            self._listen_thread.start()

    def _reduce_noise(self, x: np.ndarray) -> np.ndarray:
        try:
            import noisereduce as nr
            y = nr.reduce_noise(y=x.astype(np.float32) / 32768.0, sr=16000, stationary=True)
            return (y * 32768.0).astype(np.int16)
        except Exception as e:  # the reference also falls back to the raw chunk (:1030-1032)
            logging.warning(f"Noise reduction failed: {e}. Returning original audio.")
            return x

    # ------------------------------------------------------------------ filters (:1034-1064)
    def _apply_post_processing(self, predictions, patience, threshold, debounce_time, n_prepared_samples):
        if not patience and debounce_time <= 0:
            return
        if not threshold:
            raise ValueError("`threshold` must be provided when using `patience` or `debounce_time`.")
        if patience and debounce_time > 0:
            raise ValueError("`patience` and `debounce_time` cannot be used together.")
        for name in predictions:
            if predictions[name] == 0.0:
                continue
            hist = self.prediction_buffer[name]
            if name in patience:
                need = patience[name]
                if len(hist) < need:
                    predictions[name] = 0.0
                    continue
                # the reference slices [-(need-1):], which for need == 1 is the WHOLE buffer (:1054)
                recent = list(hist)[-(need - 1):] + [predictions[name]]
                if (np.array(recent) >= threshold[name]).sum() < need:
                    predictions[name] = 0.0
            elif debounce_time > 0 and name in threshold:
                frame_s = n_prepared_samples / 16000.0
                if frame_s <= 0:
                    continue
                k = int(np.ceil(debounce_time / frame_s))
                recent = np.array(hist)[-k:]
                if predictions[name] >= threshold[name] and (recent >= threshold[name]).any():
                    predictions[name] = 0.0
