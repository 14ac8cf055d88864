"""``StreamBank`` — the reference's per-stream interpreter state for MANY streams at once.

One ``NanoInterpreter`` of the reference serves one audio stream: a ring of the last
``clip_samples`` samples, a cumulative counter, five warm-up predictions reported as 0, a
30-deep prediction history and the patience / debounce filters
(reference nanowakeword/interpreter/nanointerpreter.py:176-183, 735-814, 1002, 1034-1064).
``StreamBank`` keeps exactly that state for ``n_streams`` independent streams: the audio rings
live on the GPU (``nww_stream_*`` in include/nww_b200.h), the O(1)-per-call bookkeeping is
vectorised numpy on the host.  Stream ``i`` of a bank behaves like its own reference
interpreter fed the same chunks.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

HISTORY = 30          # prediction_buffer depth (nanointerpreter.py:1002)
WARMUP = 5            # predictions forced to 0.0 after (re)start (:789-790)


class StreamBank:
    def __init__(self, engine, n_streams: int):
        self.engine = engine
        self.n = int(n_streams)
        engine.stream_open(self.n)
        self.raw_scores = np.zeros(self.n, np.float32)
        self.post_processed_scores = np.zeros(self.n, np.float32)
        self._hist = np.zeros((self.n, HISTORY), np.float32)   # column HISTORY-1 is the most recent
        self._hist_len = np.zeros(self.n, np.int64)

    def close(self):
        self.engine.stream_close()

    # ------------------------------------------------------------------ predict for every stream
    def push(self, chunks: np.ndarray, patience: int = 0, threshold: float = 0.0, debounce_time: float = 0.0) -> np.ndarray:
        """Give every stream ``chunks[i]`` (int16, any common length) and return the (n_streams,)
        post-processed scores, as ``predict()`` would stream by stream.  ``patience`` /
        ``threshold`` / ``debounce_time`` apply to all streams (the reference keys them by model
        name; a bank holds one model)."""
        if not isinstance(chunks, np.ndarray):
            raise ValueError("Input audio `chunks` must be a Numpy array.")
        if chunks.dtype != np.int16:
            chunks = chunks.astype(np.int16)
        raw = self.engine.stream_push_host(chunks)
        return self._finish(raw, chunks.shape[1], patience, threshold, debounce_time)

    def _finish(self, raw: np.ndarray, n_samples: int, patience, threshold, debounce_time) -> np.ndarray:
        self.raw_scores = raw.astype(np.float32, copy=True)
        final = raw.astype(np.float32, copy=True)
        final[self._hist_len < WARMUP] = 0.0
        self._post(final, patience, threshold, debounce_time, n_samples)
        self._hist[:, :-1] = self._hist[:, 1:]
        self._hist[:, -1] = final
        self._hist_len = np.minimum(self._hist_len + 1, HISTORY)
        self.post_processed_scores = final
        return final.copy()

    def _post(self, final, patience, threshold, debounce_time, n_samples):
        """Vectorised ``_apply_post_processing`` (nanointerpreter.py:1034-1064)."""
        if not patience and debounce_time <= 0:
            return
        if not threshold:
            raise ValueError("`threshold` must be provided when using `patience` or `debounce_time`.")
        if patience and debounce_time > 0:
            raise ValueError("`patience` and `debounce_time` cannot be used together.")
        live = final != 0.0
        cols = np.arange(HISTORY)[None, :]
        valid = cols >= (HISTORY - self._hist_len)[:, None]            # entries that exist in each history
        if patience:
            need = int(patience)
            short = self._hist_len < need
            # the reference slices [-(need-1):], which for need == 1 is the whole buffer (:1054)
            depth = HISTORY if need == 1 else need - 1
            recent = valid & (cols >= HISTORY - depth)
            hits = ((self._hist >= threshold) & recent).sum(1) + (final >= threshold)
            final[live & (short | (hits < need))] = 0.0
        else:
            frame_s = n_samples / 16000.0
            if frame_s <= 0:
                return
            k = int(np.ceil(debounce_time / frame_s))
            recent = valid & (cols >= HISTORY - k)
            fired = ((self._hist >= threshold) & recent).any(1)
            final[live & (final >= threshold) & fired] = 0.0

    def detected(self, threshold: float) -> np.ndarray:
        return self.post_processed_scores >= threshold

    def reset(self, ids: Optional[np.ndarray] = None):
        """``reset()`` of the listed streams (all when ``ids`` is None): audio ring, counters,
        scores and prediction history (nanointerpreter.py:719-733)."""
        self.engine.stream_reset(ids)
        sel = slice(None) if ids is None else np.asarray(ids, dtype=np.int64)
        self.raw_scores[sel] = 0.0
        self.post_processed_scores[sel] = 0.0
        self._hist[sel] = 0.0
        self._hist_len[sel] = 0


class CascadeBank:
    """Two-stage cascade (gate -> verifier) for many streams: the reference's ``load_model(cascade=True)``
    (nanointerpreter.py:476-496) evaluated per stream as ``_predict_e2e`` does (:758-769).  Both models keep their own
    ring, counters and history for every stream and both receive every chunk; a stream's verifier result is
    reported (and remembered) as 0.0 — raw score included — whenever its gate score of the same call, after the
    gate's own warm-up zeroing, is below ``gate_threshold``.  ``push`` returns the verifier's post-processed scores,
    like ``DetectionResult.score``; ``gate_scores`` holds the gate's (``DetectionResult.gate_score``).

    The verifier engine receives every chunk (its rings must advance) but SCORES only the streams whose gate fired
    (``nww_stream_push_select_host``): a gated-off stream costs the verifier one ingest step, so the bank's throughput
    scales with the gate's pass rate.  ``select_in_engine=False`` keeps the round-1 form (score everything, zero on
    the host) for A/B measurements.
    """

    def __init__(self, gate_engine, verifier_engine, n_streams: int, gate_threshold: float = 0.3, select_in_engine: bool = True):
        self.select_in_engine = bool(select_in_engine)
        self.gate = StreamBank(gate_engine, n_streams)
        self.verifier = StreamBank(verifier_engine, n_streams)
        self.n = int(n_streams)
        self.gate_threshold = float(gate_threshold)
        self.gate_scores = np.zeros(self.n, np.float32)

    def close(self):
        self.gate.close()
        self.verifier.close()

    def push(self, chunks: np.ndarray, patience: int = 0, threshold: float = 0.0, debounce_time: float = 0.0) -> np.ndarray:
        if not isinstance(chunks, np.ndarray):
            raise ValueError("Input audio `chunks` must be a Numpy array.")
        if chunks.dtype != np.int16:
            chunks = chunks.astype(np.int16)
        g = self.gate
        graw = g.engine.stream_push_host(chunks)
        gcur = graw.astype(np.float32, copy=True)
        gcur[g._hist_len < WARMUP] = 0.0                      # what `current[gate]` holds when the verifier is reached
        self.gate_scores = g._finish(graw, chunks.shape[1], 0, 0.0, 0.0)   # filters are keyed by the verifier's name
        if self.select_in_engine:
            passed = np.flatnonzero(gcur >= self.gate_threshold)
            vraw = self.verifier.engine.stream_push_host(chunks, select=passed).astype(np.float32, copy=True)
            self.last_pass_rate = passed.size / max(1, self.n)
        else:
            vraw = self.verifier.engine.stream_push_host(chunks).astype(np.float32, copy=True)
            vraw[gcur < self.gate_threshold] = 0.0            # skipped: score and raw score are 0.0 (:760-766)
        return self.verifier._finish(vraw, chunks.shape[1], patience, threshold, debounce_time)

    @property
    def raw_scores(self):
        return self.verifier.raw_scores

    def detected(self, threshold: float) -> np.ndarray:
        return self.verifier.detected(threshold)

    def reset(self, ids: Optional[np.ndarray] = None):
        self.gate.reset(ids)
        self.verifier.reset(ids)
        if ids is None:
            self.gate_scores[:] = 0.0
        else:
            self.gate_scores[np.asarray(ids, dtype=np.int64)] = 0.0
