"""Oracle of the streaming bookkeeping around the model.  TEST INFRASTRUCTURE ONLY.

Restates ``NanoInterpreter._predict_e2e`` (reference interpreter/nanointerpreter.py:735-814)
and ``_apply_post_processing`` (:1034-1064) for a single model, the way the reference does it:
a ``deque(maxlen=clip_samples)`` of Python floats (x/32768), a cumulative sample counter that
only ``reset()`` clears, the last ``clip_samples`` samples re-scored on every call once the
counter has reached ``clip_samples``, the first five outputs reported as 0.0, then
patience / debounce.  The model call is the numpy oracle (or any callable).
"""
from __future__ import annotations

from collections import deque

import numpy as np

from .heads import forward_scores


class OracleInterpreter:
    def __init__(self, sd=None, cfg=None, name="model", clip_samples=16000, score_fn=None):
        self.name = name
        self.clip = clip_samples
        if score_fn is None:
            def score_fn(clip_f32):
                # the float clip goes to the model as it is (nanointerpreter.py:771-783); for int16 input
                # float32(x) / 32768 is exact, so this equals scoring the int16 window
                return float(forward_scores(clip_f32[None, :], sd, cfg).item())
        self.score_fn = score_fn
        self.reset()

    def reset(self):
        self.buf = deque(maxlen=self.clip)
        self.buf_samples = 0
        self.prediction_buffer = deque(maxlen=30)
        self.raw_scores = {self.name: 0.0}
        self.post_processed_scores = {self.name: 0.0}

    def predict(self, x, patience={}, threshold={}, debounce_time=0.0):
        x_float = x.astype(np.float32) / 32768.0
        self.buf.extend(x_float.tolist())
        self.buf_samples += len(x)
        if self.buf_samples >= self.clip:
            clip = np.array(list(self.buf)[-self.clip:], dtype=np.float32)
            score = self.score_fn(clip)
        else:
            score = 0.0
        self.raw_scores[self.name] = score
        if len(self.prediction_buffer) < 5:
            score = 0.0
        preds = {self.name: score}
        self._post(preds, patience, threshold, debounce_time, len(x))
        self.prediction_buffer.append(preds[self.name])
        self.post_processed_scores[self.name] = preds[self.name]
        return dict(preds)

    def _post(self, preds, patience, threshold, debounce_time, n_samples):
        if not patience and debounce_time <= 0:
            return
        if (patience or debounce_time > 0) and not threshold:
            raise ValueError("`threshold` must be provided when using `patience` or `debounce_time`.")
        if patience and debounce_time > 0:
            raise ValueError("`patience` and `debounce_time` cannot be used together.")
        n = self.name
        if preds[n] == 0.0:
            return
        if n in patience:
            need = patience[n]
            if len(self.prediction_buffer) < need:
                preds[n] = 0.0
                return
            recent = np.array(list(self.prediction_buffer)[-(need - 1):] + [preds[n]])
            if (recent >= threshold[n]).sum() < need:
                preds[n] = 0.0
        elif debounce_time > 0 and n in threshold:
            dur = n_samples / 16000.0
            if dur <= 0:
                return
            k = int(np.ceil(debounce_time / dur))
            recent = np.array(self.prediction_buffer)[-k:]
            if preds[n] >= threshold[n] and (recent >= threshold[n]).any():
                preds[n] = 0.0
