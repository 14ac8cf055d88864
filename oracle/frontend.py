"""Oracle front end: int16 PCM -> log-mel (dB).  TEST INFRASTRUCTURE ONLY.

Restates, in numpy, what the reference computes with
``torchaudio.transforms.MelSpectrogram`` + ``AmplitudeToDB``
(reference nanowakeword/modules/architectures.py:830-837, 869-878) and what it
deploys as a conv1d-DFT (reference nanowakeword/_export/onnx.py:27-83):

    x = int16 / 32768                      (nanointerpreter.py:750)
    reflect-pad n_fft//2 when centred      (_export/onnx.py:70-72)
    frames of win_length at hop_length, periodic Hann, DFT of n_fft points
    power = re^2 + im^2                    (_export/onnx.py:77)
    mel   = power @ fb  (HTK, norm None)   (_export/onnx.py:81-82)
    dB    = 10*log10(max(mel, 1e-10))      (AmplitudeToDB defaults; architectures.py:837)

Two geometries are pinned:
  REF64x101  n_fft=win=400, hop 160, centred+reflect, 201 bins, 64 mels -> (64,101)
  NS40x98    frame 400 zero-extended to n_fft 512, hop 160, not centred,
             257 bins, 40 mels -> (40,98)   (the geometry BASELINE.json names)
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, asdict

import numpy as np


@dataclass(frozen=True)
class FrontendSpec:
    name: str
    sample_rate: int = 16000
    n_fft: int = 400
    win_length: int = 400
    hop_length: int = 160
    n_mels: int = 64
    center: bool = True          # reflect padding of n_fft//2 on both sides
    f_min: float = 0.0
    f_max: float = 8000.0
    amin: float = 1e-10
    clip_samples: int = 16000

    @property
    def n_freqs(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def n_frames(self) -> int:
        if self.center:
            return 1 + self.clip_samples // self.hop_length
        return 1 + (self.clip_samples - self.win_length) // self.hop_length

    def to_dict(self) -> dict:
        return asdict(self)


GEOMETRIES = {
    "REF64x101": FrontendSpec("REF64x101", n_fft=400, win_length=400, hop_length=160,
                              n_mels=64, center=True),
    "NS40x98": FrontendSpec("NS40x98", n_fft=512, win_length=400, hop_length=160,
                            n_mels=40, center=False),
}


_TABLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "nanowakeword_b200", "tables")


def reference_tables(spec: "FrontendSpec"):
    """(window, fb) float64 copies of the float32 tables torchaudio builds for a pinned
    geometry (dumped by tools/make_tables.py), or None for an unpinned geometry."""
    path = os.path.join(_TABLE_DIR, spec.name + ".npz")
    if not os.path.exists(path):
        return None
    t = np.load(path)
    if t["window"].shape != (spec.win_length,) or t["fb"].shape != (spec.n_freqs, spec.n_mels):
        return None
    return t["window"].astype(np.float64), t["fb"].astype(np.float64)


def hann_window(win_length: int) -> np.ndarray:
    """Periodic Hann, as ``torch.hann_window(win_length)`` (torchaudio Spectrogram default).

    Returned as the float32-valued table the reference stores
    (state_dict key ``model.mel_spec.spectrogram.window``), widened to float64.
    """
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    return w.astype(np.float32).astype(np.float64)


def _hz_to_mel_htk(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(spec: FrontendSpec) -> np.ndarray:
    """HTK triangular filterbank (n_freqs, n_mels), norm=None.

    Follows torchaudio.functional.melscale_fbanks as called by the reference's
    MelSpectrogram (architectures.py:830-836): bin centres linspace(0, sr//2, n_freqs),
    mel points linspace(mel(f_min), mel(f_max), n_mels+2), fb = max(0, min(down, up)).
    torchaudio evaluates this in float32; we evaluate in float64 and round the table to
    float32 (the difference, <=2e-7 absolute, is checked against the torchaudio table in
    tests/test_oracle_golden.py).
    """
    all_freqs = np.linspace(0.0, float(spec.sample_rate // 2), spec.n_freqs)
    m_pts = np.linspace(_hz_to_mel_htk(spec.f_min), _hz_to_mel_htk(spec.f_max), spec.n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]            # (n_freqs, n_mels+2)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    return fb.astype(np.float32).astype(np.float64)


def frame_signal(x: np.ndarray, spec: FrontendSpec) -> np.ndarray:
    """(B, N) float -> (B, T, win_length) frames, after the centring pad if any."""
    if spec.center:
        p = spec.n_fft // 2
        x = np.pad(x, ((0, 0), (p, p)), mode="reflect")
        # torch.stft with win_length == n_fft frames the padded signal directly.
        assert spec.win_length == spec.n_fft, "centred geometry assumes win_length == n_fft"
    n = x.shape[1]
    t = 1 + (n - spec.win_length) // spec.hop_length
    idx = (np.arange(t)[:, None] * spec.hop_length) + np.arange(spec.win_length)[None, :]
    return x[:, idx]


def _tables(spec: FrontendSpec, window=None, fb=None):
    """The reference's own float32 tables when the geometry is pinned, else computed."""
    ref = reference_tables(spec)
    if window is None:
        window = ref[0] if ref is not None else hann_window(spec.win_length)
    if fb is None:
        fb = ref[1] if ref is not None else mel_filterbank(spec)
    return np.asarray(window, dtype=np.float64), np.asarray(fb, dtype=np.float64)


def power_spectrum(pcm: np.ndarray, spec: FrontendSpec, dtype=np.float64, window=None) -> np.ndarray:
    """int16 (B, N) -> power (B, T, n_freqs)."""
    pcm = np.asarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    if pcm.dtype == np.int16:
        x = pcm.astype(dtype) / dtype(32768.0)
    else:  # already float in [-1, 1) as the reference feeds ORT
        x = pcm.astype(dtype)
    frames = frame_signal(x, spec) * _tables(spec, window)[0].astype(dtype)
    # A frame shorter than n_fft is zero-extended on the right (TF/Kaldi framing); torch
    # centres the window inside n_fft instead, which only changes the phase, not |X|^2.
    spec_c = np.fft.rfft(frames, n=spec.n_fft, axis=-1)
    re = spec_c.real.astype(dtype)
    im = spec_c.imag.astype(dtype)
    return re * re + im * im


def log_mel(pcm: np.ndarray, spec: FrontendSpec, dtype=np.float64, window=None, fb=None) -> np.ndarray:
    """int16 (B, N) -> log-mel dB (B, n_mels, T), the layout the reference's CNN sees."""
    p = power_spectrum(pcm, spec, dtype, window)
    fb = _tables(spec, window, fb)[1].astype(dtype)
    mel = p @ fb                                         # (B, T, n_mels)
    db = dtype(10.0) * np.log10(np.maximum(mel, dtype(spec.amin)))
    return np.ascontiguousarray(np.swapaxes(db, 1, 2)).astype(dtype)
