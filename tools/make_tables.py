"""Dev tool (build container): dump the front-end constant tables exactly as the reference
holds them — the float32 Hann window and HTK mel filterbank that torchaudio's
MelSpectrogram computes at construction (reference modules/architectures.py:830-836; they
travel in the model as buffers ``mel_spec.spectrogram.window`` / ``mel_spec.mel_scale.fb``).
torchaudio evaluates them in float32, so they differ from a float64 evaluation by up to
8e-6 (fb) / 2.4e-7 (window) — enough to move the log-mel by ~1e-4 dB, hence shipped as data.
"""
import os, sys
import numpy as np
import torchaudio
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.frontend import GEOMETRIES
for name, g in GEOMETRIES.items():
    m = torchaudio.transforms.MelSpectrogram(sample_rate=g.sample_rate, n_fft=g.n_fft, win_length=g.win_length,
                                             hop_length=g.hop_length, n_mels=g.n_mels, center=g.center)
    np.savez_compressed(os.path.join(ROOT, "nanowakeword_b200", "tables", name + ".npz"),
                        window=m.spectrogram.window.numpy(), fb=m.mel_scale.fb.numpy())
    print(name, "written")
