"""CPU oracle for the wake-word hot path — TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the arithmetic the reference
(arcosoph/nanowakeword v3.0.0) ships for the per-window path

    int16 PCM -> /32768 -> STFT -> |.|^2 -> mel -> 10*log10 -> head -> classifier -> sigmoid

It exists to *check* the CUDA engine.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``nanowakeword_b200/`` imports it, and the product path raises if the
CUDA library is missing instead of falling back here.

Parity pinning: the reference has no tests or golden vectors of its own,
so the oracle is pinned against outputs of the reference's own PyTorch modules
(`nanowakeword.modules.model.Model`, `torchaudio` transforms, `ONNXSafeMelSpectrogram`)
run in the build container by ``tests/golden/make_golden.py``; the resulting vectors are
committed under ``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.
The embedding-mode front end (downloaded melspectrogram.onnx / embedding_model.onnx,
`nanowakeword/interpreter/models/_registry.py:34-47`) is NOT restated: parity unpinned,
out of scope.
"""

from .frontend import FrontendSpec, GEOMETRIES, log_mel, hann_window, mel_filterbank  # noqa: F401
from .heads import forward_logits, forward_scores, head_input_from_mel  # noqa: F401
