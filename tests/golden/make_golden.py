"""Generate golden vectors by running the REFERENCE's own modules (build container only).

Usage (from the repo root, in the container that has /root/reference):
    python tests/golden/make_golden.py                 # everything
    python tests/golden/make_golden.py gru lstm rnn    # only the named heads (front-end file untouched)

Imports ``nanowakeword.modules.model.Model`` and ``nanowakeword._export.onnx`` from
/root/reference (read-only; ``torchinfo``/``matplotlib`` are stubbed because they are
only used for summaries and plots, reference modules/model.py:30-31, 428-581) plus
torchaudio, feeds them the seeded weights of ``nanowakeword_b200.synth`` and a small
fixed set of PCM windows, and stores inputs + outputs as compressed ``.npz`` under
tests/golden/.  The GPU box has no /root/reference: tests only read the ``.npz`` files.
"""
import os
import sys
import types
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    sys.path.insert(0, REF)
    for m in ("torchinfo", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["torchinfo"].summary = lambda *a, **k: None
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from nanowakeword.modules.model import Model
    from nanowakeword._export.onnx import ONNXSafeMelSpectrogram, make_onnx_safe_adaptive_pool
    return Model, ONNXSafeMelSpectrogram, make_onnx_safe_adaptive_pool


def read_wav(path):
    with wave.open(path, "rb") as f:
        assert f.getframerate() == 16000 and f.getsampwidth() == 2 and f.getnchannels() == 1
        return np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16)


def golden_pcm():
    """10 windows of 16000 int16 samples: 4 reference example WAVs (last second, as
    predict_clip scores it — nanointerpreter.py:756, 828-830), 2 full-scale uniform,
    2 speech-like Gaussian, all-zero, +-1 LSB."""
    from nanowakeword_b200.synth import synth_pcm
    wavs = ["positive/example_wakeWord.wav", "negative/jast_example.wav",
            "noise/noise-free-sound-0003.wav", "rir/Echo(rir)_Download_from_anywhere.wav"]
    rows, names = [], []
    for w in wavs:
        x = read_wav(os.path.join(REF, "examples/training_data", w))
        x = x[-16000:] if len(x) >= 16000 else np.pad(x, (16000 - len(x), 0))
        rows.append(x)
        names.append("wav:" + w)
    u = synth_pcm(2, seed=0, kind="uniform")
    g = synth_pcm(2, seed=0, kind="gauss")
    rows += [u[0], u[1], g[0], g[1], np.zeros(16000, np.int16)]
    names += ["uniform0", "uniform1", "gauss0", "gauss1", "zeros"]
    lsb = np.where(np.random.default_rng(7).random(16000) < 0.5, -1, 1).astype(np.int16)
    rows.append(lsb)
    names.append("lsb")
    return np.stack(rows).astype(np.int16), names


def main():
    import torch
    import torchaudio
    torch.set_num_threads(4)
    Model, SafeMel, safe_pool = import_reference()
    from nanowakeword_b200.synth import default_config, make_state_dict
    from oracle.frontend import GEOMETRIES

    pcm, names = golden_pcm()
    x32 = torch.from_numpy(pcm.astype(np.float32) / 32768.0)
    x64 = x32.double()
    out = {"pcm": pcm, "names": np.array(names)}

    # ---- front ends -----------------------------------------------------------------
    mels = {}
    for gname, g in GEOMETRIES.items():
        def build():
            return torchaudio.transforms.MelSpectrogram(
                sample_rate=g.sample_rate, n_fft=g.n_fft, win_length=g.win_length,
                hop_length=g.hop_length, n_mels=g.n_mels, center=g.center)
        todb = torchaudio.transforms.AmplitudeToDB()
        pre = (lambda t: t) if g.center else (
            lambda t: torch.nn.functional.pad(t, ((g.n_fft - g.win_length) // 2,) * 2))
        m32 = build()
        m64 = build().double()
        with torch.no_grad():
            mel32 = todb(m32(pre(x32)))
            mel64 = todb(m64(pre(x64)))
            safe = todb(SafeMel(build())(pre(x32)))
        assert mel64.shape[1:] == (g.n_mels, g.n_frames), mel64.shape
        out[f"{gname}.mel_ta64"] = mel64.numpy()
        out[f"{gname}.mel_ta32"] = mel32.numpy()
        out[f"{gname}.mel_convdft32"] = safe.numpy()
        out[f"{gname}.fb"] = m32.mel_scale.fb.numpy()
        out[f"{gname}.window"] = m32.spectrogram.window.numpy()
        mels[gname] = mel64
    only = sys.argv[1:]
    if not only:
        np.savez_compressed(os.path.join(HERE, "frontend.npz"), **out)
        print("frontend.npz written")

    # ---- heads (fed the float64 torchaudio mel so head errors are isolated) -----------
    for mt in ("dnn", "cnn", "tcn", "bcresnet", "crnn", "e2e_dnn", "gru", "lstm", "rnn", "quartznet", "e2e_quartznet", "e2e_cnn"):
        if only and mt not in only:
            continue
        cfg = default_config(mt)
        sd_np = make_state_dict(cfg, seed=0)
        model = Model(cfg, "golden", input_shape=tuple(cfg["input_shape"]), model_type=mt,
                      layer_dim=cfg["layer_dim"], n_blocks=cfg["n_blocks"], mode=cfg["mode"]).eval()
        sd_t = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
        if mt == "e2e_dnn":   # non-learned front-end buffers keep the module's own values
            for k in ("model.mel_spec.spectrogram.window", "model.mel_spec.mel_scale.fb"):
                sd_t[k] = model.state_dict()[k]
        model.load_state_dict(sd_t, strict=True)
        res = {}
        with torch.no_grad():
            if mt == "e2e_dnn":
                m64 = model.double()
                logits64 = m64(x64)
                res["logits64"] = logits64.numpy()
                m32 = model.float()
                logits32 = m32(x32)
                # the deployed graph: conv-DFT mel + AvgPool rewrite + sigmoid (onnx.py:157-172)
                from nanowakeword._export.onnx import replace_mel_spectrogram
                replace_mel_spectrogram(m32)

                class Wrap(torch.nn.Module):
                    def __init__(s, m):
                        super().__init__(); s.trained_model = m
                    def forward(s, x):
                        return torch.sigmoid(s.trained_model(x)).view(-1, 1, 1)
                wrapped = Wrap(m32).eval()
                safe_pool(wrapped, x32[:1].unsqueeze(1))
                res["scores32_deployed"] = wrapped(x32.unsqueeze(1)).numpy()
                res["logits32"] = logits32.numpy()
            elif mt in ("e2e_quartznet", "e2e_cnn"):   # raw audio in, no mel: the reference module is the whole graph
                res["logits64"] = model.double()(x64).numpy()
                res["emb64"] = model.double().model(x64).numpy()
                res["logits32"] = model.float()(x32).numpy()
            else:
                geom = "NS40x98"
                mel = mels[geom]
                feat = mel.transpose(1, 2).contiguous() if mt in ("dnn", "tcn", "gru", "lstm", "rnn", "quartznet") else mel
                logits64 = model.double()(feat)
                res["logits64"] = logits64.numpy()
                res["logits32"] = model.float()(feat.float()).numpy()
                # embedding before the classifier, for layer-wise triage
                res["emb64"] = model.double().model(feat).numpy()
            res["scores64"] = torch.sigmoid(torch.from_numpy(res["logits64"])).view(-1, 1, 1).numpy()
        np.savez_compressed(os.path.join(HERE, f"head_{mt}.npz"), **res)
        print(mt, "logits64", np.round(res["logits64"].ravel(), 3))


if __name__ == "__main__":
    main()
