"""CPU tests of the host layer: C-ABI symbol table, blob packing, interpreter bookkeeping
(against the oracle's restatement of the reference's streaming logic, with a fake session)."""
import ctypes
import os
import re

import numpy as np
import pytest

from nanowakeword_b200 import _lib
from nanowakeword_b200.interpreter import DetectionResult, NanoInterpreter
from nanowakeword_b200.synth import default_config, make_state_dict
from nanowakeword_b200.weights import fold_bn, pack_blob, pack_tensors
from oracle.interp import OracleInterpreter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "nww_b200.h")).read()
    declared = set(re.findall(r"\b(nww_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load_library()                      # raises if the .so is missing: no fallback
    for name in declared:
        assert getattr(lib, name) is not None
    assert ctypes.sizeof(_lib.NwwSpec) == 4 * 12 + 4 * 8


def test_create_fails_loudly_without_gpu_or_with_bad_blob():
    import torch
    lib = _lib.load_library()
    spec = _lib.NwwSpec()
    spec.struct_size = ctypes.sizeof(_lib.NwwSpec)
    h = ctypes.c_void_p()
    bad = b"not a blob at all"
    rc = lib.nww_create(ctypes.byref(spec), bad, len(bad), 0, ctypes.byref(h))
    assert rc != 0 and h.value is None
    msg = _lib.last_error(lib)
    if torch.cuda.is_available():
        assert "magic" in msg
    else:
        assert "no CUDA device" in msg and "no CPU fallback" in msg


@pytest.mark.parametrize("mt", ["dnn", "cnn", "tcn", "bcresnet", "crnn", "e2e_dnn", "gru", "lstm", "rnn", "quartznet", "e2e_quartznet", "e2e_cnn"])
def test_pack_tensors_layouts(mt):
    cfg = default_config(mt)
    sd = make_state_dict(cfg, 0)
    t = pack_tensors(sd, cfg)
    n = int(t["tail.n_layers"][0])
    assert t[f"tail.{n - 1}.W"].shape[0] == 1
    for i in range(1, n):
        assert t[f"tail.{i}.W"].shape[1] == t[f"tail.{i - 1}.W"].shape[0]
    blob = pack_blob(t)
    assert blob[:7] == b"NWWB200" and len(blob) % 4 == 0
    assert all(v.dtype in (np.float32, np.int32) for v in t.values())


def test_fold_bn_is_exact():
    rng = np.random.default_rng(0)
    sd = {"bn.weight": rng.uniform(0.5, 1.5, 4), "bn.bias": rng.normal(size=4),
          "bn.running_mean": rng.normal(size=4), "bn.running_var": rng.uniform(0.5, 2, 4)}
    w, b = rng.normal(size=(4, 3)), rng.normal(size=4)
    x = rng.normal(size=(5, 3))
    y = (x @ w.T + b - sd["bn.running_mean"]) / np.sqrt(sd["bn.running_var"] + 1e-5) * sd["bn.weight"] + sd["bn.bias"]
    w2, b2 = fold_bn(w, b, sd, "bn")
    assert np.allclose(x @ w2.T + b2, y, atol=1e-12)


class _FakeSession:
    """Session duck type whose 'model' is a cheap deterministic function of the window."""

    def __init__(self, clip=16000):
        self.clip = clip
        self.calls = 0

    def get_inputs(self):
        class A:
            name, shape = "input", ["batch_size", 16000]
        return [A()]

    @staticmethod
    def fn(pcm16):
        v = np.abs(pcm16.astype(np.float64)).mean(axis=-1) / 3000.0
        return (1.0 / (1.0 + np.exp(-(v - 1.0) * 4))).astype(np.float32)

    def run(self, _, feed):
        self.calls += 1
        x = np.asarray(feed["input"])
        assert x.dtype == np.int16 and x.shape == (1, self.clip)
        return [self.fn(x).reshape(-1, 1, 1)]


def _make_interp(**kw):
    return NanoInterpreter(["/nonexistent/wake.pt"], sessions={"wake": _FakeSession()}, **kw)


def _oracle():
    def score_fn(clip_f32):
        pcm = np.rint(clip_f32.astype(np.float64) * 32768).astype(np.int16)
        return float(_FakeSession.fn(pcm[None])[0])
    return OracleInterpreter(name="wake", score_fn=score_fn)


@pytest.mark.parametrize("chunk", [1280, 400, 16000, 20000])
def test_streaming_bookkeeping_matches_reference_logic(chunk):
    rng = np.random.default_rng(1)
    audio = np.clip(rng.normal(0, 3000, 16000 * 4), -32768, 32767).astype(np.int16)
    audio[20000:30000] //= 50
    it, orc = _make_interp(), _oracle()
    for i in range(0, len(audio), chunk):
        r = it.predict(audio[i:i + chunk])
        o = orc.predict(audio[i:i + chunk])
        assert isinstance(r, DetectionResult)
        assert r.score == pytest.approx(o["wake"], abs=1e-7)
        assert it.raw_scores["wake"] == pytest.approx(orc.raw_scores["wake"], abs=1e-7)
    assert it.score == pytest.approx(orc.post_processed_scores["wake"], abs=1e-7)


def test_patience_debounce_and_errors():
    rng = np.random.default_rng(2)
    audio = np.clip(rng.normal(0, 6000, 16000 * 3), -32768, 32767).astype(np.int16)
    for kw in ({"patience": {"wake": 3}, "threshold": {"wake": 0.5}},
               {"patience": {"wake": 1}, "threshold": {"wake": 0.5}},
               {"debounce_time": 0.5, "threshold": {"wake": 0.5}}):
        it, orc = _make_interp(), _oracle()
        for i in range(0, len(audio), 1280):
            r = it.predict(audio[i:i + 1280], **kw)
            o = orc.predict(audio[i:i + 1280], **kw)
            assert r.score == pytest.approx(o["wake"], abs=1e-7)
    it = _make_interp()
    for i in range(0, 16000 * 2, 1280):
        it.predict(audio[i:i + 1280])
    with pytest.raises(ValueError):
        it.predict(audio[:1280], patience={"wake": 2})
    with pytest.raises(ValueError):
        it.predict(audio[:1280], patience={"wake": 2}, threshold={"wake": 0.5}, debounce_time=1.0)
    with pytest.raises(ValueError):
        it.predict([1, 2, 3])
    with pytest.raises(NotImplementedError):
        _make_interp(vad_threshold=0.5)


def test_detection_result_and_properties():
    r = DetectionResult({"wake": 0.7, "gate": 0.4}, "wake", "gate", threshold=0.5)
    assert r.score == 0.7 and r.gate_score == 0.4 and r.detected and "wake" in r and r["gate"] == 0.4
    assert DetectionResult({"wake": 0.7}, "wake", None).detected is False
    assert r.get("missing", 1.5) == 1.5 and "detected=True" in repr(r)
    it = _make_interp()
    assert it.model_name == "wake" and not it.is_cascade and it.gate_name is None
    assert it.info["loaded_models"] == ["wake"] and it.detected(0.0) and not it.detected(0.1)
    it.reset()
    assert it.e2e_buffer_samples["wake"] == 0


def test_load_model_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        NanoInterpreter.load_model(str(tmp_path / "nope.onnx"))
    with pytest.raises(TypeError):
        NanoInterpreter.load_model(42)
    with pytest.raises(ValueError):
        NanoInterpreter.load_model("x.onnx", remote_pipeline="bogus")
    p = tmp_path / "m.onnx"
    p.write_bytes(b"\x08\x07")
    with pytest.raises(NotImplementedError):          # .onnx without its .pt/.json siblings
        NanoInterpreter.load_model(str(p))


def test_cascade_gating_order(tmp_path):
    class Gate(_FakeSession):
        def run(self, _, feed):
            return [np.full((1, 1, 1), 0.1, np.float32)]
    ver = _FakeSession()
    it = NanoInterpreter(["/x/wake_lite.pt", "/x/wake.pt"], sessions={"wake_lite": Gate(), "wake": ver})
    it.cascade_config = {"gate": "wake_lite", "verifier": "wake", "gate_threshold": 0.3}
    x = np.zeros(16000, np.int16)
    for _ in range(8):
        r = it.predict(x)
    assert ver.calls == 0 and r.score == 0.0 and it.model_name == "wake" and it.gate_name == "wake_lite"


class _FakeStreamEngine:
    """Host-only stand-in for Engine's stream calls: per-stream rings in numpy and a toy score
    function, so StreamBank's bookkeeping can be compared with the oracle interpreter on CPU."""

    def __init__(self, clip, score_fn):
        self.clip, self.score_fn = clip, score_fn

    def stream_open(self, n):
        self.n = n
        self.rings = np.zeros((n, self.clip), np.int16)
        self.count = np.zeros(n, np.int64)

    def stream_close(self):
        pass

    def stream_reset(self, ids=None):
        sel = slice(None) if ids is None else np.asarray(ids)
        self.rings[sel] = 0
        self.count[sel] = 0

    def stream_push_host(self, chunks):
        L = chunks.shape[1]
        self.rings = np.concatenate([self.rings, chunks], axis=1)[:, -self.clip:]
        self.count += L
        out = np.array([self.score_fn(r) for r in self.rings], np.float32)
        out[self.count < self.clip] = 0.0
        return out


@pytest.mark.parametrize("mode", ["plain", "patience", "patience1", "debounce"])
def test_stream_bank_matches_per_stream_oracle_interpreters(mode):
    from nanowakeword_b200.streams import StreamBank
    clip, n, L = 4000, 6, 500

    def score_i16(win):                      # toy "model": a smooth function of the window, in (0, 1)
        return float(1.0 / (1.0 + np.exp(-(np.abs(win.astype(np.float64)).mean() / 2000.0 - 1.5) * 3.0)))

    bank = StreamBank(_FakeStreamEngine(clip, score_i16), n)
    oracles = [OracleInterpreter(name="m", clip_samples=clip,
                                 score_fn=lambda c: score_i16(np.rint(c.astype(np.float64) * 32768.0).astype(np.int16)))
               for _ in range(n)]
    kw_bank = {"plain": {}, "patience": dict(patience=3, threshold=0.5), "patience1": dict(patience=1, threshold=0.5),
               "debounce": dict(debounce_time=0.1, threshold=0.5)}[mode]
    kw_or = {"plain": {}, "patience": dict(patience={"m": 3}, threshold={"m": 0.5}),
             "patience1": dict(patience={"m": 1}, threshold={"m": 0.5}),
             "debounce": dict(debounce_time=0.1, threshold={"m": 0.5})}[mode]
    rng = np.random.default_rng(3)
    for step in range(40):
        amp = rng.uniform(500, 9000, size=(n, 1))
        chunks = np.clip(rng.normal(0, 1, (n, L)) * amp, -32768, 32767).astype(np.int16)
        if step == 17:
            bank.reset([1, 4])
            oracles[1].reset()
            oracles[4].reset()
        if step == 29:
            bank.reset()
            for o in oracles:
                o.reset()
        got = bank.push(chunks, **kw_bank)
        for i, o in enumerate(oracles):
            want = o.predict(chunks[i], **kw_or)["m"]
            assert abs(got[i] - want) < 1e-6, (mode, step, i)
            assert abs(bank.raw_scores[i] - o.raw_scores["m"]) < 1e-6
    with pytest.raises(ValueError):
        bank.push(chunks, patience=2)                       # threshold missing
    with pytest.raises(ValueError):
        bank.push(chunks, patience=2, debounce_time=0.5, threshold=0.5)


def test_multi_layer_recurrent_heads_are_refused():
    """The engine builds single-layer GRU / LSTM / RNN heads (the reference default, n_blocks = 1); deeper stacks must
    fail loudly at pack time instead of being scored wrongly."""
    for mt in ("gru", "lstm", "rnn"):
        cfg = default_config(mt, n_blocks=2)
        sd = make_state_dict(cfg, 0)
        with pytest.raises(ValueError, match="single-layer"):
            pack_tensors(sd, cfg)


def test_recurrent_gate_matrix_layout():
    """weights.py packs each direction as one [x | 1 | pad | h] x [4H gate columns] matrix (csrc/nww_rnn.cuh): evaluate
    that matrix in numpy and compare with the oracle's GRU / LSTM."""
    from oracle.heads import _cast_sd, rnn_last_output_bidir
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    x = np.random.default_rng(3).normal(-20, 30, (4, 98, 40))
    for mt, prefix, kind in (("gru", "model.gru", "gru"), ("lstm", "model.lstm", "lstm"), ("rnn", "model.layer1", "lstm")):
        cfg = default_config(mt)
        sd = make_state_dict(cfg, 0)
        t = pack_tensors(sd, cfg)
        wf, wb = t["rnn.fwd.w"].astype(np.float64), t["rnn.bwd.w"].astype(np.float64)
        hid, kx = wf.shape[1] // 4, wb.shape[0]
        assert wf.shape[0] == kx + hid and kx % 16 == 0 and kx > 40

        def step(xt, h, c, w):
            a = np.zeros((xt.shape[0], w.shape[0]))
            a[:, :40], a[:, 40] = xt, 1.0
            if w.shape[0] > kx:
                a[:, kx:] = h
            g = a @ w
            g0, g1, g2, g3 = g[:, :hid], g[:, hid:2 * hid], g[:, 2 * hid:3 * hid], g[:, 3 * hid:]
            if kind == "lstm":
                c = sig(g1) * c + sig(g0) * np.tanh(g2)
                return sig(g3) * np.tanh(c), c
            r, z = sig(g0), sig(g1)
            hn = (1 - z) * np.tanh(g2 + r * g3) + z * h
            return hn, hn
        h = c = np.zeros((4, hid))
        for s in range(98):
            h, c = step(x[:, s], h, c, wf)
        hb, _ = step(x[:, 97], np.zeros((4, hid)), np.zeros((4, hid)), wb)
        ref = rnn_last_output_bidir(x, _cast_sd(sd, np.float64), prefix, kind)
        assert np.abs(np.concatenate([h, hb], 1) - ref).max() < 1e-6


def test_cascade_bank_matches_cascade_interpreters():
    """CascadeBank (many streams, gate -> verifier) against one cascade NanoInterpreter per stream: the verifier's
    score, raw score and history must follow nanointerpreter.py:758-769 — skipped (0.0) whenever the gate's
    warm-up-zeroed score of the same call is below the gate threshold."""
    from nanowakeword_b200.streams import CascadeBank

    class Gate(_FakeSession):
        @staticmethod
        def fn(pcm16):
            v = np.abs(pcm16.astype(np.float64)).mean(axis=-1) / 2500.0
            return (1.0 / (1.0 + np.exp(-(v - 1.2) * 3))).astype(np.float32)

    n, L, thr = 4, 4000, 0.45
    gate_eng = _FakeStreamEngine(16000, lambda w: float(Gate.fn(w[None, :])[0]))
    ver_eng = _FakeStreamEngine(16000, lambda w: float(_FakeSession.fn(w[None, :])[0]))
    bank = CascadeBank(gate_eng, ver_eng, n, gate_threshold=thr)
    interps = []
    for _ in range(n):
        it = NanoInterpreter(["/x/wake_lite.pt", "/x/wake.pt"], sessions={"wake_lite": Gate(), "wake": _FakeSession()})
        it.cascade_config = {"gate": "wake_lite", "verifier": "wake", "gate_threshold": thr}
        interps.append(it)
    rng = np.random.default_rng(5)
    skipped = passed = 0
    for step in range(30):
        amp = rng.uniform(1000, 6000, size=(n, 1))
        chunks = np.clip(rng.normal(0, 1, (n, L)) * amp, -32768, 32767).astype(np.int16)
        if step == 14:
            bank.reset([2])
            interps[2].reset()
        got = bank.push(chunks, patience=2, threshold=0.3)
        for i, it in enumerate(interps):
            r = it.predict(chunks[i], patience={"wake": 2}, threshold={"wake": 0.3})
            assert abs(got[i] - r.score) < 1e-6, (step, i, got[i], r.score)
            assert abs(bank.gate_scores[i] - r.gate_score) < 1e-6
            assert abs(bank.raw_scores[i] - it.raw_scores["wake"]) < 1e-6
            skipped += r.gate_score < thr
            passed += r.gate_score >= thr
    assert skipped > 10 and passed > 10          # both branches exercised
