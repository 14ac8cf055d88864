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
    with pytest.raises(ValueError, match="no GraphProto"):      # an .onnx path is parsed (round 2): this one has no graph
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

    scored = 0                        # streams actually scored (the selective push skips the others)

    def stream_push_host(self, chunks, select=None):
        L = chunks.shape[1]
        self.rings = np.concatenate([self.rings, chunks], axis=1)[:, -self.clip:]
        self.count += L
        ids = range(len(self.rings)) if select is None else [int(i) for i in select]
        out = np.zeros(len(self.rings), np.float32)
        for i in ids:
            out[i] = self.score_fn(self.rings[i])
        self.scored += len(ids)
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
    # the verifier engine scored only the streams whose gate fired (nww_stream_push_select_host), the gate all of them
    assert gate_eng.scored == 30 * n and ver_eng.scored == passed


# ---------------------------------------------------------------------------------------------- round 2: boundary
class _FloatAwareSession(_FakeSession):
    """Accepts what the interpreter hands over: int16, or float32 once the ring went float."""

    def run(self, _, feed):
        self.calls += 1
        x = np.asarray(feed["input"])
        self.last_dtype = x.dtype
        pcm = x if x.dtype == np.int16 else x.astype(np.float64) * 32768.0
        return [self.fn(pcm).reshape(-1, 1, 1)]


def test_pcm_ring_empty_chunk_and_float_switch():
    """deque.extend([]) is a no-op; non-int16 input follows nanointerpreter.py:750 verbatim."""
    it = NanoInterpreter(["/nonexistent/wake.pt"], sessions={"wake": _FloatAwareSession()})
    r = it.predict(np.zeros(0, np.int16))
    assert r.score == 0.0 and it.e2e_buffer_samples["wake"] == 0
    rng = np.random.default_rng(3)
    a = np.clip(rng.normal(0, 3000, 16000), -32768, 32767).astype(np.int16)
    it.predict(a)
    assert it.e2e_buffer["wake"].data.dtype == np.int16 and it.models["wake"].last_dtype == np.int16
    it.predict(a[:1280].astype(np.float64))                      # integers in a float array stay on the int16 path
    assert it.e2e_buffer["wake"].data.dtype == np.int16
    off = a[:1280].astype(np.float32) + 0.25                     # off the int16 grid: the ring goes float for good
    it.predict(off)
    ring = it.e2e_buffer["wake"].data
    assert ring.dtype == np.float32 and it.models["wake"].last_dtype == np.float32
    assert np.array_equal(ring[-1280:], off / np.float32(32768.0))
    assert np.array_equal(ring[-2560:-1280], a[:1280].astype(np.float32) / np.float32(32768.0))
    it.predict(a[:640])                                          # int16 chunks are scaled into the float ring
    assert np.array_equal(it.e2e_buffer["wake"].data[-640:], a[:640].astype(np.float32) / np.float32(32768.0))
    it.reset()
    assert it.e2e_buffer["wake"].data.dtype == np.int16 and len(it.e2e_buffer["wake"]) == 0


def test_predict_clip_wav_and_array(tmp_path):
    """predict_clip (nanointerpreter.py:816-833): WAV path or ndarray, ONE predict() on the whole clip in e2e mode."""
    import wave
    rng = np.random.default_rng(5)
    audio = np.clip(rng.normal(0, 4000, 19375), -32768, 32767).astype(np.int16)
    path = str(tmp_path / "clip.wav")
    with wave.open(path, "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000)
        f.writeframes(audio.tobytes())
    it, orc = _make_interp(), _oracle()
    out = it.predict_clip(path)
    assert isinstance(out, list) and len(out) == 1 and isinstance(out[0], DetectionResult)
    assert it.models["wake"].calls == 1
    assert it.raw_scores["wake"] == pytest.approx(orc.predict(audio) and orc.raw_scores["wake"], abs=1e-7)
    it2 = _make_interp()
    out2 = it2.predict_clip(audio)
    assert it2.raw_scores["wake"] == it.raw_scores["wake"] and out2[0].score == out[0].score
    bad = str(tmp_path / "bad.wav")
    with wave.open(bad, "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(8000)
        f.writeframes(audio.tobytes())
    with pytest.raises(ValueError, match="16kHz, 16-bit, single-channel"):
        it.predict_clip(bad)
    with pytest.raises(TypeError):
        it.predict_clip(12345)


class _FakePyAudio:
    """Stand-in for the pyaudio module: serves prepared int16 audio chunk by chunk, then raises KeyboardInterrupt
    (what Ctrl+C does to the reference's blocking loop, nanointerpreter.py:928-929)."""
    paInt16 = 8

    def __init__(self, audio, raise_at_end=True):
        self.audio, self.pos, self.raise_at_end = audio, 0, raise_at_end
        self.opened = self.closed = self.terminated = 0
        self.open_kwargs = None

    def PyAudio(self):
        return self

    def open(self, **kw):
        self.opened += 1
        self.open_kwargs = kw
        return self

    def read(self, n, exception_on_overflow=True):
        assert exception_on_overflow is False
        if self.pos + n > len(self.audio):
            if self.raise_at_end:
                raise KeyboardInterrupt
            self.pos = 0
        out = self.audio[self.pos:self.pos + n]
        self.pos += n
        return out.tobytes()

    def stop_stream(self):
        pass

    def close(self):
        self.closed += 1

    def terminate(self):
        self.terminated += 1


def test_listen_signature_and_callback_contract(monkeypatch):
    """listen() as the reference defines it (nanointerpreter.py:835-945): same parameters, on_audio -> predict ->
    on_score(verifier, gate) every chunk, on_detection(name, score) above threshold outside the cooldown, reset()
    after a detection."""
    import inspect
    import sys
    sig = inspect.signature(NanoInterpreter.listen)
    assert list(sig.parameters) == ["self", "on_detection", "threshold", "cooldown", "chunk_size", "on_score", "on_audio", "blocking"]
    assert [sig.parameters[k].default for k in ("threshold", "cooldown", "chunk_size", "blocking")] == [0.5, 1.0, 1280, True]

    rng = np.random.default_rng(9)
    loud = np.clip(rng.normal(0, 9000, 1280 * 40), -32768, 32767).astype(np.int16)     # fake model fires on loud audio
    fake = _FakePyAudio(loud)
    monkeypatch.setitem(sys.modules, "pyaudio", fake)
    it = _make_interp()
    resets, audio_chunks, scores, dets = [], [], [], []
    orig_reset = it.reset
    it.reset = lambda: (resets.append(len(scores)), orig_reset())[1]
    it.listen(on_detection=lambda name, s: dets.append((name, s, len(scores))), threshold=0.5, cooldown=0.0,
              on_score=lambda v, g: scores.append((v, g)), on_audio=lambda a: audio_chunks.append(a.copy()))
    assert fake.open_kwargs == dict(format=fake.paInt16, channels=1, rate=16000, input=True, frames_per_buffer=1280)
    assert fake.closed == 1 and fake.terminated == 1
    assert len(audio_chunks) == 40 and all(a.dtype == np.int16 and a.shape == (1280,) for a in audio_chunks)
    assert len(scores) == 40 and all(g == 0.0 for _, g in scores)
    # same chunks through an oracle interpreter that is reset after each detection -> same detections
    orc, expect = _oracle(), []
    for i in range(40):
        v = orc.predict(loud[i * 1280:(i + 1) * 1280])["wake"]
        assert scores[i][0] == pytest.approx(v, abs=1e-7)
        if v > 0.5:
            expect.append(i + 1)
            orc.reset()
    assert [d[2] for d in dets] == expect and len(expect) >= 1
    assert all(d[0] == "wake" and d[1] > 0.5 for d in dets)
    assert resets == expect                                          # reset() right after every detection
    # 13 chunks of 1280 fill the 16000-sample ring (the five warm-up zeros are spent on calls 1..5, which score 0
    # anyway): the first detection is the 13th call, and after every reset() it takes 13 calls again
    assert expect == [13, 26, 39]

    # cooldown: one detection only when the cooldown outlasts the audio
    fake2 = _FakePyAudio(loud)
    monkeypatch.setitem(sys.modules, "pyaudio", fake2)
    it2, dets2 = _make_interp(), []
    it2.listen(on_detection=lambda n, s: dets2.append(n), threshold=0.5, cooldown=3600.0)
    assert dets2 == ["wake"]

    # default on_detection prints; non-blocking mode runs on a daemon thread until stop()
    fake3 = _FakePyAudio(loud, raise_at_end=False)
    monkeypatch.setitem(sys.modules, "pyaudio", fake3)
    it3, seen = _make_interp(), []
    it3.listen(threshold=2.0, on_score=lambda v, g: seen.append(v), blocking=False)
    import time
    t0 = time.time()
    while len(seen) < 5 and time.time() - t0 < 10:
        time.sleep(0.01)
    assert it3._listen_thread is not None and it3._listen_thread.daemon
    it3.stop()
    assert it3._listen_thread is None and len(seen) >= 5 and fake3.terminated == 1


def test_listen_without_pyaudio_raises_import_error(monkeypatch):
    import sys
    monkeypatch.setitem(sys.modules, "pyaudio", None)
    with pytest.raises(ImportError, match="PyAudio is required for listen"):
        _make_interp().listen()


def test_crnn_packer_refuses_what_the_engine_does_not_build():
    """multi-layer or LSTM CRNN checkpoints must be refused, not packed as layer 0 / opaque size errors."""
    cfg = default_config("crnn")
    sd = make_state_dict(cfg, 0)
    deep = dict(sd)
    deep["model.rnn.weight_ih_l1"] = np.zeros((384, 256), np.float32)
    with pytest.raises(ValueError, match="single recurrent layer"):
        pack_tensors(deep, cfg)
    with pytest.raises(ValueError, match="crnn_rnn_type"):
        pack_tensors(sd, dict(cfg, crnn_rnn_type="lstm"))
    lstm = dict(sd)
    lstm["model.rnn.weight_hh_l0"] = np.zeros((512, 128), np.float32)
    with pytest.raises(ValueError, match=r"\(4H, H\)"):
        pack_tensors(lstm, cfg)


def test_package_exports_match_the_reference():
    """nanowakeword/__init__.py:1-5 exports NanoInterpreter, VAD, AudioFeatures: the names import; the two that are out
    of scope say so when constructed."""
    import nanowakeword_b200 as pkg
    for name in ("NanoInterpreter", "VAD", "AudioFeatures"):
        assert hasattr(pkg, name) and name in pkg.__all__
    for cls in (pkg.VAD, pkg.AudioFeatures):
        with pytest.raises(NotImplementedError, match="outside the B200 hot path"):
            cls()
