"""GPU parity, round 2: the float feed (no re-quantisation), cross-stream ordering, predict_clip /
predict_batch on the real engine.  Same tolerances as tests/test_gpu_parity.py."""
import numpy as np
import pytest

from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-3
MEL_TOL = 1e-4
RAW_HEADS = ("e2e_quartznet", "e2e_cnn")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _engine(mt, **kw):
    from nanowakeword_b200 import Engine
    cfg = default_config(mt)
    sd = make_state_dict(cfg, seed=0)
    return Engine(sd, cfg, device=0, **kw), sd, cfg


def _off_grid_audio(n, seed):
    """Quiet float audio that is NOT int16 / 32768: a few tens of LSB in amplitude scaled by an irrational-ish gain,
    so rounding it to the int16 grid changes the log-mel by far more than the tolerance; one sample at exactly
    +1.0 (which int16 cannot hold) and one beyond full scale."""
    q = (synth_pcm(n, seed=seed, kind="gauss") // 64).astype(np.float32) / np.float32(32768.0)
    x = (q * np.float32(0.7310586)).astype(np.float32)
    x[:, 5000] = 1.0
    x[:, 9000] = -1.25
    return x


@pytest.mark.parametrize("mt", ["cnn", "dnn", "tcn", "bcresnet", "crnn", "e2e_dnn", "gru", "quartznet", "e2e_quartznet", "e2e_cnn"])
def test_float_feed_is_not_requantised(torch_cuda, mt):
    """nww_run_windows_f32 (through Engine.score_device_f32): the reference's session takes float32 audio
    (nanointerpreter.py:750, 771-783).  On the int16 grid the float path must agree with the int16 path; off
    the grid it must follow the oracle fed the SAME floats, not their int16 rounding."""
    from oracle.heads import forward_scores
    eng, sd, cfg = _engine(mt)
    geom = "REF64x101" if mt == "e2e_dnn" else "NS40x98"
    # (1) on the grid
    pcm = np.concatenate([synth_pcm(20, seed=61, kind="gauss"), synth_pcm(12, seed=62, kind="uniform")])
    xf = pcm.astype(np.float32) / np.float32(32768.0)
    s_i16 = eng.score_device(torch_cuda.from_numpy(pcm).cuda()).cpu().numpy()
    s_f32 = eng.score_device_f32(torch_cuda.from_numpy(xf).cuda()).cpu().numpy()
    ref = forward_scores(pcm, sd, cfg).ravel()
    assert np.abs(s_f32 - ref).max() < SCORE_TOL
    assert np.abs(s_f32 - s_i16).max() < 1e-4
    # (2) off the grid
    x = _off_grid_audio(24, seed=63)
    ref_f, mel_f = forward_scores(x, sd, cfg, return_mel=True)
    if mt in RAW_HEADS:
        got = eng.score_device_f32(torch_cuda.from_numpy(x).cuda())
    else:
        got, extra = eng.score_device_f32(torch_cuda.from_numpy(x).cuda(), want_mel=True)
        mel = extra["mel"].cpu().numpy()
        assert np.abs(mel - mel_f).max() < MEL_TOL
        # and the int16 rounding of the same audio is a different signal at this tolerance
        from oracle.frontend import GEOMETRIES, log_mel
        xq = np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
        assert np.abs(log_mel(xq, GEOMETRIES[geom]) - mel_f).max() > 100 * MEL_TOL
    assert np.abs(got.cpu().numpy() - ref_f.ravel()).max() < SCORE_TOL


def test_session_run_float_feeds(torch_cuda):
    """B200Session.run: float32 (B, N) / (B, 1, N) on the grid takes the int16 path (bit-identical to int16 input),
    off the grid takes the float path; '+1.0' is not clipped."""
    from nanowakeword_b200 import B200Session
    from oracle.heads import forward_scores
    cfg = default_config("cnn")
    sd = make_state_dict(cfg, seed=0)
    sess = B200Session(state_dict=sd, cfg=cfg, device=0)
    pcm = synth_pcm(9, seed=71, kind="gauss")
    a = sess.run(None, {"input": pcm})[0]
    b = sess.run(None, {"input": pcm.astype(np.float32) / 32768.0})[0]
    c = sess.run(None, {"input": (pcm.astype(np.float32) / 32768.0)[:, None, :]})[0]
    assert a.shape == (9, 1, 1) and np.array_equal(a, b) and np.array_equal(a, c)
    x = _off_grid_audio(7, seed=72)
    d = sess.run(None, {"audio": x})[0]
    assert d.shape == (7, 1, 1) and d.dtype == np.float32
    assert np.abs(d.ravel() - forward_scores(x, sd, cfg).ravel()).max() < SCORE_TOL
    with pytest.raises(ValueError):
        sess.run(None, {"input": x[:, :100]})
    with pytest.raises(ValueError):
        sess.run(None, {"wrong": x})


def test_calls_on_different_cuda_streams_are_ordered(torch_cuda):
    """the engine's workspaces are shared between calls; calls enqueued on different CUDA streams (and the
    host-path calls on the engine's private stream) must behave as if issued on one stream."""
    torch = torch_cuda
    eng, sd, cfg = _engine("cnn")
    pcm_a = torch.from_numpy(synth_pcm(1500, seed=81, kind="gauss")).cuda()
    pcm_b = torch.from_numpy(synth_pcm(1500, seed=82, kind="uniform")).cuda()
    ref_a = eng.score_device(pcm_a).cpu().numpy()
    ref_b = eng.score_device(pcm_b).cpu().numpy()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        out_a = eng.score_device(pcm_a, stream=s1)
        out_b = eng.score_device(pcm_b, stream=s2)             # no synchronisation in between
        host_b = eng.score_host(pcm_b.cpu().numpy()[:700])      # private stream, right behind
        out_a2 = eng.score_device(pcm_a, stream=s2)
        torch.cuda.synchronize()
        assert np.array_equal(out_a.cpu().numpy(), ref_a)
        assert np.array_equal(out_b.cpu().numpy(), ref_b)
        assert np.array_equal(host_b, ref_b[:700])
        assert np.array_equal(out_a2.cpu().numpy(), ref_a)
    # stream rings: pushes on a caller stream, reset on the engine's stream, pushes again
    n = 64
    eng.stream_open(n)
    chunks = torch.from_numpy(synth_pcm(n, seed=83, kind="gauss")).cuda()
    for rep in range(2):
        pieces = [chunks[:, s:s + 1600].contiguous() for s in range(0, 16000, 1600)]
        torch.cuda.synchronize()                                # the slices were made on torch's current stream
        for i, piece in enumerate(pieces):
            sc = eng.stream_push_device(piece, stream=s1 if i & 1 else s2)
        torch.cuda.synchronize()
        assert np.array_equal(sc.cpu().numpy(), eng.score_device(chunks).cpu().numpy())
        eng.stream_reset()                                      # no sync needed before the next pushes
    eng.stream_close()
    with pytest.raises(ValueError):
        eng.score_device(pcm_a, out=torch.empty(10, device="cuda"))
    with pytest.raises(ValueError):
        eng.logmel_device(pcm_a[:, :100].contiguous())


def test_predict_clip_and_predict_batch_on_the_engine(torch_cuda, tmp_path, golden_frontend):
    """predict_clip (nanointerpreter.py:816-833) on the reference's example recordings (the first four golden
    windows are its example WAVs cropped / padded to one second), as WAV files and as arrays; predict_batch against
    the oracle."""
    import wave
    from nanowakeword_b200 import NanoInterpreter, save_model
    from oracle.heads import forward_scores
    from oracle.interp import OracleInterpreter
    cfg = default_config("cnn")
    sd = make_state_dict(cfg, seed=0)
    interp = NanoInterpreter.load_model(save_model(str(tmp_path / "wake.pt"), sd, cfg))
    g = golden_frontend["pcm"]
    ref = forward_scores(g, sd, cfg).ravel()
    for i in range(4):
        clip = np.concatenate([g[i], g[(i + 1) % 4][:3375]])           # 19 375 samples, like the positive example
        path = str(tmp_path / f"ex{i}.wav")
        with wave.open(path, "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000)
            f.writeframes(clip.tobytes())
        orc = OracleInterpreter(sd, cfg, name="wake")
        want = orc.predict(clip)["wake"]
        for arg in (path, clip):
            interp.reset()
            out = interp.predict_clip(arg)
            assert len(out) == 1 and abs(out[0].score - want) < SCORE_TOL
            assert abs(interp.raw_scores["wake"] - orc.raw_scores["wake"]) < SCORE_TOL
            assert interp.raw_scores["wake"] > 0.0 and out[0].score == 0.0   # scored, but still inside the warm-up
    got = interp.predict_batch(g)
    assert got.shape == (len(g),) and np.abs(got - ref).max() < SCORE_TOL
    # a float chunk that is off the int16 grid goes through the float ring and the float path
    interp.reset()
    x = _off_grid_audio(1, seed=91)[0] * np.float32(32768.0)            # predict() scales by 1/32768 itself
    orc = OracleInterpreter(sd, cfg, name="wake")
    interp.predict(x)
    orc.predict(x)
    assert interp.e2e_buffer["wake"].data.dtype == np.float32
    assert abs(interp.raw_scores["wake"] - orc.raw_scores["wake"]) < SCORE_TOL


# ------------------------------------------------------------------------ BASELINE configs #3, #4, #5 at their stated sizes
def _device_pcm(torch, n, seed):
    """Full-scale uniform int16 windows generated on the device (n x 32 kB would take minutes through numpy)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    out = torch.empty((n, 16000), dtype=torch.int16, device="cuda")
    for o in range(0, n, 16384):
        out[o:o + 16384] = torch.randint(-32768, 32768, (min(16384, n - o), 16000), generator=g, device="cuda",
                                         dtype=torch.int32).to(torch.int16)
    return out


@pytest.mark.parametrize("mt,n", [("bcresnet", 65536), ("crnn", 131072)])
def test_configs_4_and_5_per_gpu_batch_sizes(torch_cuda, mt, n):
    """BASELINE configs #4 (262 144 windows / 4 GPUs, BcResNet) and #5 (1 048 576 / 8 GPUs, CRNN-GRU) at the number of
    windows ONE GPU owns: every launch chunk of the batch is scored; a strided sample is held to the oracle and the
    size-independent properties of a per-window map are checked on the full result."""
    from oracle.heads import forward_scores
    torch = torch_cuda
    eng, sd, cfg = _engine(mt)
    pcm = _device_pcm(torch, n, seed=100 + n % 97)
    # plant known windows across chunk boundaries: duplicates must score identically wherever they land
    chunk = eng.info["chunk_windows"]
    plants = [0, 1, chunk - 1, chunk, 3 * chunk + 7, n // 2, n - chunk - 1, n - 1]
    for p in plants[1:]:
        pcm[p] = pcm[0]
    s = eng.score_device(pcm)
    torch.cuda.synchronize()
    assert s.shape == (n,) and bool(torch.isfinite(s).all()) and float(s.min()) >= 0.0 and float(s.max()) <= 1.0
    sp = s[torch.tensor(plants, device="cuda")].cpu().numpy()
    assert np.all(sp == sp[0])
    idx = np.arange(5, n, n // 24)[:24]
    ref = forward_scores(pcm[torch.from_numpy(idx).cuda()].cpu().numpy(), sd, cfg).ravel()
    assert np.abs(s.cpu().numpy()[idx] - ref).max() < SCORE_TOL
    # the same windows in a small batch give the same bits (no dependence on batch size / chunking)
    small = eng.score_device(pcm[torch.from_numpy(idx).cuda()].contiguous()).cpu().numpy()
    assert np.array_equal(small, s.cpu().numpy()[idx])
    del pcm
    torch.cuda.empty_cache()


def test_config_3_65536_streams_tcn(torch_cuda):
    """BASELINE config #3: 65 536 streams x 1280-sample steps, TCN head.  Every stream is pushed through the ring /
    incremental-mel path; a strided sample of streams is checked step by step against its own oracle interpreter
    (nanointerpreter.py:735-814), and the whole bank against batch scoring of the ring contents."""
    from nanowakeword_b200 import StreamBank
    from oracle.interp import OracleInterpreter
    torch = torch_cuda
    eng, sd, cfg = _engine("tcn")
    n, L, steps = 65536, 1280, 16
    rng = np.random.default_rng(33)
    base = np.clip(rng.normal(0, 3000, (257, L * steps)), -32768, 32767).astype(np.int16)     # 257 distinct streams, tiled
    audio = np.tile(base, (n // 257 + 1, 1))[:n]
    sample = [0, 1, 256, 257, 4143, 4144, 30000, n - 1]
    oracles = {i: OracleInterpreter(sd, cfg, name="m") for i in sample}
    bank = StreamBank(eng, n)
    for s in range(steps):
        chunks = np.ascontiguousarray(audio[:, s * L:(s + 1) * L])
        got = bank.push(chunks)
        assert got.shape == (n,)
        for i in sample:
            want = oracles[i].predict(chunks[i])["m"]
            assert abs(got[i] - want) < SCORE_TOL, (s, i, got[i], want)
            assert abs(bank.raw_scores[i] - oracles[i].raw_scores["m"]) < SCORE_TOL
        # streams fed the same audio give the same bits, wherever they sit in the bank
        assert np.array_equal(bank.raw_scores[:257], bank.raw_scores[257:514])
    assert (bank.raw_scores > 0).all()
    # ring contents == last 16000 samples: batch scoring of those windows agrees with the stream path
    tail = np.ascontiguousarray(audio[:4096, steps * L - 16000:steps * L])
    assert np.abs(eng.score_host(tail) - bank.raw_scores[:4096]).max() < 1e-5
    bank.close()


@pytest.mark.parametrize("mt,chunk_len", [("bcresnet", 1280), ("crnn", 1280), ("gru", 1280), ("lstm", 960), ("rnn", 1280),
                                          ("quartznet", 1280), ("e2e_cnn", 1600)])
def test_stream_rings_match_oracle_interpreters_other_heads(torch_cuda, golden_frontend, mt, chunk_len):
    """Stream-vs-reference-interpreter parity for the heads round 1 only checked incremental-vs-full against themselves."""
    from nanowakeword_b200 import StreamBank
    from oracle.interp import OracleInterpreter
    eng, sd, cfg = _engine(mt)
    n = 3
    g = golden_frontend["pcm"]
    audio = np.stack([np.concatenate([g[(i + k) % len(g)] for k in range(2)]) for i in range(n)])   # (n, 32000)
    bank = StreamBank(eng, n)
    oracles = [OracleInterpreter(sd, cfg, name="m") for _ in range(n)]
    n_steps = audio.shape[1] // chunk_len
    for s in range(n_steps):
        chunks = np.ascontiguousarray(audio[:, s * chunk_len:(s + 1) * chunk_len])
        if s == n_steps - 6:
            bank.reset([1])
            oracles[1].reset()
        got = bank.push(chunks)
        for i, o in enumerate(oracles):
            want = o.predict(chunks[i])["m"]
            assert abs(got[i] - want) < SCORE_TOL, (s, i, got[i], want)
            assert abs(bank.raw_scores[i] - o.raw_scores["m"]) < SCORE_TOL
    assert (bank.raw_scores[[0, 2]] > 0).all()
    bank.close()


def test_load_model_from_onnx_with_cascade(torch_cuda, tmp_path, golden_frontend):
    """The files a reference training run leaves behind (trainer.py:474-535: <name>.onnx, <name>_lite.onnx, no sidecar)
    load as a cascade (nanointerpreter.py:458-501) and stream like two oracle interpreters gated per call (:758-769)."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    from onnx_writer import write_e2e_model
    from nanowakeword_b200 import NanoInterpreter
    from oracle.interp import OracleInterpreter
    cfg_v, cfg_g = default_config("e2e_dnn"), default_config("e2e_quartznet")
    sd_v, sd_g = make_state_dict(cfg_v, seed=0), make_state_dict(cfg_g, seed=0)
    stem = str(tmp_path / "hey")
    write_e2e_model(stem + ".onnx", sd_v, cfg_v, style="folded")
    write_e2e_model(stem + "_lite.onnx", sd_g, cfg_g, style="explicit")
    # pick a gate threshold that actually splits the calls
    gate_thr = 0.9
    interp = NanoInterpreter.load_model(stem + ".onnx", cascade=True, gate_threshold=gate_thr)
    assert interp.is_cascade and interp.gate_name == "hey_lite" and interp.model_name == "hey"
    assert list(interp.models) == ["hey_lite", "hey"]
    assert interp.models["hey"].get_inputs()[0].shape == ["batch_size", 1, 16000] and interp.e2e_input_ndim["hey"] == 3
    og, ov = OracleInterpreter(sd_g, cfg_g, name="g"), OracleInterpreter(sd_v, cfg_v, name="v")
    g = golden_frontend["pcm"]
    stream = np.concatenate([g[0], g[4], g[1], g[5]])
    passed = skipped = 0
    for i in range(0, len(stream), 1280):
        chunk = stream[i:i + 1280]
        r = interp.predict(chunk)
        gs = og.predict(chunk)["g"]
        assert abs(r.gate_score - gs) < SCORE_TOL
        if og.buf_samples >= 16000 and gs < gate_thr:            # verifier skipped: 0.0, raw score 0.0, buffer still fed
            ov.buf.extend((chunk.astype(np.float32) / 32768.0).tolist())
            ov.buf_samples += len(chunk)
            ov.prediction_buffer.append(0.0)
            want, skipped = 0.0, skipped + 1
            assert interp.raw_scores["hey"] == 0.0
        else:
            want = ov.predict(chunk)["v"]
            passed += og.buf_samples >= 16000
        assert abs(r.score - want) < SCORE_TOL, (i, r.score, want)
    assert passed > 0 or skipped > 0


@pytest.mark.parametrize("mt,L", [("cnn", 1280), ("tcn", 1280), ("crnn", 1280), ("dnn", 1280), ("e2e_quartznet", 1280),
                                  ("cnn", 1000), ("tcn", 777), ("gru", 1280)])
def test_selective_stream_push_scores_only_the_listed_streams(torch_cuda, mt, L):
    """nww_stream_push_select[_host]: every stream ingests its chunk, only the listed ones are scored — and they get
    exactly the score a full push gives them (bit-identical), in any order of ids, across launch-chunk boundaries."""
    from nanowakeword_b200 import Engine
    cfg = default_config(mt)
    sd = make_state_dict(cfg, seed=0)
    full = Engine(sd, cfg, device=0, chunk_windows=37)
    sel = Engine(sd, cfg, device=0, chunk_windows=37)
    n = 101                                  # L = 1280: incremental mel ring; 1000 / 777: full windows out of the PCM ring
    full.stream_open(n)
    sel.stream_open(n)
    rng = np.random.default_rng(17)
    for step in range(17 if L >= 1000 else 25):
        chunks = np.clip(rng.normal(0, 3000, (n, L)), -32768, 32767).astype(np.int16)
        ids = rng.permutation(n)[:rng.integers(0, n + 1)] if step % 5 else np.arange(0)
        a = full.stream_push_host(chunks)
        if step % 2:
            b = sel.stream_push_host(chunks, select=ids)
        else:
            b = sel.stream_push_select_device(torch_cuda.from_numpy(chunks).cuda(),
                                              torch_cuda.from_numpy(ids.astype(np.int64)).cuda()).cpu().numpy()
        want = np.zeros(n, np.float32)
        want[ids] = a[ids]
        assert np.array_equal(b, want), (mt, step, np.abs(b - want).max())
    assert (a != 0).any()
    # a re-opened, larger bank re-allocates the selection staging
    sel.stream_open(4 * n)
    big = np.tile(chunks, (4, 1))
    got = sel.stream_push_host(big, select=np.arange(4 * n - 1, -1, -1))
    assert got.shape == (4 * n,) and np.all(got == 0.0)          # rings not full yet: every score is masked
    sel.stream_open(n)
    with pytest.raises(ValueError):
        sel.stream_push_host(chunks, select=[0, 0])
    with pytest.raises(ValueError):
        sel.stream_push_host(chunks, select=[n])
    full.stream_close()
    sel.stream_close()


def test_cascade_bank_on_engines_scales_with_the_gate(torch_cuda):
    """CascadeBank on two real engines: identical results with and without in-engine selection, and the verifier's
    kernel launches drop when the gate lets nothing through."""
    from nanowakeword_b200 import CascadeBank, Engine
    cfg_g, cfg_v = default_config("dnn"), default_config("cnn")
    sd_g, sd_v = make_state_dict(cfg_g, seed=0), make_state_dict(cfg_v, seed=0)
    n, L = 64, 1280
    rng = np.random.default_rng(23)
    audio = np.clip(rng.normal(0, 3000, (n, L * 20)), -32768, 32767).astype(np.int16)
    outs = {}
    for mode in (True, False):
        bank = CascadeBank(Engine(sd_g, cfg_g, device=0), Engine(sd_v, cfg_v, device=0), n, gate_threshold=0.2,
                           select_in_engine=mode)
        res = [bank.push(np.ascontiguousarray(audio[:, s * L:(s + 1) * L])).copy() for s in range(20)]
        outs[mode] = (np.stack(res), bank.raw_scores.copy(), bank.gate_scores.copy())
        bank.close()
    for a, b in zip(outs[True], outs[False]):
        assert np.array_equal(a, b)
    assert (outs[True][0] > 0).any() and (outs[True][0][-1] == 0).any()      # some streams pass the gate, some do not
    bank = CascadeBank(Engine(sd_g, cfg_g, device=0), Engine(sd_v, cfg_v, device=0), n, gate_threshold=2.0)   # nothing passes
    for s in range(15):
        bank.push(np.ascontiguousarray(audio[:, s * L:(s + 1) * L]))
    l0 = bank.verifier.engine.info["kernel_launches"]
    bank.push(np.ascontiguousarray(audio[:, 15 * L:16 * L]))
    assert bank.verifier.engine.info["kernel_launches"] - l0 == 1            # the fused ingest kernel only
    assert np.all(bank.raw_scores == 0.0)
    bank.close()


def test_tcn_fused_cone_launch_is_bit_identical_to_per_layer_launches(torch_cuda):
    """The cooperative single-launch form of the TCN's row-GEMM layers (grid barriers between layers, L2 loads, the
    stream-mode gather inside) against one launch per layer: same bits in batch mode (ragged chunks) and stream mode."""
    from nanowakeword_b200 import Engine
    cfg = default_config("tcn")
    sd = make_state_dict(cfg, seed=0)
    a = Engine(sd, cfg, device=0, tcn_layers="rows_fused")
    b = Engine(sd, cfg, device=0)
    pcm = np.concatenate([synth_pcm(700, seed=5, kind="gauss"), synth_pcm(337, seed=6, kind="uniform")])
    dev = torch_cuda.from_numpy(pcm).cuda()
    assert np.array_equal(a.score_device(dev).cpu().numpy(), b.score_device(dev).cpu().numpy())
    n, L = 333, 1280
    a.stream_open(n)
    b.stream_open(n)
    rng = np.random.default_rng(3)
    la = a.info["kernel_launches"]
    for step in range(15):
        chunks = np.clip(rng.normal(0, 3000, (n, L)), -32768, 32767).astype(np.int16)
        assert np.array_equal(a.stream_push_host(chunks), b.stream_push_host(chunks)), step
    assert (a.info["kernel_launches"] - la) == 15 * 4          # ingest, fused cone, dense tail, mask
    a.stream_close()
    b.stream_close()


@pytest.mark.parametrize("mt", ["e2e_dnn", "bcresnet"])
def test_fused_first_conv_is_bit_identical(torch_cuda, mt):
    """e2e_dnn: conv1 inside conv2's loader (conv3x3_umma_kernel<true>); bcresnet: front end + init conv in one stage kernel
    (bc_stage_kernel).  Both repeat the stand-alone kernels' arithmetic FMA for FMA, so scores and log-mel are identical."""
    torch = torch_cuda
    pcm = torch.from_numpy(synth_pcm(333, seed=21, kind="gauss")).cuda()
    res = []
    for fused in (True, False):
        eng, _, _ = _engine(mt, fused_first_conv=fused)
        s, extra = eng.score_device(pcm, want_mel=True)
        res.append((s.cpu().numpy().copy(), extra["mel"].cpu().numpy().copy()))
        eng.close()
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])


def test_piecewise_host_push_matches_single_piece(torch_cuda):
    """nww_stream_push_host cuts a large bank into pieces (copy of piece p + 1 behind the kernels of piece p); a stream's
    scores do not depend on how the bank was cut."""
    n = 9000                                                   # >= 2 x 4096: the piecewise path; not a multiple of the piece count
    rng = np.random.default_rng(5)
    one, _, _ = _engine("tcn", push_pieces=1)
    many, _, _ = _engine("tcn", push_pieces=3)
    one.stream_open(n)
    many.stream_open(n)
    for i in range(15):
        chunks = np.clip(rng.normal(0, 3000, (n, 1280)), -32768, 32767).astype(np.int16)
        a = one.stream_push_host(chunks)
        b = many.stream_push_host(chunks)
        assert np.array_equal(a, b), i
    assert a.max() > 0.0
    one.close()
    many.close()
