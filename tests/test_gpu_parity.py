"""GPU parity: the CUDA engine (through the C ABI) against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star): final sigmoid score within 1e-3, log-mel within 1e-4 dB,
both measured against the float64 oracle, which is itself pinned to the reference's modules.
"""
import numpy as np
import pytest

from conftest import load_golden_head
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-3
MEL_TOL = 1e-4

HEADS = ["dnn", "cnn", "tcn", "bcresnet", "crnn", "e2e_dnn", "gru", "lstm", "rnn", "quartznet", "e2e_quartznet", "e2e_cnn"]
RAW_HEADS = ("e2e_quartznet", "e2e_cnn")          # audio in, no log-mel to compare


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


@pytest.mark.parametrize("geom,mt", [("NS40x98", "cnn"), ("REF64x101", "e2e_dnn")])
def test_logmel_matches_golden_and_oracle(torch_cuda, golden_frontend, geom, mt):
    from oracle.frontend import GEOMETRIES, log_mel
    eng, _, _ = _engine(mt)
    pcm = golden_frontend["pcm"]
    dev = torch_cuda.from_numpy(pcm).cuda()
    mel = eng.logmel_device(dev).cpu().numpy()
    ref64 = golden_frontend[f"{geom}.mel_ta64"]               # torchaudio float64 (the reference's transform)
    assert mel.shape == ref64.shape
    assert np.abs(mel - ref64).max() < MEL_TOL
    assert np.abs(mel - log_mel(pcm, GEOMETRIES[geom])).max() < MEL_TOL
    mel_t = eng.logmel_device(dev, time_major=True).cpu().numpy()
    assert np.array_equal(mel_t, np.swapaxes(mel, 1, 2))
    # digital silence hits the amin clamp exactly
    zeros = list(golden_frontend["names"]).index("zeros")
    assert np.all(mel[zeros] == -100.0)


@pytest.mark.parametrize("mt", HEADS)
def test_scores_match_golden(torch_cuda, golden_frontend, mt):
    eng, sd, cfg = _engine(mt)
    g = load_golden_head(mt)
    pcm = golden_frontend["pcm"]
    scores, extra = eng.score_device(torch_cuda.from_numpy(pcm).cuda(), want_logits=True, want_emb=True)
    scores = scores.cpu().numpy()
    assert np.abs(scores - g["scores64"].ravel()).max() < SCORE_TOL
    logit_scale = np.abs(sd["classifier.3.weight"]).sum() + 1.0
    assert np.abs(extra["logits"].cpu().numpy() - g["logits64"].ravel()).max() < 2e-4 * logit_scale
    if "emb64" in g:
        emb = extra["emb"].cpu().numpy()
        assert np.abs(emb - g["emb64"]).max() < 1e-4 * max(1.0, np.abs(g["emb64"]).max())


@pytest.mark.parametrize("mt", HEADS)
def test_scores_match_oracle_on_seeded_batch(torch_cuda, mt):
    from oracle.heads import forward_scores
    eng, sd, cfg = _engine(mt)
    pcm = np.concatenate([synth_pcm(40, seed=11, kind="uniform"), synth_pcm(40, seed=12, kind="gauss"),
                          (synth_pcm(20, seed=13, kind="gauss") // 64).astype(np.int16)])
    ref, mel_ref = forward_scores(pcm, sd, cfg, return_mel=True)
    if mt in RAW_HEADS:
        scores = eng.score_device(torch_cuda.from_numpy(pcm).cuda())
        with pytest.raises(ValueError):
            eng.score_device(torch_cuda.from_numpy(pcm).cuda(), want_mel=True)
    else:
        scores, extra = eng.score_device(torch_cuda.from_numpy(pcm).cuda(), want_mel=True)
        assert np.abs(extra["mel"].cpu().numpy() - mel_ref).max() < MEL_TOL
    assert np.abs(scores.cpu().numpy() - ref.ravel()).max() < SCORE_TOL
    # host (end-to-end) path gives the same numbers as the device path
    host = eng.score_host(pcm)
    assert np.array_equal(host, scores.cpu().numpy())


def test_ragged_and_empty_batches(torch_cuda):
    from oracle.heads import forward_scores
    eng, sd, cfg = _engine("cnn", chunk_windows=37)          # force several ragged chunks
    assert eng.score_host(np.zeros((0, 16000), np.int16)).shape == (0,)
    pcm = synth_pcm(101, seed=5, kind="gauss")
    ref = forward_scores(pcm, sd, cfg).ravel()
    got = eng.score_host(pcm)
    assert np.abs(got - ref).max() < SCORE_TOL
    one = eng.score_host(pcm[:1])
    assert np.abs(one - ref[:1]).max() < SCORE_TOL
    with pytest.raises(ValueError):
        eng.score_host(pcm[:, :15999])
    with pytest.raises(ValueError):
        eng.score_host(pcm.astype(np.float32))


def test_full_size_properties(torch_cuda):
    """BASELINE config #2 size (4096 windows, CNN): batch-order independence and agreement
    between a window scored alone and inside the batch (the size-independent properties of a
    per-window map)."""
    eng, sd, cfg = _engine("cnn")
    pcm = synth_pcm(4096, seed=1234, kind="uniform")
    dev = torch_cuda.from_numpy(pcm).cuda()
    s = eng.score_device(dev).cpu().numpy()
    assert np.isfinite(s).all() and (s >= 0).all() and (s <= 1).all()
    perm = np.random.default_rng(0).permutation(4096)
    s_perm = eng.score_device(torch_cuda.from_numpy(pcm[perm]).cuda()).cpu().numpy()
    assert np.array_equal(s_perm, s[perm])
    idx = [0, 1, 147, 148, 1183, 1184, 4095]
    s_sub = eng.score_device(torch_cuda.from_numpy(pcm[idx]).cuda()).cpu().numpy()
    assert np.array_equal(s_sub, s[idx])
    from oracle.heads import forward_scores
    ref = forward_scores(pcm[idx], sd, cfg).ravel()
    assert np.abs(s_sub - ref).max() < SCORE_TOL


def test_session_duck_type_and_interpreter(torch_cuda, tmp_path, golden_frontend):
    """The reference-facing path: save artefacts like the reference's exporter, load_model(),
    predict() in 1280-sample chunks, and compare with the oracle's streaming restatement."""
    from nanowakeword_b200 import NanoInterpreter, save_model
    from oracle.interp import OracleInterpreter
    cfg = default_config("cnn")
    sd = make_state_dict(cfg, seed=0)
    path = save_model(str(tmp_path / "hey_b200.pt"), sd, cfg)
    interp = NanoInterpreter.load_model(path)
    sess = interp.models["hey_b200"]
    assert sess.get_providers() == ["B200ExecutionProvider"]
    assert sess.get_inputs()[0].name == "input" and sess.get_inputs()[0].shape[-1] == 16000
    x = golden_frontend["pcm"][0]
    out = sess.run(None, {"input": (x.astype(np.float32) / 32768.0).reshape(1, -1)})
    assert out[0].shape == (1, 1, 1) and out[0].dtype == np.float32
    stream = np.concatenate([golden_frontend["pcm"][0], golden_frontend["pcm"][1], golden_frontend["pcm"][6]])
    oracle = OracleInterpreter(sd, cfg, name="hey_b200")
    for i in range(0, len(stream), 1280):
        chunk = stream[i:i + 1280]
        r = interp.predict(chunk)
        o = oracle.predict(chunk)
        assert abs(r.score - o["hey_b200"]) < SCORE_TOL
        assert abs(interp.raw_scores["hey_b200"] - oracle.raw_scores["hey_b200"]) < SCORE_TOL
    assert interp.detected(0.0)
    interp.reset()
    assert interp.score == 0.0 and interp.e2e_buffer_samples["hey_b200"] == 0


@pytest.mark.parametrize("mt,chunk_len", [("cnn", 1280), ("cnn", 1000), ("dnn", 1280), ("tcn", 777), ("e2e_quartznet", 1001),
                                          ("e2e_dnn", 999)])       # (windows at odd ring offsets through the REF64x101 front end)
def test_stream_rings_match_oracle_interpreters(torch_cuda, golden_frontend, mt, chunk_len):
    """Multi-stream mode (nww_stream_*): every stream of a StreamBank must behave like its own
    reference interpreter (oracle/interp.py restates nanointerpreter.py:735-814) fed the same
    chunks — including chunk lengths that leave the ring's window on an odd int16 boundary."""
    from nanowakeword_b200 import StreamBank
    from oracle.heads import forward_scores
    from oracle.interp import OracleInterpreter
    eng, sd, cfg = _engine(mt)
    n = 5
    g = golden_frontend["pcm"]
    audio = np.stack([np.concatenate([g[(i + k) % len(g)] for k in range(3)]) for i in range(n)])   # (n, 48000)
    bank = StreamBank(eng, n)
    oracles = [OracleInterpreter(sd, cfg, name="m") for _ in range(n)]
    n_steps = audio.shape[1] // chunk_len
    windows, where = [], []
    for s in range(n_steps):
        chunks = np.ascontiguousarray(audio[:, s * chunk_len:(s + 1) * chunk_len])
        if s == n_steps // 2:
            bank.reset([2])
            oracles[2].reset()
        got = bank.push(chunks)
        for i, o in enumerate(oracles):
            want = o.predict(chunks[i])["m"]
            assert abs(got[i] - want) < SCORE_TOL, (s, i, got[i], want)
            assert abs(bank.raw_scores[i] - o.raw_scores["m"]) < SCORE_TOL
    bank.reset()
    assert np.all(bank.push(np.zeros((n, chunk_len), np.int16)) == 0.0)
    bank.close()


def test_stream_push_device_equals_batch_scoring(torch_cuda):
    """After each stream has received exactly one window's worth of audio, the ring path and the
    batch path see the same 16000 samples and must return bit-identical scores."""
    eng, sd, cfg = _engine("cnn")
    n = 300
    pcm = synth_pcm(n, seed=21, kind="gauss")
    batch = eng.score_device(torch_cuda.from_numpy(pcm).cuda()).cpu().numpy()
    eng.stream_open(n)
    dev = torch_cuda.from_numpy(pcm).cuda()
    for s in range(0, 16000, 3200):
        scores = eng.stream_push_device(dev[:, s:s + 3200].contiguous())
        if s + 3200 < 16000:
            assert float(scores.abs().max()) == 0.0            # not enough audio yet
    assert np.array_equal(scores.cpu().numpy(), batch)
    eng.stream_close()
    with pytest.raises(ValueError):
        eng.stream_push_host(np.zeros((n, 100), np.int16))      # closed: no streams are open


@pytest.mark.parametrize("mt", ["cnn", "dnn", "tcn", "bcresnet", "crnn", "gru", "lstm", "rnn", "quartznet"])
def test_incremental_stream_mel_is_bit_identical_to_full_recompute(torch_cuda, mt):
    """The mel ring (nww_stream_mel.cuh) only computes the frames a chunk completes; because the
    NS40x98 front end is not centred those are the same arithmetic on the same samples as in a
    full re-run of the window, so both modes must agree to the last bit — through per-stream
    resets and through a full reset."""
    from nanowakeword_b200 import Engine
    cfg = default_config(mt)
    sd = make_state_dict(cfg, seed=0)
    inc = Engine(sd, cfg, device=0)
    full = Engine(sd, cfg, device=0, stream_incremental=False)
    n, L = 37, 1280
    inc.stream_open(n)
    full.stream_open(n)
    rng = np.random.default_rng(7)
    seen_nonzero = False
    for step in range(32):
        if step == 20:
            inc.stream_reset([3, 36])
            full.stream_reset([3, 36])
        if step == 27:
            inc.stream_reset()
            full.stream_reset()
        Ls = L if step % 5 else 640                       # mixed chunk lengths, all multiples of the hop
        chunks = np.clip(rng.normal(0, 3000, (n, Ls)), -32768, 32767).astype(np.int16)
        a = inc.stream_push_host(chunks)
        b = full.stream_push_host(chunks)
        assert np.array_equal(a, b), (mt, step, np.abs(a - b).max())
        seen_nonzero |= bool((a != 0).any())
    assert seen_nonzero
    inc.stream_close()
    full.stream_close()


@pytest.mark.parametrize("mt,kw", [("cnn", dict(tensor_cores=False)), ("cnn", dict(cnn_stage="v1")),
                                   ("cnn", dict(cnn_stage="v3")), ("cnn", dict(cnn_stage="v4")), ("crnn", dict(cnn_stage="v4")),
                                   ("tcn", dict(tcn_layers="cone")), ("cnn", dict(stream_ingest="two_kernels")),
                                   ("dnn", dict(tensor_cores=False)), ("bcresnet", dict(tensor_cores=False)),
                                   ("crnn", dict(tensor_cores=False)), ("e2e_dnn", dict(tensor_cores=False)),
                                   ("tcn", dict(tensor_cores=False))])
def test_cuda_core_variants_match_golden(torch_cuda, golden_frontend, mt, kw):
    """The FP32 CUDA-core variants kept for A/B measurements (no tcgen05 dense layer / conv2 / 1x1 GEMMs)
    must meet the same tolerances as the default tensor-core paths."""
    eng, sd, cfg = _engine(mt, **kw)
    g = load_golden_head(mt)
    scores = eng.score_device(torch_cuda.from_numpy(golden_frontend["pcm"]).cuda()).cpu().numpy()
    assert np.abs(scores - g["scores64"].ravel()).max() < SCORE_TOL


def test_single_stream_with_chunks_longer_than_the_clip(torch_cuda, golden_frontend):
    """One stream, chunks of 20000 samples (> clip_samples): only the last 16000 samples of a chunk
    matter (nanointerpreter.py:751-756: the deque drops the rest), the counter passes clip_samples on the
    first push, and the score is the batch score of that tail."""
    from nanowakeword_b200 import StreamBank
    from oracle.interp import OracleInterpreter
    eng, sd, cfg = _engine("cnn")
    g = golden_frontend["pcm"]
    audio = np.concatenate([g[0], g[1], g[4], g[5], g[6]])[:80000]
    bank = StreamBank(eng, 1)
    oracle = OracleInterpreter(sd, cfg, name="m")
    for s in range(4):
        chunk = audio[s * 20000:(s + 1) * 20000]
        got = bank.push(chunk[None, :])
        want = oracle.predict(chunk)["m"]
        assert abs(got[0] - want) < SCORE_TOL
        assert abs(bank.raw_scores[0] - oracle.raw_scores["m"]) < SCORE_TOL
        tail = eng.score_host(np.ascontiguousarray(chunk[-16000:][None, :]))
        assert abs(bank.raw_scores[0] - tail[0]) < 1e-6
    bank.close()


@pytest.mark.parametrize("act", ["gelu", "silu"])
@pytest.mark.parametrize("mt", ["cnn", "dnn", "bcresnet", "crnn", "e2e_dnn"])
def test_other_activations_match_oracle(torch_cuda, mt, act):
    """model.py:81-87 lets a model use GELU (exact erf) or SiLU instead of ReLU; the fused kernels carry the
    activation as a template / runtime code, so every variant is held to the oracle."""
    from nanowakeword_b200 import Engine
    from oracle.heads import forward_scores
    cfg = default_config(mt, activation_function=act)
    sd = make_state_dict(cfg, seed=1)
    eng = Engine(sd, cfg, device=0)
    pcm = np.concatenate([synth_pcm(24, seed=31, kind="gauss"), synth_pcm(8, seed=32, kind="uniform")])
    ref = forward_scores(pcm, sd, cfg).ravel()
    got = eng.score_device(torch_cuda.from_numpy(pcm).cuda()).cpu().numpy()
    assert np.abs(got - ref).max() < SCORE_TOL, (mt, act, np.abs(got - ref).max())


@pytest.mark.parametrize("mt,kw", [("gru", dict(layer_dim=64)), ("lstm", dict(layer_dim=64)), ("gru", dict(activation_function="gelu")),
                                   ("quartznet", dict(quartznet_config=[[64, 11, 1], [64, 13, 2], [128, 17, 1], [128, 5, 1]])),
                                   ("e2e_cnn", dict(activation_function="silu")), ("e2e_cnn", dict(activation_function="gelu"))])
def test_sequence_head_variants_match_oracle(torch_cuda, mt, kw):
    """Other shapes of the §8(f) sequence heads: 64 hidden units (the second template instantiation of
    rnn_seq_kernel), a non-ReLU classifier, and a QuartzNet config with repeated blocks (identity residual), the
    e2e-style kernel sizes 11 / 13 / 17 and one size (5) that takes the generic depthwise kernel."""
    from nanowakeword_b200 import Engine
    from oracle.heads import forward_scores
    cfg = default_config(mt, **kw)
    sd = make_state_dict(cfg, seed=2)
    eng = Engine(sd, cfg, device=0)
    pcm = np.concatenate([synth_pcm(45, seed=41, kind="gauss"), synth_pcm(25, seed=42, kind="uniform")])
    ref = forward_scores(pcm, sd, cfg).ravel()
    got = eng.score_device(torch_cuda.from_numpy(pcm).cuda()).cpu().numpy()
    assert np.abs(got - ref).max() < SCORE_TOL, (mt, kw, np.abs(got - ref).max())


@pytest.mark.parametrize("mt", ["gru", "lstm", "quartznet", "e2e_quartznet", "e2e_cnn"])
def test_sequence_heads_do_not_depend_on_batch_composition(torch_cuda, mt):
    """A window's score must not depend on which tile / CTA / chunk it lands in: the recurrent kernel uses 32-row tiles
    for small batches and 128-row tiles for large ones, the row GEMM walks K in a fixed order."""
    eng, sd, cfg = _engine(mt)
    base = np.concatenate([synth_pcm(50, seed=51, kind="gauss"), synth_pcm(20, seed=52, kind="uniform")])
    small = eng.score_device(torch_cuda.from_numpy(base).cuda()).cpu().numpy()
    reps = 203 if mt in ("gru", "lstm") else 60                     # 14 210 windows: 128-row tiles, ragged last tile
    big = eng.score_device(torch_cuda.from_numpy(np.tile(base, (reps, 1))).cuda()).cpu().numpy()
    assert np.array_equal(big.reshape(reps, -1), np.tile(small, (reps, 1)))
