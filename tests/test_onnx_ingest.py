"""CPU tests of the .onnx ingestion: hand-serialised graphs in the shape the reference's
exporter writes them (tests/golden/onnx_writer.py) -> onnx_reader -> (state_dict, cfg) -> oracle -> golden scores."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_golden_head
from nanowakeword_b200.onnx_reader import load_onnx, parse_onnx
from nanowakeword_b200.session import load_artifacts
from nanowakeword_b200.synth import default_config, make_state_dict
from nanowakeword_b200.weights import pack_tensors

sys.path.insert(0, GOLDEN)
from onnx_writer import write_e2e_model, write_feature_head  # noqa: E402


@pytest.mark.parametrize("style", ["folded", "explicit"])
@pytest.mark.parametrize("mt", ["e2e_dnn", "e2e_cnn", "e2e_quartznet"])
def test_onnx_graph_round_trips_to_golden_scores(tmp_path, golden_frontend, mt, style):
    from oracle.heads import forward_scores
    cfg = default_config(mt)
    sd = make_state_dict(cfg, seed=0)                       # the weights the golden vectors were made with
    path = write_e2e_model(str(tmp_path / f"{mt}.onnx"), sd, cfg, style=style)
    m = parse_onnx(path)
    assert m.metadata == {"mode": "e2e"} and m.inputs[0] == ("input", ["batch_size", 1, 16000]) and m.opset == 17
    sd2, cfg2 = load_onnx(path)
    assert cfg2["model_type"] == mt and cfg2["activation_function"] == "relu" and cfg2["input_ndim"] == 3
    assert cfg2["embedding_dim"] == cfg["embedding_dim"]
    if mt == "e2e_quartznet":
        assert cfg2["e2e_quartznet_config"] == cfg["e2e_quartznet_config"] and cfg2["e2e_frontend_depth"] == 3
    g = load_golden_head(mt)
    got = forward_scores(golden_frontend["pcm"], sd2, cfg2).ravel()
    assert np.abs(got - g["scores64"].ravel()).max() < 1e-6
    # and the engine's packer accepts it: same kernel-ready tensors as from the state_dict, up to the BN fold's rounding
    a, b = pack_tensors(sd, cfg), pack_tensors(sd2, cfg2)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape and np.allclose(a[k], b[k], rtol=2e-6, atol=2e-7), k


@pytest.mark.parametrize("act", ["gelu", "silu"])
@pytest.mark.parametrize("mt", ["e2e_dnn", "e2e_cnn"])
def test_onnx_activation_detection(tmp_path, mt, act):
    cfg = default_config(mt, activation_function=act)
    sd = make_state_dict(cfg, seed=3)
    for ndim in (2, 3):
        path = write_e2e_model(str(tmp_path / f"m{ndim}.onnx"), sd, cfg, style="explicit", input_ndim=ndim, e2e_metadata=(ndim == 3))
        sd2, cfg2 = load_onnx(path)      # ndim 2 without metadata: the interpreter's shape heuristic (nanointerpreter.py:984-992)
        assert cfg2["activation_function"] == act and cfg2["model_type"] == mt and cfg2["input_ndim"] == ndim


def test_load_artifacts_prefers_what_the_trainer_writes(tmp_path):
    """trainer.py:474-535 leaves <name>.onnx, <name>_lite.onnx and <name>.pt (main model only), no sidecar."""
    import torch
    cfg = default_config("e2e_cnn")
    sd = make_state_dict(cfg, seed=0)
    stem = str(tmp_path / "hey")
    write_e2e_model(stem + ".onnx", sd, cfg)
    write_e2e_model(stem + "_lite.onnx", make_state_dict(cfg, seed=1), cfg)
    torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, stem + ".pt")
    for p in (stem + ".onnx", stem + ".pt", stem + "_lite.onnx"):
        sd2, cfg2 = load_artifacts(p)
        assert cfg2["model_type"] == "e2e_cnn" and "classifier.3.weight" in sd2
    os.remove(stem + ".onnx")
    with pytest.raises(NotImplementedError, match="does not identify its architecture"):
        load_artifacts(stem + ".pt")
    with pytest.raises(FileNotFoundError):
        load_artifacts(stem + ".onnx")


def test_embedding_mode_graphs_and_garbage_are_refused(tmp_path):
    p = write_feature_head(str(tmp_path / "feat.onnx"))
    with pytest.raises(NotImplementedError, match="embedding"):
        load_onnx(p)
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x00\x01garbage that is not a protobuf" * 3)
    with pytest.raises((ValueError, NotImplementedError)):
        load_onnx(str(bad))
