"""Host-side packer: reference ``state_dict`` -> the flat blob the CUDA engine consumes.

Everything that is free at load time is done here once, in float64, so the kernels see
inference-ready tensors:

* BatchNorm (eval mode, eps 1e-5) folded into the preceding conv / linear
  (CRNN ``model.cnn.*`` architectures.py:222-230, BcResNet architectures.py:627-641,
  E2E mel-CNN architectures.py:840-865 incl. the BatchNorm1d after fc1);
* Dropout dropped (identity in eval);
* conv weights re-laid out input-channel-major with the output channel contiguous, which is
  what the shared-memory broadcast loads in the kernels want;
* the dense tail (every ``Linear`` after the per-window body, the classifier of
  modules/model.py:291-296 included) expressed as one list of layers.

The blob format is documented in ``csrc/nww_blob.h``.
"""
from __future__ import annotations

import os
import struct

import numpy as np

BN_EPS = 1e-5

ARCH_IDS = {"dnn": 0, "cnn": 1, "tcn": 2, "bcresnet": 3, "crnn": 4, "e2e_dnn": 5, "gru": 6, "lstm": 7, "rnn": 7, "quartznet": 8, "e2e_quartznet": 9, "e2e_cnn": 10}
ACT_IDS = {"relu": 0, "gelu": 1, "silu": 2}
POST_NONE, POST_ACT, POST_LN_ACT = 0, 1, 2

GEOMETRY_PARAMS = {
    # name: n_fft, win_length, hop_length, n_mels, center, clip_samples
    "NS40x98": dict(n_fft=512, win_length=400, hop_length=160, n_mels=40, center=0, clip_samples=16000),
    "REF64x101": dict(n_fft=400, win_length=400, hop_length=160, n_mels=64, center=1, clip_samples=16000),
}
GEOMETRY_IDS = {"NS40x98": 0, "REF64x101": 1}

_TABLE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables")


def geometry_for(cfg: dict) -> str:
    """The front-end geometry a model is used with: the reference's own e2e_dnn carries
    REF64x101 inside the graph (architectures.py:830-836); the feature heads named by
    BASELINE.json are fed the 40x98 log-mel."""
    return cfg.get("geometry") or ("REF64x101" if cfg["model_type"] == "e2e_dnn" else "NS40x98")


def frontend_tables(geometry: str, sd: dict | None = None):
    """(window, fb) as float32 — from the model itself when it carries them, else the
    torchaudio tables shipped for the pinned geometry."""
    if sd is not None and "model.mel_spec.spectrogram.window" in sd:
        return (np.asarray(sd["model.mel_spec.spectrogram.window"], np.float32),
                np.asarray(sd["model.mel_spec.mel_scale.fb"], np.float32))
    t = np.load(os.path.join(_TABLE_DIR, geometry + ".npz"))
    return t["window"].astype(np.float32), t["fb"].astype(np.float32)


def _f64(sd, k):
    return np.asarray(sd[k], dtype=np.float64)


def fold_bn(w, b, sd, bn_prefix):
    """Fold y = BN(conv(x)) into conv weights/bias; ``w`` is (Cout, ...)."""
    g, beta = _f64(sd, bn_prefix + ".weight"), _f64(sd, bn_prefix + ".bias")
    mu, var = _f64(sd, bn_prefix + ".running_mean"), _f64(sd, bn_prefix + ".running_var")
    s = g / np.sqrt(var + BN_EPS)
    w2 = w * s.reshape((-1,) + (1,) * (w.ndim - 1))
    b0 = np.zeros(w.shape[0]) if b is None else b
    return w2, (b0 - mu) * s + beta


def _tail_layer(out, i, w, b, post, ln=None):
    out[f"tail.{i}.W"] = np.ascontiguousarray(w, dtype=np.float32)
    out[f"tail.{i}.b"] = np.ascontiguousarray(b, dtype=np.float32)
    out[f"tail.{i}.post"] = np.array([post], dtype=np.int32)
    if ln is not None:
        out[f"tail.{i}.ln_g"] = np.ascontiguousarray(ln[0], dtype=np.float32)
        out[f"tail.{i}.ln_b"] = np.ascontiguousarray(ln[1], dtype=np.float32)


def _conv3x3_ic_tap_oc(w):
    """(Cout, Cin, 3, 3) -> (Cin, 9, Cout)"""
    cout, cin = w.shape[:2]
    return np.ascontiguousarray(w.reshape(cout, cin, 9).transpose(1, 2, 0))


def pack_tensors(sd: dict, cfg: dict) -> dict[str, np.ndarray]:
    """Named, kernel-ready tensors for ``cfg['model_type']``."""
    mt = cfg["model_type"]
    geom = geometry_for(cfg)
    win, fb = frontend_tables(geom, sd)
    out: dict[str, np.ndarray] = {"frontend.window": win, "frontend.fb": fb}
    layers = []                                    # (W, b, post, ln)

    if mt == "dnn":
        layers.append((_f64(sd, "model.layer1.weight"), _f64(sd, "model.layer1.bias"), POST_LN_ACT,
                       (_f64(sd, "model.layernorm1.weight"), _f64(sd, "model.layernorm1.bias"))))
        i = 0
        while f"model.blocks.{i}.fcn_layer.weight" in sd:
            p = f"model.blocks.{i}"
            layers.append((_f64(sd, p + ".fcn_layer.weight"), _f64(sd, p + ".fcn_layer.bias"), POST_LN_ACT,
                           (_f64(sd, p + ".layer_norm.weight"), _f64(sd, p + ".layer_norm.bias"))))
            i += 1
        layers.append((_f64(sd, "model.last_layer.weight"), _f64(sd, "model.last_layer.bias"), POST_NONE, None))
    elif mt == "cnn":
        out["cnn.w1"] = _f64(sd, "model.conv1.weight").reshape(16, 9).astype(np.float32)
        out["cnn.b1"] = _f64(sd, "model.conv1.bias").astype(np.float32)
        out["cnn.w2"] = _conv3x3_ic_tap_oc(_f64(sd, "model.conv2.weight")).astype(np.float32)
        out["cnn.b2"] = _f64(sd, "model.conv2.bias").astype(np.float32)
        layers.append((_f64(sd, "model.fc1.weight"), _f64(sd, "model.fc1.bias"), POST_ACT, None))
        layers.append((_f64(sd, "model.fc2.weight"), _f64(sd, "model.fc2.bias"), POST_NONE, None))
    elif mt == "tcn":
        i = 0
        while f"model.tcn_blocks.{i}.conv1.weight" in sd:
            p = f"model.tcn_blocks.{i}"
            for cv in ("conv1", "conv2"):
                w = _f64(sd, f"{p}.{cv}.weight")                      # (Cout, Cin, k)
                out[f"tcn.{i}.{cv}.w"] = np.ascontiguousarray(w.transpose(2, 1, 0)).astype(np.float32)  # (k, Cin, Cout)
                out[f"tcn.{i}.{cv}.b"] = _f64(sd, f"{p}.{cv}.bias").astype(np.float32)
            if p + ".downsample.weight" in sd:
                w = _f64(sd, p + ".downsample.weight")[:, :, 0]       # (Cout, Cin)
                out[f"tcn.{i}.down.w"] = np.ascontiguousarray(w.T).astype(np.float32)   # (Cin, Cout)
                out[f"tcn.{i}.down.b"] = _f64(sd, p + ".downsample.bias").astype(np.float32)
            i += 1
        layers.append((_f64(sd, "model.fc.weight"), _f64(sd, "model.fc.bias"), POST_NONE, None))
    elif mt == "bcresnet":
        w, b = fold_bn(_f64(sd, "model.init_conv.0.weight"), None, sd, "model.init_conv.1")
        out["bc.init.w"] = _conv3x3_ic_tap_oc(w).astype(np.float32)               # (1, 9, 32)
        out["bc.init.b"] = b.astype(np.float32)
        for j, name in enumerate(("block1", "block2", "block3")):
            p = "model." + name
            dw = _f64(sd, p + ".depthwise.weight")                    # (C, 1, 3, 3)
            out[f"bc.{j}.dw"] = dw.reshape(dw.shape[0], 9).astype(np.float32)
            pw, pb = fold_bn(_f64(sd, p + ".pointwise.weight")[:, :, 0, 0], None, sd, p + ".bn1")
            sw, sb = fold_bn(_f64(sd, p + ".shortcut.0.weight")[:, :, 0, 0], None, sd, p + ".shortcut.1")
            out[f"bc.{j}.pw.w"] = np.ascontiguousarray(pw.T).astype(np.float32)    # (Cin, Cout)
            out[f"bc.{j}.pw.b"] = pb.astype(np.float32)
            out[f"bc.{j}.sc.w"] = np.ascontiguousarray(sw.T).astype(np.float32)    # (Cin, Cout)
            out[f"bc.{j}.sc.b"] = sb.astype(np.float32)
        layers.append((_f64(sd, "model.fc.weight"), _f64(sd, "model.fc.bias"), POST_NONE, None))
    elif mt == "crnn":
        i = 0
        while f"model.cnn.{4 * i}.weight" in sd:
            w, b = fold_bn(_f64(sd, f"model.cnn.{4 * i}.weight"), _f64(sd, f"model.cnn.{4 * i}.bias"),
                           sd, f"model.cnn.{4 * i + 1}")
            out[f"crnn.conv{i}.w"] = _conv3x3_ic_tap_oc(w).astype(np.float32)       # (Cin, 9, Cout)
            out[f"crnn.conv{i}.b"] = b.astype(np.float32)
            i += 1
        # CRNNModel builds its recurrent part with num_layers = n_blocks and rnn_type from crnn_rnn_type
        # (architectures.py:243-262, default 'lstm'); the engine holds ONE bidirectional GRU layer.  Refuse anything
        # else here instead of packing layer 0 only (silently wrong scores) or failing on an opaque size mismatch.
        if "model.rnn.weight_ih_l1" in sd:
            raise ValueError("crnn: only a single recurrent layer (n_blocks = 1) is built into the B200 engine")
        if str(cfg.get("crnn_rnn_type", "gru")).lower() != "gru":
            raise ValueError(f"crnn: crnn_rnn_type '{cfg.get('crnn_rnn_type')}' is not built into the B200 engine (gru only)")
        w_hh0 = _f64(sd, "model.rnn.weight_hh_l0")
        if w_hh0.shape[0] != 3 * w_hh0.shape[1]:
            raise ValueError(f"crnn: model.rnn.weight_hh_l0 has shape {w_hh0.shape}; a GRU layer has (3H, H) "
                             "(an LSTM checkpoint has (4H, H): crnn_rnn_type must be 'gru')")
        if "model.rnn.weight_ih_l0_reverse" not in sd:
            raise ValueError("crnn: the B200 engine builds the bidirectional GRU of CRNNModel only")
        for sfx, tag in (("", "fwd"), ("_reverse", "bwd")):
            out[f"crnn.gru.{tag}.w_ih_nk"] = _f64(sd, "model.rnn.weight_ih_l0" + sfx).astype(np.float32)  # (3H, In), dense-kernel layout
            out[f"crnn.gru.{tag}.w_ih_kn"] = np.ascontiguousarray(_f64(sd, "model.rnn.weight_ih_l0" + sfx).T).astype(np.float32)  # (In, 3H), row-GEMM layout
            out[f"crnn.gru.{tag}.w_hh"] = np.ascontiguousarray(_f64(sd, "model.rnn.weight_hh_l0" + sfx).T).astype(np.float32)  # (H, 3H)
            out[f"crnn.gru.{tag}.b_ih"] = _f64(sd, "model.rnn.bias_ih_l0" + sfx).astype(np.float32)
            out[f"crnn.gru.{tag}.b_hh"] = _f64(sd, "model.rnn.bias_hh_l0" + sfx).astype(np.float32)
        layers.append((_f64(sd, "model.fc.weight"), _f64(sd, "model.fc.bias"), POST_NONE, None))
    elif mt in ("gru", "lstm", "rnn"):
        # GRUModel / LSTMModel / RNNModel (architectures.py:129-146, 83-99, 149-161): one bidirectional layer,
        # out[:, -1, :].  Packed for csrc/nww_rnn.cuh: per direction ONE matrix whose rows are
        # [x (In) | constant 1 -> bias | zero padding to a multiple of 16 | h (H)] and whose 4H columns are all gates.
        name = {"gru": "model.gru", "lstm": "model.lstm", "rnn": "model.layer1"}[mt]
        if f"{name}.weight_ih_l1" in sd:
            raise ValueError(f"{mt}: only single-layer recurrent heads (n_blocks = 1) are built into the B200 engine")
        for sfx, tag in (("", "fwd"), ("_reverse", "bwd")):
            w_ih, w_hh = _f64(sd, f"{name}.weight_ih_l0{sfx}"), _f64(sd, f"{name}.weight_hh_l0{sfx}")
            b_ih, b_hh = _f64(sd, f"{name}.bias_ih_l0{sfx}"), _f64(sd, f"{name}.bias_hh_l0{sfx}")
            hid, n_in = w_hh.shape[1], w_ih.shape[1]
            kx = (n_in + 1 + 15) // 16 * 16
            w = np.zeros((kx + hid, 4 * hid))
            if mt == "gru":                                    # torch gate order r, z, n -> columns [r | z | n_x | n_h]
                w[:n_in, :3 * hid] = w_ih.T
                w[n_in, :2 * hid] = b_ih[:2 * hid] + b_hh[:2 * hid]
                w[n_in, 2 * hid:3 * hid] = b_ih[2 * hid:]
                w[n_in, 3 * hid:] = b_hh[2 * hid:]
                w[kx:, :2 * hid] = w_hh.T[:, :2 * hid]
                w[kx:, 3 * hid:] = w_hh.T[:, 2 * hid:]
            else:                                              # torch gate order i, f, g, o
                w[:n_in] = w_ih.T
                w[n_in] = b_ih + b_hh
                w[kx:] = w_hh.T
            # the reverse direction contributes its first step only (zero state): x rows suffice
            out[f"rnn.{tag}.w"] = np.ascontiguousarray(w if tag == "fwd" else w[:kx]).astype(np.float32)
        out["rnn.cell"] = np.array([0 if mt == "gru" else 1], dtype=np.int32)
        fc = "model.layer2" if mt == "rnn" else "model.fc"
        layers.append((_f64(sd, fc + ".weight"), _f64(sd, fc + ".bias"), POST_NONE, None))
    elif mt in ("quartznet", "e2e_quartznet", "e2e_cnn"):
        pre = "model"
        if mt in ("e2e_quartznet", "e2e_cnn"):
            # E2ERawQuartzNet (architectures.py:796-817): RawAudioFrontend (:695-714) = strided Conv1d (no bias) +
            # BatchNorm + ReLU layers straight on the audio.  On channel-last buffers a strided Conv1d is a row GEMM
            # whose row t is the CONTIGUOUS slice of k * C_in floats starting at input step stride * t - pad, so each
            # layer is packed as one [k * C_in (padded to 64)][C_out (padded to 64)] matrix with BatchNorm folded in.
            pre = "model.backbone"
            i = 0
            while f"model.frontend.conv_blocks.{3 * i}.weight" in sd:
                w, b = fold_bn(_f64(sd, f"model.frontend.conv_blocks.{3 * i}.weight"), None, sd,
                               f"model.frontend.conv_blocks.{3 * i + 1}")
                cout, cin, k = w.shape
                kk = -(-(k * cin) // 64) * 64
                g = np.zeros((kk, -(-cout // 64) * 64))
                g[:k * cin, :cout] = w.transpose(2, 1, 0).reshape(k * cin, cout)        # row = tap * C_in + channel
                out[f"raw.{i}.w"] = g.astype(np.float32)
                out[f"raw.{i}.b"] = b.astype(np.float32)
                i += 1
        if mt == "e2e_cnn":
            # RawAudioBackbone (architectures.py:738-774): the front end's (C, T) output is a one-channel image; four 3x3
            # Conv2d (no bias) + BatchNorm + activation.  conv1 (1 -> 24) is a direct kernel; conv2..4 are row GEMMs on
            # zero-padded NHWC images whose K axis is three segments (kernel rows) of 3 * C_in contiguous floats:
            # weight row = (kh * 3 + kw) * C_in + ci.
            w, b = fold_bn(_f64(sd, "model.backbone.conv1.0.weight"), None, sd, "model.backbone.conv1.1")
            out["rawcnn.conv1.w"] = np.ascontiguousarray(w.reshape(w.shape[0], 9).T).astype(np.float32)        # (9, 24)
            out["rawcnn.conv1.b"] = b.astype(np.float32)
            for j in (2, 3, 4):
                w, b = fold_bn(_f64(sd, f"model.backbone.conv{j}.0.weight"), None, sd, f"model.backbone.conv{j}.1")
                cout, cin = w.shape[:2]
                kk = -(-(9 * cin) // 64) * 64
                g = np.zeros((kk, -(-cout // 64) * 64))
                g[:9 * cin, :cout] = w.transpose(2, 3, 1, 0).reshape(9 * cin, cout)
                out[f"rawcnn.conv{j}.w"] = g.astype(np.float32)
                out[f"rawcnn.conv{j}.b"] = b.astype(np.float32)
            layers.append((_f64(sd, "model.backbone.fc.weight"), _f64(sd, "model.backbone.fc.bias"), POST_NONE, None))
        # QuartzNetModel (architectures.py:366-437).  Per block the engine runs a depthwise FIR (no bias) and ONE row
        # GEMM [dw(x) | x] @ W + b: BatchNorm folded into the pointwise / residual 1x1 weights, the depthwise bias
        # pushed through the pointwise weights into b.  Channel counts are padded so that K is a multiple of 64.
        i = 0
        while f"{pre}.quartznet_blocks.{i}.depthwise_conv.weight" in sd:
            p = f"{pre}.quartznet_blocks.{i}"
            wd, bd = _f64(sd, p + ".depthwise_conv.weight")[:, 0, :], _f64(sd, p + ".depthwise_conv.bias")    # (C, k)
            wp, bp = _f64(sd, p + ".pointwise_conv.weight")[:, :, 0], _f64(sd, p + ".pointwise_conv.bias")    # (N, C)
            c, k = wd.shape
            n_out = wp.shape[0]
            wp_f, b = fold_bn(wp, bp + wp @ bd, sd, p + ".batch_norm")
            has_res = p + ".residual_connector.0.weight" in sd
            cp = -(-c // 32) * 32 if has_res else -(-c // 64) * 64
            w = np.zeros((cp * (2 if has_res else 1), n_out))
            w[:c] = wp_f.T
            if has_res:
                wr_f, br = fold_bn(_f64(sd, p + ".residual_connector.0.weight")[:, :, 0],
                                   _f64(sd, p + ".residual_connector.0.bias"), sd, p + ".residual_connector.1")
                w[cp:cp + c] = wr_f.T
                b = b + br
            dw = np.zeros((k, cp))
            dw[:, :c] = wd.T
            out[f"qn.{i}.dw"] = dw.astype(np.float32)
            out[f"qn.{i}.w"] = w.astype(np.float32)
            out[f"qn.{i}.b"] = b.astype(np.float32)
            out[f"qn.{i}.meta"] = np.array([c, n_out, k, int(has_res)], dtype=np.int32)
            i += 1
        if mt != "e2e_cnn":
            layers.append((_f64(sd, pre + ".fc.weight"), _f64(sd, pre + ".fc.bias"), POST_NONE, None))
    elif mt == "e2e_dnn":
        for j, i in enumerate((0, 4, 8)):
            w, b = fold_bn(_f64(sd, f"model.conv_block.{i}.weight"), _f64(sd, f"model.conv_block.{i}.bias"),
                           sd, f"model.conv_block.{i + 1}")
            out[f"e2e.conv{j}.w"] = _conv3x3_ic_tap_oc(w).astype(np.float32)
            out[f"e2e.conv{j}.b"] = b.astype(np.float32)
        w, b = fold_bn(_f64(sd, "model.fc1.weight"), _f64(sd, "model.fc1.bias"), sd, "model.bn1")
        layers.append((w, b, POST_ACT, None))
        layers.append((_f64(sd, "model.out.weight"), _f64(sd, "model.out.bias"), POST_NONE, None))
    else:
        raise ValueError(f"Unsupported model_type: '{mt}'.")

    layers.append((_f64(sd, "classifier.0.weight"), _f64(sd, "classifier.0.bias"), POST_ACT, None))
    layers.append((_f64(sd, "classifier.3.weight"), _f64(sd, "classifier.3.bias"), POST_NONE, None))
    if layers[-1][0].shape[0] != 1:
        raise ValueError("only single-class (n_classes=1) models are supported")
    for i, (w, b, post, ln) in enumerate(layers):
        _tail_layer(out, i, w, b, post, ln)
    out["tail.n_layers"] = np.array([len(layers)], dtype=np.int32)
    return out


def pack_blob(tensors: dict[str, np.ndarray]) -> bytes:
    """Serialise named tensors (float32 / int32) into the engine's blob format."""
    buf = bytearray(b"NWWB200\0" + struct.pack("<II", 1, len(tensors)))
    for name, arr in tensors.items():
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.float32:
            dtype = 0
        elif arr.dtype == np.int32:
            dtype = 1
        else:
            raise TypeError(f"{name}: unsupported dtype {arr.dtype}")
        nb = name.encode()
        buf += struct.pack("<I", len(nb)) + nb
        buf += struct.pack("<II", dtype, arr.ndim) + struct.pack(f"<{arr.ndim}I", *arr.shape)
        buf += struct.pack("<Q", arr.nbytes)
        buf += b"\0" * ((-len(buf)) % 16)
        buf += arr.tobytes()
    return bytes(buf)
