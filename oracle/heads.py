"""Oracle classifier heads in numpy.  TEST INFRASTRUCTURE ONLY.

Each function restates one reference module (eval mode: Dropout = identity, BatchNorm
uses running statistics) and consumes the reference's own ``state_dict`` key names, so
the same weights drive the reference (when generating goldens), this oracle and the
CUDA engine.

Reference anchors (all under /root/reference/nanowakeword/):
  Net/FCNBlock (DNN)        modules/architectures.py:102-126
  CNNModel                  modules/architectures.py:51-80
  TCNModel/TemporalBlock    modules/architectures.py:290-362
  BcResNetModel/Block       modules/architectures.py:620-687
  CRNNModel (GRU)           modules/architectures.py:209-287
  LSTMModel                 modules/architectures.py:83-99
  GRUModel                  modules/architectures.py:129-146
  RNNModel (bi-LSTM, H=64)  modules/architectures.py:149-161
  QuartzNetModel/Block      modules/architectures.py:366-437
  RawAudioFrontend          modules/architectures.py:695-714
  E2ERawQuartzNet           modules/architectures.py:796-817
  E2ERawCNN / RawAudioBackbone  modules/architectures.py:738-793
  E2E_MelSpectrogram_CNN    modules/architectures.py:820-888
  Model.classifier/forward  modules/model.py:291-296, 562-571
  sigmoid + view(-1,1,1)    _export/onnx.py:164-172
  AdaptiveAvgPool rewrite   _export/onnx.py:96-154
"""
from __future__ import annotations

import math

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

from .frontend import GEOMETRIES, FrontendSpec, log_mel

BN_EPS = 1e-5
LN_EPS = 1e-5

# Heads whose input is (T, F) rather than (F, T)  (model.py:128-236 passes input_shape[1]
# as the feature dim for dnn/tcn; cnn/bcresnet/crnn treat input_shape as (freq, time)).
TIME_MAJOR_HEADS = ("dnn", "tcn", "gru", "lstm", "rnn", "quartznet")


# ----------------------------------------------------------------------------- primitives
def _erf(x):
    try:
        from scipy.special import erf
        return erf(x)
    except Exception:  # pragma: no cover
        return np.vectorize(math.erf)(x)


def activation(x, name: str):
    """ReLU / exact-erf GELU / SiLU, selected as in model.py:81-87."""
    name = (name or "relu").lower()
    if name == "gelu":
        return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))
    if name == "silu":
        return x / (1.0 + np.exp(-x))
    return np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def linear(x, w, b=None):
    y = x @ w.T
    return y if b is None else y + b


def layernorm(x, w, b):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)      # biased, as torch
    return (x - mu) / np.sqrt(var + LN_EPS) * w + b


def batchnorm(x, sd, prefix, axis=1):
    shape = [1] * x.ndim
    shape[axis] = -1
    g = sd[prefix + ".weight"].reshape(shape)
    b = sd[prefix + ".bias"].reshape(shape)
    m = sd[prefix + ".running_mean"].reshape(shape)
    v = sd[prefix + ".running_var"].reshape(shape)
    return (x - m) / np.sqrt(v + BN_EPS) * g + b


def conv2d(x, w, b=None, stride=(1, 1), pad=(1, 1), groups=1):
    """NCHW cross-correlation, zero padding, as torch.nn.Conv2d (im2col + one matmul)."""
    bsz, cin, _, _ = x.shape
    cout, cin_g, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad[0], pad[0]), (pad[1], pad[1])))
    win = sliding_window_view(xp, (kh, kw), axis=(2, 3))[:, :, ::stride[0], ::stride[1]]
    ho, wo = win.shape[2], win.shape[3]                  # win: (B, Cin, Ho, Wo, kh, kw)
    if groups == 1:
        cols = np.ascontiguousarray(win.transpose(0, 2, 3, 1, 4, 5)).reshape(bsz * ho * wo, cin * kh * kw)
        y = (cols @ w.reshape(cout, cin * kh * kw).T).reshape(bsz, ho, wo, cout).transpose(0, 3, 1, 2)
    else:
        assert groups == cin and cin_g == 1 and cout == cin, "only depthwise grouping is used"
        y = np.einsum("bchwij,cij->bchw", win, w[:, 0], optimize=True)
    if b is not None:
        y = y + b.reshape(1, -1, 1, 1)
    return y


def maxpool2(x):
    """MaxPool2d(kernel 2, stride 2): floor on odd sizes (49 -> 24, 101 -> 50)."""
    h2, w2 = x.shape[2] // 2, x.shape[3] // 2
    x = x[:, :, : 2 * h2, : 2 * w2]
    return x.reshape(x.shape[0], x.shape[1], h2, 2, w2, 2).max(axis=(3, 5))


def avgpool_fixed(x, out_hw):
    """AdaptiveAvgPool2d as deployed: AvgPool2d(k = in-(out-1)*(in//out), s = in//out)
    (_export/onnx.py:139-147).  For the shapes used this equals torch's adaptive bins."""
    _, _, h, w = x.shape
    oh, ow = out_hw
    sh, sw = h // oh, w // ow
    kh, kw = h - (oh - 1) * sh, w - (ow - 1) * sw
    out = np.empty(x.shape[:2] + (oh, ow), dtype=x.dtype)
    for i in range(oh):
        for j in range(ow):
            out[:, :, i, j] = x[:, :, i * sh:i * sh + kh, j * sw:j * sw + kw].mean(axis=(2, 3))
    return out


def causal_conv1d(x, w, b, dilation):
    """Conv1d(padding=(k-1)*d, dilation=d) followed by the manual chomp ``[:, :, :-pad]``
    (architectures.py:313-321): output[t] depends on x[t-(k-1)d .. t] only."""
    _, _, t = x.shape
    cout, cin, k = w.shape
    pad = (k - 1) * dilation
    xp = np.pad(x, ((0, 0), (0, 0), (pad, 0)))
    y = np.zeros((x.shape[0], cout, t), dtype=x.dtype)
    for j in range(k):
        y += np.einsum("bct,oc->bot", xp[:, :, j * dilation:j * dilation + t], w[:, :, j], optimize=True)
    return y + b.reshape(1, -1, 1)


def gru_last_output_bidir(x, sd, prefix):
    """Single-layer bidirectional GRU, batch_first; returns ``out[:, -1, :]`` (B, 2H)
    (architectures.py:279-282).  Gate order r, z, n;
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)).  At the last time index the reverse
    direction has consumed exactly one element (x[:, -1]) from a zero state."""
    def cell(xt, h, sfx):
        w_ih, w_hh = sd[f"{prefix}.weight_ih_l0{sfx}"], sd[f"{prefix}.weight_hh_l0{sfx}"]
        b_ih, b_hh = sd[f"{prefix}.bias_ih_l0{sfx}"], sd[f"{prefix}.bias_hh_l0{sfx}"]
        hsz = w_hh.shape[1]
        gi = xt @ w_ih.T + b_ih
        gh = h @ w_hh.T + b_hh
        r = sigmoid(gi[:, :hsz] + gh[:, :hsz])
        z = sigmoid(gi[:, hsz:2 * hsz] + gh[:, hsz:2 * hsz])
        n = np.tanh(gi[:, 2 * hsz:] + r * gh[:, 2 * hsz:])
        return (1.0 - z) * n + z * h

    bsz, steps, _ = x.shape
    hsz = sd[f"{prefix}.weight_hh_l0"].shape[1]
    h = np.zeros((bsz, hsz), dtype=x.dtype)
    for t in range(steps):
        h = cell(x[:, t], h, "")
    hb = cell(x[:, -1], np.zeros((bsz, hsz), dtype=x.dtype), "_reverse")
    return np.concatenate([h, hb], axis=1)


def _rnn_cell(kind, sd, prefix, layer, sfx):
    """One direction of one layer of torch.nn.GRU / torch.nn.LSTM as a step function
    ``(x_t, state) -> state`` with ``state = (h,)`` or ``(h, c)``.
    GRU gate order r, z, n;  LSTM gate order i, f, g, o  (torch docs; both zero initial state)."""
    w_ih, w_hh = sd[f"{prefix}.weight_ih_l{layer}{sfx}"], sd[f"{prefix}.weight_hh_l{layer}{sfx}"]
    b_ih, b_hh = sd[f"{prefix}.bias_ih_l{layer}{sfx}"], sd[f"{prefix}.bias_hh_l{layer}{sfx}"]
    hsz = w_hh.shape[1]

    def gru(xt, st):
        (h,) = st
        gi = xt @ w_ih.T + b_ih
        gh = h @ w_hh.T + b_hh
        r = sigmoid(gi[:, :hsz] + gh[:, :hsz])
        z = sigmoid(gi[:, hsz:2 * hsz] + gh[:, hsz:2 * hsz])
        n = np.tanh(gi[:, 2 * hsz:] + r * gh[:, 2 * hsz:])
        return ((1.0 - z) * n + z * h,)

    def lstm(xt, st):
        h, c = st
        g = xt @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i, f = sigmoid(g[:, :hsz]), sigmoid(g[:, hsz:2 * hsz])
        gg, o = np.tanh(g[:, 2 * hsz:3 * hsz]), sigmoid(g[:, 3 * hsz:])
        c = f * c + i * gg
        return (o * np.tanh(c), c)

    return (gru if kind == "gru" else lstm), hsz, (1 if kind == "gru" else 2)


def rnn_last_output_bidir(x, sd, prefix, kind):
    """Bidirectional, batch_first, n_layers >= 1 GRU / LSTM; returns ``out[:, -1, :]`` (B, 2H) of the
    top layer (architectures.py:95-96, 141-142, 157-158).  Layers below the top run both directions
    over the whole sequence (their concatenated outputs feed the next layer); at the top the reverse
    direction only needs its first step, the one that consumes x[:, -1]."""
    n_layers = 0
    while f"{prefix}.weight_ih_l{n_layers}" in sd:
        n_layers += 1
    bsz, steps, _ = x.shape
    seq = x
    for layer in range(n_layers):
        top = layer == n_layers - 1
        outs = []
        for sfx in ("", "_reverse"):
            step, hsz, n_state = _rnn_cell(kind, sd, prefix, layer, sfx)
            st = tuple(np.zeros((bsz, hsz), dtype=x.dtype) for _ in range(n_state))
            order = range(steps) if sfx == "" else range(steps - 1, -1, -1)
            if top and sfx:
                order = [steps - 1]
            hs = {}
            for t in order:
                st = step(seq[:, t], st)
                hs[t] = st[0]
            outs.append(hs)
        if top:
            return np.concatenate([outs[0][steps - 1], outs[1][steps - 1]], axis=1)
        seq = np.stack([np.concatenate([outs[0][t], outs[1][t]], axis=1) for t in range(steps)], axis=1)
    raise ValueError("no recurrent layers found under " + prefix)


# ----------------------------------------------------------------------------- backbones
def _dnn(x, sd, cfg):
    act = cfg.get("activation_function", "relu")
    h = x.reshape(x.shape[0], -1)                                   # row-major flatten of (T, F)
    h = activation(layernorm(linear(h, sd["model.layer1.weight"], sd["model.layer1.bias"]),
                             sd["model.layernorm1.weight"], sd["model.layernorm1.bias"]), act)
    i = 0
    while f"model.blocks.{i}.fcn_layer.weight" in sd:
        p = f"model.blocks.{i}"
        h = activation(layernorm(linear(h, sd[p + ".fcn_layer.weight"], sd[p + ".fcn_layer.bias"]),
                                 sd[p + ".layer_norm.weight"], sd[p + ".layer_norm.bias"]), act)
        i += 1
    return linear(h, sd["model.last_layer.weight"], sd["model.last_layer.bias"])


def _cnn(x, sd, cfg):
    act = cfg.get("activation_function", "relu")
    h = x[:, None]
    h = maxpool2(activation(conv2d(h, sd["model.conv1.weight"], sd["model.conv1.bias"]), act))
    h = maxpool2(activation(conv2d(h, sd["model.conv2.weight"], sd["model.conv2.bias"]), act))
    h = h.reshape(h.shape[0], -1)
    h = activation(linear(h, sd["model.fc1.weight"], sd["model.fc1.bias"]), act)
    return linear(h, sd["model.fc2.weight"], sd["model.fc2.bias"])


def _tcn(x, sd, cfg):
    h = np.swapaxes(x, 1, 2)                                        # (B, F, T)
    i = 0
    while f"model.tcn_blocks.{i}.conv1.weight" in sd:
        p = f"model.tcn_blocks.{i}"
        d = 2 ** i
        o = np.maximum(causal_conv1d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], d), 0)
        o = np.maximum(causal_conv1d(o, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], d), 0)
        if p + ".downsample.weight" in sd:
            res = np.einsum("bct,oc->bot", h, sd[p + ".downsample.weight"][:, :, 0]) \
                + sd[p + ".downsample.bias"].reshape(1, -1, 1)
        else:
            res = h
        h = np.maximum(o + res, 0)
        i += 1
    return linear(h[:, :, -1], sd["model.fc.weight"], sd["model.fc.bias"])


def _bcresnet(x, sd, cfg):
    act = cfg.get("activation_function", "relu")
    h = x[:, None]
    h = conv2d(h, sd["model.init_conv.0.weight"])
    h = maxpool2(activation(batchnorm(h, sd, "model.init_conv.1"), act))
    for name, stride in (("block1", (2, 2)), ("block2", (2, 2)), ("block3", (2, 1))):
        p = "model." + name
        res = conv2d(h, sd[p + ".shortcut.0.weight"], stride=stride, pad=(0, 0))
        res = batchnorm(res, sd, p + ".shortcut.1")
        o = conv2d(h, sd[p + ".depthwise.weight"], stride=stride, pad=(1, 1), groups=h.shape[1])
        o = conv2d(o, sd[p + ".pointwise.weight"], pad=(0, 0))
        o = activation(batchnorm(o, sd, p + ".bn1"), act)           # activation BEFORE the add (:646-647)
        h = o + res
    h = h.mean(axis=(2, 3))
    return linear(h, sd["model.fc.weight"], sd["model.fc.bias"])


def _crnn(x, sd, cfg):
    act = cfg.get("activation_function", "relu")
    h = x[:, None]
    i = 0
    while f"model.cnn.{4 * i}.weight" in sd:
        h = conv2d(h, sd[f"model.cnn.{4 * i}.weight"], sd[f"model.cnn.{4 * i}.bias"])
        h = maxpool2(activation(batchnorm(h, sd, f"model.cnn.{4 * i + 1}"), act))
        i += 1
    b, c, hh, w = h.shape
    seq = np.swapaxes(h.reshape(b, c * hh, w), 1, 2)                # (B, W, C*H)
    if cfg.get("crnn_rnn_type", "lstm").lower() != "gru":
        raise NotImplementedError("oracle restates the GRU variant only (BASELINE config #5)")
    last = gru_last_output_bidir(seq, sd, "model.rnn")
    return linear(last, sd["model.fc.weight"], sd["model.fc.bias"])


def _gru(x, sd, cfg):
    """GRUModel on a (T, F) sequence (architectures.py:129-146)."""
    return linear(rnn_last_output_bidir(x, sd, "model.gru", "gru"), sd["model.fc.weight"], sd["model.fc.bias"])


def _lstm(x, sd, cfg):
    """LSTMModel (architectures.py:83-99)."""
    return linear(rnn_last_output_bidir(x, sd, "model.lstm", "lstm"), sd["model.fc.weight"], sd["model.fc.bias"])


def _rnn(x, sd, cfg):
    """RNNModel: bidirectional LSTM with 64 hidden units and n_blocks layers (architectures.py:149-161)."""
    return linear(rnn_last_output_bidir(x, sd, "model.layer1", "lstm"), sd["model.layer2.weight"], sd["model.layer2.bias"])


def _quartznet(x, sd, cfg, prefix="model"):
    """QuartzNetModel on a (T, F) sequence (architectures.py:366-437): per block a depthwise Conv1d
    (padding='same': for a kernel k torch pads (k-1)//2 on the left and the rest on the right), a 1x1 Conv1d,
    BatchNorm1d, plus the residual (1x1 Conv1d + BatchNorm1d when the channel count changes, else the
    identity), ReLU; then the mean over time and a Linear."""
    h = np.swapaxes(x, 1, 2)                                        # (B, C, T)
    i = 0
    while f"{prefix}.quartznet_blocks.{i}.depthwise_conv.weight" in sd:
        p = f"{prefix}.quartznet_blocks.{i}"
        wd = sd[p + ".depthwise_conv.weight"]                       # (C, 1, k)
        k = wd.shape[2]
        left = (k - 1) // 2
        xp = np.pad(h, ((0, 0), (0, 0), (left, k - 1 - left)))
        y = np.einsum("bctk,ck->bct", sliding_window_view(xp, k, axis=2), wd[:, 0], optimize=True)
        y = y + sd[p + ".depthwise_conv.bias"].reshape(1, -1, 1)
        y = np.einsum("oc,bct->bot", sd[p + ".pointwise_conv.weight"][:, :, 0], y, optimize=True)
        y = batchnorm(y + sd[p + ".pointwise_conv.bias"].reshape(1, -1, 1), sd, p + ".batch_norm")
        if p + ".residual_connector.0.weight" in sd:
            r = np.einsum("oc,bct->bot", sd[p + ".residual_connector.0.weight"][:, :, 0], h, optimize=True)
            r = batchnorm(r + sd[p + ".residual_connector.0.bias"].reshape(1, -1, 1), sd, p + ".residual_connector.1")
        else:
            r = h
        h = np.maximum(y + r, 0)
        i += 1
    return linear(h.mean(axis=2), sd[prefix + ".fc.weight"], sd[prefix + ".fc.bias"])


def raw_audio_frontend(x, sd, prefix="model.frontend"):
    """RawAudioFrontend (architectures.py:695-714): depth x [Conv1d(k = 41 / 13, stride 16 / 4, padding k // 2,
    no bias) -> BatchNorm1d -> ReLU] on (B, 1, N) float audio; returns (B, C, T)."""
    h = x[:, None, :]
    i = 0
    while f"{prefix}.conv_blocks.{3 * i}.weight" in sd:
        w = sd[f"{prefix}.conv_blocks.{3 * i}.weight"]             # (Cout, Cin, k)
        cout, cin, k = w.shape
        stride, pad = (16, 20) if i == 0 else (4, 6)
        assert k == (41 if i == 0 else 13)
        xp = np.pad(h, ((0, 0), (0, 0), (pad, pad)))
        win = sliding_window_view(xp, k, axis=2)[:, :, ::stride]   # (B, Cin, T_out, k)
        y = np.einsum("bctk,ock->bot", win, w, optimize=True)
        h = np.maximum(batchnorm(y, sd, f"{prefix}.conv_blocks.{3 * i + 1}"), 0)
        i += 1
    return h


def _e2e_quartznet(x, sd, cfg):
    """E2ERawQuartzNet (architectures.py:796-817) on float audio (B, N) already scaled by 1 / 32768
    (nanointerpreter.py:750)."""
    h = raw_audio_frontend(x, sd)                                   # (B, C, T)
    return _quartznet(np.swapaxes(h, 1, 2), sd, cfg, prefix="model.backbone")


def _e2e_melcnn_body(mel, sd, cfg):
    act = cfg.get("activation_function", "relu")
    h = mel[:, None]
    for i, pool in ((0, True), (4, True), (8, False)):
        h = conv2d(h, sd[f"model.conv_block.{i}.weight"], sd[f"model.conv_block.{i}.bias"])
        h = activation(batchnorm(h, sd, f"model.conv_block.{i + 1}"), act)
        if pool:
            h = maxpool2(h)
    h = avgpool_fixed(h, (1, 4)).reshape(h.shape[0], -1)
    h = linear(h, sd["model.fc1.weight"], sd["model.fc1.bias"])
    h = activation(batchnorm(h, sd, "model.bn1"), act)
    return linear(h, sd["model.out.weight"], sd["model.out.bias"])


_BACKBONES = {"dnn": _dnn, "cnn": _cnn, "tcn": _tcn, "bcresnet": _bcresnet, "crnn": _crnn,
              "e2e_dnn": _e2e_melcnn_body, "gru": _gru, "lstm": _lstm, "rnn": _rnn, "quartznet": _quartznet,
              "e2e_quartznet": _e2e_quartznet}

def _e2e_cnn(x, sd, cfg):
    """E2ERawCNN (architectures.py:777-793): RawAudioFrontend (depth 2) -> (B, 1, C, T) image -> RawAudioBackbone
    (:738-774): four Conv2d 3x3 (no bias; strides (1,2), (2,2), (2,2), 1) + BatchNorm + activation, global average
    pool, Linear."""
    act = cfg.get("activation_function", "relu")
    h = raw_audio_frontend(x, sd)[:, None]                          # (B, 1, C, T)
    for name, stride in (("conv1", (1, 2)), ("conv2", (2, 2)), ("conv3", (2, 2)), ("conv4", (1, 1))):
        p = f"model.backbone.{name}"
        h = activation(batchnorm(conv2d(h, sd[p + ".0.weight"], None, stride=stride), sd, p + ".1"), act)
    return linear(h.mean(axis=(2, 3)), sd["model.backbone.fc.weight"], sd["model.backbone.fc.bias"])


_BACKBONES["e2e_cnn"] = _e2e_cnn

# Heads that consume the audio itself (no log-mel front end): float samples = int16 / 32768 (nanointerpreter.py:750)
RAW_AUDIO_HEADS = ("e2e_quartznet", "e2e_cnn")


def classifier(emb, sd, cfg):
    """Linear(E, E/2) -> act -> Linear(E/2, 1)  (model.py:291-296)."""
    h = activation(linear(emb, sd["classifier.0.weight"], sd["classifier.0.bias"]),
                   cfg.get("activation_function", "relu"))
    return linear(h, sd["classifier.3.weight"], sd["classifier.3.bias"])


def head_input_from_mel(mel, model_type: str):
    """(B, F, T) log-mel -> the layout the head consumes."""
    if model_type in TIME_MAJOR_HEADS:
        return np.ascontiguousarray(np.swapaxes(mel, 1, 2))
    return mel


def _cast_sd(sd, dtype):
    return {k: np.asarray(v).astype(dtype) for k, v in sd.items() if np.asarray(v).dtype.kind == "f"}


def embedding_from_features(x, sd, cfg, dtype=np.float64):
    """Head input (already a log-mel in the head's layout) -> embedding (B, E)."""
    sd = _cast_sd(sd, dtype)
    return _BACKBONES[cfg["model_type"]](np.asarray(x, dtype=dtype), sd, cfg)


def forward_logits(pcm, sd, cfg, frontend: FrontendSpec | str | None = None, dtype=np.float64,
                   return_mel=False):
    """int16 PCM (B, N) — or float PCM already scaled by 1/32768, as the session is fed — -> logits (B, 1):
    front end + backbone + classifier."""
    if cfg["model_type"] in RAW_AUDIO_HEADS:
        sd = _cast_sd(sd, dtype)
        pcm = np.asarray(pcm)
        # int16 -> x / 32768 (nanointerpreter.py:750); float input is what the session is fed: already scaled
        x = pcm.astype(dtype) / dtype(32768.0) if pcm.dtype == np.int16 else pcm.astype(dtype)
        logits = classifier(_BACKBONES[cfg["model_type"]](x, sd, cfg), sd, cfg)
        return (logits, None) if return_mel else logits
    if frontend is None:
        frontend = "REF64x101" if cfg["model_type"] == "e2e_dnn" else "NS40x98"
    if isinstance(frontend, str):
        frontend = GEOMETRIES[frontend]
    mel = log_mel(pcm, frontend, dtype)
    sd = _cast_sd(sd, dtype)
    emb = _BACKBONES[cfg["model_type"]](head_input_from_mel(mel, cfg["model_type"]), sd, cfg)
    logits = classifier(emb, sd, cfg)
    return (logits, mel) if return_mel else logits


def forward_scores(pcm, sd, cfg, frontend=None, dtype=np.float64, return_mel=False):
    """As deployed: sigmoid(logits).view(-1, 1, 1)  (_export/onnx.py:169-172)."""
    out = forward_logits(pcm, sd, cfg, frontend, dtype, return_mel)
    logits, mel = out if return_mel else (out, None)
    scores = sigmoid(logits).reshape(-1, 1, 1)
    return (scores, mel) if return_mel else scores
