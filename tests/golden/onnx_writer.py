"""Hand-serialised ONNX fixtures for the ``.onnx`` ingestion tests.  TEST INFRASTRUCTURE ONLY.

Neither ``onnx`` nor ``onnxscript`` / ``onnxruntime`` is installable offline, so ``torch.onnx.export`` cannot run in the
build container.  This module writes the ``ModelProto`` the reference's exporter produces
(nanowakeword/_export/onnx.py:157-221: ``InferenceWrapper(model)`` = sigmoid(logits).view(-1, 1, 1), input ``input``
``(batch_size, 1, 16000)``, output ``output``, ``metadata_props mode = e2e``) for the three e2e architectures, straight
from a reference-keyed ``state_dict``, with a minimal protobuf encoder.  Two styles, because exporter versions differ:

* ``folded``   — BatchNorm fused into the preceding Conv (what the TorchScript exporter does in eval mode), anonymous
                 initializer names (``onnx::Conv_123``), Linear as ``Gemm(transB = 1)``;
* ``explicit`` — ``BatchNormalization`` nodes kept, parameter names kept, Linear as ``MatMul`` + ``Add``.

The graphs carry the operators the reader has to see through (Unsqueeze, Pad(reflect), Transpose, Pow, Log, Mul, Clip,
Flatten, ReduceMean, Reshape ...); they are not meant to be executed.
"""
from __future__ import annotations

import math
import struct

import numpy as np

BN_EPS = 1e-5


# --------------------------------------------------------------------------------------------- protobuf encoder
def _vi(x: int) -> bytes:
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(fno: int, payload: bytes) -> bytes:
    return _vi(fno << 3 | 2) + _vi(len(payload)) + payload


def _int(fno: int, v: int) -> bytes:
    return _vi(fno << 3) + _vi(v)


def _str(fno: int, s: str) -> bytes:
    return _ld(fno, s.encode())


def tensor(name: str, arr: np.ndarray, raw=True) -> bytes:
    arr = np.asarray(arr)
    dt = {np.dtype("float32"): 1, np.dtype("int64"): 7}[arr.dtype]
    out = b"".join(_int(1, d) for d in arr.shape) + _int(2, dt)
    if raw:
        out += _ld(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    elif dt == 1:
        out += _ld(4, arr.astype("<f4").tobytes())                       # packed float_data
    else:
        out += _ld(7, b"".join(_vi(int(v)) for v in arr.ravel()))
    return out + _str(8, name)


def attr(name: str, v) -> bytes:
    out = _str(1, name)
    if isinstance(v, float):
        out += _vi(2 << 3 | 5) + struct.pack("<f", v) + _int(20, 1)
    elif isinstance(v, int):
        out += _int(3, v) + _int(20, 2)
    elif isinstance(v, (bytes, str)):
        out += _ld(4, v.encode() if isinstance(v, str) else v) + _int(20, 3)
    elif isinstance(v, (list, tuple)):
        out += _ld(8, b"".join(_vi(int(x)) for x in v)) + _int(20, 7)    # packed ints
    else:
        raise TypeError(type(v))
    return out


def node(op: str, ins, outs, name="", **attrs) -> bytes:
    out = b"".join(_str(1, i) for i in ins) + b"".join(_str(2, o) for o in outs)
    if name:
        out += _str(3, name)
    out += _str(4, op)
    for k, v in attrs.items():
        out += _ld(5, attr(k, v))
    return out


def value_info(name: str, shape) -> bytes:
    dims = b""
    for d in shape:
        dims += _ld(1, _str(2, d) if isinstance(d, str) else _int(1, d))
    ttype = _int(1, 1) + _ld(2, dims)
    return _str(1, name) + _ld(2, _ld(1, ttype))


class Graph:
    def __init__(self, style: str):
        self.style, self.nodes, self.inits, self.n = style, [], [], 0

    def tmp(self, hint="t"):
        self.n += 1
        return f"/{hint}_{self.n}"

    def init(self, name, arr, anonymous_kind=None):
        if self.style == "folded" and anonymous_kind:
            self.n += 1
            name = f"onnx::{anonymous_kind}_{self.n}"
        self.inits.append(tensor(name, np.ascontiguousarray(arr, dtype=np.float32 if np.asarray(arr).dtype.kind == "f" else np.int64),
                                 raw=(self.n % 2 == 0)))
        return name

    def op(self, op, ins, hint=None, **attrs):
        out = self.tmp(hint or op)
        self.nodes.append(node(op, ins, [out], name=out, **attrs))
        return out

    def serialize(self, in_shape, e2e=True) -> bytes:
        g = b"".join(_ld(1, n) for n in self.nodes) + _str(2, "main_graph") + b"".join(_ld(5, t) for t in self.inits)
        g += _ld(11, value_info("input", in_shape)) + _ld(12, value_info("output", ["batch_size", 1, 1]))
        m = _int(1, 8) + _str(2, "pytorch") + _str(3, "2.8.0") + _ld(7, g) + _ld(8, _str(1, "") + _int(2, 17))
        if e2e:
            m += _ld(14, _str(1, "mode") + _str(2, "e2e"))
        return m


# --------------------------------------------------------------------------------------------- building blocks
def _bn_fold(w, b, sd, p):
    s = sd[p + ".weight"].astype(np.float64) / np.sqrt(sd[p + ".running_var"].astype(np.float64) + BN_EPS)
    w2 = w.astype(np.float64) * s.reshape((-1,) + (1,) * (w.ndim - 1))
    b0 = np.zeros(w.shape[0]) if b is None else b.astype(np.float64)
    return w2, (b0 - sd[p + ".running_mean"].astype(np.float64)) * s + sd[p + ".bias"].astype(np.float64)


def conv_bn(g: Graph, x, sd, conv, bn, **attrs):
    w = sd[conv + ".weight"]
    b = sd.get(conv + ".bias")
    if g.style == "folded" and bn:
        w, b = _bn_fold(w, b, sd, bn)
        return g.op("Conv", [x, g.init(conv + ".weight", w, "Conv"), g.init(conv + ".bias", b, "Conv")], **attrs)
    ins = [x, g.init("trained_model." + conv + ".weight", w)]
    if b is not None:
        ins.append(g.init("trained_model." + conv + ".bias", b))
    y = g.op("Conv", ins, **attrs)
    if bn:
        y = g.op("BatchNormalization", [y] + [g.init(f"trained_model.{bn}.{k}", sd[f"{bn}.{k}"])
                                               for k in ("weight", "bias", "running_mean", "running_var")],
                 epsilon=float(BN_EPS), momentum=0.9)
    return y


def linear(g: Graph, x, sd, name, bn=None):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if g.style == "folded":
        y = g.op("Gemm", [x, g.init("trained_model." + name + ".weight", w), g.init("trained_model." + name + ".bias", b)],
                 alpha=1.0, beta=1.0, transB=1)
    else:
        y = g.op("MatMul", [x, g.init(name + ".weight_t", w.T, "MatMul")])
        y = g.op("Add", [g.init("trained_model." + name + ".bias", b), y])
    if bn:                                        # BatchNorm1d after a Linear is not fused by the exporter
        y = g.op("BatchNormalization", [y] + [g.init(f"trained_model.{bn}.{k}", sd[f"{bn}.{k}"])
                                               for k in ("weight", "bias", "running_mean", "running_var")],
                 epsilon=float(BN_EPS), momentum=0.9)
    return y


def activation(g: Graph, x, act):
    if act == "relu":
        return g.op("Relu", [x])
    if act == "silu":
        return g.op("Mul", [x, g.op("Sigmoid", [x])])
    # exact-erf GELU as the exporter decomposes it below opset 20
    e = g.op("Erf", [g.op("Div", [x, g.init("c_sqrt2", np.array(math.sqrt(2.0), np.float32), "Constant")])])
    return g.op("Mul", [g.op("Mul", [x, g.op("Add", [e, g.init("c_one", np.array(1.0, np.float32), "Constant")])]),
                        g.init("c_half", np.array(0.5, np.float32), "Constant")])


def head(g: Graph, emb, sd, act):
    h = activation(g, linear(g, emb, sd, "classifier.0"), act)
    logits = linear(g, h, sd, "classifier.3")
    p = g.op("Sigmoid", [logits])
    out = "output"
    g.nodes.append(node("Reshape", [p, g.init("c_shape", np.array([-1, 1, 1], np.int64), "Constant")], [out], name="/Reshape_out"))


def raw_frontend(g: Graph, x, sd, depth):
    for i in range(depth):
        k, s = (41, 16) if i == 0 else (13, 4)
        x = conv_bn(g, x, sd, f"model.frontend.conv_blocks.{3 * i}", f"model.frontend.conv_blocks.{3 * i + 1}",
                    kernel_shape=[k], strides=[s], pads=[k // 2, k // 2], dilations=[1], group=1)
        x = g.op("Relu", [x])
    return x


# --------------------------------------------------------------------------------------------- the three e2e models
def write_e2e_model(path: str, sd: dict, cfg: dict, style: str = "folded", input_ndim: int = 3, e2e_metadata: bool = True):
    sd = {k: np.asarray(v) for k, v in sd.items()}
    mt, act = cfg["model_type"], cfg.get("activation_function", "relu").lower()
    g = Graph(style)
    x = "input"
    if mt == "e2e_dnn":
        # ONNXSafeMelSpectrogram (_export/onnx.py:27-83) + AmplitudeToDB, then the Conv2d stack
        from nanowakeword_b200.weights import frontend_tables     # the torchaudio tables when the state_dict has none
        win, fb = frontend_tables("REF64x101", sd)
        win = win.astype(np.float64)
        n_fft = 400
        ang = -2.0 * np.pi * np.outer(np.arange(n_fft // 2 + 1), np.arange(n_fft)) / n_fft
        real = (np.cos(ang) * win[None, :])[:, None, :].astype(np.float32)
        imag = (np.sin(ang) * win[None, :])[:, None, :].astype(np.float32)
        if input_ndim == 2:
            x = g.op("Unsqueeze", [x, g.init("c_axes1", np.array([1], np.int64), "Constant")])
        x = g.op("Pad", [x, g.init("c_pads", np.array([0, 0, 200, 0, 0, 200], np.int64), "Constant")], mode="reflect")
        re = g.op("Conv", [x, g.init("trained_model.model.mel_spec.real_basis", real)], kernel_shape=[400], strides=[160], pads=[0, 0], dilations=[1], group=1)
        im = g.op("Conv", [x, g.init("trained_model.model.mel_spec.imag_basis", imag)], kernel_shape=[400], strides=[160], pads=[0, 0], dilations=[1], group=1)
        two = g.init("c_two", np.array(2.0, np.float32), "Constant")
        p = g.op("Add", [g.op("Pow", [re, two]), g.op("Pow", [im, two])])
        p = g.op("Transpose", [p], perm=[0, 2, 1])
        m = g.op("MatMul", [p, g.init("trained_model.model.mel_spec.mel_fb", fb)])
        m = g.op("Transpose", [m], perm=[0, 2, 1])
        m = g.op("Clip", [m, g.init("c_amin", np.array(1e-10, np.float32), "Constant")])
        db = g.op("Mul", [g.op("Log", [m]), g.init("c_db", np.array(10.0 / math.log(10.0), np.float32), "Constant")])
        x = g.op("Unsqueeze", [db, g.init("c_axes1b", np.array([1], np.int64), "Constant")])
        for j, i in enumerate((0, 4, 8)):
            x = conv_bn(g, x, sd, f"model.conv_block.{i}", f"model.conv_block.{i + 1}", kernel_shape=[3, 3], strides=[1, 1],
                        pads=[1, 1, 1, 1], dilations=[1, 1], group=1)
            x = activation(g, x, act)
            if j < 2:
                x = g.op("MaxPool", [x], kernel_shape=[2, 2], strides=[2, 2], pads=[0, 0, 0, 0])
        x = g.op("AveragePool", [x], kernel_shape=[16, 7], strides=[16, 6], pads=[0, 0, 0, 0])     # _export/onnx.py:96-154
        x = g.op("Flatten", [x], axis=1)
        x = activation(g, linear(g, x, sd, "model.fc1", bn="model.bn1"), act)
        emb = linear(g, x, sd, "model.out")
    elif mt in ("e2e_cnn", "e2e_quartznet"):
        if input_ndim == 2:
            x = g.op("Unsqueeze", [x, g.init("c_axes1", np.array([1], np.int64), "Constant")])
        depth = cfg.get("e2e_frontend_depth", 2 if mt == "e2e_cnn" else 3)
        x = raw_frontend(g, x, sd, depth)
        if mt == "e2e_cnn":
            x = g.op("Unsqueeze", [x, g.init("c_axes1b", np.array([1], np.int64), "Constant")])
            for j, st in enumerate(([1, 2], [2, 2], [2, 2], [1, 1])):
                x = conv_bn(g, x, sd, f"model.backbone.conv{j + 1}.0", f"model.backbone.conv{j + 1}.1", kernel_shape=[3, 3],
                            strides=st, pads=[1, 1, 1, 1], dilations=[1, 1], group=1)
                x = activation(g, x, act)
            x = g.op("GlobalAveragePool", [x])
            x = g.op("Reshape", [x, g.init("c_flat", np.array([0, -1], np.int64), "Constant")])
            emb = linear(g, x, sd, "model.backbone.fc")
        else:
            x = g.op("Transpose", [g.op("Transpose", [x], perm=[0, 2, 1])], perm=[0, 2, 1])    # permute, then QuartzNetModel permutes back
            i = 0
            while f"model.backbone.quartznet_blocks.{i}.depthwise_conv.weight" in sd:
                p = f"model.backbone.quartznet_blocks.{i}"
                c, _, k = sd[p + ".depthwise_conv.weight"].shape
                y = conv_bn(g, x, sd, p + ".depthwise_conv", None, kernel_shape=[k], strides=[1], pads=[k // 2, k // 2],
                            dilations=[1], group=int(c))
                y = conv_bn(g, y, sd, p + ".pointwise_conv", p + ".batch_norm", kernel_shape=[1], strides=[1], pads=[0, 0],
                            dilations=[1], group=1)
                res = x
                if p + ".residual_connector.0.weight" in sd:
                    res = conv_bn(g, x, sd, p + ".residual_connector.0", p + ".residual_connector.1", kernel_shape=[1],
                                  strides=[1], pads=[0, 0], dilations=[1], group=1)
                x = g.op("Relu", [g.op("Add", [y, res])])
                i += 1
            x = g.op("ReduceMean", [x], axes=[2], keepdims=0)
            emb = linear(g, x, sd, "model.backbone.fc")
    else:
        raise ValueError(mt)
    head(g, emb, sd, act)
    shape = ["batch_size", 1, 16000] if input_ndim == 3 else ["batch_size", 16000]
    with open(path, "wb") as f:
        f.write(g.serialize(shape, e2e=e2e_metadata))
    return path


def write_feature_head(path: str, n_frames=16, n_feat=96):
    """A stand-in for an embedding-mode head (input (B, 16, 96)): one Flatten + Gemm + Sigmoid."""
    g = Graph("folded")
    x = g.op("Flatten", ["input"], axis=1)
    sd = {"l.weight": np.zeros((1, n_frames * n_feat), np.float32), "l.bias": np.zeros(1, np.float32)}
    y = g.op("Sigmoid", [linear(g, x, sd, "l")])
    g.nodes.append(node("Reshape", [y, g.init("c_shape", np.array([-1, 1, 1], np.int64), "Constant")], ["output"]))
    with open(path, "wb") as f:
        f.write(g.serialize(["batch_size", n_frames, n_feat], e2e=False))
    return path
