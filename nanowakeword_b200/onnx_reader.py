"""``.onnx`` ingestion without the ``onnx`` package.

The reference's trainer always writes ``<name>.onnx`` (and ``<name>_lite.onnx``) but a ``.pt`` only for the main
model and never a description of the architecture (reference nanowakeword/trainer.py:474-535,
_export/onnx.py:157-221).  This module reads the ONNX file itself:

* :func:`parse_onnx` — a minimal protobuf wire-format reader for ``ModelProto``: graph nodes (op type, inputs,
  outputs, attributes), initializers, the graph input's shape and ``metadata_props`` (``mode = e2e`` is what
  _export/onnx.py:212-221 records and what the interpreter looks for, nanointerpreter.py:968-992);
* :func:`onnx_to_artifacts` — pattern-matches the graph STRUCTURE (operator sequence and weight shapes — initializer
  names differ between exporter versions, and BatchNorm may arrive folded into the preceding Conv / Gemm or as an
  explicit ``BatchNormalization`` node) onto the e2e architectures the engine builds — ``E2E_MelSpectrogram_CNN``,
  ``E2ERawCNN``, ``E2ERawQuartzNet`` (architectures.py:777-888) — and returns ``(state_dict, cfg)`` keyed like the
  reference's own ``state_dict`` (folded BatchNorms come back as identity statistics), i.e. exactly what
  ``weights.pack_tensors`` consumes for a ``.pt`` + sidecar.

Graphs of embedding-mode heads (input ``(B, 16, 96)`` features) are recognised and refused: their front end lives in
downloaded binaries outside the hot path.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------------- protobuf wire format


def _varint(buf: bytes, pos: int):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) of one message; length-delimited values are memoryview slices."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _packed_varints(v, wt):
    if wt == 0:
        return [v]
    out, pos = [], 0
    b = bytes(v)
    while pos < len(b):
        x, pos = _varint(b, pos)
        out.append(x)
    return out


def _signed(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}


def _tensor(buf: bytes):
    dims, dtype, name, raw = [], 1, "", None
    floats, int64s, int32s, doubles = [], [], [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += [_signed(x) for x in _packed_varints(v, wt)]
        elif fno == 2:
            dtype = v
        elif fno == 4:
            floats.append(np.frombuffer(bytes(v), "<f4"))
        elif fno == 5:
            int32s += [_signed(x) for x in _packed_varints(v, wt)]
        elif fno == 7:
            int64s += [_signed(x) for x in _packed_varints(v, wt)]
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 10:
            doubles.append(np.frombuffer(bytes(v), "<f8"))
    if dtype not in _DTYPES:
        raise ValueError(f"initializer '{name}': unsupported ONNX data type {dtype}")
    np_dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dt).newbyteorder("<")).astype(np_dt)
    elif floats:
        arr = np.concatenate(floats).astype(np_dt)
    elif doubles:
        arr = np.concatenate(doubles).astype(np_dt)
    elif int64s:
        arr = np.asarray(int64s, dtype=np_dt)
    else:
        arr = np.asarray(int32s, dtype=np_dt)
    return name, arr.reshape(dims) if dims else arr.reshape(())


@dataclass
class Node:
    op: str
    inputs: list
    outputs: list
    name: str = ""
    attrs: dict = field(default_factory=dict)


def _attribute(buf: bytes):
    name, val = "", None
    ints, floats = [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            val = struct.unpack("<f", bytes(v))[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = bytes(v)
        elif fno == 5:
            val = _tensor(bytes(v))[1]
        elif fno == 7:
            floats += list(np.frombuffer(bytes(v), "<f4")) if wt == 2 else [struct.unpack("<f", bytes(v))[0]]
        elif fno == 8:
            ints += [_signed(x) for x in _packed_varints(v, wt)]
    if ints:
        val = ints
    elif floats:
        val = floats
    return name, val


def _node(buf: bytes) -> Node:
    n = Node("", [], [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(v).decode())
        elif fno == 2:
            n.outputs.append(bytes(v).decode())
        elif fno == 3:
            n.name = bytes(v).decode()
        elif fno == 4:
            n.op = bytes(v).decode()
        elif fno == 5:
            k, a = _attribute(bytes(v))
            n.attrs[k] = a
    return n


def _value_info(buf: bytes):
    name, shape = "", None
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:                                            # TypeProto
            for f2, _, v2 in _fields(bytes(v)):
                if f2 == 1:                                       # tensor_type
                    for f3, _, v3 in _fields(bytes(v2)):
                        if f3 == 2:                               # shape
                            shape = []
                            for f4, _, v4 in _fields(bytes(v3)):
                                if f4 == 1:                       # dim
                                    d = None
                                    for f5, w5, v5 in _fields(bytes(v4)):
                                        if f5 == 1:
                                            d = _signed(v5)
                                        elif f5 == 2:
                                            d = bytes(v5).decode()
                                    shape.append(d)
    return name, shape


@dataclass
class OnnxModel:
    nodes: list
    initializers: dict
    inputs: list            # [(name, shape)] graph inputs that are not initializers
    outputs: list
    metadata: dict
    producer: str = ""
    opset: int = 0


def parse_onnx(path: str) -> OnnxModel:
    buf = open(path, "rb").read()
    graph, meta, producer, opset = None, {}, "", 0
    try:
        for fno, wt, v in _fields(buf):
            if fno == 7:
                graph = bytes(v)
            elif fno == 2 and wt == 2:
                producer = bytes(v).decode(errors="replace")
            elif fno == 8:
                for f2, _, v2 in _fields(bytes(v)):
                    if f2 == 2:
                        opset = max(opset, v2)
            elif fno == 14:
                k = val = ""
                for f2, _, v2 in _fields(bytes(v)):
                    if f2 == 1:
                        k = bytes(v2).decode()
                    elif f2 == 2:
                        val = bytes(v2).decode()
                meta[k] = val
    except (IndexError, struct.error) as ex:
        raise ValueError(f"{path}: not a readable ONNX ModelProto ({ex})") from ex
    if graph is None:
        raise ValueError(f"{path}: no GraphProto found (is this an ONNX file?)")
    nodes, inits, inputs, outputs = [], {}, [], []
    for fno, wt, v in _fields(graph):
        if fno == 1:
            nodes.append(_node(bytes(v)))
        elif fno == 5:
            name, arr = _tensor(bytes(v))
            inits[name] = arr
        elif fno == 11:
            inputs.append(_value_info(bytes(v)))
        elif fno == 12:
            outputs.append(_value_info(bytes(v)))
    # Constant nodes are initializers in disguise
    for n in nodes:
        if n.op == "Constant" and "value" in n.attrs and isinstance(n.attrs["value"], np.ndarray):
            inits[n.outputs[0]] = n.attrs["value"]
    inputs = [(nm, sh) for nm, sh in inputs if nm not in inits]
    return OnnxModel(nodes, inits, inputs, outputs, meta, producer, opset)


# --------------------------------------------------------------------------------------------- structure matching


@dataclass
class _Layer:
    kind: str                 # conv | linear | act | maxpool | avgpool | gap | add
    w: np.ndarray = None
    b: np.ndarray = None
    attrs: dict = field(default_factory=dict)
    src: str = ""             # tensor this layer reads (first data input)
    out: str = ""
    name: str = ""            # act name


def _fold_bn_into(layer: _Layer, scale, bias, mean, var, eps):
    s = scale.astype(np.float64) / np.sqrt(var.astype(np.float64) + eps)
    w = layer.w.astype(np.float64) * s.reshape((-1,) + (1,) * (layer.w.ndim - 1))
    b0 = np.zeros(layer.w.shape[0]) if layer.b is None else layer.b.astype(np.float64)
    layer.w, layer.b = w, (b0 - mean.astype(np.float64)) * s + bias.astype(np.float64)


def _layers(m: OnnxModel):
    """The weighted / structural operators of the graph in execution order, with BatchNorm folded into its producer
    and shape-only operators skipped (their outputs alias their inputs for the purpose of following the data)."""
    alias = {}

    def res(t):
        while t in alias:
            t = alias[t]
        return t

    layers, by_out = [], {}
    const = m.initializers
    passthrough = {"Reshape", "Transpose", "Unsqueeze", "Squeeze", "Flatten", "Identity", "Dropout", "Cast", "Pad", "Slice"}
    for n in m.nodes:
        ins = [res(t) for t in n.inputs]
        data_ins = [t for t in ins if t and t not in const]
        if n.op == "Constant":
            continue
        if n.op in passthrough:
            if data_ins:
                alias[n.outputs[0]] = data_ins[0]
            if n.op == "Pad":
                by_out[n.outputs[0]] = ("pad", n.attrs.get("mode", b"constant"))
            continue
        L = None
        if n.op == "Conv":
            w = const[n.inputs[1]]
            b = const[n.inputs[2]] if len(n.inputs) > 2 and n.inputs[2] else None
            L = _Layer("conv", w.astype(np.float64), None if b is None else b.astype(np.float64),
                       dict(strides=n.attrs.get("strides"), pads=n.attrs.get("pads"), group=n.attrs.get("group", 1),
                            dilations=n.attrs.get("dilations")))
        elif n.op == "Gemm":
            w = const[n.inputs[1]].astype(np.float64)
            if not n.attrs.get("transB", 0):
                w = w.T
            b = const[n.inputs[2]].astype(np.float64) if len(n.inputs) > 2 and n.inputs[2] else None
            L = _Layer("linear", np.ascontiguousarray(w) * float(n.attrs.get("alpha", 1.0)), b)
        elif n.op == "MatMul" and n.inputs[1] in const:
            L = _Layer("linear", np.ascontiguousarray(const[n.inputs[1]].astype(np.float64).T), None)
        elif n.op == "Add" and len(data_ins) == 1 and layers and layers[-1].kind == "linear" and layers[-1].b is None \
                and layers[-1].out == data_ins[0]:
            other = [t for t in n.inputs if t in const]
            layers[-1].b = const[other[0]].astype(np.float64).ravel()
            alias[n.outputs[0]] = layers[-1].out
            continue
        elif n.op == "BatchNormalization":
            prod = by_out.get(data_ins[0])
            sc, bi, mu, var = (const[t] for t in n.inputs[1:5])
            if isinstance(prod, _Layer) and prod.kind in ("conv", "linear"):
                _fold_bn_into(prod, sc, bi, mu, var, float(n.attrs.get("epsilon", 1e-5)))
                alias[n.outputs[0]] = prod.out
                continue
            raise NotImplementedError("BatchNormalization that does not follow a Conv / Gemm")
        elif n.op in ("Relu", "Gelu"):
            L = _Layer("act", name=n.op.lower())
        elif n.op == "Erf":
            L = _Layer("act", name="gelu")                     # x * 0.5 * (1 + erf(x / sqrt 2)): the exact-erf GELU
        elif n.op == "Sigmoid":
            L = _Layer("act", name="sigmoid")
        elif n.op == "MaxPool":
            L = _Layer("maxpool", attrs=dict(kernel=n.attrs.get("kernel_shape"), strides=n.attrs.get("strides")))
        elif n.op == "AveragePool":
            L = _Layer("avgpool", attrs=dict(kernel=n.attrs.get("kernel_shape"), strides=n.attrs.get("strides")))
        elif n.op in ("GlobalAveragePool", "ReduceMean"):
            L = _Layer("gap")
        elif n.op == "Add" and len(data_ins) == 2:
            L = _Layer("add")
        elif n.op in ("LSTM", "GRU", "LayerNormalization"):
            L = _Layer(n.op.lower())
        if L is None:                                          # arithmetic glue (Mul, Pow, Log, Clip, Div, Sqrt ...): follow the data
            if data_ins:
                alias[n.outputs[0]] = data_ins[0]
            continue
        L.src = data_ins[0] if data_ins else ""
        L.out = n.outputs[0]
        if L.kind == "add":
            L.attrs["ins"] = data_ins
        by_out[L.out] = L
        layers.append(L)
    return layers


def _identity_bn(sd, prefix, n):
    sd[prefix + ".weight"] = np.ones(n, np.float32)
    sd[prefix + ".bias"] = np.zeros(n, np.float32)
    sd[prefix + ".running_mean"] = np.zeros(n, np.float32)
    sd[prefix + ".running_var"] = np.full(n, 1.0 - 1e-5, np.float64)     # + eps = 1: the packer's fold is the identity


def _activation(layers):
    """The model's activation function: the first non-sigmoid activation after the conv stack decides; a Sigmoid that
    feeds a Mul is SiLU (x * sigmoid(x)) — only the final Sigmoid is the probability (_export/onnx.py:169-172)."""
    names = [L.name for L in layers if L.kind == "act"]
    body = names[:-1] if names and names[-1] == "sigmoid" else names
    # the raw front end always uses ReLU (architectures.py:704): look at the activations that follow it
    for nm in reversed(body):
        if nm == "gelu":
            return "gelu"
        if nm == "sigmoid":
            return "silu"
    return "relu"


def onnx_to_artifacts(m: OnnxModel):
    """(state_dict, cfg) of an e2e graph exported by the reference (see the module docstring)."""
    if not m.inputs:
        raise ValueError("ONNX graph has no input")
    in_name, in_shape = m.inputs[0]
    dims = [d for d in (in_shape or [])]
    is_e2e = m.metadata.get("mode") == "e2e"
    if not is_e2e:                                             # the interpreter's own fallback (nanointerpreter.py:984-992)
        nd = len(dims)
        is_e2e = nd <= 2 or (nd == 3 and dims[-1] not in (32, 64, 96))
    if not is_e2e:
        raise NotImplementedError(
            f"this ONNX graph takes embedding features {dims} (embedding mode); its mel / embedding front end lives in "
            "downloaded binaries (interpreter/models/_registry.py:34-47) and is outside the B200 hot path — only e2e "
            "models (metadata mode=e2e: e2e_dnn, e2e_cnn, e2e_quartznet) can be served")
    clip = dims[-1]
    if not isinstance(clip, int) or clip != 16000:
        raise NotImplementedError(f"e2e clip length {clip}: the engine is built for clip_samples = 16000")
    layers = _layers(m)
    convs = [L for L in layers if L.kind == "conv"]
    lins = [L for L in layers if L.kind == "linear"]
    if not convs or len(lins) < 3:
        raise NotImplementedError("unrecognised graph: expected a convolutional e2e model followed by dense layers")
    act = _activation(layers)
    sd: dict = {}
    cfg = {"mode": "e2e", "input_shape": [clip], "activation_function": act, "n_blocks": 1, "input_ndim": len(dims)}

    def put_linear(prefix, L):
        sd[prefix + ".weight"] = L.w.astype(np.float32)
        sd[prefix + ".bias"] = (L.b if L.b is not None else np.zeros(L.w.shape[0])).astype(np.float32)

    # classifier = the last two linears (model.py:291-296): Linear(E, E/2) -> act -> Linear(E/2, 1)
    put_linear("classifier.0", lins[-2])
    put_linear("classifier.3", lins[-1])
    if lins[-1].w.shape[0] != 1:
        raise NotImplementedError("only single-class (n_classes = 1) models are supported")
    cfg["embedding_dim"] = int(lins[-2].w.shape[1])
    body_lins = lins[:-2]

    c0 = convs[0]
    if c0.w.ndim == 3 and c0.w.shape[1] == 1 and c0.w.shape[2] == 400 and c0.w.shape[0] == 201 and len(convs) >= 5 \
            and convs[1].w.shape == c0.w.shape:
        # ---- E2E_MelSpectrogram_CNN: conv-DFT (real, imag bases: _export/onnx.py:42-63), mel matmul, 3 x Conv2d, fc1, out
        if list(c0.attrs.get("strides") or []) != [160]:
            raise NotImplementedError("e2e_dnn: STFT hop is not 160")
        mel = body_lins[0]
        if mel.w.shape != (64, 201):
            raise NotImplementedError(f"e2e_dnn: mel filterbank of shape {mel.w.T.shape}, expected (201, 64)")
        # real_basis[0] = cos(0) * window: the Hann window itself (the k = 0 row of _export/onnx.py:55)
        sd["model.mel_spec.spectrogram.window"] = c0.w[0, 0, :].astype(np.float32)
        sd["model.mel_spec.mel_scale.fb"] = np.ascontiguousarray(mel.w.T).astype(np.float32)
        c2d = convs[2:5]
        if [tuple(c.w.shape[1:]) for c in c2d] != [(1, 3, 3), (16, 3, 3), (32, 3, 3)] or len(body_lins) != 3:
            raise NotImplementedError("e2e_dnn: unexpected Conv2d stack / dense layers")
        for j, (c, i) in enumerate(zip(c2d, (0, 4, 8))):
            sd[f"model.conv_block.{i}.weight"] = c.w.astype(np.float32)
            sd[f"model.conv_block.{i}.bias"] = (c.b if c.b is not None else np.zeros(c.w.shape[0])).astype(np.float32)
            _identity_bn(sd, f"model.conv_block.{i + 1}", c.w.shape[0])
        pools = [L for L in layers if L.kind == "avgpool"]
        if not pools or list(pools[-1].attrs["kernel"]) != [16, 7] or list(pools[-1].attrs["strides"]) != [16, 6]:
            # AdaptiveAvgPool2d((1, 4)) on (16, 25) as the exporter rewrites it (_export/onnx.py:96-154)
            raise NotImplementedError("e2e_dnn: expected AvgPool2d(kernel (16, 7), stride (16, 6)) after the conv stack")
        put_linear("model.fc1", body_lins[1])
        _identity_bn(sd, "model.bn1", body_lins[1].w.shape[0])
        put_linear("model.out", body_lins[2])
        cfg["model_type"] = "e2e_dnn"
        return sd, cfg

    if c0.w.ndim == 3 and c0.w.shape[1] == 1 and c0.w.shape[2] == 41:
        # ---- RawAudioFrontend (architectures.py:695-714): Conv1d k 41 / s 16, then k 13 / s 4 layers, no bias, BN, ReLU
        depth = 1
        while depth < len(convs) and convs[depth].w.ndim == 3 and convs[depth].w.shape[2] == 13 and \
                list(convs[depth].attrs.get("strides") or []) == [4] and convs[depth].attrs.get("group", 1) == 1:
            depth += 1
        for i in range(depth):
            c = convs[i]
            want = [16] if i == 0 else [4]
            if list(c.attrs.get("strides") or []) != want:
                raise NotImplementedError("raw front end: unexpected stride")
            sd[f"model.frontend.conv_blocks.{3 * i}.weight"] = c.w.astype(np.float32)
            # the reference's layer has no bias: a folded BatchNorm leaves its shift in the Conv bias -> carry it as BN bias
            _identity_bn(sd, f"model.frontend.conv_blocks.{3 * i + 1}", c.w.shape[0])
            if c.b is not None:
                sd[f"model.frontend.conv_blocks.{3 * i + 1}.bias"] = c.b.astype(np.float32)
        cfg["e2e_frontend_channels"] = int(convs[0].w.shape[0])
        cfg["e2e_frontend_depth"] = depth
        rest = convs[depth:]
        if rest and rest[0].w.ndim == 4:
            # ---- E2ERawCNN: RawAudioBackbone (architectures.py:738-774), four 3x3 Conv2d without bias + BN + act, fc
            if len(rest) != 4 or len(body_lins) != 1 or depth != 2:
                raise NotImplementedError("e2e_cnn: expected frontend depth 2, four Conv2d layers and one fc")
            for j, c in enumerate(rest):
                sd[f"model.backbone.conv{j + 1}.0.weight"] = c.w.astype(np.float32)
                _identity_bn(sd, f"model.backbone.conv{j + 1}.1", c.w.shape[0])
                if c.b is not None:
                    sd[f"model.backbone.conv{j + 1}.1.bias"] = c.b.astype(np.float32)
            put_linear("model.backbone.fc", body_lins[0])
            cfg["model_type"] = "e2e_cnn"
            return sd, cfg
        # ---- E2ERawQuartzNet: QuartzNetModel blocks (architectures.py:366-437): depthwise, pointwise (+BN), [residual 1x1 (+BN)]
        if len(body_lins) != 1 or depth != 3:
            raise NotImplementedError("e2e_quartznet: expected frontend depth 3 and one fc layer")
        i, blk, qcfg = 0, 0, []
        while i < len(rest):
            dw = rest[i]
            if dw.attrs.get("group", 1) != dw.w.shape[0] or dw.w.shape[1] != 1:
                raise NotImplementedError("e2e_quartznet: expected a depthwise Conv1d at the start of a block")
            pw = rest[i + 1]
            p = f"model.backbone.quartznet_blocks.{blk}"
            sd[p + ".depthwise_conv.weight"] = dw.w.astype(np.float32)
            sd[p + ".depthwise_conv.bias"] = (dw.b if dw.b is not None else np.zeros(dw.w.shape[0])).astype(np.float32)
            sd[p + ".pointwise_conv.weight"] = pw.w.astype(np.float32)
            sd[p + ".pointwise_conv.bias"] = (pw.b if pw.b is not None else np.zeros(pw.w.shape[0])).astype(np.float32)
            _identity_bn(sd, p + ".batch_norm", pw.w.shape[0])
            i += 2
            cin, cout = dw.w.shape[0], pw.w.shape[0]
            if i < len(rest) and rest[i].attrs.get("group", 1) == 1 and rest[i].w.shape[2] == 1 and rest[i].w.shape[1] == cin \
                    and rest[i].w.shape[0] == cout and (cin != cout):
                rc = rest[i]
                sd[p + ".residual_connector.0.weight"] = rc.w.astype(np.float32)
                sd[p + ".residual_connector.0.bias"] = (rc.b if rc.b is not None else np.zeros(cout)).astype(np.float32)
                _identity_bn(sd, p + ".residual_connector.1", cout)
                i += 1
            qcfg.append([int(cout), int(dw.w.shape[2]), 1])
            blk += 1
        put_linear("model.backbone.fc", body_lins[0])
        cfg["model_type"] = "e2e_quartznet"
        cfg["e2e_quartznet_config"] = qcfg
        return sd, cfg

    raise NotImplementedError(
        "unrecognised e2e graph: the B200 engine builds e2e_dnn (E2E_MelSpectrogram_CNN), e2e_cnn (E2ERawCNN) and "
        "e2e_quartznet (E2ERawQuartzNet)")


def load_onnx(path: str):
    """``path`` -> (state_dict as numpy, cfg)."""
    return onnx_to_artifacts(parse_onnx(path))
