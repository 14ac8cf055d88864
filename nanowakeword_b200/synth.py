"""Seeded random-init weights in the reference's ``state_dict`` layout.

No trained nanowakeword model is reachable offline, so parity tests,
golden vectors and the benchmark all use weights drawn here from
``numpy.random.default_rng(seed)`` — reproducible on any box without shipping megabytes
of fixtures.  Key names and shapes are exactly those of
``nanowakeword.modules.model.Model(...).state_dict()`` (reference modules/model.py:67-296,
modules/architectures.py) for the default hyper-parameters; ``tests/golden/make_golden.py``
loads them into the reference with ``load_state_dict(strict=True)`` to prove it.

BatchNorm running statistics and affine terms are randomised (not the identity a fresh
module has) so that BatchNorm folding in the packer is actually exercised.
"""
from __future__ import annotations

import numpy as np

# (gain, bias) applied to classifier.3 so that logits on full-scale noise spread over
# roughly +-3 instead of collapsing to ~0 (random init) or saturating.  Produced by
# tools/calibrate_synth.py with the oracle; constants, so every box agrees.
_LOGIT_CAL = {
    "dnn": (48.1, 5.16),
    "cnn": (28.4, -6.80),
    "tcn": (7.7, -0.96),
    "bcresnet": (58.1, -7.41),
    "crnn": (292.7, 32.75),
    "e2e_dnn": (132.1, 14.09),
    "gru": (94.2, -12.01),
    "lstm": (56.6, 9.78),
    "rnn": (97.1, -18.93),
    "quartznet": (40.3, -6.82),
    "e2e_quartznet": (36.2, 2.44),
    "e2e_cnn": (100.0, 3.64),        # global average pooling flattens a random-init model: gain capped, logits spread ~ +-0.3
}

DEFAULT_INPUT_SHAPE = {
    "dnn": (98, 40), "tcn": (98, 40),                     # (T, F)
    "gru": (98, 40), "lstm": (98, 40), "rnn": (98, 40), "quartznet": (98, 40),
    "cnn": (40, 98), "bcresnet": (40, 98), "crnn": (40, 98),   # (F, T)
    "e2e_dnn": (16000,), "e2e_quartznet": (16000,), "e2e_cnn": (16000,),
}


def default_config(model_type: str, **overrides) -> dict:
    cfg = {
        "model_type": model_type,
        "input_shape": list(DEFAULT_INPUT_SHAPE[model_type]),
        "activation_function": "relu",
        "embedding_dim": 64,
        "layer_dim": 128,
        "n_blocks": 1,
        "mode": "e2e" if model_type.startswith("e2e") else "embedding",
    }
    if model_type == "tcn":
        cfg.update(tcn_channels=[64, 64, 128], tcn_kernel_size=3)
    if model_type == "crnn":
        cfg.update(crnn_cnn_channels=[16, 32, 32], crnn_rnn_type="gru")
    if model_type == "e2e_quartznet":                                              # model.py:99-131
        cfg.update(e2e_frontend_channels=32, e2e_frontend_depth=3, e2e_quartznet_config=[[64, 11, 1], [64, 13, 1], [64, 17, 1]])
    if model_type == "e2e_cnn":                                                    # model.py:109-117
        cfg.update(e2e_frontend_channels=32, e2e_frontend_depth=2)
    if model_type == "quartznet":
        cfg.update(quartznet_config=[[256, 33, 1], [256, 33, 1], [512, 39, 1]])      # model.py:240
    cfg.update(overrides)
    return cfg


class _Gen:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.sd: dict[str, np.ndarray] = {}

    def uniform(self, name, shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        self.sd[name] = self.rng.uniform(-b, b, size=shape).astype(np.float32)

    def conv(self, prefix, cout, cin_g, kh, kw, bias=True):
        fan_in = cin_g * kh * kw
        self.uniform(prefix + ".weight", (cout, cin_g, kh, kw), fan_in)
        if bias:
            self.uniform(prefix + ".bias", (cout,), fan_in)

    def conv1d(self, prefix, cout, cin, k):
        self.uniform(prefix + ".weight", (cout, cin, k), cin * k)
        self.uniform(prefix + ".bias", (cout,), cin * k)

    def linear(self, prefix, out_f, in_f):
        self.uniform(prefix + ".weight", (out_f, in_f), in_f)
        self.uniform(prefix + ".bias", (out_f,), in_f)

    def norm_affine(self, prefix, n):
        self.sd[prefix + ".weight"] = self.rng.uniform(0.5, 1.5, n).astype(np.float32)
        self.sd[prefix + ".bias"] = self.rng.normal(0.0, 0.1, n).astype(np.float32)

    def batchnorm(self, prefix, n):
        self.norm_affine(prefix, n)
        self.sd[prefix + ".running_mean"] = self.rng.normal(0.0, 0.1, n).astype(np.float32)
        self.sd[prefix + ".running_var"] = self.rng.uniform(0.5, 2.0, n).astype(np.float32)
        self.sd[prefix + ".num_batches_tracked"] = np.zeros((), dtype=np.int64)


def _conv_out_hw(h, w, n_pools):
    for _ in range(n_pools):
        h, w = h // 2, w // 2
    return h, w


def make_state_dict(cfg: dict, seed: int = 0) -> dict[str, np.ndarray]:
    """Random weights for ``cfg`` (see :func:`default_config`) keyed like the reference."""
    mt = cfg["model_type"]
    emb = cfg.get("embedding_dim", 64)
    ld = cfg.get("layer_dim", 128)
    nb = cfg.get("n_blocks", 1)
    shape = tuple(cfg["input_shape"])
    g = _Gen(seed)

    if mt == "dnn":                                   # architectures.py:102-126
        g.linear("model.layer1", ld, shape[0] * shape[1])
        g.norm_affine("model.layernorm1", ld)
        for i in range(nb):
            g.linear(f"model.blocks.{i}.fcn_layer", ld, ld)
            g.norm_affine(f"model.blocks.{i}.layer_norm", ld)
        g.linear("model.last_layer", emb, ld)
    elif mt == "cnn":                                 # architectures.py:51-80
        g.conv("model.conv1", 16, 1, 3, 3)
        g.conv("model.conv2", 32, 16, 3, 3)
        h, w = _conv_out_hw(shape[0], shape[1], 2)
        g.linear("model.fc1", 128, 32 * h * w)
        g.linear("model.fc2", emb, 128)
    elif mt == "tcn":                                 # architectures.py:290-362
        chans = cfg.get("tcn_channels", [64, 64, 128])
        k = cfg.get("tcn_kernel_size", 3)
        cin = shape[1]
        for i, c in enumerate(chans):
            p = f"model.tcn_blocks.{i}"
            g.conv1d(p + ".conv1", c, cin, k)
            g.conv1d(p + ".conv2", c, c, k)
            if cin != c:
                g.conv1d(p + ".downsample", c, cin, 1)
            cin = c
        g.linear("model.fc", emb, chans[-1])
    elif mt == "bcresnet":                            # architectures.py:620-687
        g.conv("model.init_conv.0", 32, 1, 3, 3, bias=False)
        g.batchnorm("model.init_conv.1", 32)
        for name, cin, cout in (("block1", 32, 64), ("block2", 64, 128), ("block3", 128, 256)):
            p = "model." + name
            g.conv(p + ".depthwise", cin, 1, 3, 3, bias=False)
            g.conv(p + ".pointwise", cout, cin, 1, 1, bias=False)
            g.batchnorm(p + ".bn1", cout)
            g.conv(p + ".shortcut.0", cout, cin, 1, 1, bias=False)
            g.batchnorm(p + ".shortcut.1", cout)
        g.linear("model.fc", emb, 256)
    elif mt == "crnn":                                # architectures.py:209-287
        chans = cfg.get("crnn_cnn_channels", [16, 32, 32])
        if cfg.get("crnn_rnn_type", "lstm").lower() != "gru":
            raise NotImplementedError("only the GRU variant of CRNN is supported")
        cin = 1
        for i, c in enumerate(chans):
            g.conv(f"model.cnn.{4 * i}", c, cin, 3, 3)
            g.batchnorm(f"model.cnn.{4 * i + 1}", c)
            cin = c
        h, _ = _conv_out_hw(shape[0], shape[1], len(chans))
        rnn_in, hid = chans[-1] * h, ld
        for sfx in ("", "_reverse"):
            for nm, sh in (("weight_ih_l0", (3 * hid, rnn_in)), ("weight_hh_l0", (3 * hid, hid)),
                           ("bias_ih_l0", (3 * hid,)), ("bias_hh_l0", (3 * hid,))):
                g.uniform(f"model.rnn.{nm}{sfx}", sh, hid)
        g.linear("model.fc", emb, 2 * hid)
    elif mt in ("gru", "lstm", "rnn"):                # architectures.py:129-146, 83-99, 149-161
        name, gates, hid = {"gru": ("gru", 3, ld), "lstm": ("lstm", 4, ld), "rnn": ("layer1", 4, 64)}[mt]
        for layer in range(nb):
            rnn_in = shape[1] if layer == 0 else 2 * hid
            for sfx in ("", "_reverse"):
                for nm, sh in ((f"weight_ih_l{layer}", (gates * hid, rnn_in)), (f"weight_hh_l{layer}", (gates * hid, hid)),
                               (f"bias_ih_l{layer}", (gates * hid,)), (f"bias_hh_l{layer}", (gates * hid,))):
                    g.uniform(f"model.{name}.{nm}{sfx}", sh, hid)
        g.linear("model.layer2" if mt == "rnn" else "model.fc", emb, 2 * hid)
    elif mt in ("quartznet", "e2e_quartznet"):        # architectures.py:366-437, 695-714, 796-817
        pre = "model"
        if mt == "e2e_quartznet":
            pre, cin = "model.backbone", 1
            for layer in range(cfg.get("e2e_frontend_depth", 3)):
                cout, k = cfg.get("e2e_frontend_channels", 32) * 2 ** layer, 41 if layer == 0 else 13
                g.uniform(f"model.frontend.conv_blocks.{3 * layer}.weight", (cout, cin, k), cin * k)
                if layer == 0:     # audio is in [-1, 1): a x100 first layer lets it move the BatchNorm-ed ReLUs, as a trained one does
                    g.sd["model.frontend.conv_blocks.0.weight"] = (g.sd["model.frontend.conv_blocks.0.weight"] * 100.0).astype(np.float32)
                g.batchnorm(f"model.frontend.conv_blocks.{3 * layer + 1}", cout)
                cin = cout
            qcfg = cfg.get("e2e_quartznet_config", [[64, 11, 1], [64, 13, 1], [64, 17, 1]])
        else:
            cin, qcfg = shape[1], cfg.get("quartznet_config", [[256, 33, 1], [256, 33, 1], [512, 39, 1]])
        i = 0
        for channels, k, reps in qcfg:
            for _ in range(reps):
                p = f"{pre}.quartznet_blocks.{i}"
                g.conv1d(p + ".depthwise_conv", cin, 1, k)
                g.conv1d(p + ".pointwise_conv", channels, cin, 1)
                g.batchnorm(p + ".batch_norm", channels)
                if cin != channels:
                    g.conv1d(p + ".residual_connector.0", channels, cin, 1)
                    g.batchnorm(p + ".residual_connector.1", channels)
                cin = channels
                i += 1
        g.linear(pre + ".fc", emb, cin)
    elif mt == "e2e_cnn":                             # architectures.py:738-793
        cin = 1
        for layer in range(cfg.get("e2e_frontend_depth", 2)):
            cout, k = cfg.get("e2e_frontend_channels", 32) * 2 ** layer, 41 if layer == 0 else 13
            g.uniform(f"model.frontend.conv_blocks.{3 * layer}.weight", (cout, cin, k), cin * k)
            g.batchnorm(f"model.frontend.conv_blocks.{3 * layer + 1}", cout)
            if layer == 0:     # as for e2e_quartznet: lets [-1, 1) audio move the BatchNorm-ed ReLUs
                g.sd["model.frontend.conv_blocks.0.weight"] = (g.sd["model.frontend.conv_blocks.0.weight"] * 100.0).astype(np.float32)
            cin = cout
        for name, ci, co in (("conv1", 1, 24), ("conv2", 24, 48), ("conv3", 48, 64), ("conv4", 64, 96)):
            g.conv(f"model.backbone.{name}.0", co, ci, 3, 3, bias=False)
            g.batchnorm(f"model.backbone.{name}.1", co)
        g.linear("model.backbone.fc", emb, 96)
    elif mt == "e2e_dnn":                             # architectures.py:820-888
        for i, (cin, cout) in zip((0, 4, 8), ((1, 16), (16, 32), (32, 64))):
            g.conv(f"model.conv_block.{i}", cout, cin, 3, 3)
            g.batchnorm(f"model.conv_block.{i + 1}", cout)
        g.linear("model.fc1", 128, 256)
        g.batchnorm("model.bn1", 128)
        g.linear("model.out", emb, 128)
    else:
        raise ValueError(f"Unsupported model_type: '{mt}'.")

    g.linear("classifier.0", emb // 2, emb)           # model.py:291-296
    g.linear("classifier.3", 1, emb // 2)
    gain, bias = _LOGIT_CAL.get(mt, (1.0, 0.0))
    g.sd["classifier.3.weight"] = (g.sd["classifier.3.weight"] * gain).astype(np.float32)
    g.sd["classifier.3.bias"] = (g.sd["classifier.3.bias"] * gain + bias).astype(np.float32)
    return g.sd


def synth_pcm(n_windows: int, clip_samples: int = 16000, seed: int = 0, kind: str = "uniform") -> np.ndarray:
    """Synthetic int16 windows: full-scale uniform noise, or a
    speech-like Gaussian (sigma 3000, clipped)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(-32768, 32768, size=(n_windows, clip_samples), dtype=np.int16)
    if kind == "gauss":
        x = rng.normal(0.0, 3000.0, size=(n_windows, clip_samples))
        return np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    raise ValueError(kind)
