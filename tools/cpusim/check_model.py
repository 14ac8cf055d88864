"""DEVELOPER TOOL: host-thread simulation of a head's kernels vs the oracle."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nanowakeword_b200.synth import default_config, make_state_dict
from nanowakeword_b200.weights import pack_blob, pack_tensors
from oracle.heads import forward_logits, embedding_from_features, head_input_from_mel
from oracle.frontend import GEOMETRIES, log_mel
sim = sys.argv[1]
arch = "cnn" if sim == "cnn2" else sim
pcm = np.load(os.path.join(ROOT, "tests/golden/frontend.npz"))["pcm"]
cfg = default_config(arch); sd = make_state_dict(cfg, 0)
d = tempfile.mkdtemp()
open(d + "/blob.bin", "wb").write(pack_blob(pack_tensors(sd, cfg)))
pcm.tofile(d + "/pcm.i16")
subprocess.check_call(["/tmp/sim_cnn2", d] if sim == "cnn2" else ["/tmp/sim_model", arch, d])
logits, mel = forward_logits(pcm, sd, cfg, return_mel=True)
got = np.fromfile(d + "/logits.f32", np.float32)
gmel = np.fromfile(d + "/mel.f32", np.float32).reshape(mel.shape)
emb = embedding_from_features(head_input_from_mel(mel, arch), sd, cfg)
gemb = np.fromfile(d + "/emb.f32", np.float32).reshape(emb.shape)
print("mel err", np.abs(gmel - mel).max())
print("emb err", np.abs(gemb - emb).max(), "emb scale", np.abs(emb).max())
print("logits", np.round(logits.ravel(), 3)); print("got   ", np.round(got, 3))
sc = 1 / (1 + np.exp(-logits.ravel())); gs = np.fromfile(d + "/scores.f32", np.float32)
print("score err", np.abs(sc - gs).max())
