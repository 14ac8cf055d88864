import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
cfg = default_config("tcn"); sd = make_state_dict(cfg, 0)
pcm = torch.from_numpy(synth_pcm(3000, seed=5, kind="gauss")).cuda()
r = {}
for k in ("rows", "cone"):
    eng = Engine(sd, cfg, tcn_layers=k); r[k] = eng.score_device(pcm).cpu().numpy().copy(); eng.close()
d = np.abs(r["rows"] - r["cone"])
print("identical", np.array_equal(r["rows"], r["cone"]), "max diff", d.max(), "n diff", int((d > 0).sum()))
