"""Dev probe: time the v3 CNN stage of an experimental library build (results are NOT valid scores)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nanowakeword_b200 import _lib
_lib.load_library(sys.argv[1]); _lib._lib = _lib.load_library(sys.argv[1])
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
cfg = default_config("cnn"); eng = Engine(make_state_dict(cfg, 0), cfg, cnn_stage="v3")
pcm = torch.from_numpy(synth_pcm(4096, seed=1234)).cuda(); out = torch.empty(4096, device="cuda")
for _ in range(3): eng.score_device(pcm, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): eng.score_device(pcm, out=out)
b.record(); torch.cuda.synchronize()
print(sys.argv[1], f"{a.elapsed_time(b) / 10:.3f} ms")
