import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
cfg = default_config(sys.argv[1]); eng = Engine(make_state_dict(cfg, 0), cfg)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
pcm = torch.from_numpy(synth_pcm(n, seed=1234)).cuda(); out = torch.empty(n, device="cuda")
for _ in range(2): eng.score_device(pcm, out=out)
torch.cuda.synchronize()
