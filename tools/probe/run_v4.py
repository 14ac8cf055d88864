import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
cfg = default_config("cnn"); eng = Engine(make_state_dict(cfg, 0), cfg, cnn_stage="v4", split_per_sm=int(sys.argv[1]) if len(sys.argv) > 1 else 7)
pcm = torch.from_numpy(synth_pcm(4096, seed=1234)).cuda(); out = torch.empty(4096, device="cuda")
for _ in range(3): eng.score_device(pcm, out=out)
torch.cuda.synchronize()
