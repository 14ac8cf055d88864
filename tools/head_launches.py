"""Dev probe: one scoring call of a head at a given chunk size (run under ncu for a per-kernel breakdown)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
mt, B, cw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = default_config(mt); sd = make_state_dict(cfg, 0)
eng = Engine(sd, cfg, chunk_windows=cw)
pcm = torch.from_numpy(synth_pcm(B, seed=1234)).cuda()
for _ in range(2):
    eng.score_device(pcm)
torch.cuda.synchronize()
