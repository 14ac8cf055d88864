"""Short profiling target: a few score_device calls of one head (for ncu -k regex:...)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
mt = sys.argv[1] if len(sys.argv) > 1 else "cnn"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = default_config(mt); sd = make_state_dict(cfg, 0)
eng = Engine(sd, cfg)
pcm = torch.from_numpy(synth_pcm(B, seed=1234)).cuda()
out = torch.empty(B, dtype=torch.float32, device="cuda")
for _ in range(reps):
    eng.score_device(pcm, out=out)
torch.cuda.synchronize()
print("done", float(out.mean()))
