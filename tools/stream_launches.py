"""Dev probe (GPU box, under ncu): ONE push of n streams after the rings are full — for a per-kernel launch list."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict
n = int(sys.argv[1]); mt = sys.argv[2]
cfg = default_config(mt); eng = Engine(make_state_dict(cfg, 0), cfg)
rng = np.random.default_rng(0)
ch = torch.from_numpy(np.clip(rng.normal(0, 3000, (n, 1280)), -32768, 32767).astype(np.int16)).cuda()
out = torch.empty(n, dtype=torch.float32, device="cuda")
eng.stream_open(n)
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 15):
    eng.stream_push_device(ch, out=out)
torch.cuda.synchronize()
print("launches", eng.info["kernel_launches"])
