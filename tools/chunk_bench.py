"""Dev probe (GPU box): full-path windows/s of a head for several internal chunk sizes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
B = 4096
pcm = torch.from_numpy(synth_pcm(B, seed=1234)).cuda()
out = torch.empty(B, dtype=torch.float32, device="cuda")
for mt in sys.argv[1].split(","):
    cfg = default_config(mt); sd = make_state_dict(cfg, 0)
    for cw in (0, 592, 1184, 2368, 4144):
        eng = Engine(sd, cfg, chunk_windows=cw)
        for _ in range(2): eng.score_device(pcm, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): eng.score_device(pcm, out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{mt:9s} chunk {eng.info['chunk_windows']:5d}: {ms:8.3f} ms  {B / ms * 1e3 / 1e6:6.3f} Mwin/s")
        eng.close()
