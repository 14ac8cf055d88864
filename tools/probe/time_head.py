"""Dev probe (GPU box): windows/s of one model type, resident input, CUDA-event timed; per-kernel split with --split."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
mt = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = default_config(mt); eng = Engine(make_state_dict(cfg, 0), cfg)
pcm = torch.from_numpy(synth_pcm(n, seed=1234)).cuda(); out = torch.empty(n, device="cuda")
for _ in range(3): eng.score_device(pcm, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): eng.score_device(pcm, out=out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"{mt}: {ms:.4f} ms per {n} -> {n / ms * 1e3 / 1e6:.3f} M windows/s", flush=True)
