import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict
cfg = default_config("tcn"); sd = make_state_dict(cfg, 0)
rng = np.random.default_rng(0)
for n in (1, 64, 1024, 2048, 4096, 16384):
    ch = torch.from_numpy(np.clip(rng.normal(0, 3000, (n, 1280)), -32768, 32767).astype(np.int16)).cuda()
    out = torch.empty(n, device="cuda")
    for k in ("auto", "cone"):
        eng = Engine(sd, cfg, **({} if k == "auto" else dict(tcn_layers="cone")))
        eng.stream_open(n)
        for _ in range(16): eng.stream_push_device(ch, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): eng.stream_push_device(ch, out=out)
        b.record(); torch.cuda.synchronize()
        print(n, k, f"{a.elapsed_time(b) / 20 * 1e3:.1f} us per push", flush=True)
        eng.close()
