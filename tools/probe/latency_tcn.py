import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
cfg = default_config("tcn"); sd = make_state_dict(cfg, 0)
for kw in (dict(), dict(tcn_layers="rows_fused"), dict(tcn_layers="cone")):
    eng = Engine(sd, cfg, **kw)
    for n in (1, 16, 256, 1024):
        pin = torch.from_numpy(synth_pcm(n, seed=3)).pin_memory().numpy()
        for _ in range(30): eng.score_host(pin)
        ts = []
        for _ in range(300):
            t0 = time.perf_counter(); eng.score_host(pin); ts.append(time.perf_counter() - t0)
        ts = np.sort(np.array(ts)) * 1e6
        print(kw, n, f"p50 {ts[150]:.1f} us", flush=True)
    eng.close()
