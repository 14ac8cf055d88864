"""Dev probe (GPU box): single-window latency of B200Session.run / Engine.score_host (config #1: B = 1, DNN head)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
for mt in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["dnn", "cnn", "tcn"]):
    cfg = default_config(mt); eng = Engine(make_state_dict(cfg, 0), cfg)
    pcm = synth_pcm(1, seed=3)
    pin = torch.from_numpy(pcm).pin_memory().numpy()
    dev = torch.from_numpy(pcm).cuda(); out = torch.empty(1, device="cuda")
    for _ in range(50): eng.score_host(pin)
    ts = []
    for _ in range(500):
        t0 = time.perf_counter(); eng.score_host(pin); ts.append(time.perf_counter() - t0)
    ts = np.sort(np.array(ts)) * 1e6
    for _ in range(50): eng.score_device(dev, out=out)
    torch.cuda.synchronize()
    td = []
    for _ in range(500):
        t0 = time.perf_counter(); eng.score_device(dev, out=out); torch.cuda.synchronize(); td.append(time.perf_counter() - t0)
    td = np.sort(np.array(td)) * 1e6
    print(f"{mt}: host path p50 {ts[250]:.1f} us p99 {ts[494]:.1f} us | device path (+sync) p50 {td[250]:.1f} us p99 {td[494]:.1f} us | launches/call {eng.info['kernel_launches'] / 1100:.1f}", flush=True)
    eng.close()
