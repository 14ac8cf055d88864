"""Dev probe (GPU box): a small scoring call per model type, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_probe.py [heads]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

heads = sys.argv[1].split(",") if len(sys.argv) > 1 else ["gru", "lstm", "rnn", "quartznet", "e2e_quartznet", "e2e_cnn"]
pcm = synth_pcm(70, seed=3, kind="gauss")
for mt in heads:
    cfg = default_config(mt)
    eng = Engine(make_state_dict(cfg, 0), cfg)
    s = eng.score_device(torch.from_numpy(pcm).cuda()).cpu().numpy()
    eng.stream_open(9)
    for i in range(14):
        t = eng.stream_push_host(pcm[:9, (i % 12) * 1280:(i % 12 + 1) * 1280].copy())
    eng.stream_close()
    print(mt, "ok", float(s.mean()), float(t.mean()), flush=True)
    eng.close()
