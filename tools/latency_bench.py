"""Dev probe (GPU box): single-window latency through the reference-facing calls (BASELINE config #1 shape)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import B200Session, NanoInterpreter
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

for mt in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["dnn", "cnn", "tcn", "e2e_dnn"]):
    cfg = default_config(mt); sd = make_state_dict(cfg, 0)
    sess = B200Session(state_dict=sd, cfg=cfg)
    x = synth_pcm(1, seed=3)
    for _ in range(50): sess.run(None, {"input": x})
    t0 = time.perf_counter()
    n = 500
    for _ in range(n): sess.run(None, {"input": x})
    run_us = (time.perf_counter() - t0) / n * 1e6
    interp = NanoInterpreter(wakeword_models=["m.pt"], sessions={"m": sess})
    chunks = synth_pcm(1, seed=4).reshape(-1)
    for i in range(0, 16000, 1280): interp.predict(chunks[i:i + 1280])
    t0 = time.perf_counter()
    for k in range(n): interp.predict(chunks[(k % 12) * 1280:(k % 12) * 1280 + 1280])
    pred_us = (time.perf_counter() - t0) / n * 1e6
    print(f"{mt:8s} session.run(1 window) {run_us:7.1f} us   NanoInterpreter.predict(1280-sample chunk) {pred_us:7.1f} us")
