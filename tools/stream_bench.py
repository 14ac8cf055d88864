"""Dev timing probe (GPU box): stream-steps/s of the multi-stream path, incremental mel ring vs full recompute."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L = 1280
heads = sys.argv[2].split(",") if len(sys.argv) > 2 else ["cnn", "dnn", "tcn", "bcresnet", "crnn"]
rng = np.random.default_rng(0)
chunks = [torch.from_numpy(np.clip(rng.normal(0, 3000, (n, L)), -32768, 32767).astype(np.int16)).cuda() for _ in range(4)]
for mt in heads:
    cfg = default_config(mt); sd = make_state_dict(cfg, 0)
    for inc in (True, False):
        eng = Engine(sd, cfg, stream_incremental=inc)
        eng.stream_open(n)
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        for i in range(14):                                  # fill the rings (13 pushes reach 16000 samples)
            eng.stream_push_device(chunks[i % 4], out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        a.record()
        for i in range(iters):
            eng.stream_push_device(chunks[i % 4], out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        print(f"{mt:9s} {'incremental' if inc else 'full       '} {n} streams x {L}: {ms:8.3f} ms/step  {n / ms * 1e3 / 1e6:7.3f} M stream-steps/s  mean score {float(out.mean()):.4f}")
        eng.stream_close(); eng.close()
