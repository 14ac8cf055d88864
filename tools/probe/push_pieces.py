"""Dev probe (GPU box): nww_stream_push_host on pinned chunks, by number of pieces the bank is cut into."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mt = sys.argv[2] if len(sys.argv) > 2 else "tcn"
rng = np.random.default_rng(0)
ch = [torch.from_numpy(np.clip(rng.normal(0, 3000, (n, 1280)), -32768, 32767).astype(np.int16)).pin_memory() for _ in range(2)]
cfg = default_config(mt); sd = make_state_dict(cfg, 0)
for pieces in (1, 2, 3, 4, 6, 8):
    eng = Engine(sd, cfg, push_pieces=pieces)
    eng.stream_open(n)
    for i in range(15): eng.stream_push_host(ch[i & 1].numpy())
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for i in range(10): eng.stream_push_host(ch[i & 1].numpy())
        best = min(best, time.perf_counter() - t0)
    print(f"{mt} pieces {pieces}: {best / 10 * 1e3:.3f} ms per push  {n * 10 / best / 1e6:.2f} M stream-steps/s", flush=True)
    eng.close()
