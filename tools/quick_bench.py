"""Dev timing probe (GPU box): per-kernel device times for the main stages."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pcm = torch.from_numpy(synth_pcm(B, seed=1234)).cuda()
print(torch.cuda.get_device_name(0), "B =", B)
_h = torch.from_numpy(synth_pcm(B, seed=1)).pin_memory(); _d = torch.empty_like(_h, device="cuda")
ms = timeit(lambda: _d.copy_(_h, non_blocking=True))
print(f"H2D pinned {_h.numel() * 2 / 1e6:.0f} MB: {ms:.3f} ms = {_h.numel() * 2 / ms / 1e6:.1f} GB/s")
for mt in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["cnn", "dnn"]):
    cfg = default_config(mt); sd = make_state_dict(cfg, 0)
    for prec in ("fp64", "v1", "fp32"):
        try:
            eng = Engine(sd, cfg, cnn_stage="v1") if prec == "v1" else Engine(sd, cfg, frontend_precision=prec)
        except Exception as ex:
            print(mt, prec, "engine:", ex); continue
        ms = timeit(lambda: eng.logmel_device(pcm))
        print(f"{mt:9s} {prec} frontend-only  {ms:8.3f} ms  {B / ms * 1e3 / 1e6:7.3f} Mwin/s")
        if prec in ("fp64", "v1"):
            out = torch.empty(B, dtype=torch.float32, device="cuda")
            ms = timeit(lambda: eng.score_device(pcm, out=out))
            print(f"{mt:9s} {prec} full path      {ms:8.3f} ms  {B / ms * 1e3 / 1e6:7.3f} Mwin/s  launches={eng.info['kernel_launches']}")
            host = torch.from_numpy(synth_pcm(B, seed=1234)).pin_memory()
            hs = torch.empty(B, dtype=torch.float32).pin_memory()
            eng.score_host_ptr(host.data_ptr(), B, hs.data_ptr())
            t0 = time.perf_counter()
            for _ in range(5): eng.score_host_ptr(host.data_ptr(), B, hs.data_ptr())
            dt = (time.perf_counter() - t0) / 5
            print(f"{mt:9s} {prec} host e2e       {dt * 1e3:8.3f} ms  {B / dt / 1e6:7.3f} Mwin/s")
