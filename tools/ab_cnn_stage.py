"""Dev probe (GPU box): the CNN stage variants against the phase-serial one (v2): bit-identity and timing."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["v2", "v4"]
for mt in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["cnn", "crnn"]):
    cfg = default_config(mt); sd = make_state_dict(cfg, 0)
    pcm = torch.from_numpy(synth_pcm(B, seed=1234)).cuda()
    res = {}
    for v in variants:
        name, _, per = v.partition(":")
        eng = Engine(sd, cfg, cnn_stage=name, split_per_sm=int(per or 0))
        out = torch.empty(B, dtype=torch.float32, device="cuda")
        small, ex = eng.score_device(pcm[:300].contiguous(), want_mel=True)
        torch.cuda.synchronize()
        for _ in range(3): eng.score_device(pcm, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.score_device(pcm, out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        res[v] = (out.cpu().numpy().copy(), small.cpu().numpy().copy(), ex["mel"].cpu().numpy().copy())
        print(f"{mt} {v}: {ms:.3f} ms  {B / ms * 1e3 / 1e6:.3f} Mwin/s", flush=True)
        eng.close()
    ref = variants[0]
    for v in variants[1:]:
        same = [np.array_equal(res[ref][i], res[v][i]) for i in range(3)]
        d = [float(np.abs(res[ref][i] - res[v][i]).max()) for i in range(3)]
        print(f"  {v} vs {ref}: scores / scores[:300] / mel[:300] identical = {same}  max diff = {d}")
