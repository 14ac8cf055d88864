"""Dev probe (GPU box): small scoring calls per model type and engine variant, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_probe.py [heads]
Round 2 adds: float feeds, selective pushes, the fused ingest kernel, the pipelined / split CNN stages, the fused TCN launch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

heads = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cnn", "dnn", "tcn", "bcresnet", "crnn", "e2e_dnn", "gru", "lstm", "rnn",
                                                         "quartznet", "e2e_quartznet", "e2e_cnn"]
variants = {"cnn": [dict(), dict(cnn_stage="v3"), dict(cnn_stage="v4", split_per_sm=1)], "crnn": [dict(), dict(cnn_stage="v4")],
            "tcn": [dict(), dict(tcn_layers="rows_fused"), dict(tcn_layers="cone")]}
pcm = synth_pcm(70, seed=3, kind="gauss")
for mt in heads:
    cfg = default_config(mt)
    for kw in variants.get(mt, [dict()]):
        eng = Engine(make_state_dict(cfg, 0), cfg, **kw)
        dev = torch.from_numpy(pcm).cuda()
        s = eng.score_device(dev).cpu().numpy()
        f = eng.score_device_f32(torch.from_numpy(pcm.astype(np.float32) / 32768.0 * 0.73).cuda()).cpu().numpy()
        eng.stream_open(9)
        for i in range(14):
            chunk = pcm[:9, (i % 12) * 1280:(i % 12 + 1) * 1280].copy()
            t = eng.stream_push_host(chunk, select=[1, 7, 4] if i % 3 == 2 else None)
        t2 = eng.stream_push_host(pcm[:9, :1000].copy(), select=[0, 8])          # not a multiple of the hop: full-window path
        eng.stream_close()
        print(mt, kw, "ok", float(s.mean()), float(f.mean()), float(t.mean()), float(t2.mean()), flush=True)
        eng.close()
