import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
from oracle.heads import forward_scores, forward_logits
mt = sys.argv[1]; B = int(sys.argv[2])
cfg = default_config(mt); sd = make_state_dict(cfg, 0)
pcm = synth_pcm(B, seed=5, kind="gauss")
kw = dict(a.split("=") for a in sys.argv[3:])
eng = Engine(sd, cfg, **kw)
print("engine", mt, kw, eng.info)
dev = torch.from_numpy(pcm).cuda()
scores, ex = eng.score_device(dev, want_mel=True, want_logits=True, want_emb=True)
torch.cuda.synchronize()
n = min(B, 64)
logits, mel = forward_logits(pcm[:n], sd, cfg, return_mel=True)
print("mel err", np.abs(ex["mel"].cpu().numpy()[:n] - mel).max())
print("logit err", np.abs(ex["logits"].cpu().numpy()[:n] - logits.ravel()).max(), "scale", np.abs(sd["classifier.3.weight"]).sum())
s = 1 / (1 + np.exp(-logits.ravel()))
print("score err", np.abs(scores.cpu().numpy()[:n] - s).max())
idx = np.arange(B - n, B)
logits2 = forward_logits(pcm[idx], sd, cfg).ravel()
print("tail-of-batch logit err", np.abs(ex["logits"].cpu().numpy()[idx] - logits2).max())
