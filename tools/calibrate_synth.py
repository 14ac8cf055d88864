"""Dev tool: derive the (gain, bias) table in nanowakeword_b200/synth.py::_LOGIT_CAL.

Runs the oracle (float64) on a mixed calibration batch with uncalibrated seed-0 weights and
prints gain/bias so that logits have mean 0 and standard deviation 2.5.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import synth
from oracle.heads import forward_logits

pcm = np.concatenate([synth.synth_pcm(24, seed=99, kind="uniform"),
                      synth.synth_pcm(24, seed=99, kind="gauss"),
                      (synth.synth_pcm(16, seed=98, kind="gauss") // 8).astype(np.int16)])
for mt in synth._LOGIT_CAL:
    synth._LOGIT_CAL[mt] = (1.0, 0.0)
for mt in list(synth._LOGIT_CAL):
    cfg = synth.default_config(mt)
    sd = synth.make_state_dict(cfg, seed=0)
    z = forward_logits(pcm, sd, cfg).ravel()
    gain = 2.5 / z.std()
    print(f'    "{mt}": ({gain:.1f}, {-gain * z.mean():.2f}),   # raw mean {z.mean():+.4f} std {z.std():.4f}')
