"""DEVELOPER TOOL: compare the host-thread simulation of frontend_kernel with the oracle."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.frontend import GEOMETRIES, log_mel
g = np.load(os.path.join(ROOT, "tests/golden/frontend.npz"))
pcm = g["pcm"]
for geom, short in (("NS40x98", "NS"), ("REF64x101", "REF")):
    spec = GEOMETRIES[geom]
    t = np.load(os.path.join(ROOT, "nanowakeword_b200/tables", geom + ".npz"))
    ref = log_mel(pcm, spec, np.float64)
    for prec in ("f32", "f64"):
        d = tempfile.mkdtemp()
        pcm.tofile(d + "/pcm.i16"); t["window"].astype(np.float32).tofile(d + "/window.f32")
        t["fb"].astype(np.float32).tofile(d + "/fb.f32")
        subprocess.check_call(["/tmp/sim_frontend", short, prec, d])
        mel = np.fromfile(d + "/mel.f32", dtype=np.float32).reshape(ref.shape)
        err = np.abs(mel - ref).reshape(len(pcm), -1).max(1)
        print(geom, prec, "max err per window:", " ".join(f"{e:.1e}" for e in err))
