"""GPU-box helper: text digest of an ncu report (the same summary tools/summarize_ncu.py writes), so that only the text has to
travel back:  python tools/ncu_digest.py gpurun_out/x.ncu-rep gpurun_out/x_ncu_full.txt"""
import importlib.util, os, sys
spec = importlib.util.spec_from_file_location("s", os.path.join(os.path.dirname(os.path.abspath(__file__)), "summarize_ncu.py"))
m = importlib.util.module_from_spec(spec)
argv, sys.argv = sys.argv, ["x"]
try:
    spec.loader.exec_module(m)
except SystemExit:
    pass
except IndexError:
    pass
sys.argv = argv
m.full(sys.argv[1], sys.argv[2])
