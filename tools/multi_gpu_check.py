"""GPU box, N ranks under torchrun: NCCL scatter of PCM from rank 0 -> per-rank engines -> gather of scores,
checked against rank 0 scoring everything alone.
    torchrun ... tools/multi_gpu_check.py [model_type = cnn] [windows per rank = 4096]
(BASELINE config #4: bcresnet 65536 on 4 GPUs; #5: crnn 131072 on 8 GPUs)"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanowakeword_b200 import Engine
from nanowakeword_b200.sharding import ShardedScorer
from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
MT = sys.argv[1] if len(sys.argv) > 1 else "cnn"
PER = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = default_config(MT); sd = make_state_dict(cfg, 0)
eng = Engine(sd, cfg, device=lr)
n = PER * world + 77                        # ragged on purpose
root = torch.from_numpy(synth_pcm(n, seed=99)).to(dev) if rank == 0 else None
sc = ShardedScorer(lambda x: eng.score_device(x), 16000, rank, world, dev)
for _ in range(2):
    out = sc.score_from_root(root, n)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = sc.score_from_root(root, n)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
for pieces in (2, 4, 8):
    for _ in range(2):
        out_p = sc.score_from_root_pipelined(root, n, n_chunks=pieces)
    torch.cuda.synchronize(); dist.barrier()
    a.record()
    for _ in range(5):
        out_p = sc.score_from_root_pipelined(root, n, n_chunks=pieces)
    b.record(); torch.cuda.synchronize()
    ms_p = a.elapsed_time(b) / 5
    if rank == 0:
        print(f"{MT} world {world}: pipelined ({pieces} pieces, transfers on a side stream) {ms_p:.3f} ms = {n / ms_p * 1e3 / 1e6:.3f} M windows/s; "
              f"identical: {bool(torch.equal(out_p, out))}")
        assert torch.equal(out_p, out)
if rank == 0:
    ref = eng.score_device(root)
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    print(f"{MT} world {world}: scatter+score+gather of {n} windows {ms:.3f} ms = {n / ms * 1e3 / 1e6:.3f} M windows/s; identical to single-GPU scores: {same}")
    assert same
dist.destroy_process_group()
