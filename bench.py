#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 wake-word engine.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Metric (BASELINE.json): 1-second 16 kHz windows scored per second.  Workload = BASELINE.json
configs[1]: batch = 4096 synthetic int16 windows per GPU, NS40x98 front end (frame 400 ->
FFT 512, hop 160, 40 mels) + CNN head, random-init weights (seeded), weak scaling.
One step = one pass of the hot path (PCM -> log-mel -> CNN -> classifier -> sigmoid) over one
batch; `value` has the PCM resident in HBM, `e2e` goes through B200Session.run() with pinned
HOST buffers (H2D of the batch and D2H of the scores inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

WINDOWS_PER_GPU = 4096
CLIP = 16000
MODEL = "cnn"
GEOMETRY = "NS40x98"
N_ROTATE = 4                      # distinct input batches cycled between steps (4 x 131 MB > 126 MB L2)
METRIC = "1-s 16kHz windows/sec"
UNIT = "windows/s"

# Algorithmic work per window (1 MAC = 2 flop):
FLOP_FRONTEND = 1.3e6             # 98 x rFFT-512 + power + sparse-triangular mel + log
FLOP_CNN_CONV = 1.13e6 + 9.03e6   # conv1 + conv2 (stage A together with the front end)
FLOP_CNN_TAIL = 1.97e6 + 0.02e6   # fc1, fc2, classifier (stage B)
BYTES_IN, BYTES_OUT = CLIP * 2, 4
# FP64 thread-instructions per window of the v2 front end: 49 packed FFT-512 x (64 x (pass 1 + pass 2) + 32 x pass 3)
FP64_OPS_PER_WINDOW = 49 * (64 * (118 + 88) + 32 * 190)


def workload_config(n_gpus):
    return {
        "workload": "configs[1]: batch=4096 synthetic 16kHz windows, CNN head, fused STFT->mel->CNN (NS40x98)",
        "windows_per_gpu": WINDOWS_PER_GPU,
        "global_batch": WINDOWS_PER_GPU * n_gpus,
        "clip_samples": CLIP,
        "geometry": GEOMETRY,
        "head": MODEL,
        "frontend_precision": "fp64 FFT/power, fp32 mel/log; conv2 bf16x3 split + fc1 tf32x3 split on tcgen05, fp32 accumulate",
        "weights": "random init, numpy default_rng(0) (nanowakeword_b200.synth)",
        "l2_policy": f"{N_ROTATE} distinct {WINDOWS_PER_GPU * CLIP * 2 / 1e6:.0f} MB input batches per GPU cycled between steps (each > 126 MB L2)",
        "parallelism": f"dp{n_gpus} (windows sharded, weights replicated, NCCL gather of scores)",
    }


# --------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_throughput(n_windows, threads, seed=1234, _cache={}):
    """Time the oracle port (numpy float32, the reference's arithmetic) on `threads` host threads.
    Only the scoring is timed (inputs and weights are built before the clock starts); every worker thread runs
    with ONE BLAS / OpenMP thread (threadpoolctl), so `threads` is the number of cores really used — the same in
    this arm and in the GPU arm's cpu_baseline leg."""
    from concurrent.futures import ThreadPoolExecutor
    from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm
    from oracle.heads import forward_scores
    from threadpoolctl import threadpool_limits
    if "sd" not in _cache:
        _cache["cfg"] = default_config(MODEL)
        _cache["sd"] = make_state_dict(_cache["cfg"], 0)
    cfg, sd = _cache["cfg"], _cache["sd"]
    key = ("pcm", n_windows, seed)
    if key not in _cache:
        _cache[key] = synth_pcm(n_windows, CLIP, seed=seed)
    pcm = _cache[key]
    chunk = 16
    parts = [pcm[i:i + chunk] for i in range(0, n_windows, chunk)]

    def work(p):
        return forward_scores(p, sd, cfg, GEOMETRY, np.float32)

    with threadpool_limits(limits=1):
        if threads == 1:
            t0 = time.perf_counter()
            out = [work(p) for p in parts]
            dt = time.perf_counter() - t0
        else:
            with ThreadPoolExecutor(max_workers=threads) as ex:
                list(ex.map(work, parts[:threads]))          # start the workers before the clock
                t0 = time.perf_counter()
                out = list(ex.map(work, parts))
                dt = time.perf_counter() - t0
    return n_windows / dt, dt, np.concatenate(out)


def cpu_baseline_sample(cores, seconds):
    """Bounded sample of the bench workload on the host: ~`seconds` of all-core work + a short 1-thread leg
    (the reference pins its session to one thread, nanointerpreter.py:955-959)."""
    thr, _, _ = cpu_oracle_throughput(16 * cores, cores)
    n_s = int(min(WINDOWS_PER_GPU, max(16 * cores, thr * seconds)))
    n_s -= n_s % 16
    passes = max(1, int(round(thr * seconds / n_s)))
    tot = 0.0
    for _ in range(passes):
        _, dt, _ = cpu_oracle_throughput(n_s, cores)
        tot += dt
    n_1 = 64
    thr1, dt1, _ = cpu_oracle_throughput(n_1, 1)
    return {"value": n_s * passes / tot, "unit": UNIT, "cores": cores, "kind": "port",
            "one_thread": {"value": thr1, "windows": n_1, "seconds": dt1},
            "sample": f"{passes} pass(es) over {n_s} windows of the same workload ({tot:.1f} s of CPU work, scoring only), "
                      f"numpy float32 oracle port, {cores} worker threads x 16-window chunks, 1 BLAS thread per worker"}, n_s, passes, tot


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    thr, _, _ = cpu_oracle_throughput(16 * cores, cores)
    per_step = int(min(WINDOWS_PER_GPU, max(16 * cores, thr * 90.0 / max(1, total_steps))))
    per_step -= per_step % 16
    for _ in range(args.warmup):
        cpu_oracle_throughput(per_step, cores)
    dt = 0.0
    for _ in range(args.steps):
        dt += cpu_oracle_throughput(per_step, cores)[1]
    value = per_step * args.steps / dt
    thr1, dt1, _ = cpu_oracle_throughput(64, 1)
    sample = (f"{per_step} of the {WINDOWS_PER_GPU} windows per step (bounded sample, scoring only), numpy float32 oracle port, "
              f"{cores} worker threads x 16-window chunks, 1 BLAS thread per worker")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "one_thread": {"value": thr1, "windows": 64, "seconds": dt1}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = pure-Python package on onnxruntime CPU (not installable offline); its per-window arithmetic "
                "is timed through the oracle port on the host cores",
    }
    emit(line)


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(prefix="nww_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- GPU arm
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained"), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, None, "fallback (B200_PROFILING.md)"


def load_ncu_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per window of the dominant kernel, from the committed digest of its
    `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/summarize_ncu.py from the .ncu-rep)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        for k, v in json.load(open(p)).items():
            if k.startswith(kernel_prefix):
                return v
    except Exception:
        pass
    return None


def pin_to_gpu_numa(local_rank):
    """Run this rank (and first-touch its pinned buffers) on the CPUs NVML reports as local to its GPU."""
    info = {"cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"
        try:
            info["numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            pass
    except Exception as ex:
        info["error"] = repr(ex)[:80]
    return info


def time_device(torch, fn, iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from nanowakeword_b200 import _lib, B200Session, Engine
    from nanowakeword_b200.sharding import ShardedScorer, gather_scores, partition
    from nanowakeword_b200.synth import default_config, make_state_dict, synth_pcm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback (use --impl reference for the CPU arm)")
    numa = pin_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    cfg = default_config(MODEL)
    sd = make_state_dict(cfg, 0)
    sess = B200Session(state_dict=sd, cfg=cfg, device=local_rank)
    eng = sess.engine
    n_local = WINDOWS_PER_GPU
    n_total = n_local * world

    # inputs: N_ROTATE distinct batches per rank, resident in HBM and mirrored in pinned host memory
    host_batches = [torch.from_numpy(synth_pcm(n_local, CLIP, seed=1234 + 97 * rank + i)).pin_memory() for i in range(N_ROTATE)]
    dev_batches = [h.to(dev) for h in host_batches]
    scores = torch.empty(n_local, dtype=torch.float32, device=dev)

    def step(i):
        eng.score_device(dev_batches[i % N_ROTATE], out=scores)
        if world > 1:
            return gather_scores(scores, n_total, rank, world)
        return scores

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (the contract's timed region) ------------------------------
    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    eng.set_profiling(True)
    eng.get_profile()
    launches0 = eng.info["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    prof = eng.get_profile()
    eng.set_profiling(False)
    launches = eng.info["kernel_launches"] - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = n_total * args.steps / (elapsed_ms * 1e-3)

    # ---- the same step sustained for >= 2 s (the 20-step region above is ~20 ms at boost clocks) ---------
    sustained = None
    if not args.no_sustained:
        n_sus = int(min(20000, max(200, 2500.0 / max(1e-3, elapsed_ms / args.steps))))
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_sus):
            step(i)
        s1.record()
        barrier()
        sus_ms = max_over_ranks(s0.elapsed_time(s1))
        sustained = {"value": n_total * n_sus / (sus_ms * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": sus_ms * 1e-3,
                     "clocks": sampler2.stop() if rank == 0 else None}

    # ---- end to end through the session duck type, host buffers --------------------------------
    host_np = [h.numpy() for h in host_batches]

    def e2e_step(i):
        return sess.run(None, {"input": host_np[i % N_ROTATE]})[0]

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_total * args.steps / e2e_s
    assert out.shape == (n_local, 1, 1)

    # ---- host link: pinned H2D rate of one rank alone and of all ranks together (the e2e ceiling at N > 1) -------
    h2d = None
    if not args.no_streams:
        dst = torch.empty_like(dev_batches[0])
        def h2d_rate():
            for _ in range(2):
                dst.copy_(host_batches[0], non_blocking=True)
            torch.cuda.synchronize()
            ms = time_device(torch, lambda i: dst.copy_(host_batches[i % N_ROTATE], non_blocking=True), 8)
            return n_local * CLIP * 2 / (ms * 1e-3) / 1e9
        barrier()
        alone = h2d_rate() if rank == 0 else 0.0
        barrier()
        together = h2d_rate()
        if world > 1:
            t = torch.tensor([together], dtype=torch.float64, device=dev)
            mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            h2d = {"rank0_alone_GBps": alone, "all_ranks_min_GBps": float(mn.item()), "all_ranks_sum_GBps": float(sm.item())}
        else:
            h2d = {"rank0_alone_GBps": alone}
        h2d["e2e_ceiling_windows_per_s"] = (h2d.get("all_ranks_sum_GBps", alone)) * 1e9 / (CLIP * 2)
        h2d["numa"] = numa

    # ---- stream mode: BASELINE config #3 (65 536 streams x TCN x 1280-sample steps) per GPU ---------------------
    # device = chunks resident in HBM; e2e = nww_stream_push_host on pinned host chunks (2 560 B per score over PCIe)
    streams = None
    if not args.no_streams:
        try:
            ns, L = 65536, 1280
            cfg_t = default_config("tcn")
            eng_t = Engine(make_state_dict(cfg_t, 0), cfg_t, device=local_rank)
            rng = np.random.default_rng(7 + rank)
            ch_host = [torch.from_numpy(np.clip(rng.normal(0, 3000, (ns, L)), -32768, 32767).astype(np.int16)).pin_memory()
                       for _ in range(2)]
            ch_dev = [c.to(dev) for c in ch_host]
            out_dev = torch.empty(ns, dtype=torch.float32, device=dev)
            eng_t.stream_open(ns)
            for i in range(14):                                  # 13 pushes fill the 16000-sample rings
                eng_t.stream_push_device(ch_dev[i & 1], out=out_dev)
            l0 = eng_t.info["kernel_launches"]
            barrier()
            ms = max_over_ranks(time_device(torch, lambda i: eng_t.stream_push_device(ch_dev[i & 1], out=out_dev), 10))
            per_push = (eng_t.info["kernel_launches"] - l0) / 10
            host_s = 1e9
            barrier()
            for _ in range(3):                                   # best of three: the host side of the box is noisy
                t0 = time.perf_counter()
                for i in range(10):
                    eng_t.stream_push_host(ch_host[i & 1].numpy())
                host_s = min(host_s, time.perf_counter() - t0)
            host_s = max_over_ranks(host_s)
            eng_t.stream_close()
            eng_t.close()
            streams = {"workload": f"configs[2]: {ns} streams per GPU x {L}-sample steps, TCN head (bf16 hi/lo split operands), "
                                   "incremental mel ring: one score per stream per step",
                       "value": ns * world / (ms * 1e-3), "e2e": ns * world * 10 / host_s, "unit": "stream-steps/s",
                       "kernel_launches_per_push": per_push,
                       "h2d_bytes_per_step": ns * world * L * 2, "d2h_bytes_per_step": ns * world * 4}
        except Exception as ex:                                  # never let a secondary block break the contract line
            streams = {"error": repr(ex)}

    # ---- N > 1: BASELINE configs #4 / #5 with the audio on ONE rank: NCCL scatter + score + gather in the timed region ----
    root_ingest = None
    if world > 1 and not args.no_streams:
        root_ingest = {}
        for tag, mt, per_gpu, pieces in (("cfg4_bcresnet", "bcresnet", 65536, 8), ("cfg5_crnn", "crnn", 131072, 4)):
            try:
                cfg_r = default_config(mt)
                eng_r = Engine(make_state_dict(cfg_r, 0), cfg_r, device=local_rank)
                n_all = per_gpu * world
                root = None
                if rank == 0:
                    g = torch.Generator(device=dev)
                    g.manual_seed(4321)
                    root = torch.empty((n_all, CLIP), dtype=torch.int16, device=dev)
                    for o in range(0, n_all, 16384):             # full-scale uniform int16, generated on the device
                        root[o:o + 16384] = torch.randint(-32768, 32768, (min(16384, n_all - o), CLIP), generator=g,
                                                          device=dev, dtype=torch.int32).to(torch.int16)
                scorer = ShardedScorer(lambda x: eng_r.score_device(x), CLIP, rank, world, dev)
                res = {}
                for mode, fn in (("sequential", lambda: scorer.score_from_root(root, n_all)),
                                 ("pipelined", lambda: scorer.score_from_root_pipelined(root, n_all, pieces))):
                    got = fn()                                   # warm-up (also sizes NCCL's buffers)
                    barrier()
                    ms = max_over_ranks(time_device(torch, lambda i: fn(), 2))
                    res[mode] = {"value": n_all / (ms * 1e-3), "ms": ms,
                                 "root_egress_GBps": (n_all - per_gpu) * CLIP * 2 / (ms * 1e-3) / 1e9}
                if rank == 0:
                    idx = torch.arange(0, n_all, n_all // 64, device=dev)
                    same = torch.equal(got[idx], eng_r.score_device(root[idx].contiguous()))
                    res.update({"workload": f"{n_all} windows on rank 0 -> {world} GPUs ({per_gpu} per GPU), {mt} head; "
                                            "scatter (grouped ncclSend/Recv) + score + gather (all_gather) timed together",
                                "exact_config": (tag == "cfg4_bcresnet" and world == 4) or (tag == "cfg5_crnn" and world == 8),
                                "unit": UNIT, "pieces": pieces, "nvlink_egress_peak_GBps": 900.0,
                                "scores_equal_single_gpu_on_64_strided_windows": bool(same)})
                root_ingest[tag] = res
                del root
                eng_r.close()
                torch.cuda.empty_cache()
            except Exception as ex:
                root_ingest[tag] = {"error": repr(ex)}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    parity = None          # filled by the cpu_baseline leg below (the only place this arm touches oracle/)

    hbm_peak, bf16_peak, bf16_sus, peak_src = load_peaks()
    a_ms = prof["stage_a_ms"] / max(1, prof["stage_a_spans"])            # average launch of the dominant kernel
    a_win = prof["stage_a_windows"] / max(1, prof["stage_a_spans"])
    share = prof["stage_a_ms"] / max(1e-9, prof["stage_a_ms"] + prof["stage_b_ms"])
    alg_bytes = BYTES_IN + BYTES_OUT                                     # 32 000 B PCM in + 4 B score out
    ach_gbs = a_win * alg_bytes / (a_ms * 1e-3) / 1e9
    sm_mhz = (clocks.get("sm_mhz") if clocks else None) or 1965.0
    cyc_per_window = a_ms * 1e-3 * sm_mhz * 1e6 * 148 / max(1.0, a_win)  # SM-cycles one window occupies one SM
    lib = _lib.load_library()
    import ctypes
    pipe_peaks = {}
    for kind, nm in ((0, "fp32_fma_tflops"), (1, "fp64_fma_tflops")):
        v = ctypes.c_double()
        rc = lib.nww_microbench(local_rank, kind, ctypes.byref(v))
        pipe_peaks[nm] = float(v.value) if rc == 0 else 0.0
    kernel_name = "cnn3_stage_kernel" if eng.cnn_stage == "v3" else "cnn2_stage_kernel"
    traffic = load_ncu_traffic(kernel_name)
    stage_s = a_ms * 1e-3
    tensor_tflops = a_win * 9.03e6 / stage_s / 1e12                      # conv2: the stage kernel's tensor-pipe work (as written)
    fp64_tflops = a_win * FP64_OPS_PER_WINDOW * 2 / stage_s / 1e12       # thread-instructions counted as FMAs (upper bound)
    fp32_tflops = a_win * (1.13e6 + 0.2e6) / stage_s / 1e12              # conv1 + sparse mel + log
    roofline = {
        "kernel": f"{kernel_name} (TMA PCM staging + FP64 FFT front end + conv1 + tcgen05 conv2, one window per CTA iteration)",
        # the path is compute-bound; the roofline is bound to flops, on the pipe that executes the
        # majority of the window's flops as written (conv2 + fc1 = 11.0 of 13.5 MFLOP: the tensor pipe)
        "bound": "tensor", "achieved": a_win * (FLOP_FRONTEND + FLOP_CNN_CONV) / stage_s / 1e12, "peak": bf16_peak,
        "unit": "TFLOP/s", "frac": a_win * (FLOP_FRONTEND + FLOP_CNN_CONV) / stage_s / 1e12 / bf16_peak,
        "peak_source": peak_src + " bf16 burst (the kernel is timed alone, launch by launch)",
        "traffic": traffic["dram_bytes_per_window"] * a_win if traffic else None,
        "traffic_source": traffic["source"] if traffic else "no committed ncu digest for this kernel",
        "avg_launch_ms": a_ms, "windows_per_launch": a_win, "share_of_step": share,
        "algorithmic_flop_per_window": FLOP_FRONTEND + FLOP_CNN_CONV,
        "algorithmic_bytes_per_window": alg_bytes,
        "sm_cycles_per_window": cyc_per_window,
        "pipes": {
            "hbm": {"achieved_GBps": ach_gbs, "peak_GBps": hbm_peak, "frac": ach_gbs / hbm_peak},
            "tensor_bf16": {"achieved_tflops": tensor_tflops, "peak_tflops": bf16_peak, "frac": tensor_tflops / bf16_peak,
                            "what": "conv2 as written (9.03 MFLOP/window); executed as 3 bf16 split products = 3x the MMA work"},
            "fp64": {"achieved_tflops": fp64_tflops, "peak_tflops": pipe_peaks["fp64_fma_tflops"],
                     "frac": fp64_tflops / max(1e-9, pipe_peaks["fp64_fma_tflops"]),
                     "what": f"{FP64_OPS_PER_WINDOW} FP64 thread-instructions per window (49 packed FFT-512 + power), peak = nww_microbench DFMA"},
            "fp32": {"achieved_tflops": fp32_tflops, "peak_tflops": pipe_peaks["fp32_fma_tflops"],
                     "frac": fp32_tflops / max(1e-9, pipe_peaks["fp32_fma_tflops"]),
                     "what": "conv1 + sparse mel + log (1.33 MFLOP/window), peak = nww_microbench FFMA"},
        },
        "note": "the kernel alternates between a front-end phase that sits on the SM's shared-memory data pipe (16 warp-private "
                "FP64 FFTs: 580 cycles per FFT against ~550 shared-memory wavefronts, tools/probe/fft_probe.cu) and a conv phase bound "
                "by barriers and tensor-core operand fetch; fractions per pipe are given so the distance to each roof is explicit",
    }
    if traffic and traffic.get("smem_wavefronts_per_window"):
        # one wavefront per cycle per SM is the shared-memory data pipe's roof (ncu l1tex__data_pipe_lsu_wavefronts_mem_shared)
        roofline["pipes"]["shared_memory"] = {
            "wavefronts_per_window": traffic["smem_wavefronts_per_window"], "cycles_per_window": cyc_per_window,
            "frac": traffic["smem_wavefronts_per_window"] / cyc_per_window,
            "what": "ncu digest (profiles/ncu_traffic.json) / live cycles per window; ~27 k of the wavefronts belong to the FFT phase, "
                    "which runs at ~95 % of this roof while it lasts"}
        roofline["pipes"]["issue"] = {"warp_instructions_per_window": traffic.get("warp_instructions_per_window"),
                                      "frac": (traffic.get("warp_instructions_per_window") or 0) / 4.0 / cyc_per_window,
                                      "what": "4 issue slots per cycle per SM"}

    # ---- secondary: the other model types the engine builds, same 4096-window batch resident in HBM -----------------
    other = None
    if world == 1 and not args.no_streams:
        other = {}
        for mt in ("dnn", "tcn", "bcresnet", "crnn", "e2e_dnn", "gru", "lstm", "rnn", "quartznet", "e2e_quartznet", "e2e_cnn"):
            try:
                cfg_o = default_config(mt)
                eng_o = Engine(make_state_dict(cfg_o, 0), cfg_o, device=local_rank)
                out_o = torch.empty(WINDOWS_PER_GPU, dtype=torch.float32, device=dev)
                for _ in range(2):
                    eng_o.score_device(dev_batches[0], out=out_o)
                torch.cuda.synchronize()
                ms = time_device(torch, lambda i: eng_o.score_device(dev_batches[i % N_ROTATE], out=out_o), 5)
                other[mt] = round(WINDOWS_PER_GPU / (ms * 1e-3), 1)
                eng_o.close()
            except Exception as ex:                              # never let the secondary block break the contract line
                other[mt] = repr(ex)
        other = {"workload": f"batch={WINDOWS_PER_GPU} windows resident in HBM, full path per model type (reference model_type names)",
                 "unit": UNIT, "values": other}

    # ---- secondary: the verifier of a cascade (nanointerpreter.py:758-769) scores only the streams whose gate fired ------
    # every stream ingests its chunk; the table shows how the verifier's push rate follows the gate's pass rate
    cascade = None
    if world == 1 and not args.no_streams:
        try:
            ns, L = 16384, 1280
            cfg_v = default_config("cnn")
            eng_v = Engine(make_state_dict(cfg_v, 0), cfg_v, device=local_rank)
            rng = np.random.default_rng(11)
            ch = torch.from_numpy(np.clip(rng.normal(0, 3000, (ns, L)), -32768, 32767).astype(np.int16)).to(dev)
            out_v = torch.empty(ns, dtype=torch.float32, device=dev)
            eng_v.stream_open(ns)
            for _ in range(14):
                eng_v.stream_push_device(ch, out=out_v)
            rates = {}
            for frac in (0.0, 0.01, 0.1, 0.5, 1.0):
                k = int(round(frac * ns))
                ids = torch.from_numpy(np.sort(rng.choice(ns, size=k, replace=False)).astype(np.int64)).to(dev)
                for _ in range(2):
                    eng_v.stream_push_select_device(ch, ids, out=out_v)
                ms = time_device(torch, lambda i: eng_v.stream_push_select_device(ch, ids, out=out_v), 10)
                rates[f"{frac:g}"] = round(ns / (ms * 1e-3), 1)
            eng_v.stream_close()
            eng_v.close()
            cascade = {"workload": f"{ns} streams x {L}-sample steps through nww_stream_push_select on the verifier (CNN head): all streams "
                                   "ingest, the listed fraction is scored", "unit": "stream-steps/s", "by_pass_rate": rates}
        except Exception as ex:
            cascade = {"error": repr(ex)}

    # ---- secondary: BASELINE configs[0] — one window, DNN head, through host buffers (what one predict() call costs) ----------
    latency = None
    if world == 1 and not args.no_streams:
        try:
            cfg_l = default_config("dnn")
            eng_l = Engine(make_state_dict(cfg_l, 0), cfg_l, device=local_rank)
            one = torch.from_numpy(synth_pcm(1, seed=3)).pin_memory().numpy()
            for _ in range(50):
                eng_l.score_host(one)
            ts = []
            for _ in range(400):
                t0 = time.perf_counter()
                eng_l.score_host(one)
                ts.append(time.perf_counter() - t0)
            ts = np.sort(np.asarray(ts)) * 1e6
            eng_l.close()
            latency = {"workload": "configs[0]: batch=1, DNN head, Engine.score_host (pinned int16 window in, score out, synchronous)",
                       "p50_us": round(float(ts[200]), 1), "p99_us": round(float(ts[395]), 1), "calls": 400}
        except Exception as ex:
            latency = {"error": repr(ex)}

    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # cpu_baseline leg: the oracle as the checker (parity spot check, outside every timed region) and as the CPU arm
        from oracle.heads import forward_scores
        k = 32
        ref, mel_ref = forward_scores(host_np[0][:k], sd, cfg, GEOMETRY, np.float64, return_mel=True)
        got, extra = eng.score_device(dev_batches[0][:k].contiguous(), want_mel=True)
        torch.cuda.synchronize()
        parity = {"score_max_abs_err": float(np.abs(got.cpu().numpy() - ref.ravel()).max()),
                  "mel_max_abs_err_db": float(np.abs(extra["mel"].cpu().numpy() - mel_ref).max()),
                  "windows_checked": k, "against": "float64 oracle (pinned to the reference's modules)"}
        try:
            os.sched_setaffinity(0, range(cores))               # the CPU leg uses every host core again
        except Exception:
            pass
        cpu = cpu_baseline_sample(cores, 15.0)[0]

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_total * CLIP * 2, "d2h_bytes_per_step": n_total * 4,
                "api": "B200Session.run(None, {'input': int16 (4096,16000) pinned host array}) per rank", "host_link": h2d},
        "gpu_launches": int(launches), "clocks": clocks, "sustained": sustained, "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity, "streams_cfg3": streams, "cascade_verifier": cascade, "latency_cfg1": latency, "root_ingest": root_ingest,
        "other_models": other,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Keep stdout for the ONE JSON line: anything else that writes to fd 1 (NCCL's version banner, library
    chatter) is sent to stderr for the duration of the run."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-streams", action="store_true", help="skip the secondary blocks (streams, root ingest, other models, host link)")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        _quiet_stdout()
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it, one rank per GPU
        port = 29500 + (os.getpid() % 1000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    _quiet_stdout()
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
