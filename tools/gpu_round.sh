#!/bin/bash
# One GPU-box visit: triage, parity tests, bench, launch list, one full ncu capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/smi.txt
for v in v2 v1; do
  timeout 300 python tools/debug_gpu.py cnn 300 cnn_stage=$v > gpurun_out/debug_cnn_$v.log 2>&1; echo "rc=$?" >> gpurun_out/debug_cnn_$v.log
done
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_bench.py 4096 cnn,dnn > gpurun_out/quick_bench.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-streams > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-cnn2_stage} -s 4 -c 2 -f -o gpurun_out/prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-streams > gpurun_out/ncu_full.log 2>&1
tail -4 gpurun_out/debug_cnn_v2.log; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/quick_bench.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
