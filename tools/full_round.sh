mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python tools/chunk_bench.py cnn,dnn,tcn,bcresnet,crnn,e2e_dnn,gru,lstm,rnn,quartznet,e2e_quartznet,e2e_cnn 2>&1 | grep -v " 592\| 1184\| 2368" > gpurun_out/heads.log
timeout 300 python tools/stream_bench.py 65536 crnn,quartznet 2>&1 | grep incremental > gpurun_out/streams2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cat gpurun_out/heads.log gpurun_out/streams2.log; tail -2 gpurun_out/smoke.log
