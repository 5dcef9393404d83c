# GPU tests on the wave policy, decode-only chunk sweep of the host pipeline, default bench
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02n_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02n_pytest.log; tail -5 gpurun_out/r02n_pytest.log
for dual in 1 0; do for chunk in 2048 4096 6144 8192; do
  echo "decode-only dual=$dual chunk=$chunk" >> gpurun_out/r02n_e2e_sweep.log
  LDPC_HOST_DUAL=$dual LDPC_HOST_CHUNK=$chunk python tools/e2e_trace.py packed 2>&1 | grep untraced >> gpurun_out/r02n_e2e_sweep.log
done; done
for chunk in 1024 1536 2048 3072; do
  echo "forward dual=1 chunk=$chunk" >> gpurun_out/r02n_e2e_sweep.log
  LDPC_HOST_CHUNK=$chunk python tools/e2e_trace.py post packed 2>&1 | grep untraced >> gpurun_out/r02n_e2e_sweep.log
done; cat gpurun_out/r02n_e2e_sweep.log
python tools/e2e_trace.py post packed 2> gpurun_out/r02n_trace_post.log; tail -40 gpurun_out/r02n_trace_post.log
python tools/latency_probe.py > gpurun_out/r02n_latency.log 2>&1; cat gpurun_out/r02n_latency.log
( time python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err ) 2> gpurun_out/r02n_bench.time; tail -3 gpurun_out/r02n_bench.time; tail -3 gpurun_out/r02n_bench.err
