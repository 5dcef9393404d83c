# full GPU test suite, resident-threshold latency probe, host-pipeline chunk sweep with posteriors, compaction knobs
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02m_pytest.log; tail -5 gpurun_out/r02m_pytest.log
python tools/resident_latency_probe.py > gpurun_out/r02m_resident_latency.jsonl 2> gpurun_out/r02m_resident_latency.err; tail -40 gpurun_out/r02m_resident_latency.jsonl
for dual in 1 0; do for chunk in 2048 4096 6144 8192; do
  echo "dual=$dual chunk=$chunk" >> gpurun_out/r02m_e2e_sweep.log
  LDPC_HOST_DUAL=$dual LDPC_HOST_CHUNK=$chunk python tools/e2e_trace.py post packed 2>&1 | grep untraced >> gpurun_out/r02m_e2e_sweep.log
done; done; cat gpurun_out/r02m_e2e_sweep.log
python tools/mc_knob_probe.py > gpurun_out/r02m_mc_knobs.jsonl 2> gpurun_out/r02m_mc_knobs.err; cat gpurun_out/r02m_mc_knobs.jsonl
