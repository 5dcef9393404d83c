# resident decode, second version: parity tests, latency A/B against the per-iteration path, U variants
set -x
python -m pytest tests/test_gpu_resident.py -x -q > gpurun_out/r02s_pytest_res.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02s_pytest_res.log; tail -8 gpurun_out/r02s_pytest_res.log
for v in "" resu0 resu4; do
  echo "variant=${v:-default}" >> gpurun_out/r02s_resident_one.log
  for k in "dvbs2 n2d2 148" "dvbs2 rcq 148" "qc wrcq1 148" "r504 n2d2 4096" "dvbs2 n2d2 1"; do
    if [ -n "$v" ]; then LDPC_B200_LIB=$PWD/tuning/libldpc_b200_$v.so python tools/resident_one.py $k; else python tools/resident_one.py $k; fi >> gpurun_out/r02s_resident_one.log 2>&1
  done
done; cat gpurun_out/r02s_resident_one.log
python tools/resident_latency_probe.py > gpurun_out/r02s_resident_latency.jsonl 2> gpurun_out/r02s_resident_latency.err; cat gpurun_out/r02s_resident_latency.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02s_pytest.log; tail -5 gpurun_out/r02s_pytest.log
