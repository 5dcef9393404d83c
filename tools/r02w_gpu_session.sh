# resident decode with 32-node tiles: parity, speed, ncu, latency table; then the whole suite
set -x
python -m pytest tests/test_gpu_resident.py -x -q > gpurun_out/r02w_pytest_res.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02w_pytest_res.log; tail -8 gpurun_out/r02w_pytest_res.log
for k in "dvbs2 n2d2 148" "dvbs2 rcq 148" "qc wrcq1 148" "r504 n2d2 4096" "dvbs2 n2d2 1"; do python tools/resident_one.py $k; done 2>&1 | grep -v "^+" > gpurun_out/r02w_resident_one.log; cat gpurun_out/r02w_resident_one.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resident_decode -s 3 -c 1 -f -o gpurun_out/res_dvbs2 python tools/resident_one.py dvbs2 n2d2 148 > gpurun_out/r02w_ncu_res.log 2>&1
ncu -i gpurun_out/res_dvbs2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_resident_v4_n2d2_dvbs2_148frames_raw.csv 2>/dev/null
rm -f gpurun_out/res_dvbs2.ncu-rep
python tools/resident_latency_probe.py > gpurun_out/r02w_resident_latency.jsonl 2> gpurun_out/r02w_resident_latency.err; cat gpurun_out/r02w_resident_latency.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r02w_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02w_pytest.log; tail -5 gpurun_out/r02w_pytest.log
