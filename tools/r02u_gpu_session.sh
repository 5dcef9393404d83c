# full GPU tests with the final resident decode + policy, ncu evidence of it, default bench
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02u_pytest.log; tail -5 gpurun_out/r02u_pytest.log
python tools/resident_one.py dvbs2 n2d2 148 > gpurun_out/r02u_resident_one.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resident_decode -s 3 -c 1 -f -o gpurun_out/res_dvbs2 python tools/resident_one.py dvbs2 n2d2 148 > gpurun_out/r02u_ncu_res.log 2>&1
ncu -i gpurun_out/res_dvbs2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_resident_n2d2_dvbs2_148frames_raw.csv 2>/dev/null
rm -f gpurun_out/res_dvbs2.ncu-rep
python tools/resident_one.py r504 n2d2 4096 >> gpurun_out/r02u_resident_one.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resident_decode -s 3 -c 1 -f -o gpurun_out/res_r504 python tools/resident_one.py r504 n2d2 4096 >> gpurun_out/r02u_ncu_res.log 2>&1
ncu -i gpurun_out/res_r504.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_resident_n2d2_r504_4096frames_raw.csv 2>/dev/null
rm -f gpurun_out/res_r504.ncu-rep
cat gpurun_out/r02u_resident_one.log
python tools/latency_probe.py > gpurun_out/r02u_latency.log 2>&1; cat gpurun_out/r02u_latency.log
( time python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err ) 2> gpurun_out/r02u_bench.time; tail -3 gpurun_out/r02u_bench.time; tail -3 gpurun_out/r02u_bench.err
