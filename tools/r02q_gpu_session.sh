# ncu evidence for the CTA-resident decode; decode-only host pipeline trace; bench after the simulator / gc changes
set -x
python tools/resident_one.py dvbs2 n2d2 148 > gpurun_out/r02q_resident_one.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:resident_decode -s 3 -c 1 -f -o gpurun_out/res_dvbs2 python tools/resident_one.py dvbs2 n2d2 148 > gpurun_out/r02q_ncu_res.log 2>&1
ncu -i gpurun_out/res_dvbs2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_resident_n2d2_dvbs2_148frames_raw.csv 2>/dev/null
ncu -i gpurun_out/res_dvbs2.ncu-rep --page source --csv > gpurun_out/r02_ncu_resident_n2d2_dvbs2_src.csv 2>/dev/null
rm -f gpurun_out/res_dvbs2.ncu-rep
python tools/resident_one.py r504 n2d2 4096 >> gpurun_out/r02q_resident_one.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:resident_decode -s 3 -c 1 -f -o gpurun_out/res_r504 python tools/resident_one.py r504 n2d2 4096 >> gpurun_out/r02q_ncu_res.log 2>&1
ncu -i gpurun_out/res_r504.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_resident_n2d2_r504_4096frames_raw.csv 2>/dev/null
ncu -i gpurun_out/res_r504.ncu-rep --page source --csv > gpurun_out/r02_ncu_resident_n2d2_r504_src.csv 2>/dev/null
rm -f gpurun_out/res_r504.ncu-rep
cat gpurun_out/r02q_resident_one.log; tail -3 gpurun_out/r02q_ncu_res.log
python tools/e2e_trace.py packed 2> gpurun_out/r02q_trace_decode_only.log; tail -25 gpurun_out/r02q_trace_decode_only.log
( time python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err ) 2> gpurun_out/r02q_bench.time; tail -3 gpurun_out/r02q_bench.time; tail -3 gpurun_out/r02q_bench.err
