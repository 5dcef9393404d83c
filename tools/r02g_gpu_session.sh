set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02g_pytest.log; tail -30 gpurun_out/r02g_pytest.log
python tools/e2e_trace.py post packed 2> gpurun_out/r02g_trace_post.log; tail -13 gpurun_out/r02g_trace_post.log
( time python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err ) 2> gpurun_out/r02g_bench.time; tail -3 gpurun_out/r02g_bench.time; tail -5 gpurun_out/r02g_bench.err
for k in basic n2d2 rcq; do python tools/small_one.py $k 1048576; done > gpurun_out/r02g_small_one.log 2>&1; cat gpurun_out/r02g_small_one.log
python tools/small_one.py basic 1048576 > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_decode -s 2 -c 1 -f -o gpurun_out/small_basic python tools/small_one.py basic 1048576 > gpurun_out/r02g_ncu_small.log 2>&1
ncu -i gpurun_out/small_basic.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_small_basic_h74_1048576frames_raw.csv 2>/dev/null
ncu -i gpurun_out/small_basic.ncu-rep --page source --csv > gpurun_out/r02_ncu_small_basic_src.csv 2>/dev/null
rm -f gpurun_out/small_basic.ncu-rep; ls -la gpurun_out | tail -5
