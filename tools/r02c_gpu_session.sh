# round-2 GPU session: tests, default bench, reference arm, launch list (ncu), full captures exported as raw CSV pages
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02c_pytest.log; tail -30 gpurun_out/r02c_pytest.log
( time python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err ) 2> gpurun_out/r02c_bench.time; tail -3 gpurun_out/r02c_bench.time; tail -5 gpurun_out/r02c_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_ref.json 2> gpurun_out/r02c_ref.err ) 2> gpurun_out/r02c_ref.time; tail -3 gpurun_out/r02c_ref.time
python tools/small_code_probe.py > gpurun_out/r02c_small_probe.log 2>&1; tail -20 gpurun_out/r02c_small_probe.log
