# on-chip decode with in-place messages: parity, C1 speed, full suite, bench
set -x
python -m pytest tests/test_gpu_small.py -x -q > gpurun_out/r02v_pytest_small.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02v_pytest_small.log; tail -6 gpurun_out/r02v_pytest_small.log
for k in "basic 4194304" "n2d2 4194304" "rcq 4194304"; do python tools/small_one.py $k; done > gpurun_out/r02v_small_one.log 2>&1; cat gpurun_out/r02v_small_one.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02v_pytest.log; tail -5 gpurun_out/r02v_pytest.log
( time python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err ) 2> gpurun_out/r02v_bench.time; tail -3 gpurun_out/r02v_bench.time; tail -3 gpurun_out/r02v_bench.err
