set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02h_pytest.log; tail -30 gpurun_out/r02h_pytest.log
for k in "basic 4194304" "basic 1048576" "n2d2 4194304"; do python tools/small_one.py $k; done > gpurun_out/r02h_small_one.log 2>&1; cat gpurun_out/r02h_small_one.log
( time python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err ) 2> gpurun_out/r02h_bench.time; tail -3 gpurun_out/r02h_bench.time; tail -5 gpurun_out/r02h_bench.err
