set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02d_pytest.log; tail -30 gpurun_out/r02d_pytest.log
( time python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err ) 2> gpurun_out/r02d_bench.time; tail -3 gpurun_out/r02d_bench.time; tail -5 gpurun_out/r02d_bench.err
python tools/small_code_probe.py > gpurun_out/r02d_small_probe.log 2>&1; tail -20 gpurun_out/r02d_small_probe.log
