# full GPU test suite + default bench + reference arm on the current tree
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02l_pytest.log; tail -5 gpurun_out/r02l_pytest.log
( time python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err ) 2> gpurun_out/r02l_bench.time; tail -3 gpurun_out/r02l_bench.time; tail -3 gpurun_out/r02l_bench.err
( time python bench.py --impl reference > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err ) 2> gpurun_out/r02l_bench_ref.time; tail -3 gpurun_out/r02l_bench_ref.time
python tools/small_code_probe.py > gpurun_out/r02l_small_probe.log 2>&1; tail -20 gpurun_out/r02l_small_probe.log
