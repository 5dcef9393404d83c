set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02bp_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02bp_pytest.log; tail -4 gpurun_out/r02bp_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02bp_smoke.log 2>&1; tail -2 gpurun_out/r02bp_smoke.log
( time python bench.py --impl reference > gpurun_out/r02bp_bench_ref.json 2> gpurun_out/r02bp_bench_ref.err ) 2> gpurun_out/r02bp_bench_ref.time; tail -3 gpurun_out/r02bp_bench_ref.time
( time python bench.py > gpurun_out/r02bp_bench.json 2> gpurun_out/r02bp_bench.err ) 2> gpurun_out/r02bp_bench.time; tail -3 gpurun_out/r02bp_bench.time; tail -3 gpurun_out/r02bp_bench.err
