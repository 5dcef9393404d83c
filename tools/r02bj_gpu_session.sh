set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02bj_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02bj_pytest.log; tail -4 gpurun_out/r02bj_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02bj_smoke.log 2>&1; tail -2 gpurun_out/r02bj_smoke.log
( time python bench.py > gpurun_out/r02bj_bench.json 2> gpurun_out/r02bj_bench.err ) 2> gpurun_out/r02bj_bench.time; tail -3 gpurun_out/r02bj_bench.time; tail -3 gpurun_out/r02bj_bench.err
