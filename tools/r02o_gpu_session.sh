# GPU tests (per-iteration kernels by default in the suites + resident policy test), default bench with the new chunk defaults
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02o_pytest.log; tail -5 gpurun_out/r02o_pytest.log
( time python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err ) 2> gpurun_out/r02o_bench.time; tail -3 gpurun_out/r02o_bench.time; tail -3 gpurun_out/r02o_bench.err
python tools/e2e_trace.py post packed 2> gpurun_out/r02o_trace_post.log; tail -45 gpurun_out/r02o_trace_post.log
