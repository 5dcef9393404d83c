set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02e_pytest.log; tail -30 gpurun_out/r02e_pytest.log
( time python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err ) 2> gpurun_out/r02e_bench.time; tail -3 gpurun_out/r02e_bench.time; tail -5 gpurun_out/r02e_bench.err
python tools/small_code_probe.py > gpurun_out/r02e_small_probe.log 2>&1; tail -20 gpurun_out/r02e_small_probe.log
python tests/pcie_probe.py > gpurun_out/r02e_pcie.log 2>&1; cat gpurun_out/r02e_pcie.log
