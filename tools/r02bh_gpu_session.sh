python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py -x -q -k "layered or fullsize" > gpurun_out/r02bh_pytest.log 2>&1; tail -2 gpurun_out/r02bh_pytest.log
python tools/layered_probe.py chain > gpurun_out/r02bh_layered_chain.log 2>&1; cat gpurun_out/r02bh_layered_chain.log
python tools/layered_qc_probe.py 32768,131072 1 > gpurun_out/r02bh_layered_qc.log 2>&1; cat gpurun_out/r02bh_layered_qc.log
