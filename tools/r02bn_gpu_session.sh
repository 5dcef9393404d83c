python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py -x -q -k "layered or fullsize" > gpurun_out/r02bn_pytest.log 2>&1; tail -3 gpurun_out/r02bn_pytest.log
for f in 1 0; do echo "fuse $f" >> gpurun_out/r02bn_layered_qc.log; LDPC_LAYERED_FUSE_HARD=$f python tools/layered_qc_probe.py 32768,131072 1 >> gpurun_out/r02bn_layered_qc.log 2>&1; done
cat gpurun_out/r02bn_layered_qc.log
