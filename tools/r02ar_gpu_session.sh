set -x
LDPC_LAYERED_V4_FRAMES=128 python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py -x -q -k "layered or fullsize" > gpurun_out/r02ar_pytest_v4.log 2>&1; tail -3 gpurun_out/r02ar_pytest_v4.log
python tools/layered_probe.py chain > gpurun_out/r02ar_layered_default.log 2>&1; cat gpurun_out/r02ar_layered_default.log
LDPC_LAYERED_V4_FRAMES=65536 python tools/layered_probe.py chain > gpurun_out/r02ar_layered_v4.log 2>&1; cat gpurun_out/r02ar_layered_v4.log
