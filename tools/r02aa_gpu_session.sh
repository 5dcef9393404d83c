set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02aa_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02aa_pytest.log; tail -4 gpurun_out/r02aa_pytest.log
python tools/train_probe.py 8192 > gpurun_out/r02aa_train_probe.jsonl 2> gpurun_out/r02aa_train_probe.err; cat gpurun_out/r02aa_train_probe.jsonl
( time python bench.py > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err ) 2> gpurun_out/r02aa_bench.time; tail -3 gpurun_out/r02aa_bench.time; tail -3 gpurun_out/r02aa_bench.err
