set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02ab_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02ab_pytest.log; tail -4 gpurun_out/r02ab_pytest.log
python tools/mc_knob_probe.py > gpurun_out/r02ab_mc_knobs.jsonl 2> gpurun_out/r02ab_mc_knobs.err; cat gpurun_out/r02ab_mc_knobs.jsonl
python tools/r02_probe.py > gpurun_out/r02ab_probe.jsonl 2> gpurun_out/r02ab_probe.err; cat gpurun_out/r02ab_probe.jsonl
( time python bench.py > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err ) 2> gpurun_out/r02ab_bench.time; tail -3 gpurun_out/r02ab_bench.time; tail -3 gpurun_out/r02ab_bench.err
