set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02as_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02as_pytest.log; tail -4 gpurun_out/r02as_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02as_smoke.log 2>&1; tail -2 gpurun_out/r02as_smoke.log
( time python bench.py --impl reference > gpurun_out/r02as_bench_ref.json 2> gpurun_out/r02as_bench_ref.err ) 2> gpurun_out/r02as_bench_ref.time; tail -3 gpurun_out/r02as_bench_ref.time
( time python bench.py > gpurun_out/r02as_bench.json 2> gpurun_out/r02as_bench.err ) 2> gpurun_out/r02as_bench.time; tail -3 gpurun_out/r02as_bench.time; tail -3 gpurun_out/r02as_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_steps2_final.csv $CMD > gpurun_out/r02as_ncu_launch.log 2>&1; tail -2 gpurun_out/r02as_ncu_launch.log
