set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02b_pytest.log; tail -3 gpurun_out/r02b_pytest.log
( time python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err ) 2> gpurun_out/r02b_bench.time; tail -3 gpurun_out/r02b_bench.time; tail -5 gpurun_out/r02b_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_ref.json 2> gpurun_out/r02b_ref.err ) 2> gpurun_out/r02b_ref.time; tail -3 gpurun_out/r02b_ref.time
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs"
$CMD > gpurun_out/r02b_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_launch.log 2>&1
for cfg in "n2d2 dvbs2" "rcq dvbs2" "wrcq1 qc"; do
  set -- $cfg
  CMD="python bench.py --decoder $1 --code $2 --frames 65536 --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs --decode-only"
  $CMD > gpurun_out/r02b_plain_$1_$2.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k regex:"cn_kernel|vn_kernel|cn_wide_kernel" -s 22 -c 2 -f -o gpurun_out/r02_ncu_full_$1_$2_65536frames $CMD > gpurun_out/r02b_ncu_$1_$2.log 2>&1
  tail -2 gpurun_out/r02b_ncu_$1_$2.log
done
ls -la gpurun_out/*.ncu-rep
