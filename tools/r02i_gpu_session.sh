# tests + ncu evidence of round 2 (raw pages exported on the box: the .ncu-rep files are too large to bring back)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo pytest_rc=$? >> gpurun_out/r02i_pytest.log; tail -30 gpurun_out/r02i_pytest.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs"
$CMD > gpurun_out/r02i_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv $CMD > gpurun_out/r02i_ncu_launch.log 2>&1
cap() {  # name decoder code skip count extra-args
  CMD="python bench.py --decoder $2 --code $3 --frames 65536 --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs $6"
  $CMD > gpurun_out/r02i_plain_$1.log 2>&1 && timeout 500 ncu --set full --clock-control none -k regex:"cn_kernel|vn_kernel|cn_wide_kernel" -s $4 -c $5 -f -o gpurun_out/cap_$1 $CMD > gpurun_out/r02i_ncu_$1.log 2>&1
  ncu -i gpurun_out/cap_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_$2_$3_65536frames_raw.csv 2>/dev/null
  rm -f gpurun_out/cap_$1.ncu-rep; tail -2 gpurun_out/r02i_ncu_$1.log
}
cap fwd n2d2 dvbs2 90 3 ""
cap rcq rcq dvbs2 62 2 "--decode-only"
cap wrcq1 wrcq1 qc 62 2 "--decode-only"
ls -la gpurun_out | tail -8
