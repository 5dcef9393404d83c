python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r02bf_pytest_train.log 2>&1; tail -2 gpurun_out/r02bf_pytest_train.log
for v in default m5 m4 u2 nopf; do
  echo "variant $v" >> gpurun_out/r02bf_train_variants.log
  if [ $v = default ]; then python tools/train_probe.py 8192 2>>gpurun_out/r02bf_train.err >> gpurun_out/r02bf_train_variants.log
  else LDPC_B200_LIB=tuning/libldpc_b200_$v.so python tools/train_probe.py 8192 2>>gpurun_out/r02bf_train.err | head -1 >> gpurun_out/r02bf_train_variants.log; fi
done
cat gpurun_out/r02bf_train_variants.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:train_bwd --csv --log-file gpurun_out/r02bf_launches_default.csv python tools/train_one.py 8192 > gpurun_out/r02bf_ncu.log 2>&1
