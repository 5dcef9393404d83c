set -x
python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r02bb_pytest_train.log 2>&1; tail -5 gpurun_out/r02bb_pytest_train.log
for v in default tu2 tu8; do
  echo "variant $v" >> gpurun_out/r02bb_train_variants.log
  if [ $v = default ]; then python tools/train_probe.py 8192 >> gpurun_out/r02bb_train_variants.log 2>gpurun_out/r02bb_train.err
  else LDPC_B200_LIB=tuning/libldpc_b200_$v.so python tools/train_probe.py 8192 >> gpurun_out/r02bb_train_variants.log 2>>gpurun_out/r02bb_train.err; fi
done
cat gpurun_out/r02bb_train_variants.log
