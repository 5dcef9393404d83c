for v in default pf m6 m8 pfm6 pfm8u2; do
  echo "variant $v" >> gpurun_out/r02bc_train_variants.log
  if [ $v = default ]; then python tools/train_probe.py 8192 2>>gpurun_out/r02bc_train.err | head -1 >> gpurun_out/r02bc_train_variants.log
  else LDPC_B200_LIB=tuning/libldpc_b200_$v.so python tools/train_probe.py 8192 2>>gpurun_out/r02bc_train.err | head -1 >> gpurun_out/r02bc_train_variants.log; fi
done
cat gpurun_out/r02bc_train_variants.log
