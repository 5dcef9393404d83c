for v in default pf1c4 pf1c8 pf0c8 pf1c16; do
  echo "variant $v" >> gpurun_out/r02bg_level_variants.log
  if [ $v = default ]; then python tools/layered_qc_probe.py 32768,131072 1 >> gpurun_out/r02bg_level_variants.log 2>>gpurun_out/r02bg.err
  else LDPC_B200_LIB=tuning/libldpc_b200_$v.so python tools/layered_qc_probe.py 32768,131072 1 >> gpurun_out/r02bg_level_variants.log 2>>gpurun_out/r02bg.err; fi
done
cat gpurun_out/r02bg_level_variants.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02bg_launches_layered_qc_131072.csv python tools/layered_qc_one.py 131072 > gpurun_out/r02bg_ncu.log 2>&1
