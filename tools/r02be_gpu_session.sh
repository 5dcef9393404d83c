for v in default pfm6 pf2m6 pf3m6 pf2m5 pf1m7; do
  echo "variant $v" >> gpurun_out/r02be_train_variants.log
  if [ $v = default ]; then python tools/train_probe.py 8192 2>>gpurun_out/r02be_train.err | head -1 >> gpurun_out/r02be_train_variants.log
  else
    LDPC_B200_LIB=tuning/libldpc_b200_$v.so python tools/train_probe.py 8192 2>>gpurun_out/r02be_train.err | head -1 >> gpurun_out/r02be_train_variants.log
    LDPC_B200_LIB=tuning/libldpc_b200_$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:train_bwd_cn --csv --log-file gpurun_out/r02be_launches_$v.csv python tools/train_one.py 8192 > gpurun_out/r02be_ncu.log 2>&1
  fi
done
cat gpurun_out/r02be_train_variants.log
LDPC_B200_LIB=tuning/libldpc_b200_pfm6.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:train_bwd_cn -s 12 -c 1 -f -o gpurun_out/train_cn python tools/train_one.py 8192 >> gpurun_out/r02be_ncu.log 2>&1
ncu -i gpurun_out/train_cn.ncu-rep --page raw --csv > gpurun_out/r02be_ncu_full_train_bwd_cn_raw.csv 2>/dev/null
ncu -i gpurun_out/train_cn.ncu-rep --page source --csv > gpurun_out/r02be_ncu_train_bwd_cn_src.csv 2>/dev/null
rm -f gpurun_out/train_cn.ncu-rep
