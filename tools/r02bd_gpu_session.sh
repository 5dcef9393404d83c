# per-kernel times of one training step (default build and the prefetch / 6-CTA variant), one full page of the check side
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:train_ --csv --log-file gpurun_out/r02bd_train_launches_default.csv python tools/train_one.py 8192 > gpurun_out/r02bd_ncu.log 2>&1
LDPC_B200_LIB=tuning/libldpc_b200_pfm6.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:train_ --csv --log-file gpurun_out/r02bd_train_launches_pfm6.csv python tools/train_one.py 8192 >> gpurun_out/r02bd_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:train_bwd_cn -s 24 -c 1 -f -o gpurun_out/train_cn python tools/train_one.py 8192 >> gpurun_out/r02bd_ncu.log 2>&1
ncu -i gpurun_out/train_cn.ncu-rep --page raw --csv > gpurun_out/r02bd_ncu_full_train_bwd_cn_raw.csv 2>/dev/null
ncu -i gpurun_out/train_cn.ncu-rep --page source --csv > gpurun_out/r02bd_ncu_train_bwd_cn_src.csv 2>/dev/null
rm -f gpurun_out/train_cn.ncu-rep
tail -3 gpurun_out/r02bd_ncu.log
