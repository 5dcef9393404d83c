# full ncu page of the check-side training backward kernel in its final form (one launch at t > 0, 8192 frames)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:train_bwd_cn -s 12 -c 1 -f -o gpurun_out/train_cn python tools/train_one.py 8192 > gpurun_out/r02bu_ncu.log 2>&1
ncu -i gpurun_out/train_cn.ncu-rep --page raw --csv > gpurun_out/r02bu_ncu_full_train_bwd_cn_v2_8192frames_raw.csv 2>/dev/null
ncu -i gpurun_out/train_cn.ncu-rep --page source --csv > gpurun_out/r02bu_ncu_train_bwd_cn_v2_src.csv 2>/dev/null
rm -f gpurun_out/train_cn.ncu-rep; tail -2 gpurun_out/r02bu_ncu.log
