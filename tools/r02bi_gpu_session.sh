for cfg in "1 0" "0 0" "65536 131072" "65536 65536"; do
  set -- $cfg
  echo "V2 from $1, V4 from $2" >> gpurun_out/r02bi_layered_v.log
  LDPC_LAYERED_V2_FRAMES=$1 LDPC_LAYERED_V4_FRAMES=$2 python tools/layered_probe.py chain 2>&1 | grep -v " 8192 " >> gpurun_out/r02bi_layered_v.log
done
cat gpurun_out/r02bi_layered_v.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:layered_pipe -s 3 -c 1 -f -o gpurun_out/pipe python tools/layered_one.py 131072 > gpurun_out/r02bi_ncu.log 2>&1
ncu -i gpurun_out/pipe.ncu-rep --page raw --csv > gpurun_out/r02bi_ncu_full_layered_pipe_v2_dvbs2_131072frames_raw.csv 2>/dev/null
ncu -i gpurun_out/pipe.ncu-rep --page source --csv > gpurun_out/r02bi_ncu_layered_pipe_src.csv 2>/dev/null
rm -f gpurun_out/pipe.ncu-rep; tail -2 gpurun_out/r02bi_ncu.log
