for r in 1 2; do
echo "prev build" >> gpurun_out/r02bo_layered_qc.log; LDPC_B200_LIB=tuning/libldpc_b200_prev.so python tools/layered_qc_probe.py 32768,131072 1 >> gpurun_out/r02bo_layered_qc.log 2>&1
for f in 0 1; do echo "fuse $f" >> gpurun_out/r02bo_layered_qc.log; LDPC_LAYERED_FUSE_HARD=$f python tools/layered_qc_probe.py 32768,131072 1 >> gpurun_out/r02bo_layered_qc.log 2>&1; done
done
cat gpurun_out/r02bo_layered_qc.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
