# 8-GPU bench at HEAD (device-resident, e2e through one host, Monte-Carlo sweep + parity, training leg)
set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02bt_bench_8gpu.json 2> gpurun_out/r02bt_bench_8gpu.err ) 2> gpurun_out/r02bt_bench_8gpu.time; tail -3 gpurun_out/r02bt_bench_8gpu.time; tail -5 gpurun_out/r02bt_bench_8gpu.err
