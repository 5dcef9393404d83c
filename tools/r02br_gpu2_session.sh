# 2-GPU box: smoke(), NCCL tests, 2-rank bench, reference arm under torchrun
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02br_smoke.log 2>&1; tail -3 gpurun_out/r02br_smoke.log
python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r02br_pytest_dist.log 2>&1; echo rc=$? >> gpurun_out/r02br_pytest_dist.log; tail -3 gpurun_out/r02br_pytest_dist.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02br_bench_2gpu.json 2> gpurun_out/r02br_bench_2gpu.err ) 2> gpurun_out/r02br_bench_2gpu.time; tail -3 gpurun_out/r02br_bench_2gpu.time; tail -3 gpurun_out/r02br_bench_2gpu.err
