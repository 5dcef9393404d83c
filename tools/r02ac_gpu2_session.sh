# 2-GPU box: smoke(), NCCL tests, 2-rank bench, reference arm under torchrun
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02ac_smoke.log 2>&1; tail -3 gpurun_out/r02ac_smoke.log
python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r02ac_pytest_dist.log 2>&1; echo rc=$? >> gpurun_out/r02ac_pytest_dist.log; tail -3 gpurun_out/r02ac_pytest_dist.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02ac_bench_2gpu.json 2> gpurun_out/r02ac_bench_2gpu.err ) 2> gpurun_out/r02ac_bench_2gpu.time; tail -3 gpurun_out/r02ac_bench_2gpu.time; tail -3 gpurun_out/r02ac_bench_2gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02ac_bench_ref_2gpu.json 2> gpurun_out/r02ac_bench_ref_2gpu.err ) 2> gpurun_out/r02ac_bench_ref_2gpu.time; tail -3 gpurun_out/r02ac_bench_ref_2gpu.time; cat gpurun_out/r02ac_bench_ref_2gpu.json | cut -c1-400
