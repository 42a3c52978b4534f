# Multi-GPU evidence (gpurun --gpus N): bit-identity of the gathered film, then the bench line at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/r03g_mgpu_check_n$N.log 2>&1; tail -5 gpurun_out/r03g_mgpu_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r03g_bench_n$N.json 2> gpurun_out/r03g_bench_n$N.err; tail -c 1800 gpurun_out/r03g_bench_n$N.json; tail -3 gpurun_out/r03g_bench_n$N.err
