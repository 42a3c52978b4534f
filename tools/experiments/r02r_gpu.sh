mkdir -p gpurun_out
{
echo "== cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
echo "== cornell rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
echo "== cornell single"; timeout 150 python tests/perf_probe.py --mode single --frames 3 --spp 16 2>&1 | tail -1
echo "== inst:1000 hero"; timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== soup"; timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
} > gpurun_out/r02r_horizon_skip.txt 2>&1
cat gpurun_out/r02r_horizon_skip.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
