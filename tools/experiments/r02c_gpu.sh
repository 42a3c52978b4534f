mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02c_tests.log 2>&1; tail -3 gpurun_out/r02c_tests.log
{
echo "== cornell hero default"; python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -4
echo "== cornell hero default, split shadow trace"; VKRT_TRACE_SPLIT=1 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
echo "== cornell hero default counted"; python tests/perf_probe.py --mode hero --frames 2 --spp 4 --count 2>&1 | tail -1
echo "== soup default"; python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -4
echo "== soup default, split"; VKRT_TRACE_SPLIT=1 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -2
echo "== inst 1000"; python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -2
} > gpurun_out/r02c_probes.txt 2>&1
cat gpurun_out/r02c_probes.txt
C5_SAMPLES=256 bash tools/run_configs.sh r02c > gpurun_out/r02c_configs_cli.txt 2>&1; tail -60 gpurun_out/r02c_configs_cli.txt
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; cut -c1-400 gpurun_out/r02c_bench.json
bash tools/ncu_round.sh r02c > gpurun_out/r02c_ncu.log 2>&1
ls -la gpurun_out | tail -20
