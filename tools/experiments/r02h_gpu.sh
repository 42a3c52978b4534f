mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02h_tests.log 2>&1; tail -4 gpurun_out/r02h_tests.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; cut -c1-300 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
bash tools/ncu_round.sh r02h > gpurun_out/r02h_ncu.log 2>&1
C5_SAMPLES=256 bash tools/run_configs.sh r02h > gpurun_out/r02h_configs_cli.txt 2>&1; grep -E "^==>|Mpaths|rebuild" gpurun_out/r02h_configs_cli.txt
