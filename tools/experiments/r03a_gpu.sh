# Final single-GPU evidence of round 2: tests, bench line, BASELINE configs through the CLI, ncu launch list + captures + traffic.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r03a_tests.log 2>&1; tail -3 gpurun_out/r03a_tests.log
python bench.py > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err; cut -c1-600 gpurun_out/r03a_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r03a_bench_reference.json 2> gpurun_out/r03a_bench_reference.err; cut -c1-400 gpurun_out/r03a_bench_reference.json
C5_SAMPLES=256 bash tools/run_configs.sh r03a > gpurun_out/r03a_configs_cli.txt 2>&1; grep -E "rebuild|Mpaths" gpurun_out/r03a_configs_cli.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03a_smoke.log 2>&1; tail -2 gpurun_out/r03a_smoke.log
bash tools/ncu_round.sh r03a > gpurun_out/r03a_ncu.log 2>&1
ls -la gpurun_out | grep r03a
