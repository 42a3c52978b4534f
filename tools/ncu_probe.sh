# ncu captures of the two hot kernels on the bench workload (run under gpurun; outputs land in gpurun_out/)
set -x
TAG=${1:-r01b}
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 1 -o gpurun_out/${TAG}_shade_hero python tests/perf_probe.py --mode hero --frames 1 --spp 8 > gpurun_out/ncu_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/${TAG}_trace_hero python tests/perf_probe.py --mode hero --frames 1 --spp 8 > gpurun_out/ncu_trace.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_hero.csv python tests/perf_probe.py --mode hero --frames 2 --spp 16 > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out
