# ncu evidence for one round (run under gpurun, one GPU): launch list of the bench command + full captures of the two hot kernels.
# Usage: bash tools/ncu_round.sh TAG     -> gpurun_out/TAG_*
set -x
TAG=${1:-r01d}
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 1 -o gpurun_out/${TAG}_shade_hero python tests/perf_probe.py --mode hero --frames 1 --spp 8 > gpurun_out/ncu_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/${TAG}_trace_hero python tests/perf_probe.py --mode hero --frames 1 --spp 8 > gpurun_out/ncu_trace.log 2>&1
ls -la gpurun_out
# DRAM traffic of every k_shade / k_trace launch of whole bench frames -> profiles/traffic.json (tools/ncu_traffic.py)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_shade|k_trace" -c 200 --csv --log-file gpurun_out/${TAG}_traffic_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-visit-counts > gpurun_out/${TAG}_ncu_traffic.log 2>&1
