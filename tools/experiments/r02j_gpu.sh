mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q -m gpu > gpurun_out/r02j_tests.log 2>&1; tail -3 gpurun_out/r02j_tests.log
{
echo "== main cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
echo "== main soup"; timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
for v in rf16 rf20 rf28 pv4 pv12 pv16 mb5 mb8; do
  echo "== $v cornell hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
  echo "== $v soup"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
done
} > gpurun_out/r02j_trace_ab.txt 2>&1
cat gpurun_out/r02j_trace_ab.txt
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 1 -o gpurun_out/r02j_shade_hero python tests/perf_probe.py --mode hero --frames 1 --spp 8 > gpurun_out/ncu_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/r02j_trace_hero python tests/perf_probe.py --mode hero --frames 1 --spp 8 > gpurun_out/ncu_trace.log 2>&1
ls -la gpurun_out | grep r02j
