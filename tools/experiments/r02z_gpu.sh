mkdir -p gpurun_out
{
echo "== main cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== main soup"; timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
echo "== main inst"; timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
for v in mb5 mb7 mb8; do
  echo "== $v cornell hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
  echo "== $v soup"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
  echo "== $v inst:1000 hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
for r in 16 20 24 28; do
  echo "== refill $r cornell hero"; VKRT_TRACE_REFILL=$r timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
  echo "== refill $r soup"; VKRT_TRACE_REFILL=$r timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
done
} > gpurun_out/r02z_trace_occupancy.txt 2>&1
cat gpurun_out/r02z_trace_occupancy.txt
