mkdir -p gpurun_out
{
echo "== main cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== main cornell rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
for v in b192x3 b128x5 b128x4 b320x2; do
  echo "== $v cornell hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
  echo "== $v cornell rgb"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
done
} > gpurun_out/r02v_shade_shapes.txt 2>&1
cat gpurun_out/r02v_shade_shapes.txt
