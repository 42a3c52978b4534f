mkdir -p gpurun_out
{
echo "== main cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
echo "== main cornell rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
for v in shpf nobar finebar shpfnobar; do
  echo "== $v cornell hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
  echo "== $v cornell rgb"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
  echo "== $v inst:1000 hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
} > gpurun_out/r02m_shade_ab.txt 2>&1
cat gpurun_out/r02m_shade_ab.txt
