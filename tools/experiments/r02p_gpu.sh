mkdir -p gpurun_out
{
echo "== main cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== main soup"; timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
echo "== main inst"; timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
for v in chunk64 chunk128 chunk256; do
  echo "== $v cornell hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
  echo "== $v soup"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
  echo "== $v inst:1000 hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
echo "== chunk64 parity"; VKRT_CUDA_LIB=variants/chunk64/libvkrt_cuda.so python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cornell or soup or determin or trace_rays or instanced" 2>&1 | tail -2
} > gpurun_out/r02p_trace_chunk_ab.txt 2>&1
cat gpurun_out/r02p_trace_chunk_ab.txt
