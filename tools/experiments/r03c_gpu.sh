mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== main cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== topcache cornell hero"; VKRT_CUDA_LIB=variants/topcache/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
echo "== main soup"; timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
echo "== topcache soup"; VKRT_CUDA_LIB=variants/topcache/libvkrt_cuda.so timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
echo "== main inst"; timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== topcache inst"; VKRT_CUDA_LIB=variants/topcache/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== topcache parity"; VKRT_CUDA_LIB=variants/topcache/libvkrt_cuda.so python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
} > gpurun_out/r03c_top_cache.txt 2>&1
cat gpurun_out/r03c_top_cache.txt
