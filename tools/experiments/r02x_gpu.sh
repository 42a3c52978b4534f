mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== main cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== idxahead cornell hero"; VKRT_CUDA_LIB=variants/idxahead/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
echo "== main cornell rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
echo "== idxahead cornell rgb"; VKRT_CUDA_LIB=variants/idxahead/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
echo "== main inst hero"; timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== idxahead inst hero"; VKRT_CUDA_LIB=variants/idxahead/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
} > gpurun_out/r02x_index_ahead.txt 2>&1
cat gpurun_out/r02x_index_ahead.txt
