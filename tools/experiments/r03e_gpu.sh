mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== main (interval lut) cornell hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== nolut cornell hero"; VKRT_CUDA_LIB=variants/nolut/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
echo "== main cornell single"; timeout 150 python tests/perf_probe.py --mode single --frames 3 --spp 16 2>&1 | tail -1
echo "== nolut cornell single"; VKRT_CUDA_LIB=variants/nolut/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode single --frames 3 --spp 16 2>&1 | tail -1
echo "== main inst hero"; timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
echo "== nolut inst hero"; VKRT_CUDA_LIB=variants/nolut/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
} > gpurun_out/r03e_interval_lut.txt 2>&1
cat gpurun_out/r03e_interval_lut.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -2
