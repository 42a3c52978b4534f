mkdir -p gpurun_out
{
for v in main stateonly b128x4 so192x3 so128x4; do
  if [ "$v" = main ]; then unset VKRT_CUDA_LIB; else export VKRT_CUDA_LIB=$PWD/variants/$v/libvkrt_cuda.so; fi
  echo "== $v hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
  echo "== $v rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
done
unset VKRT_CUDA_LIB
} > gpurun_out/r02f_shade_ab.txt 2>&1
cat gpurun_out/r02f_shade_ab.txt
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x > gpurun_out/r02f_tests.log 2>&1; tail -4 gpurun_out/r02f_tests.log
python -m pytest tests/test_gpu_reference.py -q -m gpu -k "lobes or closures or alpha or bundled" >> gpurun_out/r02f_tests.log 2>&1; tail -3 gpurun_out/r02f_tests.log
