mkdir -p gpurun_out
{
for v in main stride16 nobarrier finebarrier; do
  if [ "$v" = main ]; then unset VKRT_CUDA_LIB; else export VKRT_CUDA_LIB=$PWD/variants/$v/libvkrt_cuda.so; fi
  echo "== $v hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
  echo "== $v rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
done
unset VKRT_CUDA_LIB
} > gpurun_out/r02g_shade_ab.txt 2>&1
cat gpurun_out/r02g_shade_ab.txt
