mkdir -p gpurun_out
{
for v in main smemstate smem384 smem256x2 b384; do
  if [ "$v" = main ]; then unset VKRT_CUDA_LIB; else export VKRT_CUDA_LIB=$PWD/variants/$v/libvkrt_cuda.so; fi
  echo "== $v hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
  echo "== $v rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
done
unset VKRT_CUDA_LIB
} > gpurun_out/r02e_shade_ab.txt 2>&1
cat gpurun_out/r02e_shade_ab.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02e_launches_soup_build.csv vkrt_b200/vkrt --soup 10000000 --spectral 0 --render-width 640 --render-height 360 --render-samples 1 > gpurun_out/r02e_soup_build.log 2>&1
python tools/launch_summary.py gpurun_out/r02e_launches_soup_build.csv > gpurun_out/r02e_launches_soup_build_summary.txt 2>&1; head -40 gpurun_out/r02e_launches_soup_build_summary.txt
python -m pytest tests/test_gpu_reference.py -q -m gpu -k "saved or closures" > gpurun_out/r02e_tests.log 2>&1; tail -4 gpurun_out/r02e_tests.log
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "material_edits or default_accel" >> gpurun_out/r02e_tests.log 2>&1; tail -3 gpurun_out/r02e_tests.log
