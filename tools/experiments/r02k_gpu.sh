mkdir -p gpurun_out
{
for sc in "--mode hero --spp 16" "--scene inst:1000 --mode hero --spp 16" "--scene soup:1000000 --mode rgb --spp 8" "--scene soup:3000000 --mode rgb --spp 4" "--scene soup:10000000 --mode rgb --spp 4"; do
  for r in 20 24; do
    echo "== refill $r $sc"; VKRT_TRACE_REFILL=$r timeout 200 python tests/perf_probe.py $sc --frames 3 2>&1 | tail -1
  done
done
for v in ploc16 ploc25 ploc40; do
  echo "== $v cornell hero"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 --count 2>&1 | tail -1
  echo "== $v inst:1000"; VKRT_CUDA_LIB=variants/$v/libvkrt_cuda.so timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
echo "== main counted"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 --count 2>&1 | tail -1
} > gpurun_out/r02k_sweeps.txt 2>&1
cat gpurun_out/r02k_sweeps.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02k_tests.log 2>&1; tail -3 gpurun_out/r02k_tests.log
