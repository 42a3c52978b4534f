mkdir -p gpurun_out
{
for m in 0 1 2 3; do
  echo "== VKRT_RAY_SORT=$m cornell hero"; VKRT_RAY_SORT=$m timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
  echo "== VKRT_RAY_SORT=$m soup"; VKRT_RAY_SORT=$m timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
done
echo "== VKRT_RAY_SORT=3 parity subset"; VKRT_RAY_SORT=3 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cornell_rgb or soup_ids or determin" 2>&1 | tail -2
} > gpurun_out/r02i_raysort_ab.txt 2>&1
cat gpurun_out/r02i_raysort_ab.txt
