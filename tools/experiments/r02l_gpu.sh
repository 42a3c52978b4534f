mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q -m gpu > gpurun_out/r02l_tests.log 2>&1; tail -5 gpurun_out/r02l_tests.log
{
for off in 0 1; do
  echo "== NO_MEMO=$off cornell hero"; VKRT_NO_SPECTRAL_MEMO=$off timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
  echo "== NO_MEMO=$off cornell single"; VKRT_NO_SPECTRAL_MEMO=$off timeout 150 python tests/perf_probe.py --mode single --frames 3 --spp 16 2>&1 | tail -1
  echo "== NO_MEMO=$off inst:1000 hero"; VKRT_NO_SPECTRAL_MEMO=$off timeout 150 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 2>&1 | tail -1
done
echo "== cornell rgb"; timeout 150 python tests/perf_probe.py --mode rgb --frames 3 --spp 16 2>&1 | tail -1
echo "== soup"; timeout 200 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -1
} > gpurun_out/r02l_memo_ab.txt 2>&1
cat gpurun_out/r02l_memo_ab.txt
