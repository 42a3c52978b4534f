# A/B of libvkrt_cuda.so variants on the two traversal probes (run under gpurun): tools/ab_trace.sh v1 v2 ...  ("main" = the in-tree build)
for v in "$@"; do
  if [ "$v" = main ]; then unset VKRT_CUDA_LIB; else export VKRT_CUDA_LIB=$PWD/variants/$v/libvkrt_cuda.so; fi
  echo "== $v hero"; timeout 150 python tests/perf_probe.py --mode hero --frames 3 --spp 16 2>&1 | tail -2
  echo "== $v soup"; timeout 150 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 2>&1 | tail -2
done
