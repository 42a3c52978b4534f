# A/B of the queue cache hints (variants/sh, tr, both: tools/build_variant.sh NAME -DSHADE_STREAM_HINTS / -DTRACE_STREAM_HINTS at the time; the hints are now the default) and of the
# persisting-L2 window over the rgb2spec cells (VKRT_L2_PERSIST=1), then a parity subset on the variant with both hints. Run under gpurun.
mkdir -p gpurun_out
OUT=gpurun_out/r03i_hints_ab.txt
: > $OUT
probe() {  # name lib l2 mode
  if [ "$2" = main ]; then unset VKRT_CUDA_LIB; else export VKRT_CUDA_LIB=$PWD/variants/$2/libvkrt_cuda.so; fi
  echo "== $1 ($4)" >> $OUT
  VKRT_L2_PERSIST=$3 timeout 30 python tests/perf_probe.py --mode $4 --frames 4 --spp 16 2>&1 | tail -3 >> $OUT
}
probe main main 0 hero
probe both both 0 hero
probe shade-hints sh 0 hero
probe trace-hints tr 0 hero
probe main+l2persist main 1 hero
probe both+l2persist both 1 hero
export VKRT_CUDA_LIB=$PWD/variants/both/libvkrt_cuda.so
timeout 70 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cornell_spectral or cornell_rgb or glass_medium or running_mean or spectral_memo" > gpurun_out/r03i_hints_parity.log 2>&1
tail -2 gpurun_out/r03i_hints_parity.log >> $OUT
unset VKRT_CUDA_LIB
probe main main 0 rgb
probe both both 0 rgb
cat $OUT
