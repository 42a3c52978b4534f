mkdir -p gpurun_out
export VKRT_BUILD_VERBOSE=1
{
echo "== cornell hero default"; python tests/perf_probe.py --mode hero --frames 2 --spp 16 2>&1 | grep -E "vkrt build|frame 1|build:"
echo "== cornell hero two-level"; python tests/perf_probe.py --mode hero --frames 2 --spp 16 --flags 8 2>&1 | grep -E "vkrt build|frame 1|build:"
echo "== soup default"; python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 2 --spp 4 2>&1 | grep -E "vkrt build|frame 1|build:"
echo "== inst 1000"; python tests/perf_probe.py --scene inst:1000 --mode hero --frames 2 --spp 16 2>&1 | grep -E "vkrt build|frame 1|build:"
echo "== C3 cli"; vkrt_b200/vkrt --soup 10000000 --spectral 0 --render-width 1920 --render-height 1080 --render-samples 64 2>&1 | tail -6
echo "== C4 cli"; vkrt_b200/vkrt --instanced 1000 assets/models/suzanne.glb --spectral 2 --render-width 1920 --render-height 1080 --render-samples 64 2>&1 | tail -8
echo "== caustics cli"; vkrt_b200/vkrt --scene assets/scenes/caustics.json --render-width 960 --render-height 540 --render-samples 64 2>&1 | tail -6
} > gpurun_out/r02d_builder.txt 2>&1
cat gpurun_out/r02d_builder.txt
unset VKRT_BUILD_VERBOSE
python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r02d_tests.log 2>&1; tail -4 gpurun_out/r02d_tests.log
