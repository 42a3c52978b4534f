# BASELINE.json configs C1..C5 through the C host CLI (vkrt_b200/vkrt) on one GPU; run under gpurun. Logs + PNGs land in gpurun_out/.
set -x
V="vkrt_b200/vkrt ${BUILDER:+--builder $BUILDER}"
TAG=${1:-r02}
O=gpurun_out/${TAG}_configs
mkdir -p $O
$V --scene assets/scenes/cornell.json --spectral 0 --render-width 512 --render-height 512 --render-samples 64 --render-output $O/c1_cornell_rgb_512.png > $O/c1.log 2>&1
$V --scene assets/scenes/cornell.json --spectral 2 --render-width 1920 --render-height 1080 --render-samples 1024 --render-output $O/c2_cornell_hero_1080p.png > $O/c2.log 2>&1
$V --soup 10000000 --spectral 0 --render-width 1920 --render-height 1080 --render-samples 256 --render-output $O/c3_soup10m_1080p.png > $O/c3.log 2>&1
$V --instanced 1000 assets/models/suzanne.glb --spectral 2 --render-width 1920 --render-height 1080 --render-samples 256 --render-output $O/c4_suzanne1000_hero_1080p.png > $O/c4.log 2>&1
$V --scene assets/scenes/cornell.json --spectral 2 --render-width 3840 --render-height 2160 --render-samples ${C5_SAMPLES:-4096} > $O/c5.log 2>&1
$V --scene assets/scenes/prism.json --render-width 960 --render-height 540 --render-samples 256 --render-output $O/prism_540p.png > $O/prism.log 2>&1
$V --scene assets/scenes/caustics.json --render-width 960 --render-height 540 --render-samples 256 --render-output $O/caustics_540p.png > $O/caustics.log 2>&1
tail -n 4 $O/*.log
