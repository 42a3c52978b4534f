# A/B of the hierarchy builders (run under gpurun): LBVH (flag 128) against PLOC (default) on the traversal probes, with visit counters.
for f in 128 0; do
  for acc in 8 16; do
    echo "== builder flag $f, accel flag $acc: cornell hero"; timeout 200 python tests/perf_probe.py --mode hero --frames 3 --spp 16 --flags $((f + acc)) 2>&1 | tail -3
    echo "== builder flag $f, accel flag $acc: cornell hero, counted"; timeout 200 python tests/perf_probe.py --mode hero --frames 2 --spp 4 --count --flags $((f + acc)) 2>&1 | tail -1
  done
  echo "== builder flag $f: soup 10M"; timeout 300 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 3 --spp 4 --flags $f 2>&1 | tail -4
  echo "== builder flag $f: soup 10M counted"; timeout 300 python tests/perf_probe.py --scene soup:10000000 --mode rgb --frames 2 --spp 2 --count --flags $f 2>&1 | tail -1
  echo "== builder flag $f: inst 1000"; timeout 200 python tests/perf_probe.py --scene inst:1000 --mode hero --frames 3 --spp 16 --flags $f 2>&1 | tail -3
done
