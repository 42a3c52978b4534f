#!/usr/bin/env python3
"""First-light GPU check (run under gpurun): oracle vs CUDA on the procedural Cornell scene.
Writes a summary to gpurun_out/first_light.json and PNGs for eyeballing."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H

hr = H.hr
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = {}
scene = hr.cornell_scene()
W = Hh = 256
prep = scene.prepare(W, Hh)
sd = prep["sceneData"].copy()
sd["samplesPerPixel"] = 8

o = H.OracleBackend()
o.upload(prep); o.resize(W, Hh)
g = H.CudaBackend()
g.upload(prep); g.resize(W, Hh)
bs = g.build_stats
out["build"] = dict(ms=bs.buildMs, blas=bs.uniqueGeometries, nodes=bs.bvh8NodeCount, tris=bs.triangleCount)
print("build", out["build"], flush=True)

o.trace_primary(sd); g.trace_primary(sd)
for which, nm in ((H.AOV_HITID_CENTER, "center"), (H.AOV_HITID_S0, "s0")):
    a, b = o.read(which), g.read(which)
    same = np.all(a == b, axis=-1)
    out["hitid_" + nm] = dict(match=float(same.mean()), mismatches=int((~same).sum()))
    print("hitid", nm, out["hitid_" + nm], flush=True)
ta, tb = o.read(H.AOV_HIT_TUV), g.read(H.AOV_HIT_TUV)
out["tuv_bitexact"] = float(np.all(ta.view(np.uint32) == tb.view(np.uint32), axis=-1).mean())
print("tuv bit-exact fraction", out["tuv_bitexact"], flush=True)

t = time.time(); o.render(sd, frames=2); out["oracle_s"] = time.time() - t
t = time.time(); g.render(sd, frames=2); out["cuda_s"] = time.time() - t
fs = g.frame_stats[-1]
out["frame"] = dict(ms=fs.frameMs, launches=fs.kernelLaunches, paths=fs.paths, ext=fs.extensionRays, shadow=fs.shadowRays)
print("frame", out["frame"], flush=True)
for which, nm in ((H.AOV_ACCUM, "accum"), (H.AOV_ALBEDO, "albedo"), (H.AOV_NORMAL, "normal"), (H.AOV_OUTPUT, "output")):
    a, b = o.read(which), g.read(which)
    out["cmp_" + nm] = H.compare_images(a.astype(np.float32), b.astype(np.float32), nm)
    print(out["cmp_" + nm], flush=True)
H.write_png(os.path.join(ROOT, "gpurun_out/first_light_cuda.png"), (g.read(H.AOV_OUTPUT)[..., :3] >> 8).astype(np.uint8))
H.write_png(os.path.join(ROOT, "gpurun_out/first_light_oracle.png"), (o.read(H.AOV_OUTPUT)[..., :3] >> 8).astype(np.uint8))
json.dump(out, open(os.path.join(ROOT, "gpurun_out/first_light.json"), "w"), indent=1)
