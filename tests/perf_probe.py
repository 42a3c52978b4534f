"""Development probe (not a test, not the bench): renders the bundled cornell scene through the C ABI at bench-like
sizes and prints per-frame device timings and ray counts. Lives under tests/ because it prepares the scene with the
oracle-side host reference (oracle/host_ref.py), which product code may not import.

  python tests/perf_probe.py [--w 1920 --h 1080 --spp 16 --frames 3 --mode hero|single|rgb --count --scene cornell|soup:N|inst:N]
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import harness as H  # noqa: E402
import scenes  # noqa: E402

hr = H.hr


def build_scene(name, w, h):
    if name == "cornell":
        sc = hr.load_scene_json(os.path.join(H.ROOT, "assets", "scenes", "cornell.json"))
        return sc
    if name.startswith("soup:"):
        return hr.soup_scene(int(name.split(":")[1]))
    if name.startswith("json:"):
        return hr.load_scene_json(os.path.join(H.ROOT, "assets", "scenes", name.split(":")[1] + ".json"))
    raise SystemExit("unknown scene " + name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--mode", default="hero")
    ap.add_argument("--scene", default="cornell")
    ap.add_argument("--count", action="store_true")
    ap.add_argument("--max-paths", type=int, default=0)
    ap.add_argument("--png", default="")
    ap.add_argument("--flags", type=int, default=0, help="extra VKRT_CUDA_FLAG_* bits (2 = no material sort, 8 = two-level, 16 = flat)")
    a = ap.parse_args()

    t0 = time.time()
    if a.scene.startswith("inst:"):
        prep = scenes.instanced(a.w, a.h, count=int(a.scene.split(":")[1]), spp=a.spp)
    else:
        sc = build_scene(a.scene, a.w, a.h)
        sc.settings.render_mode = 0 if a.mode == "rgb" else 1
        sc.settings.spectral_sampling = 1 if a.mode == "hero" else 0
        prep = sc.prepare(a.w, a.h)
    prep["sceneData"]["samplesPerPixel"] = a.spp
    print("scene prep %.2fs: %d verts %d indices %d instances %d emissive tris" % (
        time.time() - t0, len(prep["vertices"]), len(prep["indices"]), len(prep["meshInfos"]), prep["lights"]["triangleCount"]), flush=True)
    g = H.CudaBackend(flags=(1 if a.count else 0) | 4 | a.flags, max_paths=a.max_paths or a.w * a.h * a.spp)
    g.upload(prep, rgb2spec=scenes.rgb2spec() if a.mode != "rgb" else None)
    bs = g.build_stats
    print("build: %.3f ms (blas %.3f tlas %.3f) geometries %d instances %d tris %d nodes %d bytes %d" % (
        bs.buildMs, bs.blasMs, bs.tlasMs, bs.uniqueGeometries, bs.instanceCount, bs.triangleCount, bs.bvh8NodeCount, bs.accelBytes))
    g.resize(a.w, a.h)
    sd = prep["sceneData"].copy()
    for f in range(a.frames):
        sd["frameNumber"] = f
        t = time.time()
        g.render_frame(sd)
        wall = (time.time() - t) * 1e3
        st = g.frame_stats[-1]
        rays = st.extensionRays + st.shadowRays
        msg = "frame %d: %.2f ms (wall %.2f) trace %.2f shade %.2f launches %d paths %d ext %d shadow %d -> %.1f Mpaths/s, %.1f Mrays/s(trace-only)" % (
            f, st.frameMs, wall, st.traceMs, st.shadeMs, st.kernelLaunches, st.paths, st.extensionRays, st.shadowRays,
            st.paths / st.frameMs / 1e3, rays / max(st.traceMs, 1e-6) / 1e3)
        if a.count:
            msg += " nodes/ray %.1f tris/ray %.1f inst/ray %.2f" % (st.nodesVisited / max(rays, 1), st.trianglesTested / max(rays, 1), st.instancesEntered / max(rays, 1))
        print(msg, flush=True)
    if a.png:
        out = g.read(H.AOV_OUTPUT)
        H.write_png(a.png, (out[..., :3] >> 8).astype(np.uint8))
        print("wrote", a.png)


if __name__ == "__main__":
    main()
