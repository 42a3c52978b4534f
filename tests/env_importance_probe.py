"""Development probe (not a test): error of the default estimator and of the environment-importance-sampling extension
(VKRT_CUDA_FLAG_ENV_IMPORTANCE) at equal sample counts on tests/scenes.py:sunlit, against a long default render.

  python tests/env_importance_probe.py [--mode rgb|hero] [--lamp]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="rgb")
    ap.add_argument("--lamp", action="store_true")
    ap.add_argument("--max-depth", type=int, default=0)
    ap.add_argument("--floor-only", action="store_true")
    ap.add_argument("--roughness", type=float, default=-1.0)
    ap.add_argument("--w", type=int, default=192)
    ap.add_argument("--h", type=int, default=128)
    a = ap.parse_args()
    prep = scenes.sunlit(a.w, a.h, spp=64, lamp=a.lamp, balls=not a.floor_only)
    table = None
    if a.max_depth:
        prep["sceneData"]["rrMaxDepth"] = a.max_depth
    if a.roughness >= 0:
        prep["materials"]["roughness"][:] = a.roughness
    if a.mode != "rgb":
        prep["sceneData"]["packedRenderSettings"] = H.hr.pack_render_settings(0, 1, 1)
        table = scenes.rgb2spec()

    def render(flags, frames):
        g = H.CudaBackend(flags=flags | 4)
        g.upload(prep, rgb2spec=table)
        g.resize(a.w, a.h)
        t0 = time.time()
        g.render(prep["sceneData"], frames=frames)
        img = g.read(H.AOV_ACCUM)[..., :3].astype(np.float64)
        dt = time.time() - t0
        g.close()
        return img, dt

    ref, _ = render(0, 128)  # 8192 spp, the reference's estimator
    print("mode %s lamp %d: reference mean %s" % (a.mode, a.lamp, ref.mean(axis=(0, 1))))
    for frames in (1, 4, 16):
        d, td = render(0, frames)
        e, te = render(64, frames)
        rd = np.sqrt(((d - ref) ** 2).mean()) / ref.mean()
        re_ = np.sqrt(((e - ref) ** 2).mean()) / ref.mean()
        print("%5d spp: default rRMSE %.4f (%.0f ms)  env-importance rRMSE %.4f (%.0f ms)  mean ratio %s (default %s)" % (
            64 * frames, rd, td * 1e3, re_, te * 1e3, np.round(e.mean(axis=(0, 1)) / ref.mean(axis=(0, 1)), 4),
            np.round(d.mean(axis=(0, 1)) / ref.mean(axis=(0, 1)), 4)))
    # where is the difference? rows of the image (top = sky, bottom = floor)
    e, _ = render(64, 64)
    for name, sl in (("sky rows 0-15", slice(0, a.h // 8)), ("upper", slice(a.h // 8, a.h // 2)), ("lower", slice(a.h // 2, a.h))):
        print("  %-14s IS/default mean ratio %s" % (name, np.round(e[sl].mean(axis=(0, 1)) / ref[sl].mean(axis=(0, 1)), 4)))


if __name__ == "__main__":
    main()
