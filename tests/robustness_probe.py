"""Development probe (not a test): hostile but well-formed ABI inputs must neither fault the device nor hang.
  python tests/robustness_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402
import scenes  # noqa: E402


def attempt(label, prep, accel, sd=None):
    g = H.CudaBackend(flags=accel)
    try:
        g.upload(prep)
        g.resize(64, 64)
        g.render(sd if sd is not None else prep["sceneData"], frames=2)
        img = g.read(H.AOV_ACCUM)
        print("%-28s accel %2d rendered; finite pixels %.4f" % (label, accel, np.isfinite(img[..., :3]).all(axis=-1).mean()), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%-28s accel %2d refused: %s" % (label, accel, str(e)[:110]), flush=True)
    g.close()


for accel in (8, 16):
    for label, val in (("nan world matrix", np.nan), ("inf world matrix", np.inf), ("zero-scale instance", 0.0)):
        prep = scenes.cornell(64, 64, spp=2)
        w = np.array(prep["world3x4"], dtype=np.float32).copy()
        if label.startswith("zero"):
            w.reshape(-1, 12)[2, [0, 1, 2, 4, 5, 6, 8, 9, 10]] = 0.0
        else:
            w.reshape(-1, 12)[2, 3] = val
        prep["world3x4"] = w
        attempt(label, prep, accel)
    prep = scenes.cornell(64, 64, spp=2)
    sd = prep["sceneData"].copy()
    vi = np.array(sd["viewInverse"], dtype=np.float32).copy()
    vi.reshape(-1)[12] = np.nan
    sd["viewInverse"] = vi
    attempt("nan camera", prep, accel, sd)
    sd = prep["sceneData"].copy()
    sd["samplesPerPixel"] = 0
    attempt("0 spp", prep, accel, sd)
    sd = prep["sceneData"].copy()
    sd["rrMaxDepth"] = 0
    attempt("max depth 0", prep, accel, sd)
    sd = prep["sceneData"].copy()
    sd["rrMaxDepth"] = 1000
    sd["rrMinDepth"] = 999
    attempt("max depth 1000", prep, accel, sd)
print("probe finished")
