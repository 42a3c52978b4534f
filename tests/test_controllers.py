"""Host-side feedback controllers (vkrt_b200/host/controllers.c) against a numpy restatement of the reference's arithmetic
(src/core/scene/timing.c:121-178 updateAutoSPP, src/core/scene/exposure.c:14-67,139-147,187-222), and their wiring into the frame
protocol on the GPU."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import harness as H

f32 = np.float32


def _lib():
    from vkrt_b200 import host
    lib = host.load_host_library()
    lib.vkrtAutoSPPStep.restype = C.c_uint32
    lib.vkrtAutoSPPStep.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_uint32]
    lib.vkrtAutoExposureStep.restype = C.c_int
    lib.vkrtAutoExposureStep.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_uint32, C.c_float, C.POINTER(C.c_float)]
    lib.vkrtAutoExposureProbePixels.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    return lib


def ref_auto_spp(control, target, measured, spp):
    """timing.c:121-178 in fp32."""
    if target <= 0 or measured <= 0:
        return control, spp
    spp = max(spp, 1)
    sppf = f32(spp)
    per = f32(measured) / sppf
    control = per if control <= 0 else f32(f32(control * f32(1 - f32(0.35))) + f32(per * f32(0.35)))
    desired = f32(f32(f32(target) * f32(0.90)) / control)
    desired = min(max(desired, f32(1.0)), f32(2048.0))
    delta = f32(desired - sppf)
    deadband = max(f32(sppf * f32(0.18 if delta > 0 else 0.08)), f32(1.0))
    if abs(delta) <= deadband:
        return control, spp
    if delta > 0:
        nxt = int(math.floor(min(desired, f32(math.ceil(f32(sppf * f32(1.25)))))))
        if nxt <= spp and spp < 2048:
            nxt = spp + 1
    else:
        nxt = int(math.ceil(max(desired, f32(math.floor(f32(sppf * f32(0.60)))))))
        if nxt >= spp and spp > 1:
            nxt = spp - 1
    return control, min(max(nxt, 1), 2048)


def test_auto_spp_step_matches_the_reference_arithmetic():
    lib = _lib()
    rng = np.random.default_rng(5)
    for trial in range(40):
        ms_per_spp = float(rng.uniform(0.05, 30.0))
        target = float(1000.0 / rng.choice([30, 60, 120, 240]))
        spp, control = int(rng.integers(1, 64)), f32(0.0)
        c_control = C.c_float(0.0)
        for step in range(60):
            measured = f32(ms_per_spp * spp * rng.uniform(0.9, 1.1))
            control, want = ref_auto_spp(control, f32(target), measured, spp)
            got = lib.vkrtAutoSPPStep(C.byref(c_control), C.c_float(target), C.c_float(float(measured)), C.c_uint32(spp))
            assert got == want, (trial, step, spp, got, want)
            assert abs(c_control.value - float(control)) <= 1e-6 * max(1.0, float(control))
            spp = got
        # converged to the budget within the controller's dead band (8 % of spp, at least one sample per pixel)
        assert spp <= 2 or ms_per_spp * (spp - 1) <= target * 1.1   # (the dead band of one sample keeps 2 from ever dropping to 1, as upstream)
    # degenerate inputs leave the state alone
    c = C.c_float(0.0)
    assert lib.vkrtAutoSPPStep(C.byref(c), C.c_float(0.0), C.c_float(5.0), C.c_uint32(7)) == 7 and c.value == 0.0
    assert lib.vkrtAutoSPPStep(C.byref(c), C.c_float(16.0), C.c_float(0.0), C.c_uint32(7)) == 7 and c.value == 0.0
    assert lib.vkrtAutoSPPStep(C.byref(c), C.c_float(16.0), C.c_float(1e-3), C.c_uint32(2048)) == 2048


def test_auto_exposure_step_and_probe_grid():
    lib = _lib()
    rng = np.random.default_rng(6)
    filt, c_filt, exposure = f32(0.0), C.c_float(0.0), f32(1.0)
    for step in range(30):
        s = rng.uniform(0.0, 4.0, (256, 4)).astype(np.float32)
        s[rng.integers(0, 256, 5), 0] = np.nan                    # non-finite samples are skipped (exposure.c:49)
        lum = (f32(0.2126) * s[:, 0] + f32(0.7152) * s[:, 1] + f32(0.0722) * s[:, 2]).astype(np.float32)
        ok = np.isfinite(lum)
        acc = f32(0.0)
        for v in lum[ok]:
            acc = f32(acc + v)
        avg = f32(acc / f32(ok.sum()))
        filt = avg if filt <= 0 else f32(f32(filt * f32(1 - f32(0.18))) + f32(avg * f32(0.18)))
        want = f32(f32(0.18) / max(f32(np.power(filt, f32(0.65))), f32(1e-4)))
        out = C.c_float(-1.0)
        changed = lib.vkrtAutoExposureStep(C.byref(c_filt), s.ctypes.data, 256, C.c_float(float(exposure)), C.byref(out))
        assert abs(c_filt.value - float(filt)) <= 2e-6 * float(filt)
        assert changed == (1 if abs(float(exposure) - float(want)) >= 1e-4 else 0)
        if changed:
            assert abs(out.value - float(want)) <= 4e-6 * float(want)
            exposure = f32(out.value)
    black = np.zeros((256, 4), np.float32)
    assert lib.vkrtAutoExposureStep(C.byref(c_filt), black.ctypes.data, 256, C.c_float(1.0), C.byref(out)) == 0   # no light: keep the exposure
    xy = np.zeros((256, 2), np.uint32)
    lib.vkrtAutoExposureProbePixels(1920, 1080, xy.ctypes.data)
    gx, gy = np.meshgrid(np.arange(16), np.arange(16))
    assert np.array_equal(xy[:, 0].reshape(16, 16), ((2 * gx + 1) * 1920) // 32) and np.array_equal(xy[:, 1].reshape(16, 16), ((2 * gy + 1) * 1080) // 32)
    lib.vkrtAutoExposureProbePixels(3, 1, xy.ctypes.data)
    assert xy[:, 0].max() == 2 and xy[:, 1].max() == 0


@pytest.mark.gpu
def test_auto_spp_and_auto_exposure_in_the_frame_loop():
    from vkrt_b200 import host
    w, h = 640, 360
    hs = host.Host(width=w, height=h)
    lib = hs.lib
    hs.load_scene(os.path.join(H.ROOT, "assets", "scenes", "cornell.json"))
    hs.set_render_mode(0)
    hs.set_samples_per_pixel(1)
    assert lib.VKRT_setAutoSPPTargetFPS(hs.h, C.c_uint32(1000)) == 0 and hs.scene_settings().autoSPPTargetFPS == 360   # clamped (settings.c:73-74)
    assert lib.VKRT_setAutoSPPTargetFPS(hs.h, C.c_uint32(30)) == 0
    assert lib.VKRT_setAutoSPPEnabled(hs.h, C.c_uint8(1)) == 0
    assert lib.VKRT_setAutoExposureEnabled(hs.h, C.c_uint8(1)) == 0
    hs.start_render(w, h, 0)
    spps, total = [], 0
    for _ in range(60):
        spp_before = hs.scene_settings().samplesPerPixel
        hs.draw()
        total += spp_before
        spps.append(spp_before)
        assert hs.render_status().totalSamples == total          # the frame's own spp is what gets counted (frame.c:373-383)
    st = hs.last_frame_stats()
    assert spps[0] == 1 and max(spps) > 4                        # grew from the reset value towards the 33 ms budget
    assert all(b <= math.ceil(a * 1.25) for a, b in zip(spps, spps[1:]))   # at most +25 % per frame
    assert st.frameMs <= 33.4 * 1.1, (spps[-5:], st.frameMs)
    # exposure follows key / L^0.65 of the probe luminance
    lib.vkrt_cuda_read_accum_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    xy = np.zeros((256, 2), np.uint32)
    lib.vkrtAutoExposureProbePixels(w, h, xy.ctypes.data)
    import vkrt_b200
    cl = vkrt_b200.load_library()
    cl.vkrt_cuda_read_accum_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    samples = np.zeros((256, 4), np.float32)
    assert cl.vkrt_cuda_read_accum_samples(C.c_void_p(hs.cuda_context()), xy.ctypes.data, 256, samples.ctypes.data) == 0
    accum = np.zeros((h, w, 4), np.float32)
    assert cl.vkrt_cuda_read_aov(C.c_void_p(hs.cuda_context()), C.c_int(0), accum.ctypes.data_as(C.c_void_p), C.c_size_t(accum.nbytes)) == 0
    assert np.array_equal(samples, accum[xy[:, 1], xy[:, 0]])   # the probe reads exactly those accumulation texels
    lum = float((0.2126 * samples[:, 0] + 0.7152 * samples[:, 1] + 0.0722 * samples[:, 2]).mean())
    want = 0.18 / lum ** 0.65
    assert abs(hs.scene_settings().exposure - want) / want < 0.05   # filtered over the last frames of a converging image
    hs.close()
