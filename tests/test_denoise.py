"""The save-time denoise stage (feature AOVs -> Open Image Denoise) against the reference's OWN source.

Open Image Denoise is a third-party neural filter that is not installed here; what the stage owns is everything around it. Both sides
run over the same stand-in library (oracle/ref_host/fake_oidn.c: a deterministic function in which every attached image and every
filter parameter is visible, plus failure injection and a call log):

  reference : /root/reference/src/core/utility/{denoise.c, export/image.c} compiled where they lie into oracle/_ref/libvkrt_refexport.so
              (prepareLinearRenderOutput, convertLinearToDisplayRGBA16), linked against the stand-in;
  product   : vkrt_b200/host/{denoise.c, export.c} in libvkrt_host.so, which binds the library at run time (VKRT_OIDN_LIBRARY).

Outputs must be byte-identical and the logged filter calls equal, for healthy runs, every failure mode, missing features and debug views.
No GPU is involved: these are the host stages VKRT_saveRenderImageEx runs between the film read-backs and the file writers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H
import refpin

FAKE = os.path.join(refpin.REF_DIR, "libfake_oidn.so")


def _fake():
    if not os.path.exists(FAKE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(H.ROOT, "oracle"), "fakeoidn"])
    return FAKE


@pytest.fixture()
def libs(monkeypatch, tmp_path):
    ref = refpin._load("libvkrt_refexport.so")
    host = C.CDLL(os.path.join(H.ROOT, "vkrt_b200", "libvkrt_host.so"))
    host.vkrtHostPrepareLinearOutput.restype = C.c_int
    host.vkrtHostPrepareLinearOutput.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    host.vkrtHostLinearToDisplay16.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_void_p]
    ref.refexport_prepare_linear.restype = C.c_int
    ref.refexport_prepare_linear.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    ref.refexport_linear_to_display.restype = C.c_int
    ref.refexport_linear_to_display.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_void_p]
    monkeypatch.setenv("VKRT_OIDN_LIBRARY", _fake())
    monkeypatch.delenv("FAKE_OIDN_FAIL", raising=False)
    host.vkrtHostResetDenoiser()
    yield ref, host, monkeypatch, tmp_path
    host.vkrtHostResetDenoiser()


def _film(w, h, seed, coverage=True, hostile=True):
    """An accumulation buffer and two RGBA16F feature AOVs with everything the preparation code branches on."""
    rng = np.random.default_rng(seed)
    beauty = (rng.random((h, w, 4), dtype=np.float32) * 4.0).astype(np.float32)
    beauty[..., 3] = 64.0   # sample count
    albedo = rng.random((h, w, 4)).astype(np.float16)
    normal = (rng.random((h, w, 4)) * 2.0 - 1.0).astype(np.float16)
    albedo[..., 3] = 1.0
    normal[..., 3] = 1.0
    if hostile:
        beauty[0, 0, 0] = np.nan
        beauty[1, 2, 1] = np.inf
        beauty[2, 1, 2] = -np.inf
        beauty[3, 3, :3] = -0.25
        albedo[0, 1, :3] = -0.5                      # clamped to 0
        albedo[1, 1, 3] = 0.0                        # no weight: cleared
        albedo[2, 2, 3] = -1.0
        albedo[3, 0, 0] = np.float16(6.0e-8)         # subnormal half
        albedo[3, 1, 1] = np.inf
        normal[0, 2, :3] = 0.0                       # zero-length normal
        normal[1, 3, :3] = np.float16(1e-7)          # below the normalisation threshold
        normal[2, 0, 3] = 0.0
        normal[3, 2, :3] = (3.0, -4.0, 12.0)
    if not coverage:
        albedo[..., 3] = 0.0
        normal[..., 3] = -2.0
    return beauty, albedo.view(np.uint16).copy(), normal.view(np.uint16).copy()


def _both(ref, host, beauty, albedo, normal, spectral, debug, denoise, fallback, log_dir, monkeypatch, features=True):
    h, w = beauty.shape[:2]
    outs, rcs, logs = [], [], []
    for side in ("reference", "product"):
        log = os.path.join(str(log_dir), side + ".log")
        if os.path.exists(log):
            os.remove(log)
        monkeypatch.setenv("FAKE_OIDN_LOG", log)
        a = albedo.ctypes.data_as(C.c_void_p) if features else None
        n = normal.ctypes.data_as(C.c_void_p) if features else None
        if side == "reference":
            out = np.zeros((h, w, 4), np.float32)
            rc = ref.refexport_prepare_linear(beauty.ctypes.data_as(C.c_void_p), a, n, w, h, 1 if spectral else 0, debug, int(denoise), int(fallback), out.ctypes.data_as(C.c_void_p))
        else:
            out = beauty.copy()
            note = C.create_string_buffer(256)
            # the product decides "denoise and no debug view" in VKRT_saveRenderImageEx before it calls the stage
            rc = host.vkrtHostPrepareLinearOutput(out.ctypes.data_as(C.c_void_p), a, n, w, h, int(spectral), int(denoise and debug == 0), int(fallback), note, 256)
        outs.append(out)
        rcs.append(rc)
        logs.append(open(log).read() if os.path.exists(log) else "")
    return outs, rcs, logs


@pytest.mark.parametrize("spectral", [False, True])
@pytest.mark.parametrize("size", [(16, 12), (33, 7)])
def test_denoised_output_is_byte_identical_to_the_reference_stage(libs, spectral, size):
    ref, host, mp, tmp = libs
    beauty, albedo, normal = _film(size[0], size[1], seed=7 + size[0])
    outs, rcs, logs = _both(ref, host, beauty, albedo, normal, spectral, 0, True, True, tmp, mp)
    assert rcs == [1, 1]
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    # three filter runs, in the reference's order and with its parameters: albedo prefilter, normal prefilter, beauty with clean features
    assert logs[0] == logs[1] and logs[0].count("\n") == 3
    lines = logs[0].splitlines()
    assert lines[0].startswith("RT main=albedo albedo=0 normal=0") and "hdr=0 srgb=0 cleanAux=0 quality=6" in lines[0]
    assert lines[1].startswith("RT main=normal")
    assert lines[2].startswith("RT main=color albedo=1 normal=1") and "hdr=1 srgb=0 cleanAux=1 quality=6" in lines[2]
    assert np.isfinite(outs[1]).all() and np.all(outs[1][..., 3] == 1.0)
    assert not np.array_equal(outs[1][..., :3], beauty[..., :3])   # the filter ran


def test_without_feature_coverage_or_without_features_only_the_beauty_is_filtered(libs):
    ref, host, mp, tmp = libs
    beauty, albedo, normal = _film(12, 9, seed=3, coverage=False)
    for features in (True, False):
        outs, rcs, logs = _both(ref, host, beauty, albedo, normal, False, 0, True, True, tmp, mp, features=features)
        assert rcs == [1, 1] and np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
        assert logs[0] == logs[1] and logs[0].splitlines() == [logs[0].splitlines()[0]] and "main=color albedo=0 normal=0" in logs[0] and "cleanAux=0" in logs[0]


@pytest.mark.parametrize("fail", ["execute", "prefilter", "device", "read"])
@pytest.mark.parametrize("fallback", [True, False])
def test_failure_modes_follow_the_reference(libs, fail, fallback):
    ref, host, mp, tmp = libs
    beauty, albedo, normal = _film(10, 8, seed=11)
    mp.setenv("FAKE_OIDN_FAIL", fail)
    outs, rcs, logs = _both(ref, host, beauty, albedo, normal, True, 0, True, fallback, tmp, mp)
    assert rcs[0] == rcs[1] and logs[0] == logs[1]
    if fail == "prefilter":     # raw features, cleanAux off, the beauty pass still runs
        assert rcs == [1, 1] and "main=color albedo=1 normal=1" in logs[0] and "cleanAux=0" in logs[0].splitlines()[-1]
    else:
        assert rcs == ([1, 1] if fallback else [0, 0])
    if rcs[0]:
        assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_debug_views_and_disabled_denoiser_leave_the_image_raw(libs):
    ref, host, mp, tmp = libs
    beauty, albedo, normal = _film(9, 9, seed=5)
    for debug, denoise in ((3, True), (0, False)):
        outs, rcs, logs = _both(ref, host, beauty, albedo, normal, True, debug, denoise, True, tmp, mp)
        assert rcs == [1, 1] and logs == ["", ""]
        assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_a_missing_library_saves_the_raw_image_with_a_reason(libs):
    """The reference links OIDN at build time; the product binds it at run time. Without the library the stage must behave like a filter
    that failed: raw image when a fallback is allowed (what the reference produces when device creation fails), an error otherwise."""
    ref, host, mp, tmp = libs
    beauty, albedo, normal = _film(8, 8, seed=2)
    mp.setenv("FAKE_OIDN_FAIL", "device")
    expected = np.zeros_like(beauty)
    assert ref.refexport_prepare_linear(beauty.ctypes.data_as(C.c_void_p), albedo.ctypes.data_as(C.c_void_p), normal.ctypes.data_as(C.c_void_p), 8, 8, 0, 0, 1, 1,
                                        expected.ctypes.data_as(C.c_void_p)) == 1
    mp.delenv("FAKE_OIDN_FAIL")
    mp.setenv("VKRT_OIDN_LIBRARY", "/nonexistent/libOpenImageDenoise.so.2")
    host.vkrtHostResetDenoiser()
    for fallback, want in ((1, 1), (0, 0)):
        out = beauty.copy()
        note = C.create_string_buffer(256)
        rc = host.vkrtHostPrepareLinearOutput(out.ctypes.data_as(C.c_void_p), albedo.ctypes.data_as(C.c_void_p), normal.ctypes.data_as(C.c_void_p), 8, 8, 0, 1, fallback, note, 256)
        assert rc == want and b"Open Image Denoise is not installed" in note.value
        if rc:
            assert np.array_equal(out.view(np.uint32), expected.view(np.uint32))


@pytest.mark.parametrize("tone,debug", [(1, 0), (0, 0), (1, 7)])
@pytest.mark.parametrize("exposure", [1.0, 0.37, 5.5, -1.0, float("nan")])
def test_cpu_tone_map_of_the_denoised_image(libs, tone, debug, exposure):
    ref, host, mp, tmp = libs
    rng = np.random.default_rng(17)
    w, h = 64, 48
    linear = np.concatenate([rng.random((h, w, 3), dtype=np.float32) ** 4 * 8.0, np.ones((h, w, 1), np.float32)], axis=2).astype(np.float32)
    linear[0, :8, 0] = [0.0, 1e-7, 0.0031308, 0.0031309, 1.0, 65504.0, -0.5, 0.5]
    a, b = np.zeros((h, w, 4), np.uint16), np.zeros((h, w, 4), np.uint16)
    assert ref.refexport_linear_to_display(linear.ctypes.data_as(C.c_void_p), w, h, tone, exposure, debug, a.ctypes.data_as(C.c_void_p)) == 1
    host.vkrtHostLinearToDisplay16(linear.ctypes.data_as(C.c_void_p), w, h, tone, exposure, debug, b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, b)
