"""Analytic checks of the oracle's renderer (CPU, seconds): cases whose answer is known in closed form, so that the checker itself
is checked independently of the CUDA path (SURVEY §8c v-vi)."""
import ctypes as C

import numpy as np
import pytest

import harness as H
import scenes

hr = H.hr


def _render(scene, w, h, spp, frames=1, spectral=None):
    prep = scene.prepare(w, h)
    prep["sceneData"]["samplesPerPixel"] = spp
    o = H.OracleBackend()
    o.upload(prep, rgb2spec=scenes.rgb2spec() if spectral is not None else None)
    o.resize(w, h)
    o.render(prep["sceneData"], frames=frames)
    return o.read(H.AOV_ACCUM)


def _material(**kw):
    m = hr.default_material()
    for k, v in kw.items():
        m[k] = v
    return hr.sanitize_material(m)


def test_constant_environment_is_returned_exactly():
    st = hr.Settings(environment_color=(0.3, 0.5, 0.7), environment_strength=2.0)
    mats = np.zeros(1, hr.MATERIAL)
    mats[0] = hr.default_material()
    acc = _render(hr.Scene(meshes=[], materials=mats, settings=st), 16, 12, 4)
    assert np.allclose(acc[..., :3], np.array([0.6, 1.0, 1.4], np.float32), rtol=1e-6)
    assert np.all(acc[..., 3] == 4.0)


def test_emitter_seen_directly_has_its_radiance():
    """A quad light filling the view, black environment: every path adds the emission once at depth 0 (MIS weight 1) and nothing
    afterwards (a light sample on the emitter's own plane has cos = 0, light_sampling.slang:52)."""
    q = hr.quad_mesh("light", 50.0)
    q.material_index = 1
    mats = np.zeros(2, hr.MATERIAL)
    mats[0] = hr.default_material()
    mats[1] = _material(emissionLuminance=5.0, emissionColor=(1.0, 0.5, 0.25), baseColor=(0, 0, 0), specular=0.0, roughness=1.0)
    st = hr.Settings(camera_pos=(0, 0, 3), camera_target=(0, 0, 0), camera_up=(0, 1, 0), environment_strength=0.0)
    acc = _render(hr.Scene(meshes=[q], materials=mats, settings=st), 16, 16, 8)
    assert np.allclose(acc[..., :3], np.array([5.0, 2.5, 1.25], np.float32), rtol=1e-5)


def test_lambert_plane_under_uniform_sky_reflects_albedo_times_radiance():
    """Pure Lambert (roughness 1, specular 0) convex receiver under a constant environment: cosine sampling makes every sample
    return albedo * L exactly, so the estimate has no variance."""
    q = hr.quad_mesh("floor", 200.0)
    q.material_index = 1
    mats = np.zeros(2, hr.MATERIAL)
    mats[0] = hr.default_material()
    mats[1] = _material(baseColor=(0.5, 0.25, 0.75), roughness=1.0, specular=0.0, metallic=0.0)
    st = hr.Settings(camera_pos=(0, 0, 2), camera_target=(0, 0, 0), camera_up=(0, 1, 0), environment_color=(1, 1, 1), environment_strength=1.5)
    acc = _render(hr.Scene(meshes=[q], materials=mats, settings=st), 12, 12, 16)
    assert np.allclose(acc[..., :3], np.array([0.75, 0.375, 1.125], np.float32), rtol=2e-3)


def test_nee_only_plus_bsdf_only_equals_full():
    means = {}
    for mode in (0, 4, 5):
        sc = hr.cornell_scene()
        sc.settings.debug_mode = mode
        means[mode] = _render(sc, 24, 24, 48, frames=2)[..., :3].astype(np.float64).mean()
    assert abs(means[4] + means[5] - means[0]) < 0.05 * means[0], means


@pytest.mark.parametrize("sampling", [0, 1])
def test_spectral_accumulation_matches_quadrature(sampling):
    """Spectral modes accumulate XYZ = integral of S(lambda) * cmf(lambda) / CIE_Y_INTEGRAL, estimated with uniformly sampled
    (van-der-Corput rotated) wavelengths (wavelength.slang:29-47, spectral.slang:99-113). For a constant white environment the
    integral is computed here by 0.25 nm quadrature over the oracle's own rgb2spec spectrum and CIE fit; the rendered mean must agree
    to 1.5 % (262 k independent wavelength samples, relative standard error ~0.3 %)."""
    st = hr.Settings(environment_color=(1, 1, 1), environment_strength=1.0, render_mode=1, spectral_sampling=sampling)
    mats = np.zeros(1, hr.MATERIAL)
    mats[0] = hr.default_material()
    scene = hr.Scene(meshes=[], materials=mats, settings=st)
    acc = _render(scene, 32, 32, 64, frames=4, spectral=True)
    xyz = acc[..., :3].reshape(-1, 3).astype(np.float64).mean(axis=0)
    lib = H.oracle_lib()
    o = H.OracleBackend()
    o.upload(scene.prepare(8, 8), rgb2spec=scenes.rgb2spec())
    white = np.ones(3, np.float32)
    lam = np.arange(360.0, 830.0, 0.25) + 0.125
    want = np.zeros(3)
    cmf = np.zeros(3, np.float32)
    for l in lam:
        s_l = lib.oracle_rgb2spec_eval(o.ctx, white.ctypes.data_as(C.c_void_p), C.c_float(l))
        lib.oracle_spectral_xyz(C.c_float(l), cmf.ctypes.data_as(C.c_void_p))
        want += s_l * cmf.astype(np.float64) * 0.25
    want /= 106.9461715
    assert np.allclose(xyz, want, rtol=0.015), (xyz, want)
    # and the white point lands near (1,1,1) linear sRGB after the film's Bradford E->D65 step (fit + table error: within 10 %)
    rgb = np.zeros(3, np.float32)
    x32 = xyz.astype(np.float32)
    lib.oracle_xyz_to_srgb(x32.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p))
    assert np.allclose(rgb, 1.0, atol=0.1), rgb
