"""Self-consistency of the oracle's restatement of vkrt's layered principled BSDF (oracle/shading.h; reference: src/shaders/bsdf/**),
the checks SURVEY §8c(iv) and (vi) ask for in place of upstream unit tests that do not exist:
  * a sampled direction carries weight == f |cos| / pdf and the pdf that eval reports for it;
  * every lobe mix integrates its pdf to 1 over the sphere (the part sampled below the horizon is lost for opaque closures);
  * white furnace: no closure reflects more than it receives;
  * rgb2spec: RGB -> spectrum -> XYZ -> linear sRGB round trip stays within a stated bound."""
import ctypes as C

import numpy as np
import pytest

import harness as H
import scenes

hr = H.hr


def _lib():
    lib = H.oracle_lib()
    lib.oracle_bsdf_eval.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p]
    lib.oracle_bsdf_eval.restype = None
    lib.oracle_bsdf_sample.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p]
    lib.oracle_bsdf_sample.restype = None
    return lib


@pytest.fixture(scope="module")
def oracle():
    o = H.OracleBackend()
    payload, info = scenes.rgb2spec()
    o.check(o.f("set_rgb2spec")(o.ctx, payload.ctypes.data_as(C.c_void_p), C.c_uint32(len(payload)), info), "set_rgb2spec")
    return o


def _material(**kw):
    m = hr.default_material()
    for k, v in kw.items():
        m[k] = v
    return np.ascontiguousarray(hr.sanitize_material(m))


MATERIALS = {
    "lambert": dict(baseColor=(1, 1, 1), roughness=1.0, specular=0.0),
    "default": dict(baseColor=(1, 1, 1)),
    "rough_metal": dict(baseColor=(1, 1, 1), metallic=1.0, roughness=0.5),
    "glossy_plastic": dict(baseColor=(1, 1, 1), roughness=0.25, specular=1.0),
    "clearcoat": dict(baseColor=(1, 1, 1), roughness=0.6, clearcoat=1.0, clearcoatGloss=0.8),
    "sheen": dict(baseColor=(1, 1, 1), roughness=0.8, sheenTintWeight=(1, 1, 1, 1.0), sheenRoughness=0.5),
    "oren_nayar": dict(baseColor=(1, 1, 1), roughness=1.0, diffuseRoughness=0.7, specular=0.0),
}


def _eval(lib, o, m, wo, wi):
    out = np.zeros(4, np.float32)
    lib.oracle_bsdf_eval(o.ctx, m.ctypes.data, np.asarray(wo, np.float32).ctypes.data, np.asarray(wi, np.float32).ctypes.data, 1, out.ctypes.data)
    return out


def _sphere_quadrature(n_theta=96, n_phi=192):
    ct = (np.arange(n_theta) + 0.5) / n_theta * 2 - 1           # uniform in cos(theta): equal solid angle cells
    ph = (np.arange(n_phi) + 0.5) / n_phi * 2 * np.pi
    c, p = np.meshgrid(ct, ph, indexing="ij")
    s = np.sqrt(1 - c * c)
    d = np.stack([s * np.cos(p), s * np.sin(p), c], -1).reshape(-1, 3).astype(np.float32)
    return d, 4 * np.pi / len(d)


@pytest.mark.parametrize("name", sorted(MATERIALS))
def test_sample_weight_and_pdf_agree_with_eval(oracle, name):
    lib = _lib()
    m = _material(**MATERIALS[name])
    rng = np.random.default_rng(11)
    checked = 0
    for trial in range(300):
        ct = rng.uniform(0.1, 1.0)
        wo = np.array([np.sqrt(1 - ct * ct), 0.0, ct], np.float32)
        seed = C.c_uint32(int(rng.integers(1, 2 ** 31)))
        out = np.zeros(8, np.float32)
        lib.oracle_bsdf_sample(oracle.ctx, m.ctypes.data, wo.ctypes.data, 1, C.byref(seed), out.ctypes.data)
        wi, weight, pdf = out[:3], out[3:6], out[6]
        if pdf <= 0 or wi[2] <= 1e-3:
            continue
        e = _eval(lib, oracle, m, wo, wi)
        assert abs(e[3] - pdf) <= 2e-3 * max(pdf, 1.0), (name, trial, e[3], pdf)
        want = e[:3] * abs(wi[2]) / e[3]
        assert np.allclose(weight, want, rtol=2e-3, atol=1e-5), (name, trial, weight, want)
        checked += 1
    assert checked > 150


@pytest.mark.parametrize("name", sorted(MATERIALS))
@pytest.mark.parametrize("cos_o", [0.95, 0.5, 0.15])
def test_pdf_integrates_to_one_and_white_furnace_holds(oracle, name, cos_o):
    lib = _lib()
    m = _material(**MATERIALS[name])
    wo = np.array([np.sqrt(1 - cos_o * cos_o), 0.0, cos_o], np.float32)
    dirs, dw = _sphere_quadrature()
    vals = np.array([_eval(lib, oracle, m, wo, d) for d in dirs])
    pdf_integral = float(vals[:, 3].sum() * dw)
    albedo = (vals[:, :3] * np.abs(dirs[:, 2:3])).sum(0) * dw
    # sampled directions that leave through the surface are rejected, so an opaque closure's pdf integrates to <= 1; GGX lobes at
    # grazing angles lose the most. The quadrature itself is good to ~1 % for roughness >= 0.25.
    assert 0.55 <= pdf_integral <= 1.02, (name, cos_o, pdf_integral)
    assert float(albedo.max()) <= 1.03, (name, cos_o, albedo)         # energy conservation with white parameters
    assert float(albedo.min()) >= 0.0


def test_lambert_is_reciprocal_and_normalised(oracle):
    lib = _lib()
    m = _material(**MATERIALS["lambert"])
    a, b = np.array([0.3, 0.1, 0.948], np.float32), np.array([-0.6, 0.2, 0.774], np.float32)
    a /= np.linalg.norm(a)
    b /= np.linalg.norm(b)
    fab, fba = _eval(lib, oracle, m, a, b), _eval(lib, oracle, m, b, a)
    assert np.allclose(fab[:3], 1 / np.pi, rtol=1e-5) and np.allclose(fab[:3], fba[:3], rtol=1e-6)
    assert np.isclose(fab[3], b[2] / np.pi, rtol=1e-5)                # cosine-weighted pdf


def test_rgb2spec_round_trip_error_bound(oracle):
    """RGB -> rgb2spec spectrum -> integrate against the CIE fit under equal-energy white -> Bradford E->D65 + XYZ->sRGB
    (utility/spectral.slang:14-136). The table (vkrt_b200/host/tools/rgb2spec_opt.c, the published Jakob-Hanika optimiser: upstream's
    srgb.coeff blob is missing) is fitted for exactly this pipeline: colours come back within 1e-3. The one exception is the achromatic
    corner above 0.995, where the ideal coefficients are infinite and the optimiser's |c| <= 200 clamp leaves a 6 % tint (the runtime of
    the reference has no grey special case either, rgb2spec.slang:33-81): bounded at 7e-2 here so that a change is noticed."""
    lib = H.oracle_lib()
    lam = np.linspace(360.0, 830.0, 941)
    cmf = np.zeros((len(lam), 3), np.float32)
    tmp = np.zeros(3, np.float32)
    for i, l in enumerate(lam):
        lib.oracle_spectral_xyz(C.c_float(l), tmp.ctypes.data_as(C.c_void_p))
        cmf[i] = tmp
    y_norm = cmf[:, 1].sum()
    rng = np.random.default_rng(2)
    colours = np.vstack([rng.uniform(0.05, 0.95, (24, 3)), [[0.99, 0.99, 0.99], [0.5, 0.5, 0.5], [0.9, 0.1, 0.1], [0.1, 0.8, 0.2], [0.1, 0.2, 0.9],
                                                             [1.0, 0.5, 0.2], [1.0, 1.0, 0.5], [1, 1, 1]]]).astype(np.float32)
    worst = white = 0.0
    for rgb in colours:
        spec = np.array([lib.oracle_rgb2spec_eval(oracle.ctx, rgb.ctypes.data_as(C.c_void_p), C.c_float(l)) for l in lam], np.float32)
        xyz = (spec[:, None] * cmf).sum(0) / y_norm
        back = np.zeros(3, np.float32)
        lib.oracle_xyz_to_srgb(xyz.astype(np.float32).ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p))
        err = float(np.abs(back - rgb).max())
        if rgb.min() >= 1.0:
            white = err
        else:
            worst = max(worst, err)
    assert worst <= 1e-3, worst
    assert white <= 7e-2, white
