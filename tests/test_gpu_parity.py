"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (DESIGN.md "Parity"):
  * integer / index work is bit-exact: primary-hit {instance, primitive} ids, and also the hit's t/u/v bits, because the
    camera, instance-transform and ray/triangle arithmetic are pinned to identical IEEE roundings on both sides;
  * radiance is floating point through libm/CUDA transcendentals (sin, cos, exp, log, pow, acos, atan2) that differ by an
    ulp and get amplified chaotically along a path, so images are compared per pixel with a relative tolerance and a bounded
    outlier fraction: at least 99.5 % of pixels within 1e-3 relative, and image RMSE <= 2 % of the mean radiance; the tone-mapped
    display images additionally agree to a mean LDR-FLIP of 0.002 (measured ~1e-4; tests/flip.py).
"""
import numpy as np
import pytest

import harness as H
import flip
import scenes

pytestmark = pytest.mark.gpu

# both acceleration-structure variants must give the same bits: BLAS per geometry + TLAS (flag 8), one flat BVH (flag 16)
ACCEL = pytest.mark.parametrize("accel", [8, 16], ids=["two_level", "flat"])


def _check_ids(o, g, sd):
    o.trace_primary(sd)
    g.trace_primary(sd)
    for which in (H.AOV_HITID_CENTER, H.AOV_HITID_S0):
        a, b = o.read(which), g.read(which)
        assert np.array_equal(a, b), "hit ids differ in %d pixels" % int((a != b).any(axis=-1).sum())
    ta, tb = o.read(H.AOV_HIT_TUV), g.read(H.AOV_HIT_TUV)
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))


def _check_images(o, g, frac=0.995, rel_rmse=0.02, flip_max=0.002):
    a, b = o.read(H.AOV_ACCUM), g.read(H.AOV_ACCUM)
    assert np.array_equal(a[..., 3], b[..., 3])  # sample counts are integers
    c = H.compare_images(a[..., :3], b[..., :3])
    assert 1.0 - c["frac_rel_gt_1e3"] >= frac, c
    assert c["rmse"] <= rel_rmse * max(c["mean_a"], 1e-6), c
    # feature AOVs are resolved at the first non-specular hit: essentially exact
    for which in (H.AOV_ALBEDO, H.AOV_NORMAL):
        fa, fb = o.read(which).astype(np.float32), g.read(which).astype(np.float32)
        assert (np.abs(fa - fb) > 2e-3).mean() < 0.005
    oa, ob = o.read(H.AOV_OUTPUT).astype(np.int64), g.read(H.AOV_OUTPUT).astype(np.int64)
    assert (np.abs(oa - ob) > 64).mean() < 0.01  # 16-bit display values
    # perceptual difference of the display images (north_star: "within a stated RMSE/FLIP tolerance"): mean LDR-FLIP (tests/flip.py)
    f = flip.mean_flip(oa[..., :3] / 65535.0, ob[..., :3] / 65535.0)
    print("mean FLIP %.5f (bound %.3f)" % (f, flip_max))
    assert f <= flip_max, f


@ACCEL
def test_cornell_rgb_ids_and_images(accel):
    w = h = 128
    prep = scenes.cornell(w, h, spp=4)
    o, g = scenes.both_backends(prep, w, h, flags=accel)
    assert g.build_stats.flat == (1 if accel == 16 else 0)
    _check_ids(o, g, prep["sceneData"])
    o.render(prep["sceneData"], frames=3)
    g.render(prep["sceneData"], frames=3)
    _check_images(o, g)


@ACCEL
def test_cornell_glass_medium_and_unsupported_transmission(accel):
    w = h = 96
    prep = scenes.cornell(w, h, spp=4, glass=True)
    prep["materials"][7]["absorptionCoefficient"] = 2.0
    prep["materials"][7]["attenuationColor"] = (0.4, 0.8, 0.9)
    o, g = scenes.both_backends(prep, w, h, flags=accel)
    _check_ids(o, g, prep["sceneData"])
    o.render(prep["sceneData"], frames=2)
    g.render(prep["sceneData"], frames=2)
    _check_images(o, g, frac=0.99, rel_rmse=0.05)


@ACCEL
def test_instanced_two_level_bvh(accel):
    w, h = 160, 96
    prep = scenes.instanced(w, h, count=64, spp=2)
    o, g = scenes.both_backends(prep, w, h, flags=accel)
    assert g.build_stats.uniqueGeometries == 3 and g.build_stats.instanceCount == 66
    _check_ids(o, g, prep["sceneData"])
    o.render(prep["sceneData"], frames=2)
    g.render(prep["sceneData"], frames=2)
    _check_images(o, g, frac=0.99, rel_rmse=0.05)


@ACCEL
def test_soup_ids_and_random_rays(accel):
    w, h = 128, 96
    prep = scenes.soup(60000, w, h, spp=1)
    o, g = scenes.both_backends(prep, w, h, flags=accel)
    _check_ids(o, g, prep["sceneData"])
    rng = np.random.default_rng(9)
    n = 50000
    org = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays = np.concatenate([org, np.full((n, 1), 1e-3, np.float32), d, np.full((n, 1), 1e4, np.float32)], axis=1)
    assert np.array_equal(o.trace_rays(rays), g.trace_rays(rays))
    rays[:, 7] = 0.5
    assert np.array_equal(o.trace_rays(rays, any_hit=True)[:, 0], g.trace_rays(rays, any_hit=True)[:, 0])


@pytest.mark.parametrize("mode", ["rgb", "hero"])
def test_textured_materials_alpha_cutout_and_environment_map(mode):
    """All four texture formats, the three wrap modes, KHR_texture_transform, normal / metallic-roughness / emissive maps, a stochastic
    alpha-MASK cut-out inside the traversal and a lat-long environment map (material/textures.slang, light/environment.slang)."""
    w, h = 144, 96
    prep = scenes.textured(w, h, spp=4)
    spectral = mode != "rgb"
    if spectral:
        prep["sceneData"]["packedRenderSettings"] = H.hr.pack_render_settings(0, 1, 1)   # tone mapping none, spectral, hero
    o, g = scenes.both_backends(prep, w, h, spectral=spectral)
    _check_ids(o, g, prep["sceneData"])
    o.render(prep["sceneData"], frames=2)
    g.render(prep["sceneData"], frames=2)
    _check_images(o, g, frac=0.985, rel_rmse=0.08)


def test_thousand_instances_two_level_bvh():
    """BASELINE config C4's shape: 1000 rotated, non-uniformly scaled (some mirrored) instances of one geometry under a TLAS; ids and
    a spectral-hero frame against the oracle."""
    w, h = 320, 192
    prep = scenes.instanced(w, h, count=1000, spp=2)
    prep["sceneData"]["packedRenderSettings"] = H.hr.pack_render_settings(0, 1, 1)
    o, g = scenes.both_backends(prep, w, h, spectral=True, flags=8)
    assert g.build_stats.instanceCount == 1002 and g.build_stats.flat == 0
    _check_ids(o, g, prep["sceneData"])
    o.render(prep["sceneData"], frames=1)
    g.render(prep["sceneData"], frames=1)
    # 2 spp on glossy metal: one sample whose lobe choice flips on a last-bit difference is a 10-unit firefly in a 0.36-mean image,
    # so the RMSE bound is looser here; the per-pixel agreement fraction is the sharp criterion
    _check_images(o, g, frac=0.985, rel_rmse=0.2, flip_max=0.01)   # measured 0.0068 (the firefly), 1e-4 everywhere else


def test_large_soup_builder_and_traversal():
    """A 2 M-triangle soup (BASELINE config C3 at a fifth of its size: the CPU side of a 10 M build would dominate the suite): the
    radix sort, the hierarchy and the collapse run with millions of primitives and a BVH far larger than L2; every primary-hit id of a
    960x540 frame and 200 k random closest-hit / any-hit rays must equal the oracle's BVH2."""
    w, h = 960, 540
    prep = scenes.soup(2000000, w, h, spp=1)
    o, g = scenes.both_backends(prep, w, h)
    assert g.build_stats.triangleCount == 2000002 and g.build_stats.flat == 1
    _check_ids(o, g, prep["sceneData"])
    rng = np.random.default_rng(10)
    n = 200000
    org = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[: n // 100, 0] = 0.0                      # rays parallel to a slab axis (the reciprocal clamp of the node test)
    d[n // 100: n // 50, 1] = 1e-12
    rays = np.concatenate([org, np.full((n, 1), 1e-3, np.float32), d, np.full((n, 1), 1e4, np.float32)], axis=1)
    assert np.array_equal(o.trace_rays(rays), g.trace_rays(rays))
    rays[:, 7] = 0.25
    assert np.array_equal(o.trace_rays(rays, any_hit=True)[:, 0], g.trace_rays(rays, any_hit=True)[:, 0])


@pytest.mark.parametrize("debug_mode", [1, 2, 3, 4, 5, 12, 13, 14, 15, 16])
def test_debug_views(debug_mode):
    w = h = 64
    prep = scenes.cornell(w, h, spp=2, debug_mode=debug_mode)
    o, g = scenes.both_backends(prep, w, h)
    o.render(prep["sceneData"], frames=1)
    g.render(prep["sceneData"], frames=1)
    a, b = o.read(H.AOV_ACCUM), g.read(H.AOV_ACCUM)
    c = H.compare_images(a[..., :3], b[..., :3])
    assert c["frac_rel_gt_1e2"] < 0.01, c


@pytest.mark.parametrize("debug_mode", [6, 7, 8, 9, 10, 11])
def test_texture_map_debug_views(debug_mode):
    """Debug views 6 - 11 on the textured stage: the selection mask (no selection: the 0.05 grey of debug.slang) and the base-colour /
    metallic / roughness / normal-map / emissive-map views (integrator/path/debug.slang:31-64 through material/textures.slang), which the
    cornell scene of test_debug_views cannot exercise. Debug pixels carry weight 0 (writeDebugFrameOutputs), the others the sample count."""
    w, h = 96, 64
    prep = scenes.textured(w, h, spp=2)
    prep["sceneData"]["debugMode"] = debug_mode
    o, g = scenes.both_backends(prep, w, h)
    o.render(prep["sceneData"], frames=1)
    g.render(prep["sceneData"], frames=1)
    a, b = o.read(H.AOV_ACCUM), g.read(H.AOV_ACCUM)
    assert np.array_equal(a[..., 3], b[..., 3])
    assert (a[..., 3] == 0).mean() > 0.5          # most of the frame is a debug pixel
    c = H.compare_images(a[..., :3], b[..., :3])
    assert c["frac_rel_gt_1e2"] < 0.01, c


def test_nee_only_plus_bsdf_only_equals_full():
    """The reference's own validation views (constants.h:39-40): with MIS the two techniques partition the estimator,
    so E[NEE-only] + E[BSDF-only] = E[full]. Checked on the GPU alone at moderate spp."""
    w = h = 48
    means = {}
    for mode in (0, 4, 5):
        prep = scenes.cornell(w, h, spp=64, debug_mode=mode)
        g = H.CudaBackend()
        g.upload(prep)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=4)
        means[mode] = g.read(H.AOV_ACCUM)[..., :3].astype(np.float64).mean()
        g.close()
    assert abs(means[4] + means[5] - means[0]) < 0.03 * means[0], means


def test_accumulation_is_a_running_mean_and_deterministic():
    w = h = 64
    prep = scenes.cornell(w, h, spp=4)
    g1, g2 = H.CudaBackend(), H.CudaBackend(max_paths=w * h * 2)  # second context forces 2 sample chunks per frame
    for g in (g1, g2):
        g.upload(prep)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=3)
    a, b = g1.read(H.AOV_ACCUM), g2.read(H.AOV_ACCUM)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "chunked and unchunked frames must be bit-identical"
    assert np.all(a[..., 3] == 12.0)
    g1.reset()
    g1.render(prep["sceneData"], frames=3)
    assert np.array_equal(a.view(np.uint32), g1.read(H.AOV_ACCUM).view(np.uint32)), "re-render must be bit-identical"


@pytest.mark.parametrize("scene", ["cornell_glass", "lobes", "textured"])
@pytest.mark.parametrize("spectral", [1, 2])
def test_spectral_memo_changes_no_bit(scene, spectral, monkeypatch):
    """k_spectral_memo looks the rgb2spec table up once per material / light constant (diffuse colour, emission) instead of once per path
    vertex; an entry is used only when its key equals the colour at hand bit for bit. Frames with the memo and with every lookup going to
    the table (VKRT_NO_SPECTRAL_MEMO=1, read at context creation) must be bit-identical, textured materials (memo misses) included."""
    w = h = 96
    prep = {"cornell_glass": lambda: scenes.cornell(w, h, spp=4, glass=True), "lobes": lambda: scenes.lobes(w, h, spp=4),
            "textured": lambda: scenes.textured(w, h, spp=4)}[scene]()
    prep["sceneData"]["packedRenderSettings"] = H.hr.pack_render_settings(0, 1, 1 if spectral == 2 else 0)
    table = scenes.rgb2spec()
    images = []
    for off in ("0", "1"):
        monkeypatch.setenv("VKRT_NO_SPECTRAL_MEMO", off)
        g = H.CudaBackend()
        g.upload(prep, rgb2spec=table)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=2)
        images.append(g.read(H.AOV_ACCUM))
        g.close()
    assert float(images[0][..., :3].mean()) > 0.0
    assert np.array_equal(images[0].view(np.uint32), images[1].view(np.uint32))


def test_full_size_frame_properties():
    """BASELINE config C2 at full size (1920x1080, spectral hero, 16 spp per frame = 33.2 M paths), where the oracle would take minutes:
    size-independent properties instead. (1) a frame rendered in three sample chunks equals the unchunked frame bit for bit; (2) the
    union of two tile-partitioned ranks equals it bit for bit; (3) sample counts are exact; (4) the 2 M primary-hit ids of the 1080p frame equal
    the oracle's (primary visibility alone is cheap enough on the CPU)."""
    w, h, spp = 1920, 1080, 16
    prep = scenes.cornell(w, h, spp=spp)
    prep["sceneData"]["packedRenderSettings"] = H.hr.pack_render_settings(0, 1, 1)
    table = scenes.rgb2spec()
    full = H.CudaBackend(max_paths=w * h * spp)
    full.upload(prep, rgb2spec=table)
    full.resize(w, h)
    full.render(prep["sceneData"], frames=2)
    ref = full.read(H.AOV_ACCUM)
    assert np.all(ref[..., 3] == 2.0 * spp)
    assert np.isfinite(ref).all() and float(ref[..., :3].mean()) > 0.0
    chunked = H.CudaBackend(max_paths=w * h * 6)     # 16 spp in chunks of 6 + 6 + 4
    chunked.upload(prep, rgb2spec=table)
    chunked.resize(w, h)
    chunked.render(prep["sceneData"], frames=2)
    assert np.array_equal(ref.view(np.uint32), chunked.read(H.AOV_ACCUM).view(np.uint32))
    chunked.close()
    acc = np.zeros_like(ref)
    for r in range(2):
        g = H.CudaBackend(rank=r, world_size=2, max_paths=w * h * spp // 2 + 65536)
        g.upload(prep, rgb2spec=table)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=2)
        acc += g.read(H.AOV_ACCUM)
        g.close()
    assert np.array_equal(ref.view(np.uint32), acc.view(np.uint32))
    gpu_ids = full.read(H.AOV_HITID_CENTER)
    full.close()
    o = H.OracleBackend()
    o.upload(prep)
    o.resize(w, h)
    o.trace_primary(prep["sceneData"])
    assert np.array_equal(gpu_ids, o.read(H.AOV_HITID_CENTER))


def test_tile_partition_matches_single_gpu():
    """Two ranks on one device: the union of their tile-compact films equals the single-rank image bit for bit."""
    w, h = 200, 120  # not a multiple of the tile size
    prep = scenes.cornell(w, h, spp=2)
    full = H.CudaBackend()
    full.upload(prep)
    full.resize(w, h)
    full.render(prep["sceneData"], frames=2)
    ref = full.read(H.AOV_ACCUM)
    acc = np.zeros_like(ref)
    for r in range(2):
        g = H.CudaBackend(rank=r, world_size=2)
        g.upload(prep)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=2)
        part = g.read(H.AOV_ACCUM)  # pixels of other ranks read back as zero
        assert np.all((part[..., 3] == 0) | (part[..., 3] == 4))
        acc += part
        g.close()
    assert np.array_equal(acc.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("sampling", [0, 1])
def test_cornell_spectral(sampling):
    """renderMode = spectral, single-wavelength (0) and hero-wavelength (1) sampling; accumulation holds XYZ."""
    w = h = 96
    prep = scenes.cornell(w, h, spp=4, render_mode=1, spectral_sampling=sampling)
    o, g = scenes.both_backends(prep, w, h, spectral=True)
    _check_ids(o, g, prep["sceneData"])
    o.render(prep["sceneData"], frames=2)
    g.render(prep["sceneData"], frames=2)
    _check_images(o, g, frac=0.99, rel_rmse=0.05)


def test_default_accel_choice():
    """vkrt_cuda_build_accel: with the default builder (radix tree and PLOC, lower surface-area cost kept) every scene whose instanced
    triangles fit without multiplying memory gets ONE flat BVH (cornell included: its wall-sized triangles make the cost comparison
    keep the PLOC hierarchy); heavily instanced scenes keep BLAS per geometry + TLAS. With the radix tree alone (flag 128) separated
    instances (cornell) keep their own BLASes and only stacked meshes (the 16-mesh soup, overlap depth 16) are flattened."""
    g = H.CudaBackend()
    g.upload(scenes.cornell(32, 32))
    assert g.build_stats.flat == 1 and g.build_stats.plocHierarchies == 1
    g.close()
    g = H.CudaBackend()
    g.upload(scenes.instanced(32, 32, count=1000))      # 1000 x 576 triangles = 0.6 M instanced triangles: below the 4 Mi flattening limit
    assert g.build_stats.flat == 1                       # (1000 x suzanne.glb = 63 M stays two-level: tests/test_gpu_reference.py, config C4)
    g.close()
    g = H.CudaBackend(flags=128)
    g.upload(scenes.cornell(32, 32))
    assert g.build_stats.flat == 0 and g.build_stats.plocHierarchies == 0
    g.close()
    g = H.CudaBackend(flags=128)
    g.upload(scenes.soup(20000, 32, 32))
    assert g.build_stats.flat == 1
    g.close()


@pytest.mark.parametrize("builder", [0, 128, 256], ids=["best", "lbvh", "ploc"])
@ACCEL
def test_every_builder_gives_the_same_hits(builder, accel):
    """The accepted hit does not depend on the hierarchy (closest hit with id tie-break): radix tree, PLOC and best-of-two must return the
    oracle's ids and t/u/v bits on a scene with wall-sized and tiny triangles, on instances and on a soup."""
    for prep, w, h in ((scenes.cornell(96, 96, spp=1), 96, 96), (scenes.instanced(96, 64, count=64, spp=1), 96, 64), (scenes.soup(30000, 96, 64, spp=1), 96, 64)):
        o, g = scenes.both_backends(prep, w, h, flags=accel | builder)
        _check_ids(o, g, prep["sceneData"])
        g.close()


@pytest.mark.parametrize("accel", [8, 16])
def test_deep_stack_kernel_is_bit_identical(accel):
    """The traversal stack size follows the depth of the built trees (api.cu: stackNeed); the deep-tree instantiation of
    k_trace (VKRT_CUDA_FLAG_DEEP_STACK = 32 forces it) must give the same hits and the same film, bit for bit."""
    w, h = 128, 96
    prep = scenes.instanced(w, h, count=64, spp=2)
    films, ids = [], []
    for extra in (0, 32):
        g = H.CudaBackend(flags=accel | extra)
        g.upload(prep)
        g.resize(w, h)
        g.trace_primary(prep["sceneData"])
        ids.append(g.read(H.AOV_HITID_CENTER).copy())
        g.render(prep["sceneData"], frames=2)
        films.append(g.read(H.AOV_ACCUM).copy())
        g.close()
    assert np.array_equal(ids[0], ids[1])
    assert np.array_equal(films[0].view(np.uint32), films[1].view(np.uint32))


@pytest.mark.parametrize("mode", ["rgb", "hero"])
def test_converged_images_agree_across_sample_sets(mode):
    """The other tests compare the two implementations on IDENTICAL samples. Here the sample sets are disjoint (the seed is a function of
    pixel, frame and sample index: the oracle renders frames 1000..1015, the GPU frames 0..63), so agreement is statistical: the
    converged images must coincide. Bounds: whole-image mean within 1 %, 8x8 block means within 4 % RMS of the mean."""
    w, h = 96, 64
    prep = scenes.cornell(w, h, spp=16)
    spectral = mode != "rgb"
    if spectral:
        prep["sceneData"]["packedRenderSettings"] = H.hr.pack_render_settings(0, 1, 1)
    o, g = scenes.both_backends(prep, w, h, spectral=spectral)
    o.render(prep["sceneData"], frames=16, first_frame=1000)   # 256 spp on the CPU
    g.render(prep["sceneData"], frames=64, first_frame=0)      # 1024 spp on the GPU
    a, b = o.read(H.AOV_ACCUM)[..., :3].astype(np.float64), g.read(H.AOV_ACCUM)[..., :3].astype(np.float64)
    assert np.isfinite(a).all() and np.isfinite(b).all()
    ma, mb = a.mean(), b.mean()
    assert abs(ma - mb) <= 0.01 * ma, (ma, mb)
    ba, bb = _block_mean(a), _block_mean(b)
    rms = np.sqrt(((ba - bb) ** 2).mean()) / ma
    print("block-mean rRMS %.4f" % rms)
    assert rms < 0.04, rms


def _block_mean(img, b=8):
    h, w, c = img.shape
    return img[:h // b * b, :w // b * b].reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))


@pytest.mark.parametrize("mode", ["rgb", "hero"])
@pytest.mark.parametrize("lamp", [False, True], ids=["env_only", "env_and_lamp"])
def test_environment_importance_sampling_extension(mode, lamp):
    """VKRT_CUDA_FLAG_ENV_IMPORTANCE (64) is an estimator the reference does not have (it reads the environment on a miss only), so it
    is checked as an estimator: the same expectation as the default path (which the other tests pin against the oracle) - block
    means of a long default render agree with a short importance-sampled one - and a much smaller error at equal sample count under
    a small bright sun. The default path must not notice the flag when the scene has no environment texture."""
    w, h = 96, 64
    prep = scenes.sunlit(w, h, spp=64, lamp=lamp)
    spectral = mode != "rgb"
    if spectral:
        prep["sceneData"]["packedRenderSettings"] = H.hr.pack_render_settings(0, 1, 1)
    table = scenes.rgb2spec() if spectral else None

    def render(flags, frames):
        g = H.CudaBackend(flags=flags)
        g.upload(prep, rgb2spec=table)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=frames)
        img = g.read(H.AOV_ACCUM)[..., :3].astype(np.float64)
        g.close()
        return img

    long_default = render(0, 48)    # 3072 spp with the reference's estimator
    short_default = render(0, 1)    # 64 spp
    short_is = render(64, 1)        # 64 spp, environment importance sampling
    long_is = render(64, 4)         # 256 spp
    assert np.isfinite(long_is).all() and np.isfinite(short_is).all()
    # same expectation: whole-image mean within 1 % (measured: 0.1 %, profiles/r01l_env_importance.txt), 8x8 block means within 6 % on average
    ma, mb = long_default.mean(axis=(0, 1)), long_is.mean(axis=(0, 1))
    assert np.all(np.abs(ma - mb) <= 0.01 * ma), (ma, mb)
    ba, bb = _block_mean(long_default), _block_mean(long_is)
    rel = np.abs(ba - bb) / (ba + 0.02 * ma)
    assert rel.mean() < 0.06, rel.mean()
    # smaller error at equal cost (against the 3072-spp image)
    err_default = np.sqrt(((short_default - long_default) ** 2).mean())
    err_is = np.sqrt(((short_is - long_default) ** 2).mean())
    assert err_is < 0.5 * err_default, (err_is, err_default)


def test_environment_importance_sampling_on_a_large_map():
    """A 1024 x 512 environment (524 288 texels) whose sun sits at texel indices >= 98 304: the texel draw multiplies 48 random bits by the
    texel count, which needs the high half of a 128-bit product - in 64 bits it wraps for every map above 65 536 texels and the draw never
    leaves the first 65 536 (the sun would be reached through alias links only, against a density that assumes the full table: a dark,
    biased image). Same expectation as the default estimator, and the variance gain must survive the larger table."""
    w, h = 96, 64
    prep = scenes.sunlit(w, h, spp=64, env_scale=16)

    def render(flags, frames):
        g = H.CudaBackend(flags=flags)
        g.upload(prep)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=frames)
        img = g.read(H.AOV_ACCUM)[..., :3].astype(np.float64)
        g.close()
        return img

    long_default, short_default = render(0, 48), render(0, 1)
    long_is, short_is = render(64, 4), render(64, 1)
    assert np.isfinite(long_is).all()
    ma, mb = long_default.mean(axis=(0, 1)), long_is.mean(axis=(0, 1))
    assert np.all(np.abs(ma - mb) <= 0.015 * ma), (ma, mb)
    err_default = np.sqrt(((short_default - long_default) ** 2).mean())
    err_is = np.sqrt(((short_is - long_default) ** 2).mean())
    assert err_is < 0.5 * err_default, (err_is, err_default)


def test_environment_importance_flag_is_inert_without_environment_texture():
    w, h = 96, 64
    prep = scenes.cornell(w, h, spp=4)
    films = []
    for flags in (0, 64):
        g = H.CudaBackend(flags=flags)
        g.upload(prep)
        g.resize(w, h)
        g.render(prep["sceneData"], frames=2)
        films.append(g.read(H.AOV_ACCUM).copy())
        g.close()
    assert np.array_equal(films[0].view(np.uint32), films[1].view(np.uint32))


def test_out_of_range_index_is_refused_at_build():
    """The ABI copies caller buffers in; an index that does not address a vertex of its geometry is refused by vkrt_cuda_build_accel
    (VKRT_ERROR_INVALID_ARGUMENT) before any kernel dereferences it."""
    prep = scenes.cornell(32, 32)
    prep["indices"] = prep["indices"].copy()
    prep["indices"][7] = 0x7FFFFFF0
    g = H.CudaBackend()
    with pytest.raises(Exception, match="build_accel"):
        g.upload(prep)
    g.close()
    g = H.CudaBackend()
    g.upload(scenes.cornell(32, 32))   # the intact scene builds
    assert g.build_stats.triangleCount > 0
    g.close()


def test_inconsistent_light_tables_are_refused():
    """vkrt_cuda_set_lights checks every link of the alias / emissive tables (the light sampler follows them blindly), and a frame may
    not name more emissive meshes than were uploaded."""
    import copy
    good = scenes.cornell(32, 32)
    for field, value in (("triAliasIdx", 1 << 30), ("meshAliasIdx", 7)):
        prep = copy.deepcopy(good)
        prep["lights"][field] = prep["lights"][field].copy()
        prep["lights"][field][0] = value
        g = H.CudaBackend()
        with pytest.raises(Exception, match="set_lights"):
            g.upload(prep)
        g.close()
    g = H.CudaBackend()
    g.upload(good)
    g.resize(32, 32)
    sd = good["sceneData"].copy()
    sd["emissiveMeshCount"] = 5
    with pytest.raises(Exception):
        g.render(sd, frames=1)
    g.render(good["sceneData"], frames=1)   # the consistent frame still renders
    g.close()


@ACCEL
@pytest.mark.parametrize("bad", [np.nan, np.inf, 3.0e38], ids=["nan", "inf", "huge"])
def test_triangles_with_non_finite_vertices_are_inactive(accel, bad):
    """Vulkan acceleration structures (what the reference builds, accel/blas.c) treat a triangle with a NaN vertex as inactive; here every
    triangle with a non-finite or overflowing coordinate collapses to a point inside the builder: it is never hit, everything else
    renders exactly as if that triangle were degenerate, and the context survives (a NaN vertex used to fault the build kernels)."""
    w = h = 64
    good = scenes.cornell(w, h, spp=2)
    tri = 40                                              # a triangle of the scene, made inactive through one of its vertices
    mesh = next(m for m in good["meshInfos"] if m["indexBase"] <= 3 * tri < m["indexBase"] + m["indexCount"])   # indices are local to a geometry
    first, last = int(mesh["indexBase"]) // 3, int(mesh["indexBase"] + mesh["indexCount"]) // 3
    vi = int(good["indices"][3 * tri]) + int(mesh["vertexBase"])
    films, ids = [], []
    for variant in ("non_finite", "degenerate"):
        prep = dict(good)
        if variant == "non_finite":
            v = good["vertices"].copy()
            v["position"][vi, 1] = bad
            prep["vertices"] = v
        # reference image: the same triangles degenerate (all three corners on one vertex) in a scene without the bad value
        idx = good["indices"].copy()
        users = first + np.nonzero((idx.reshape(-1, 3)[first:last] == idx[3 * tri]).any(axis=1))[0]
        if variant == "degenerate":
            for t in users:
                keep = [c for c in idx[3 * t:3 * t + 3] if c != idx[3 * tri]]
                idx[3 * t:3 * t + 3] = keep[0] if keep else idx[3 * tri]
            prep["indices"] = idx
        g = H.CudaBackend(flags=accel)
        g.upload(prep)
        g.resize(w, h)
        g.trace_primary(prep["sceneData"])
        ids.append(g.read(H.AOV_HITID_CENTER).copy())
        g.render(prep["sceneData"], frames=2)
        films.append(g.read(H.AOV_ACCUM).copy())
        g.close()
    assert np.isfinite(films[0]).all()
    assert np.array_equal(ids[0], ids[1])
    assert np.array_equal(films[0].view(np.uint32), films[1].view(np.uint32))


def test_dispersive_glass_hero_collapse():
    """Rough glass with an Abbe number: hero paths collapse to one wavelength on refraction (spectral_hero/transport.slang:77-87)."""
    w = h = 96
    prep = scenes.cornell(w, h, spp=4, glass=True, render_mode=1, spectral_sampling=1)
    prep["materials"][7]["abbeNumber"] = 25.0
    prep["materials"][7]["absorptionCoefficient"] = 1.5
    prep["materials"][7]["attenuationColor"] = (0.9, 0.6, 0.3)
    o, g = scenes.both_backends(prep, w, h, spectral=True)
    o.render(prep["sceneData"], frames=2)
    g.render(prep["sceneData"], frames=2)
    _check_images(o, g, frac=0.985, rel_rmse=0.08)


def test_spectral_without_table_fails():
    w = h = 32
    prep = scenes.cornell(w, h, spp=1, render_mode=1)
    g = H.CudaBackend()
    g.upload(prep)
    g.resize(w, h)
    with pytest.raises(Exception):
        g.render(prep["sceneData"], frames=1)


def test_material_edits_keep_the_bvh_unless_it_depends_on_them():
    """a colour / roughness edit must not pay a BVH rebuild. vkrt_cuda_build_accel returns the previous build unless geometry,
    instance matrices / sharing / any-hit flags or a material's "transmits" bit changed; the frame after the edit still matches the
    oracle with the same edit."""
    w = h = 64
    prep = scenes.cornell(w, h, spp=2)
    o, g = scenes.both_backends(prep, w, h)
    first = g.build_stats.buildMs
    mats = prep["materials"].copy()
    mats[3]["baseColor"] = (0.1, 0.9, 0.2)
    mats[3]["roughness"] = 0.2
    for b in (o, g):
        m = np.ascontiguousarray(mats)
        b.check(b.f("set_materials")(b.ctx, m.ctypes.data_as(H.C.c_void_p), H.C.c_uint32(len(m))), "set_materials")
        b.build_accel()
    assert g.build_stats.buildMs == first, "a colour edit rebuilt the BVH"
    o.render(prep["sceneData"], frames=2)
    g.render(prep["sceneData"], frames=2)
    _check_images(o, g)
    mats[3]["transmission"] = 1.0          # now shadow rays must treat that instance as transmissive: the instance records change
    m = np.ascontiguousarray(mats)
    g.check(g.f("set_materials")(g.ctx, m.ctypes.data_as(H.C.c_void_p), H.C.c_uint32(len(m))), "set_materials")
    g.build_accel()
    assert g.build_stats.buildMs != first, "a transmission edit must rebuild the instance records"
    g.close()
