"""Scene fixtures shared by the parity tests (built with the oracle-side host reference, tests only)."""
import os

import numpy as np

import harness as H

hr = H.hr


def cornell(w, h, spp=4, glass=False, **settings):
    scene = hr.cornell_scene(with_glass=glass)
    for k, v in settings.items():
        setattr(scene.settings, k, v)
    prep = scene.prepare(w, h)
    prep["sceneData"]["samplesPerPixel"] = spp
    return prep


def soup(triangles, w, h, spp=2):
    scene = hr.soup_scene(triangles)
    prep = scene.prepare(w, h)
    prep["sceneData"]["samplesPerPixel"] = spp
    return prep


def instanced(w, h, count=64, spp=2, seed=0x5EED0002):
    """Jittered grid of instances of one sphere geometry with random Euler rotations and non-uniform scales
    (exercises the lossy PRS decomposition, the two-level BVH and negative scales)."""
    rng = np.random.default_rng(seed)
    base = hr.uv_sphere_mesh("blob", 1.0, 24, 12)
    mats = np.zeros(6, hr.MATERIAL)
    mats[:] = hr.default_material()
    for k, (col, rough, metal) in enumerate([((0.8, 0.3, 0.2), 0.6, 0.0), ((0.2, 0.6, 0.8), 0.2, 1.0), ((0.9, 0.9, 0.9), 0.05, 0.0), ((0.3, 0.8, 0.3), 1.0, 0.0)]):
        m = hr.default_material()
        m["baseColor"], m["roughness"], m["metallic"] = col, rough, metal
        mats[1 + k] = hr.sanitize_material(m)
    lm = hr.default_material()
    lm["emissionLuminance"] = 20.0
    mats[5] = hr.sanitize_material(lm)
    meshes = []
    side = int(round(count ** (1 / 3.0))) or 1
    n = 0
    for ix in range(side):
        for iy in range(side):
            for iz in range(side):
                if n >= count:
                    break
                m = hr.HostMesh(name="i%d" % n, vertices=base.vertices, indices=base.indices)
                pos = (np.array([ix, iy, iz]) + 0.5) / side * 2 - 1 + rng.uniform(-0.1, 0.1, 3)
                scale = rng.uniform(0.3, 0.6, 3) / side
                if n % 7 == 3:
                    scale[0] = -scale[0]
                m.world = hr.build_mesh_transform(pos, rng.uniform(-180, 180, 3), scale)
                m.material_index = 1 + n % 4
                meshes.append(m)
                n += 1
    light = hr.quad_mesh("light", 0.6)
    light.world = hr.build_mesh_transform((0, 0, 1.6), (180, 0, 0), (1, 1, 1))
    light.material_index = 5
    meshes.append(light)
    floor = hr.quad_mesh("floor", 3.0)
    floor.world = hr.build_mesh_transform((0, 0, -1.2), (0, 0, 0), (1, 1, 1))
    meshes.append(floor)
    st = hr.Settings(camera_pos=(2.2, -3.0, 1.4), camera_target=(0, 0, -0.1), vfov=40.0)
    scene = hr.Scene(meshes=meshes, materials=mats, settings=st)
    prep = scene.prepare(w, h)
    prep["sceneData"]["samplesPerPixel"] = spp
    return prep


def rgb2spec():
    path = os.path.join(H.ROOT, "assets", "rgb2spec", "srgb.coeff")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(H.ROOT, "vkrt_b200"), "rgb2spec"])
    return H.load_rgb2spec(path)


def both_backends(prep, w, h, spectral=False, **cuda_kw):
    table = rgb2spec() if spectral else None
    o, g = H.OracleBackend(), H.CudaBackend(**cuda_kw)
    for b in (o, g):
        b.upload(prep, rgb2spec=table)
        b.resize(w, h)
    return o, g
