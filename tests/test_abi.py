"""The C-ABI library must load on a CPU-only box, export every symbol include/vkrt_cuda.h declares, and refuse to run
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import conftest

ROOT = conftest.ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vkrt_cuda.h")).read()
    return sorted(set(re.findall(r"VKRT_CUDA_API\s+[\w\s\*]+?\b(vkrt_cuda_\w+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for must in ("vkrt_cuda_create", "vkrt_cuda_destroy", "vkrt_cuda_set_geometry", "vkrt_cuda_set_instances", "vkrt_cuda_set_materials",
                 "vkrt_cuda_set_lights", "vkrt_cuda_set_textures", "vkrt_cuda_set_rgb2spec", "vkrt_cuda_build_accel", "vkrt_cuda_resize",
                 "vkrt_cuda_reset_accumulation", "vkrt_cuda_render_frame", "vkrt_cuda_gather", "vkrt_cuda_read_aov"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import vkrt_b200
    lib = vkrt_b200.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert set(vkrt_b200.EXPORTS) == set(_declared_symbols())


def test_header_compiles_as_c99(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "vkrt_cuda.h"\nint main(void){ return sizeof(SceneData) == 240 && sizeof(Material) == 272 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


@pytest.mark.skipif(conftest.cuda_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    import vkrt_b200
    with pytest.raises(vkrt_b200.VkrtError) as e:
        vkrt_b200.CudaContext()
    assert e.value.code == -5


def test_missing_extension_raises(tmp_path):
    import vkrt_b200
    with pytest.raises(ImportError):
        vkrt_b200.load_library(str(tmp_path / "nope.so"))


def test_ctypes_mirrors_have_the_c_struct_sizes(tmp_path):
    """Every ctypes.Structure the Python callers (bindings, tests, bench) pass by pointer must have the size and the field offsets of the C
    struct it mirrors: a library call that fills a shorter Python buffer writes past its end (found with an AddressSanitizer build of
    libvkrt_host.so: VKRT_MeshSnapshot is 16-byte aligned through Material, 624 bytes, its first mirror was 620)."""
    import subprocess

    import vkrt_b200
    from vkrt_b200 import host
    import test_images
    pairs = [  # (C type, header, ctypes class, fields whose offset is compared)
        ("vkrt_cuda_create_info", "vkrt_cuda.h", vkrt_b200.CreateInfo, ["flags"]),
        ("vkrt_cuda_build_stats", "vkrt_cuda.h", vkrt_b200.BuildStats, ["triangleCount", "accelBytes", "plocHierarchies"]),
        ("vkrt_cuda_frame_stats", "vkrt_cuda.h", vkrt_b200.FrameStats, ["paths", "instancesEntered", "shadeKernelMs"]),
        ("RGB2SpecTableInfo", "vkrt_cuda.h", vkrt_b200.RGB2SpecInfo, ["dataOffset"]),
        ("vkrt_cuda_texture", "vkrt_cuda.h", vkrt_b200.TextureDesc, ["colorSpace"]),
        ("VKRT_CreateInfo", "vkrt_host.h", host.CreateInfo, ["preferredDeviceName", "cudaFlags", "hostOnly"]),
        ("VKRT_PreparedScene", "vkrt_host.h", host.PreparedScene, ["materials", "triAliasIdx", "sceneData"]),
        ("VKRT_OfflineRenderResult", "vkrt_host.h", host.OfflineRenderResult, ["samples", "mpathsPerSecond", "shadowRays"]),
        ("Camera", "vkrt_host.h", host.Camera, ["vfov"]),
        ("VKRT_SceneSettingsSnapshot", "vkrt_host.h", host.SceneSettings, ["exposure", "autoSPPTargetFPS", "environmentTextureIndex", "selectedMeshIndex"]),
        ("VKRT_RenderStatusSnapshot", "vkrt_host.h", host.RenderStatus, ["totalSamples", "renderTargetSamples", "displayFrameTimeMs"]),
        ("VKRT_LoadedImage", "vkrt_host.h", test_images.LoadedImage, ["colorSpace"]),
        ("VKRT_TextureUpload", "vkrt_host.h", test_images.TextureUpload, ["colorSpace"]),
        ("VKRT_TextureSnapshot", "vkrt_host.h", test_images.TextureSnapshot, ["name"]),
    ]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vkrt_cuda.h"', '#include "vkrt_host.h"', '#include "vkrt_closure.h"', 'int main(void) {']
    for ctype, _, _, fields in pairs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (ctype, ctype))
        for f in fields:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (ctype, f, ctype, f))
    for ctype in ("VKRT_MeshSnapshot", "VKRT_MaterialSnapshot", "vkrt_closure_query", "vkrt_closure_result", "Material", "MeshInfo", "SceneData"):
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (ctype, ctype))
    lines.append('return 0; }')
    src, exe = tmp_path / "sizes.c", tmp_path / "sizes"
    src.write_text("\n".join(lines) + "\n")
    subprocess.check_call(["gcc", "-std=gnu11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    c = dict((k, int(v)) for k, v in (ln.split() for ln in subprocess.check_output([str(exe)]).decode().splitlines()))
    for ctype, _, cls, fields in pairs:
        assert C.sizeof(cls) == c[ctype], (ctype, C.sizeof(cls), c[ctype])
        for f in fields:
            assert getattr(cls, f).offset == c["%s.%s" % (ctype, f)], (ctype, f)
    # mirrors declared inside tests as byte arrays / numpy dtypes
    assert c["VKRT_MeshSnapshot"] == 624 and c["VKRT_MaterialSnapshot"] == 544
    import harness as H
    assert H.hr.MATERIAL.itemsize == c["Material"] == 272 and H.hr.MESH_INFO.itemsize == c["MeshInfo"] == 80 and H.hr.SCENE_DATA.itemsize == c["SceneData"] == 240
    import refpin
    assert refpin.CLOSURE_QUERY.itemsize == c["vkrt_closure_query"] and refpin.CLOSURE_RESULT.itemsize == c["vkrt_closure_result"]


def test_hot_kernels_keep_their_register_and_spill_budget():
    """Build-time guard of the occupancy the measurements rely on: k_trace at <= 72 registers (7 blocks x 128 threads per SM) with
    no more than a few spilled words, k_shade at <= 128 registers (2 blocks x 256 threads) with the spill it had when it was measured.
    Read from the ptxas log the Makefile keeps (vkrt_b200/build/wavefront.ptxas.log, written by build())."""
    log = os.path.join(ROOT, "vkrt_b200", "build", "wavefront.ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log: run __graft_entry__.build()")
    text = open(log).read()
    found = {}
    for m in re.finditer(r"Compiling entry function '(\w+)' for 'sm_100a'.*?(\d+) bytes spill stores.*?Used (\d+) registers", text, flags=re.S):
        found[m.group(1)] = (int(m.group(3)), int(m.group(2)))
    budget = {   # mangled entry: (registers, spill-store bytes)
        "_ZN2vk7k_traceILb0ELb1ELi40EEEvNS_11TraceParamsE": (72, 32),     # flat (single-level) walk, default stack
        "_ZN2vk7k_traceILb0ELb0ELi40EEEvNS_11TraceParamsE": (72, 64),     # two-level walk
        "_ZN2vk7k_shadeILi0ELb0EEEvNS_11FrameParamsEj": (128, 128),       # RGB
        "_ZN2vk7k_shadeILi1ELb0EEEvNS_11FrameParamsEj": (128, 160),       # spectral single
        "_ZN2vk7k_shadeILi2ELb0EEEvNS_11FrameParamsEj": (128, 288),       # spectral hero
    }
    for name, (regs, spill) in budget.items():
        assert name in found, name
        assert found[name][0] <= regs and found[name][1] <= spill, (name, found[name])
