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
