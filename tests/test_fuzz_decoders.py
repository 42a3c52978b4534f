"""Robustness of the file readers against corrupt input: texture decoders, .glb importer, vkrt.scene reader.
Texture decoders (vkrt_b200/host/image_decode.c parses untrusted PNG / JPEG / EXR input;
the reference delegates this to libspng / libjpeg-turbo / tinyexr). tests/fuzz/fuzz_image_decode.c is built with AddressSanitizer +
UndefinedBehaviorSanitizer and fed mutated copies of the committed fixture files: every input must decode or be rejected cleanly.
A short deterministic pass runs here; longer passes: `fuzz_image_decode <seed> <iterations> files...`."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fuzzer(tmp_path_factory):
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    d = tmp_path_factory.mktemp("fuzz")
    exe = str(d / "fuzz_image_decode")
    cmd = ["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=gnu11", "-o", exe,
           os.path.join(ROOT, "tests", "fuzz", "fuzz_image_decode.c"), os.path.join(ROOT, "vkrt_b200", "host", "image_decode.c"),
           "-I" + os.path.join(ROOT, "include"), "-lz", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-300:])
    files = {}
    gold = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))
    for k in gold.files:
        if k.startswith("file_"):
            p = str(d / k[5:])
            with open(p, "wb") as f:
                f.write(gold[k].tobytes())
            files.setdefault(k[5:].split("_")[0], []).append(p)
    return exe, files


def _run(cmd, env):
    """LeakSanitizer needs ptrace; where the sandbox forbids it the pass is repeated without leak detection (ASan + UBSan stay on)."""
    r = subprocess.run(cmd, capture_output=True, encoding="utf-8", errors="replace", env=env, timeout=600)
    if "LeakSanitizer has encountered a fatal error" in r.stderr:
        env = dict(env, ASAN_OPTIONS=env["ASAN_OPTIONS"].replace("detect_leaks=1", "detect_leaks=0"))
        r = subprocess.run(cmd, capture_output=True, encoding="utf-8", errors="replace", env=env, timeout=600)
    return r


@pytest.mark.parametrize("codec", ["png", "jpeg", "exr"])
def test_mutated_files_decode_or_fail_cleanly(fuzzer, codec):
    exe, files = fuzzer
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:allocator_may_return_null=1:max_allocation_size_mb=1024",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = _run([exe, "7", "400"] + sorted(files[codec]), env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "runtime error" not in r.stderr and "ERROR: AddressSanitizer" not in r.stderr, r.stderr[-2000:]
    assert r.stdout.startswith("decoded ")


HOST_SRCS = ["api.c", "scene_prep.c", "scene_file.c", "gltf_import.c", "hjson.c", "export.c", "image_decode.c", "jpeg_encode.c", "controllers.c", "denoise.c"]
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:allocator_may_return_null=1:max_allocation_size_mb=1024",
           UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")


def _build_host_fuzzer(tmp, name):
    """The importer and the scene reader sit on top of the whole C host, so these two link all of it (sanitized) against libvkrt_cuda.so;
    nothing on these paths touches the device (host-only handle)."""
    lib_dir = os.path.join(ROOT, "vkrt_b200")
    if not shutil.which("gcc") or not os.path.exists(os.path.join(lib_dir, "libvkrt_cuda.so")):
        pytest.skip("gcc or libvkrt_cuda.so not available")
    exe = str(tmp / name)
    cmd = ["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=gnu11", "-o", exe,
           os.path.join(ROOT, "tests", "fuzz", name + ".c")] + [os.path.join(lib_dir, "host", f) for f in HOST_SRCS] + [
           "-I" + os.path.join(ROOT, "include"), "-L" + lib_dir, "-lvkrt_cuda", "-lz", "-lm", "-ldl", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-300:])
    return exe


def _clean(r):
    assert r.returncode == 0, r.stderr[-2000:]
    assert "runtime error" not in r.stderr and "ERROR: AddressSanitizer" not in r.stderr and "LeakSanitizer" not in r.stderr, r.stderr[-2000:]


def test_mutated_glb_files_import_or_fail_cleanly(tmp_path):
    """Every accepted import must also be self-consistent (indices below the vertex count — the reference rejects the primitive otherwise,
    src/app/mesh/loader.c:1859 — and material indices inside the material list): the fuzzer aborts on a violation."""
    exe = _build_host_fuzzer(tmp_path, "fuzz_gltf_import")
    models = [os.path.join(ROOT, "assets", "models", m) for m in ("cube.glb", "plane.glb", "prism.glb")]
    r = _run([exe, "11", "400", str(tmp_path / "scratch.glb")] + models, ENV)
    _clean(r)
    assert r.stdout.startswith("imported ")


def test_mutated_scene_files_load_or_fail_cleanly(tmp_path):
    exe = _build_host_fuzzer(tmp_path, "fuzz_scene_file")
    (tmp_path / "scenes").mkdir()
    os.symlink(os.path.join(ROOT, "assets", "models"), str(tmp_path / "models"))   # the scene imports ../models/*.glb
    r = _run([exe, "13", "800", str(tmp_path / "scenes" / "scratch.json"), os.path.join(ROOT, "assets", "scenes", "prism.json")], ENV)
    _clean(r)
    assert r.stdout.strip().splitlines()[-1].startswith("loaded ")
