"""Robustness of the texture decoders against corrupt files (vkrt_b200/host/image_decode.c parses untrusted PNG / JPEG / EXR input;
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


@pytest.mark.parametrize("codec", ["png", "jpeg", "exr"])
def test_mutated_files_decode_or_fail_cleanly(fuzzer, codec):
    exe, files = fuzzer
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:allocator_may_return_null=1:max_allocation_size_mb=1024",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([exe, "7", "400"] + sorted(files[codec]), capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "runtime error" not in r.stderr and "ERROR: AddressSanitizer" not in r.stderr, r.stderr[-2000:]
    assert r.stdout.startswith("decoded ")
