"""bench.py's JSON contract, checked on the CPU: the reference arm (the reference's shaders from oracle/_ref on the host cores) prints one line with the
keys the driver reads, and our own arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--ref-step-seconds", "0.3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-1500:]                     # stdout carries exactly the JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spectral 1080p Mpaths/s" and d["unit"] == "Mpaths/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # "reference" = the reference's own shaders from oracle/_ref (built wherever the reference checkout or its transliterated unit exists), else the port
    want_kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "shaders_gen.inc")) else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["sample"] and cb["value"] == pytest.approx(d["value"])
    e = d["e2e"]
    assert e["value"] == pytest.approx(d["value"]) and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_our_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())    # no number without a GPU


def test_derived_roofline_fields_are_plain_arithmetic_on_the_measured_line():
    """bench.derived_roofline_fields (bytes at the algorithmic and the DRAM level, build fraction) on the numbers of a committed
    bench line: DRAM GB/s = ncu bytes per launch / live launch time, build GB/s = 520 B x triangles / build time, both over the measured peak."""
    sys.path.insert(0, ROOT)
    import bench
    d = json.loads(open(os.path.join(ROOT, "profiles", "r03f_bench_n1.json")).read().strip().splitlines()[-1])
    roofline, c3, peak = d["roofline"], d["config"]["c3"], d["roofline"]["peak"]
    bench.derived_roofline_fields(roofline, c3, peak)
    dram = roofline["dram"]
    assert dram["GBps"] == pytest.approx(roofline["traffic"] / 1e9 / (roofline["avg_launch_ms"] * 1e-3))
    assert dram["frac"] == pytest.approx(dram["GBps"] / peak) and 0.0 < dram["frac"] < roofline["frac"] < 1.0
    assert dram["algorithmic_over_dram"] == pytest.approx(roofline["algorithmic_bytes_per_launch"] / roofline["traffic"])
    assert c3["build_achieved_GBps_algorithmic"] == pytest.approx(c3["triangles"] * 520 / 1e9 / (c3["bvh_build_ms"] * 1e-3))
    assert c3["build_roofline_frac"] == pytest.approx(c3["build_achieved_GBps_algorithmic"] / peak) and 0.0 < c3["build_roofline_frac"] < 1.0
    json.dumps(d)   # still serialisable
    # missing inputs leave the line untouched
    bench.derived_roofline_fields(None, None, peak)
    r2 = {"traffic": None, "avg_launch_ms": 1.0, "algorithmic_bytes_per_launch": 1.0}
    bench.derived_roofline_fields(r2, {"triangles": 10, "bvh_build_ms": 0.0}, peak)
    assert "dram" not in r2
