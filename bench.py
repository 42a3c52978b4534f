#!/usr/bin/env python3
"""bench.py — the headline measurement: spectral (hero-wavelength) 1080p path tracing throughput in Mpaths/s.

Workload (BASELINE.json configs[1]): the bundled cornell scene (assets/scenes/cornell.json, bunny.glb standing in
for the missing dragon.glb), 1920x1080, renderMode = spectral, spectralSamplingMode = hero, depth 4..8, NEE + MIS on, a procedural
lat-long HDR sky as environment map (the bundled .exr environments are missing blobs upstream). One STEP = one frame
(VKRT_draw) of `spp_per_step` samples per pixel over the whole image = W*H*spp camera paths; 1024 spp = 32 such steps at N = 1
(32 spp per step: the frame's fixed cost — kernel tails, 45 launches — is ~1.2 ms, B200 probe: 8 / 16 / 32 / 64 spp per frame = 480 / 497 / 502 / 508 Mpaths/s).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's algorithm on the host CPU (the oracle restatement;
                                                           the reference itself needs Vulkan + slangc and cannot be built here)

Multi-GPU: the image is partitioned into interleaved 32x32 tiles (vkrt_b200/csrc/tiles.h), the scene is replicated, there is no
data-path collective; the film is gathered to rank 0 over NCCL once at the end of the timed region. Per-GPU work is kept fixed as N
grows (spp_per_step = 32*N over the same 1080p image => "weak" scaling, value = all paths of all ranks / max-over-ranks time).

Timing: `value` = device time between two CUDA events recorded on the library's own stream (vkrt_cuda_timer_begin/end) around K
asynchronously enqueued frames, inputs (scene, BVH, film) resident in HBM, barrier + synchronize on both sides, max over ranks.
`e2e` = the same K frames through the public host API (VKRT_draw: SceneData comes from host memory every frame) plus a device->host
read of the accumulation image into pinned memory every step (after an NCCL gather for N > 1), host wall clock, max over ranks.
The per-step working set (wavefront queues, ~28 GB) is far larger than the 126 MB L2, so no explicit L2 flush is needed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = os.path.join(ROOT, "assets", "scenes", "cornell.json")
TABLE = os.path.join(ROOT, "assets", "rgb2spec", "srgb.coeff")
METRIC = "spectral 1080p Mpaths/s"
FLAG_COUNT_RAYS, FLAG_STAGE_TIMING = 1, 4

# Algorithmic bytes (DESIGN.md "Roofline accounting"; sizes are the actual struct sizes of vkrt_b200/csrc/wavefront.cuh, accel.cuh)
B_EXT_RAY = 52          # trace: read rayO 16 + rayD 16, write hitA 16 + hitB 4
B_SHADOW_RAY = 40       # trace: read shO 16 + shD 16 + shTarget 8 (contribution/radiance only touched when unoccluded)
B_NODE, B_TRI, B_INSTANCE = 80, 48, 64
B_SHADE_READ_HERO = 32 + 16 + 16 + 16 + 16 + 16    # ray, hitA, meta {record, rng, flags, hit v}, thr4, heroMisc, techPdf
B_SHADE_READ_FRESH = 32                             # thr4 + techPdf are constants at depth 0 and are not read for camera paths
B_SHADE_SURFACE = 12 + 144 + 80 + 32 + 272           # indices, 3 ShaderVertex, MeshInfo, MeshTrig, Material
B_SHADE_WRITE_PATH_HERO = 32 + 16 + 16 + 16 + 16 + 16 + 16   # ray, meta, thr4, heroMisc, techPdf, prevVertexTechPdf, prevBsdfTechPdf
B_SHADE_WRITE_SHADOW = 16 + 16 + 16 + 8 + 4
B_SHADE_FEATURES = 36                                 # featA, featB, follow at depth 0
B_BUILD_TRIANGLE = 520   # BVH build (DESIGN.md "Build"): 48 vertices + 12 key/index + 8 sort passes x 24 + 2 x 64 binary nodes + 64 refit + ~27 BVH8 + 48 re-pack


def derived_roofline_fields(roofline, c3, peak):
    """SURVEY 8(d) asks for the traversal / shading bytes at two levels and for the build's fraction. Pure arithmetic on numbers measured
    elsewhere in this run (nothing is timed here): `roofline.dram` = the ncu DRAM bytes per launch of profiles/traffic.json over this run's live
    launch time, next to the algorithmic figure (their ratio is what L1 / L2 serve); `c3.build_*` = the 10 M-triangle build against the HBM peak."""
    if roofline and roofline.get("traffic") and roofline.get("avg_launch_ms", 0) > 0 and peak > 0:
        gbps = roofline["traffic"] / 1e9 / (roofline["avg_launch_ms"] * 1e-3)
        roofline["dram"] = {"GBps": gbps, "frac": gbps / peak, "algorithmic_over_dram": roofline["algorithmic_bytes_per_launch"] / roofline["traffic"],
                            "note": "DRAM bytes per launch from the ncu capture named in traffic_source (32 spp per step), divided by this run's live launch time"}
    if c3 and c3.get("bvh_build_ms", 0) > 0 and peak > 0:
        gbps = c3["triangles"] * B_BUILD_TRIANGLE / 1e9 / (c3["bvh_build_ms"] * 1e-3)
        c3["build_algorithmic_bytes_per_triangle"] = B_BUILD_TRIANGLE
        c3["build_achieved_GBps_algorithmic"] = gbps
        c3["build_roofline_frac"] = gbps / peak


def procedural_sky(width=1024, height=512):
    """Lat-long HDR environment in the reference's convention (light/environment.slang:9-14): u = frac(phi/2pi + 0.5),
    v = theta/pi with theta measured from +Z. Horizon-to-zenith gradient plus a soft sun disk."""
    v = (np.arange(height, dtype=np.float32) + 0.5) / height
    u = (np.arange(width, dtype=np.float32) + 0.5) / width
    theta = v[:, None] * np.pi
    phi = (u[None, :] - 0.5) * 2.0 * np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta) * np.ones_like(phi)], axis=-1)
    up = np.clip(d[..., 2], 0.0, 1.0)
    sky = np.array([0.35, 0.55, 1.0], np.float32) * (0.25 + 0.75 * up[..., None]) + np.array([0.9, 0.8, 0.7], np.float32) * (1.0 - up[..., None]) ** 4 * 0.6
    ground = np.array([0.12, 0.10, 0.08], np.float32)
    img = np.where(d[..., 2:3] >= 0.0, sky, ground)
    el, az = np.radians(40.0), np.radians(-60.0)
    sun = np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], np.float32)
    cosang = np.clip((d * sun).sum(-1), -1.0, 1.0)
    img = img + np.array([1.0, 0.92, 0.8], np.float32) * 25.0 * np.exp(-((np.arccos(cosang) / np.radians(6.0)) ** 2))[..., None]
    out = np.ones((height, width, 4), np.float32)
    out[..., :3] = img
    return out


def setup_scene(host_mod, width, height, spp, **kw):
    hs = host_mod.Host(width=width, height=height, **kw)
    hs.load_scene(SCENE)
    hs.set_render_mode(1)                  # VKRT_RENDER_MODE_SPECTRAL
    hs.set_spectral_sampling_mode(1)       # VKRT_SPECTRAL_SAMPLING_MODE_HERO
    hs.set_environment_texture(procedural_sky())
    hs.set_environment_light((1.0, 1.0, 1.0), 1.0)
    hs.load_rgb2spec(TABLE)
    hs.set_samples_per_pixel(spp)
    hs.start_render(width, height, 0xFFFFFFFF)   # like the reference's offline loop: VKRT_startRender(w, h, UINT32_MAX), benchmark.c:213
    return hs


def ensure_built():
    table_missing = not os.path.exists(TABLE)
    if table_missing or not os.path.exists(os.path.join(ROOT, "vkrt_b200", "libvkrt_host.so")):
        import __graft_entry__ as g
        g.build()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while a timed region runs (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.lines, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if not any(a <= ts <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(f) > 3 + k and f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (oracle/oracle.cpp, a line-by-line CPU restatement of the reference's shaders with a
# BVH2 traverser) on the host cores. This is the only place bench.py touches oracle/.
# ------------------------------------------------------------------------------------------------------------------------------
def cpu_tag():
    """Names this host's CPU (model + feature flags): the -march=native oracle build is per machine, and oracle/_build travels from the
    build container to the GPU box."""
    import hashlib
    model, flags = "", ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and not model:
                model = line.split(":", 1)[1].strip()
            elif line.startswith("flags") and not flags:
                flags = " ".join(sorted(line.split(":", 1)[1].split()))
            if model and flags:
                break
    except OSError:
        pass
    return hashlib.sha1((model + "|" + flags).encode()).hexdigest()[:12]


class OracleRunner:
    """The CPU arm. kind = "reference": the reference's own shaders (src/shaders/**/*.slang, transliterated to C++ by oracle/ref_slang when
    `make -C oracle ref` ran where the reference checkout exists; oracle/_ref/shaders_gen.inc travels to the GPU box) over the oracle's
    BVH2 traversal and texture sampler, which stand in for the Vulkan driver the reference gets them from. kind = "port": the oracle's
    restatement of the same shaders, used only when oracle/_ref is absent. Both are rebuilt -O3 -march=native for the machine they run on."""

    def __init__(self, prep, width, height, threads):
        tag = cpu_tag()
        odir = os.path.join(ROOT, "oracle")
        self.kind = "port"
        if os.path.exists(os.path.join(odir, "_ref", "shaders_gen.inc")):
            try:
                subprocess.check_call(["make", "-s", "-C", odir, "refnative", "NATIVE_TAG=" + tag])
                self.lib = C.CDLL(os.path.join(odir, "_ref", "libvkrt_refshade_native_%s.so" % tag))
                self.kind = "reference"
            except (subprocess.CalledProcessError, OSError) as e:
                print("bench.py: reference-shader library unavailable (%s); falling back to the oracle port" % e, file=sys.stderr)
        if self.kind == "port":
            subprocess.check_call(["make", "-s", "-C", odir, "native", "NATIVE_TAG=" + tag])
            self.lib = C.CDLL(os.path.join(odir, "_build", "liboracle_native_%s.so" % tag))
        self.ctx = C.c_void_p()
        assert self.lib.oracle_create(C.byref(self.ctx)) == 0
        self.threads = threads or self.lib.oracle_max_threads()
        self.lib.oracle_set_threads(self.ctx, self.threads)
        p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)  # noqa: E731
        self._keep = [np.ascontiguousarray(v) for v in prep.values() if isinstance(v, np.ndarray)]
        ok = self.lib.oracle_set_geometry(self.ctx, p(prep["vertices"]), C.c_uint32(len(prep["vertices"])), p(prep["indices"]), C.c_uint32(len(prep["indices"])))
        ok |= self.lib.oracle_set_instances(self.ctx, p(prep["meshInfos"]), p(prep["world3x4"]), p(prep["geometrySource"]), p(prep["alphaTested"]),
                                            C.c_uint32(len(prep["meshInfos"])))
        ok |= self.lib.oracle_set_materials(self.ctx, p(prep["materials"]), C.c_uint32(len(prep["materials"])))
        nm, nt = len(prep["emissiveMeshes"]), len(prep["emissiveTriangles"])
        pad = lambda a, dt: a if len(a) else np.zeros(1, dt)  # noqa: E731
        ok |= self.lib.oracle_set_lights(self.ctx, p(pad(prep["emissiveMeshes"], prep["emissiveMeshes"].dtype)), C.c_uint32(nm),
                                         p(pad(prep["emissiveTriangles"], prep["emissiveTriangles"].dtype)), C.c_uint32(nt), p(pad(prep["meshAliasQ"], np.float32)),
                                         p(pad(prep["meshAliasIdx"], np.uint32)), p(pad(prep["triAliasQ"], np.float32)), p(pad(prep["triAliasIdx"], np.uint32)))
        assert ok == 0, "oracle scene upload failed"
        self.width, self.height = width, height

    def set_textures(self, textures):
        class Tex(C.Structure):
            _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("colorSpace", C.c_uint32)]
        arr = (Tex * len(textures))()
        for i, t in enumerate(textures):
            self._keep.append(t)
            arr[i] = Tex(t.ctypes.data, t.shape[1], t.shape[0], 3, 1)
        assert self.lib.oracle_set_textures(self.ctx, arr, C.c_uint32(len(textures))) == 0

    def set_rgb2spec(self, path):
        raw = open(path, "rb").read()
        res = int(np.frombuffer(raw, "<u4", 1, 4)[0])
        payload = np.frombuffer(raw, "<f4", res + 9 * res ** 3, 8).copy()
        self._keep.append(payload)

        class Info(C.Structure):
            _fields_ = [("res", C.c_uint32), ("scaleOffset", C.c_uint32), ("dataOffset", C.c_uint32)]
        assert self.lib.oracle_set_rgb2spec(self.ctx, payload.ctypes.data_as(C.c_void_p), C.c_uint32(len(payload)), Info(res, 0, res)) == 0

    def finish(self):
        assert self.lib.oracle_build_accel(self.ctx) == 0
        assert self.lib.oracle_resize(self.ctx, C.c_uint32(self.width), C.c_uint32(self.height)) == 0

    def render_bands(self, sd_bytes, frame, spp, bands, rows_per_band):
        """Renders `bands` bands of `rows_per_band` rows spread evenly over the image; returns (paths, seconds)."""
        sd = np.frombuffer(sd_bytes, dtype=np.uint8).copy()
        sd[128:132] = np.frombuffer(np.uint32(frame).tobytes(), np.uint8)
        sd[132:136] = np.frombuffer(np.uint32(spp).tobytes(), np.uint8)
        rays = (C.c_uint64 * 2)()
        t0 = time.perf_counter()
        paths = 0
        for b in range(bands):
            y0 = int((b + 0.5) * self.height / bands - rows_per_band / 2)
            y0 = max(0, min(self.height - rows_per_band, y0))
            if self.kind == "reference":
                rc = self.lib.refshade_render_frame_rows(self.ctx, sd.ctypes.data_as(C.c_void_p), C.c_uint32(y0), C.c_uint32(y0 + rows_per_band))
            else:
                rc = self.lib.oracle_render_frame_rows(self.ctx, sd.ctypes.data_as(C.c_void_p), C.c_uint32(y0), C.c_uint32(y0 + rows_per_band), rays)
            assert rc == 0, "CPU render failed"
            paths += rows_per_band * self.width * spp
        return paths, time.perf_counter() - t0


CPU_NOTE = {
    "reference": "the reference's own Slang shaders (raygen / hit / miss / any-hit, BSDF, NEE, MIS, film) transliterated to C++ and compiled -O3 "
                 "-march=native (oracle/ref_slang), std::thread over rows; BVH2 traversal + bilinear sampler are the oracle's stand-ins for the "
                 "Vulkan driver (the reference repository holds no traversal code); lavapipe itself is not installable here",
    "port": "CPU oracle restatement of the reference shaders (oracle/oracle.cpp, BVH2, std::thread over rows); oracle/_ref was not available",
}


def measure_c3(host_mod, device, triangles):
    """BASELINE.json configs[2] in the same run: the C host's procedural triangle soup, BVH build time (second build: the first one of a
    process also allocates the builder's scratch), traversal rate of RGB frames at 1080p with CUDA events around the k_trace launches,
    and the visit counts of one instrumented frame."""
    w, h, spp = 1920, 1080, 8
    out = {"triangles": triangles, "frame": "%dx%d RGB, %d spp per frame, depth 4..8, NEE+MIS" % (w, h, spp)}
    hs = host_mod.Host(width=w, height=h, device=device, max_paths=w * h * spp + 65536, cuda_flags=FLAG_STAGE_TIMING)
    hs.generate_soup(triangles, 1)
    hs.set_render_mode(0)
    hs.set_samples_per_pixel(spp)
    hs.start_render(w, h, 0xFFFFFFFF)
    hs.draw()                      # upload + first build + one warm-up frame
    import vkrt_b200
    lib = vkrt_b200.load_library()
    again = vkrt_b200.BuildStats()
    lib.vkrt_cuda_invalidate_accel(C.c_void_p(hs.cuda_context()))
    assert lib.vkrt_cuda_build_accel(C.c_void_p(hs.cuda_context()), C.byref(again)) == 0
    out.update({"bvh_build_ms": again.buildMs, "build_Mtris_per_s": again.triangleCount / max(again.buildMs, 1e-6) / 1e3, "bvh8_nodes": int(again.bvh8NodeCount),
                "accel_MB": again.accelBytes / 1e6, "hierarchy": "PLOC" if again.plocHierarchies else "radix tree (lower surface-area cost than PLOC)"})
    rays = tms = fms = paths = 0.0
    for _ in range(3):
        hs.draw()
        st = hs.last_frame_stats()
        rays += st.extensionRays + st.shadowRays
        tms += st.traceMs
        fms += st.frameMs
        paths += st.paths
    out.update({"mrays_per_s_trace": rays / max(tms, 1e-9) / 1e3, "mpaths_per_s": paths / max(fms, 1e-9) / 1e3, "trace_share_of_frame": tms / max(fms, 1e-9)})
    hs.close()
    hc = host_mod.Host(width=w, height=h, device=device, max_paths=w * h * 2 + 65536, cuda_flags=FLAG_COUNT_RAYS)
    hc.generate_soup(triangles, 1)
    hc.set_render_mode(0)
    hc.set_samples_per_pixel(2)
    hc.start_render(w, h, 0xFFFFFFFF)
    hc.draw()
    cs = hc.last_frame_stats()
    r = max(cs.extensionRays + cs.shadowRays, 1)
    out["visits"] = {"nodes_per_ray": cs.nodesVisited / r, "triangles_per_ray": cs.trianglesTested / r}
    out["achieved_GBps_algorithmic"] = out["mrays_per_s_trace"] * 1e6 * (0.58 * B_EXT_RAY + 0.42 * B_SHADOW_RAY + out["visits"]["nodes_per_ray"] * B_NODE +
                                                                          out["visits"]["triangles_per_ray"] * B_TRI) / 1e9
    hc.close()
    return out


def make_oracle(width, height, threads):
    from vkrt_b200 import host
    hs = setup_scene(host, width, height, 1, host_only=True)   # the C host prepares the scene; no device involved
    prep = hs.prepare_scene()
    orc = OracleRunner(prep, width, height, threads)
    orc.set_textures([procedural_sky()])
    orc.set_rgb2spec(TABLE)
    orc.finish()
    sd = prep["sceneData"].tobytes()
    hs.close()
    return orc, sd


def cpu_sample(orc, sd, budget_s, frame=0):
    """Bounded sample: 4 bands x (8 rows per thread) x 1 spp at a time until `budget_s` seconds of CPU work have been done."""
    paths, secs, n = 0, 0.0, 0
    rows = min(8 * orc.threads, orc.height // 4)
    while secs < budget_s:
        p, s = orc.render_bands(sd, frame + n, 1, 4, rows)
        paths, secs, n = paths + p, secs + s, n + 1
    return paths, secs, "%d x (4 bands x %d rows x %d px x 1 spp) of the 1080p spectral-hero frame" % (n, rows, orc.width)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ensure_built()
    w, h = args.width, args.height
    orc, sd = make_oracle(w, h, 0)
    per_step_budget = args.ref_step_seconds
    for k in range(args.warmup):
        cpu_sample(orc, sd, min(1.0, per_step_budget), frame=k)
    paths, secs, sample = 0, 0.0, ""
    for k in range(args.steps):
        p, s, sample = cpu_sample(orc, sd, per_step_budget, frame=100 + k)
        paths, secs = paths + p, secs + s
    value = paths / secs / 1e6
    line = {"metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": "cornell.json (bunny) 1920x1080 spectral hero, procedural HDR sky, depth 4..8, NEE+MIS; each step = a bounded sample of that frame",
                       "note": CPU_NOTE[orc.kind]},
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": orc.threads, "kind": orc.kind, "sample": "per step: " + sample},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    # fd 1 carries exactly one line (the JSON result): libraries that print banners to stdout while they initialise (NCCL's version
    # line) are sent to stderr; the real stdout comes back for the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return _run_ours(args, real_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)


def _run_ours(args, real_stdout):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ensure_built()
    if world > 1:
        dist.barrier()

    import vkrt_b200
    from vkrt_b200 import host
    lib = vkrt_b200.load_library()
    w, h = args.width, args.height
    spp = args.spp * world   # fixed per-GPU work: the image is split N ways, the samples per step grow N-fold
    t_setup = time.perf_counter()
    hs = setup_scene(host, w, h, spp, device=local_rank, rank=rank, world_size=world, max_paths=(w * h * args.spp * 11) // 10, cuda_flags=FLAG_STAGE_TIMING)
    hs.update_scene()    # scene upload + BVH build (outside the timed steps, like model load)
    setup_s = time.perf_counter() - t_setup
    ctx = C.c_void_p(hs.cuda_context())
    bs = hs.build_stats()
    prep = hs.prepare_scene()
    scene_bytes = sum(prep[k].nbytes for k in ("vertices", "indices", "meshInfos", "world3x4", "materials", "emissiveMeshes", "emissiveTriangles",
                                               "meshAliasQ", "meshAliasIdx", "triAliasQ", "triAliasIdx")) + os.path.getsize(TABLE) + 1024 * 512 * 16
    if world > 1:   # NCCL communicator of the library itself (film gather): unique id from rank 0, broadcast with torch.distributed
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident = torch.frombuffer(bytearray(vkrt_b200.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(ident, 0)
        buf = C.create_string_buffer(bytes(ident.cpu().numpy().tobytes()), 128)
        rc = lib.vkrt_cuda_comm_init(ctx, buf)
        assert rc == 0, (lib.vkrt_cuda_last_error(ctx) or b"").decode()

    def check(rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, (lib.vkrt_cuda_last_error(ctx) or b"").decode()))

    def barrier():
        check(lib.vkrt_cuda_sync(ctx), "sync")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sd = np.frombuffer(prep["sceneData"].tobytes(), dtype=np.uint8).copy()
    frame_counter = [0]

    def enqueue_frame():
        sd[128:132] = np.frombuffer(np.uint32(frame_counter[0]).tobytes(), np.uint8)
        frame_counter[0] += 1
        check(lib.vkrt_cuda_render_frame_async(ctx, sd.ctypes.data_as(C.c_void_p)), "render_frame_async")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    windows = []

    # ---- value: K frames, inputs resident, device events on the library's stream ----
    for _ in range(args.warmup):
        enqueue_frame()
    gather_ms = C.c_float(0.0)
    if world > 1:   # the first NCCL exchange sets up the peer connections (hundreds of ms at 8 ranks): that belongs to the warm-up
        check(lib.vkrt_cuda_gather(ctx, C.byref(gather_ms)), "gather (warm-up)")
    barrier()
    t0 = time.time()
    check(lib.vkrt_cuda_timer_begin(ctx), "timer_begin")
    for _ in range(args.steps):
        enqueue_frame()
    if world > 1:
        check(lib.vkrt_cuda_gather(ctx, C.byref(gather_ms)), "gather")
    ms = C.c_float()
    check(lib.vkrt_cuda_timer_end(ctx, C.byref(ms)), "timer_end")
    barrier()
    windows.append((t0, time.time()))
    device_ms = max_over_ranks(float(ms.value))
    total_paths = float(w) * h * spp * args.steps   # all ranks together: every pixel of the image, spp samples, K steps
    value = total_paths / (device_ms * 1e-3) / 1e6

    # ---- e2e: the same through the host API, with a device->host read of the result every step ----
    hs.lib.VKRT_invalidateAccumulation(hs.h)
    pinned = torch.empty((h, w, 4), dtype=torch.float32, pin_memory=True)
    pinned_ptr = C.c_void_p(pinned.data_ptr())
    launches = trace_launches = shade_launches_proper = 0
    trace_ms = shade_ms = frame_ms = shade_kernel_ms = 0.0
    ext = sh = local_paths = 0
    for _ in range(args.warmup):   # warm-up steps are whole steps: draw + gather + device->host read (the first read into a fresh pinned
        hs.draw()                  # buffer costs ~60 ms once)
        if world > 1:
            check(lib.vkrt_cuda_gather(ctx, C.byref(gather_ms)), "gather")
        if rank == 0:
            check(lib.vkrt_cuda_read_aov(ctx, C.c_int(0), pinned_ptr, C.c_size_t(w * h * 16)), "read_aov")
    barrier()
    t0 = time.time()
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        hs.draw()
        st = hs.last_frame_stats()
        launches += st.kernelLaunches
        trace_launches += st.traceLaunches
        trace_ms += st.traceMs
        shade_ms += st.shadeMs
        shade_kernel_ms += st.shadeKernelMs
        shade_launches_proper += st.shadeLaunches
        frame_ms += st.frameMs
        ext += st.extensionRays
        sh += st.shadowRays
        local_paths += st.paths
        if world > 1:
            check(lib.vkrt_cuda_gather(ctx, C.byref(gather_ms)), "gather")
        if rank == 0:
            check(lib.vkrt_cuda_read_aov(ctx, C.c_int(0), pinned_ptr, C.c_size_t(w * h * 16)), "read_aov")
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - tw0)
    windows.append((t0, time.time()))
    e2e_value = total_paths / e2e_s / 1e6
    clocks = None
    if rank == 0:
        time.sleep(0.25)
        sampler.stop()
        clocks = sampler.summary(windows)
    all_ext, all_sh = sum_over_ranks(float(ext)), sum_over_ranks(float(sh))
    all_launches = int(sum_over_ranks(float(launches)))

    # ---- roofline of the dominant kernel (by measured time), from the per-launch events of the e2e region on rank 0 ----
    # ---- strong scaling (N > 1): the SAME 32-spp step split over the N GPUs' tiles, device-timed like `value` ----
    strong = None
    if world > 1:
        hs.lib.VKRT_invalidateAccumulation(hs.h)
        sd_strong = sd.copy()
        sd_strong[132:136] = np.frombuffer(np.uint32(args.spp).tobytes(), np.uint8)

        def enqueue_strong():
            sd_strong[128:132] = np.frombuffer(np.uint32(frame_counter[0]).tobytes(), np.uint8)
            frame_counter[0] += 1
            check(lib.vkrt_cuda_render_frame_async(ctx, sd_strong.ctypes.data_as(C.c_void_p)), "render_frame_async")
        for _ in range(args.warmup):
            enqueue_strong()
        barrier()
        check(lib.vkrt_cuda_timer_begin(ctx), "timer_begin")
        for _ in range(args.steps):
            enqueue_strong()
        check(lib.vkrt_cuda_gather(ctx, C.byref(gather_ms)), "gather")
        ms2 = C.c_float()
        check(lib.vkrt_cuda_timer_end(ctx, C.byref(ms2)), "timer_end")
        barrier()
        strong_ms = max_over_ranks(float(ms2.value))
        strong = {"value": float(w) * h * args.spp * args.steps / (strong_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": strong_ms / args.steps,
                  "spp_per_step": args.spp, "note": "total work fixed: every GPU renders its interleaved tiles of the same %d-spp 1080p step; one NCCL gather inside the timed region" % args.spp}

    roofline = cpu_base = kernels = c3 = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        # per-ray BVH visit counts from one instrumented frame (same kernels with counters, untimed, reduced spp)
        visits = None
        if not args.no_visit_counts:   # (every N: rank 0 renders one small full-frame probe on its own GPU; the BVH is the same on all ranks)
            hc = setup_scene(host, w, h, 2, device=local_rank, max_paths=w * h * 2 + 65536, cuda_flags=FLAG_COUNT_RAYS)
            hc.draw()
            cs = hc.last_frame_stats()
            r = max(cs.extensionRays + cs.shadowRays, 1)
            visits = {"nodes_per_ray": cs.nodesVisited / r, "triangles_per_ray": cs.trianglesTested / r, "instances_per_ray": cs.instancesEntered / r}
            hc.close()
        other_launches = launches - trace_launches - shade_launches_proper   # raygen, material sort, film
        n_shade = shade_launches_proper              # k_shade launches proper, counted by the library (one per depth and sample chunk)
        shade_bytes = (ext * (B_SHADE_READ_HERO + B_SHADE_SURFACE) - local_paths * B_SHADE_READ_FRESH + max(ext - local_paths, 0) * B_SHADE_WRITE_PATH_HERO +
                       sh * B_SHADE_WRITE_SHADOW + local_paths * B_SHADE_FEATURES)
        trace_bytes = ext * B_EXT_RAY + sh * B_SHADOW_RAY
        if visits:
            trace_bytes += (ext + sh) * (visits["nodes_per_ray"] * B_NODE + visits["triangles_per_ray"] * B_TRI + visits["instances_per_ray"] * B_INSTANCE)
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except (OSError, ValueError):
            pass
        kernels = {
            "k_shade<hero>": {"launches": n_shade, "ms_total": shade_kernel_ms,
                              "note": "CUDA events around the k_shade launches alone; raygen + material sort + film (%d launches) took %.2f ms more" % (
                                  other_launches, shade_ms - shade_kernel_ms),
                              "algorithmic_GB": shade_bytes / 1e9, "achieved_GBps": shade_bytes / 1e9 / max(shade_kernel_ms * 1e-3, 1e-9)},
            "k_trace": {"launches": trace_launches, "ms_total": trace_ms, "algorithmic_GB": trace_bytes / 1e9, "achieved_GBps": trace_bytes / 1e9 / max(trace_ms * 1e-3, 1e-9),
                        "Mrays_per_s": (ext + sh) / max(trace_ms * 1e-3, 1e-9) / 1e6, "visits": visits},
        }
        dom = "k_shade<hero>" if shade_kernel_ms >= trace_ms else "k_trace"
        kd = kernels[dom]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kd["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": kd["achieved_GBps"] / peak,
                    "traffic": traffic.get(dom), "traffic_source": traffic.get("_source"), "peak_source": peak_src, "avg_launch_ms": kd["ms_total"] / max(kd["launches"], 1),
                    "algorithmic_bytes_per_launch": kd["algorithmic_GB"] * 1e9 / max(kd["launches"], 1),
                    "share_of_step": kd["ms_total"] / max(frame_ms, 1e-9)}
        if world == 1 and not args.no_c3:
            c3 = measure_c3(host, local_rank, args.c3_triangles)
        if world == 1 and not args.no_cpu_baseline:
            orc, osd = make_oracle(w, h, 0)
            cpu_sample(orc, osd, 1.0)
            p, s, sample = cpu_sample(orc, osd, args.cpu_seconds, frame=10)
            cpu_base = {"value": p / s / 1e6, "unit": "Mpaths/s", "cores": orc.threads, "kind": orc.kind, "sample": sample, "note": CPU_NOTE[orc.kind]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": device_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cornell.json (bunny for the missing dragon) %dx%d spectral hero, procedural HDR sky env map, depth 4..8, NEE+MIS" % (w, h),
                       "spp_per_step": spp, "paths_per_step": w * h * spp, "parallelism": "tiles32x32 x%d, scene replicated" % world,
                       "l2": "per-step working set (wavefront queues ~%.1f GB) >> 126 MB L2, no flush needed" % (w * h * args.spp * 420 / 1e9),
                       "setup_s": setup_s, "scene_upload_bytes": scene_bytes, "bvh_build_ms": bs.buildMs, "triangles": int(bs.triangleCount),
                       "mrays_per_s": (all_ext + all_sh) / e2e_s / 1e6, "rays_per_path": (all_ext + all_sh) / max(total_paths, 1),
                       "gather_ms": float(gather_ms.value)},
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": 240, "d2h_bytes_per_step": w * h * 16, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": all_launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
        }
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        try:
            if args.spp == 32 or not (roofline or {}).get("traffic"):
                derived_roofline_fields(roofline, c3, peak)
            else:   # traffic.json was captured at 32 spp per step: per-launch DRAM bytes of another step size are not comparable
                derived_roofline_fields(None, c3, peak)
        except Exception as e:   # derived, optional fields must never cost the bench line
            sys.stderr.write("derived roofline fields skipped: %r\n" % (e,))
        if c3:
            line["config"]["c3"] = c3
        if strong:
            line["config"]["strong"] = strong
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    hs.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=32, help="samples per pixel per step and per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work spent on the cpu_baseline sample")
    ap.add_argument("--ref-step-seconds", type=float, default=3.0, help="CPU work per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-visit-counts", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the 10 M-triangle traversal block (config.c3)")
    ap.add_argument("--c3-triangles", type=int, default=10_000_000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
