"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py

Every rank renders its interleaved tiles of the bundled cornell scene (C host -> C ABI), the film is gathered to rank 0 with the
library's NCCL exchange (vkrt_cuda_gather), and rank 0 compares the assembled image bit for bit with a single-GPU render of the
same frames: the tile partition moves no floating-point value across ranks, so the result must be identical."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vkrt_b200  # noqa: E402
from vkrt_b200 import host  # noqa: E402


def render(w, h, spp, frames, spectral, **kw):
    hs = host.Host(width=w, height=h, **kw)
    hs.load_scene(os.path.join(ROOT, "assets", "scenes", "cornell.json"))
    if spectral:
        hs.set_render_mode(1)
        hs.set_spectral_sampling_mode(1)
        hs.load_rgb2spec(os.path.join(ROOT, "assets", "rgb2spec", "srgb.coeff"))
    hs.set_samples_per_pixel(spp)
    hs.start_render(w, h, spp * frames)
    return hs


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = vkrt_b200.load_library()
    w, h, spp, frames = 333, 190, 3, 2   # not a multiple of the tile size
    ok = True
    for spectral in (False, True):
        hs = render(w, h, spp, frames, spectral, device=local, rank=rank, world_size=world)
        hs.update_scene()
        ctx = C.c_void_p(hs.cuda_context())
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident = torch.frombuffer(bytearray(vkrt_b200.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(ident, 0)
        assert lib.vkrt_cuda_comm_init(ctx, C.create_string_buffer(ident.cpu().numpy().tobytes(), 128)) == 0
        for _ in range(frames):
            hs.draw()
        ms = C.c_float()
        rc = lib.vkrt_cuda_gather(ctx, C.byref(ms))
        assert rc == 0, (lib.vkrt_cuda_last_error(ctx) or b"").decode()
        if rank == 0:
            got = {}
            for which, dt, nc in ((0, np.float32, 4), (1, np.uint16, 4), (3, np.uint16, 4)):
                out = np.zeros((h, w, nc), dt)
                assert lib.vkrt_cuda_read_aov(ctx, C.c_int(which), out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)) == 0
                got[which] = out
            one = render(w, h, spp, frames, spectral, device=local)
            for _ in range(frames):
                one.draw()
            c1 = C.c_void_p(one.cuda_context())
            for which, dt, nc in ((0, np.float32, 4), (1, np.uint16, 4), (3, np.uint16, 4)):
                ref = np.zeros((h, w, nc), dt)
                assert lib.vkrt_cuda_read_aov(c1, C.c_int(which), ref.ctypes.data_as(C.c_void_p), C.c_size_t(ref.nbytes)) == 0
                same = np.array_equal(got[which].view(np.uint8), ref.view(np.uint8))
                print("mgpu_check world=%d spectral=%d aov=%d gather=%.3f ms: %s" % (world, spectral, which, ms.value, "bit-identical" if same else "MISMATCH"), flush=True)
                ok &= same
            assert np.all(got[0][..., 3] == spp * frames)
            one.close()
        hs.close()
        dist.barrier()
    if rank == 0:
        print("mgpu_check:", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
