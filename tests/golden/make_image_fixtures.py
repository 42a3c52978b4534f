#!/usr/bin/env python3
"""Generates tests/golden/images.npz: small PNG / JPEG / EXR files (as byte arrays) next to the pixels that the reference's decoders
produce for them. The reference decodes PNG with libspng, JPEG with libjpeg-turbo (TJFLAG_ACCURATEDCT) and EXR with tinyexr
(src/core/utility/image.c:123-330, exr.cpp); Pillow decodes JPEG with the same libjpeg-turbo (islow IDCT, fancy upsampling) and OpenCV
writes/reads EXR with OpenEXR, so their outputs pin vkrt_b200/host/image_decode.c without either library being needed at test time.

  python tests/golden/make_image_fixtures.py        (needs Pillow and opencv-python; run where they exist, commit the .npz)
"""
import io
import os

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20240611)
h, w = 23, 37
y, x = np.mgrid[0:h, 0:w]
rgb = np.stack([128 + 100 * np.sin(x / 5.0) * np.cos(y / 4.0), 128 + 90 * np.cos(x / 3.0 + y / 6.0), x * 255.0 / w], -1) + rng.normal(0, 6, (h, w, 3))
rgb = np.clip(rgb, 0, 255).astype(np.uint8)
rgba = np.dstack([rgb, 255 - rgb[..., :1]])
out = {}


def put(name, blob, pixels):
    out["file_" + name] = np.frombuffer(blob, np.uint8)
    out["want_" + name] = pixels


# ---- JPEG: 4:4:4, 4:2:2, 4:2:0, grey; odd sizes; a restart interval ----
for tag, mode, sub, q, size in [("444_q90", "RGB", 0, 90, (w, h)), ("422_q75", "RGB", 1, 75, (w, h)), ("420_q50", "RGB", 2, 50, (w, h)),
                                ("420_17x9", "RGB", 2, 95, (17, 9)), ("420_1x1", "RGB", 2, 80, (1, 1)), ("grey_q85", "L", 0, 85, (w, h)),
                                ("422_8x24", "RGB", 1, 100, (8, 24))]:
    im = Image.fromarray(rgb).convert(mode).resize(size)
    buf = io.BytesIO()
    kw = dict(quality=q)
    if mode == "RGB":
        kw["subsampling"] = sub
    im.save(buf, "JPEG", **kw)
    want = np.array(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    put("jpeg_" + tag, buf.getvalue(), np.dstack([want, np.full(want.shape[:2] + (1,), 255, np.uint8)]))

# ---- PNG: every colour type; 16-bit; palette with tRNS (the reference does not apply tRNS: alpha stays opaque) ----
for tag, mode in [("rgb8", "RGB"), ("rgba8", "RGBA"), ("grey8", "L"), ("greyalpha8", "LA"), ("palette", "P"), ("bilevel", "1")]:
    im = Image.fromarray(rgba, "RGBA").convert(mode)
    buf = io.BytesIO()
    im.save(buf, "PNG")
    want = np.array(Image.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
    if mode == "P":
        want[..., 3] = 255
    put("png_" + tag, buf.getvalue(), want)
g16 = (np.array(Image.fromarray(rgb).convert("L")).astype(np.uint16) * 257 + 3)
buf = io.BytesIO()
Image.fromarray(g16).save(buf, "PNG")
put("png_grey16", buf.getvalue(), np.dstack([g16, g16, g16, np.full_like(g16, 65535)]))
ok, enc = cv2.imencode(".png", (rgba.astype(np.uint16) * 257)[..., [2, 1, 0, 3]])
put("png_rgba16", enc.tobytes(), rgba.astype(np.uint16) * 257)

# ---- EXR: float / half, NONE / RLE / ZIPS / ZIP / PIZ, RGB / RGBA / Y; one PXR24 file (unsupported: must be rejected) ----
# a 70-row smooth image spans three 32-line PIZ blocks and exercises the 14-bit wavelet and the run-length symbol
hdr = (rng.random((h, w, 3)) * 10).astype(np.float32)
hdr[3, 4] = [1e4, 0.0, 1e-5]
for tag, typ, comp, chans in [("float_none_rgb", cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_NO, 3),
                              ("float_zip_rgba", cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_ZIP, 4),
                              ("float_rle_y", cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_RLE, 1),
                              ("half_zips_rgb", cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_ZIPS, 3),
                              ("half_zip_rgba", cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_ZIP, 4),
                              ("float_piz_rgb", cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION_PIZ, 3),
                              ("half_piz_rgba", cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_PIZ, 4),
                              ("half_pxr24_rgb", cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION_PXR24, 3)]:
    src = hdr if chans == 3 else (np.dstack([hdr, hdr[..., :1] * 0.1]) if chans == 4 else hdr[..., 0])
    bgr = src[..., ::-1] if chans == 3 else (src[..., [2, 1, 0, 3]] if chans == 4 else src)
    ok, enc = cv2.imencode(".exr", bgr, [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, comp])
    want = np.ones((h, w, 4), np.float32)
    if chans == 1:
        want[..., 0] = want[..., 1] = want[..., 2] = src
    else:
        want[..., :chans] = src
    if typ == cv2.IMWRITE_EXR_TYPE_HALF:
        want = want.astype(np.float16)
    put("exr_" + tag, enc.tobytes(), want)

yy, xx = np.mgrid[0:70, 0:45]
smooth = np.dstack([xx / 44.0, yy / 69.0, np.full(xx.shape, 0.25)]).astype(np.float32)
ok, enc = cv2.imencode(".exr", smooth[..., ::-1], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ])
want = np.ones((70, 45, 4), np.float16)
want[..., :3] = smooth
put("exr_half_piz_smooth", enc.tobytes(), want)

np.savez_compressed(os.path.join(HERE, "images.npz"), **out)
print("wrote", os.path.join(HERE, "images.npz"), sum(v.nbytes for v in out.values()), "bytes in", len(out), "arrays")
