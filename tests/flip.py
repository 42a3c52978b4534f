"""LDR-FLIP, the perceptual image-difference measure north_star names next to RMSE, restated from its publication (Andersson, Nilsson,
Akenine-Moller, Oskarsson, Astrom, Fairchild: "FLIP: A Difference Evaluator for Alternating Images", HPG 2020) — the published package
is not in this image and there is no network. Inputs are sRGB images in [0, 1], HxWx3; the result is the per-pixel error map in [0, 1]
(0 = identical). Constants are the paper's (Table 1 and section 4): contrast-sensitivity Gaussians for the achromatic / red-green /
blue-yellow channels, Hunt adjustment, HyAB distance with the 0.7 exponent and the (0.4, 0.95) remapping, edge / point detectors of
width 0.082 degrees, feature exponent 0.5; 67 pixels per degree (0.7 m from a 0.7 m wide 3840-pixel monitor) unless stated otherwise.
Test infrastructure only."""
import numpy as np
from scipy.ndimage import correlate

_RGB2XYZ = np.array([[10135552 / 24577794, 8788810 / 24577794, 4435075 / 24577794],
                     [2613072 / 12288897, 8788810 / 12288897, 887015 / 12288897],
                     [1425312 / 73733382, 8788810 / 73733382, 70074185 / 73733382]])
_XYZ2RGB = np.linalg.inv(_RGB2XYZ)
_WHITE = _RGB2XYZ @ np.ones(3)


def _srgb_to_linear(c):
    c = np.clip(c, 0.0, 1.0)
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def _linrgb_to_ycxcz(rgb):
    xyz = rgb @ _RGB2XYZ.T / _WHITE
    y = 116.0 * xyz[..., 1] - 16.0
    return np.stack([y, 500.0 * (xyz[..., 0] - xyz[..., 1]), 200.0 * (xyz[..., 1] - xyz[..., 2])], axis=-1)


def _ycxcz_to_linrgb(ycc):
    y = (ycc[..., 0] + 16.0) / 116.0
    xyz = np.stack([y + ycc[..., 1] / 500.0, y, y - ycc[..., 2] / 200.0], axis=-1) * _WHITE
    return xyz @ _XYZ2RGB.T


def _linrgb_to_lab(rgb):
    xyz = rgb @ _RGB2XYZ.T / _WHITE
    d = 6.0 / 29.0
    f = np.where(xyz > d ** 3, np.cbrt(np.maximum(xyz, 0.0)), xyz / (3 * d * d) + 4.0 / 29.0)
    return np.stack([116.0 * f[..., 1] - 16.0, 500.0 * (f[..., 0] - f[..., 1]), 200.0 * (f[..., 1] - f[..., 2])], axis=-1)


def _csf_kernel(ppd, a1, b1, a2, b2, radius):
    x = np.arange(-radius, radius + 1) / ppd
    d = x[None, :] ** 2 + x[:, None] ** 2
    g = a1 * np.sqrt(np.pi / b1) * np.exp(-np.pi ** 2 * d / b1) + a2 * np.sqrt(np.pi / b2) * np.exp(-np.pi ** 2 * d / b2)
    return g / g.sum()


def _spatial_filter(ycc, ppd):
    params = {"A": (1.0, 0.0047, 0.0, 1e-5), "RG": (1.0, 0.0053, 0.0, 1e-5), "BY": (34.1, 0.04, 13.5, 0.025)}
    radius = int(np.ceil(3.0 * np.sqrt(0.04 / (2.0 * np.pi ** 2)) * ppd))
    out = np.stack([correlate(ycc[..., k], _csf_kernel(ppd, *params[name], radius), mode="nearest") for k, name in enumerate(("A", "RG", "BY"))], axis=-1)
    return np.clip(_ycxcz_to_linrgb(out), 0.0, 1.0)


def _hunt(lab):
    out = lab.copy()
    out[..., 1] *= 0.01 * lab[..., 0]
    out[..., 2] *= 0.01 * lab[..., 0]
    return out


def _hyab(a, b):
    d = a - b
    return np.abs(d[..., 0]) + np.sqrt(d[..., 1] ** 2 + d[..., 2] ** 2)


def _feature_kernels(ppd):
    sd = 0.5 * 0.082 * ppd
    radius = int(np.ceil(3.0 * sd))
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    xx, yy = np.meshgrid(x, x)
    g = np.exp(-(xx ** 2 + yy ** 2) / (2.0 * sd * sd))
    edge = -xx * g
    point = (xx ** 2 / (sd * sd) - 1.0) * g
    kernels = []
    for k in (edge, point):
        neg, pos = -k[k < 0].sum(), k[k > 0].sum()
        kernels.append(np.where(k < 0, k / neg, k / pos))
    return kernels


def _feature_norm(y, kernel):
    fx = correlate(y, kernel, mode="nearest")
    fy = correlate(y, kernel.T, mode="nearest")
    return np.sqrt(fx * fx + fy * fy)


def flip_map(reference_srgb, test_srgb, ppd=67.0):
    ref = _linrgb_to_ycxcz(_srgb_to_linear(np.asarray(reference_srgb, np.float64)))
    tst = _linrgb_to_ycxcz(_srgb_to_linear(np.asarray(test_srgb, np.float64)))
    # colour pipeline
    qc, pc, pt = 0.7, 0.4, 0.95
    lab_r = _hunt(_linrgb_to_lab(_spatial_filter(ref, ppd)))
    lab_t = _hunt(_linrgb_to_lab(_spatial_filter(tst, ppd)))
    green = _hunt(_linrgb_to_lab(np.array([[[0.0, 1.0, 0.0]]])))
    blue = _hunt(_linrgb_to_lab(np.array([[[0.0, 0.0, 1.0]]])))
    cmax = float(_hyab(green, blue)[0, 0]) ** qc
    e = _hyab(lab_r, lab_t) ** qc
    limit = pc * cmax
    colour = np.where(e < limit, pt / limit * e, pt + (e - limit) / (cmax - limit) * (1.0 - pt))
    # feature pipeline
    yr, yt = (ref[..., 0] + 16.0) / 116.0, (tst[..., 0] + 16.0) / 116.0
    edge, point = _feature_kernels(ppd)
    feature = np.maximum(np.abs(_feature_norm(yr, edge) - _feature_norm(yt, edge)), np.abs(_feature_norm(yr, point) - _feature_norm(yt, point)))
    feature = np.clip(feature / np.sqrt(2.0), 0.0, 1.0) ** 0.5
    return np.clip(colour, 0.0, 1.0) ** (1.0 - feature)


def mean_flip(reference_srgb, test_srgb, ppd=67.0):
    return float(flip_map(reference_srgb, test_srgb, ppd).mean())
