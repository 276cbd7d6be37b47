"""Image comparison metrics shared by the parity tests (SURVEY.md §8c tolerances)."""
import numpy as np


def rgbe_roundtrip(img):
    """Quantise like stbi_write_hdr (truncating 8-bit mantissa, shared exponent) and decode like stbi_loadf.
    The golden images went through exactly this, which biases them by about -0.3 %."""
    rgb = np.asarray(img[..., :3], np.float32)
    m = rgb.max(axis=-1)
    mant, e = np.frexp(m)
    scale = np.where(m < 1e-32, 0.0, mant * 256.0 / np.maximum(m, 1e-38)).astype(np.float32)
    q = np.floor(rgb * scale[..., None]).clip(0, 255)
    out = q * np.ldexp(1.0, e - 8)[..., None].astype(np.float32)
    out[m < 1e-32] = 0
    return out.astype(np.float32)


def mse(a, b):
    return float(np.mean((np.asarray(a[..., :3], np.float64) - np.asarray(b[..., :3], np.float64)) ** 2))


def luminance(img):
    return 0.2126 * img[..., 0] + 0.7152 * img[..., 1] + 0.0722 * img[..., 2]


def mean_lum_ratio(a, b):
    return float(luminance(a).mean() / max(luminance(b).mean(), 1e-12))


def box3(img):
    p = np.pad(img, ((1, 1), (1, 1), (0, 0)), mode="edge")
    h, w = img.shape[:2]
    acc = np.zeros_like(img, dtype=np.float64)
    for dy in range(3):
        for dx in range(3):
            acc += p[dy:dy + h, dx:dx + w]
    return acc / 9.0


def p99_rel_err(a, b, floor=0.02):
    """99th percentile of the per-pixel relative luminance error after a 3x3 box filter."""
    la, lb = luminance(box3(a[..., :3])), luminance(box3(b[..., :3]))
    rel = np.abs(la - lb) / np.maximum(lb, floor)
    return float(np.percentile(rel, 99))


def firefly_mask(ref, factor=3.0, floor=0.25):
    """Pixels of a (golden) image that are isolated spikes: luminance above `factor` times the median of their 3x3
    neighbourhood and more than `floor` above it.  A converged render has none; a 1024-2048 spp golden with a point light inside
    a scattering medium has a few hundred (paths that scatter next to the light: the 1 / d^2 of lightSampling.glsl:20-31)."""
    lum = luminance(np.asarray(ref[..., :3], np.float64))
    p = np.pad(lum, 1, mode="edge")
    h, w = lum.shape
    stack = np.stack([p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)])
    med = np.median(stack, axis=0)
    return (lum > factor * med) & (lum - med > floor)


def mse_masked(a, b, mask):
    """MSE over the pixels NOT in mask"""
    d = (np.asarray(a[..., :3], np.float64) - np.asarray(b[..., :3], np.float64)) ** 2
    keep = ~mask
    return float(d[keep].mean()) if keep.any() else 0.0
