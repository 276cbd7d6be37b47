"""Output stage of the path: storeToDisk's PNG branch - optional v * 2^exposure, clamp to [0, 1],
linear -> sRGB, uchar(255 x) truncation, RGBA - against a numpy restatement of
/root/reference/src/lib/vengine/core/ImageUtils.cpp:12-21, 34-76 and VulkanRendererPathTracing.cpp:958-975.
The PNG files are read back with an independent decoder (PIL), not with the host library's own."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_png_bytes(img, exposure):
    """numpy float32 restatement of applyExposure (ImageUtils.cpp:39-49: first min(channels, 3) channels times pow(2, exposure), only when
    exposure != 0: VulkanRendererPathTracing.cpp:972-974) + writeToDisk PNG (ImageUtils.cpp:58-65: ALL channels, alpha included, go
    through clamp, linearToSRGB (:12-21) and a truncating cast)."""
    x = np.array(img, np.float32, copy=True)
    if exposure != 0.0:
        x[..., :3] = x[..., :3] * np.float32(np.power(np.float32(2.0), np.float32(exposure)))
    x = np.clip(x, np.float32(0.0), np.float32(1.0))
    lo = x * np.float32(12.92)
    hi = np.float32(1.055) * np.power(x, np.float32(1.0 / 2.4), dtype=np.float32) - np.float32(0.055)
    s = np.where(x <= np.float32(0.0031308), lo, hi).astype(np.float32)
    return (np.float32(255.0) * s).astype(np.uint8)  # truncation, like static_cast<unsigned char>


def read_png(path):
    from PIL import Image
    im = Image.open(path)
    assert im.mode == "RGBA"
    return np.asarray(im)


def test_linear_to_srgb_known_answers():
    """hand-computed values of the reference's arithmetic, including its quirk: 1.055f - 0.055f rounds to 0.99999994 in
    single precision, so a saturated channel (and alpha = 1) is written as 254, not 255"""
    v = np.array([[[0.0, 0.0031308, 0.5, 1.0]]], np.float32)
    b = reference_png_bytes(v, 0.0)[0, 0]
    assert list(b) == [0, 10, 187, 254]
    assert reference_png_bytes(np.array([[[0.25, 2.0, -1.0, 1.0]]], np.float32), 1.0)[0, 0].tolist() == [187, 254, 0, 254]


@pytest.mark.parametrize("exposure", [0.0, 1.5, -2.0])
def test_png_writer_matches_reference_arithmetic(capi, tmp_path, exposure):
    rng = np.random.default_rng(3)
    h, w = 37, 53  # not multiples of anything: PNG rows / filter bytes must still line up
    img = rng.gamma(0.6, 0.7, (h, w, 4)).astype(np.float32)
    img[..., 3] = 1.0
    img[0, 0, :3] = (0.0, 0.0031308, 0.0031309)  # both sides of the sRGB knee
    img[0, 1, :3] = (-0.5, 1.0, 7.0)             # clamp below / at / above
    img[0, 2, :3] = (1e-9, 0.999999, np.float32(1.0) / np.float32(2.0 ** exposure) if exposure else 1.0)
    out = str(tmp_path / "img")
    capi.write_image(out, img, "png", exposure)
    got = read_png(out + ".png")
    want = reference_png_bytes(img, exposure)
    assert got.shape == want.shape == (h, w, 4)
    # pow() of numpy and of libm may differ in the last bit, which the truncation can turn into one level: never more, and rarely
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1 and np.mean(d > 0) < 2e-3, (int(d.max()), float(np.mean(d > 0)))
    assert np.all(got[..., 3] == 254)  # alpha = 1 goes through linearToSRGB too (ImageUtils.cpp:60-62)
    assert got[0, 1, 0] == 0 and got[0, 1, 2] == 254 and (exposure < 0 or got[0, 1, 1] == 254)
    assert got[0, 2, 2] == 254  # a value that exposure lifts exactly to 1
    # exposure leaves the file untouched when it is exactly 0 and is NOT applied to HDR output
    pos = np.clip(img, 0, None)  # (radiance is never negative; RGBE has no sign)
    capi.write_image(out, pos, "hdr", exposure)
    back = capi.read_hdr(out + ".hdr")
    from imgmetrics import rgbe_roundtrip
    assert np.allclose(back[..., :3], rgbe_roundtrip(pos), rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_render_writes_png_with_exposure(capi, tmp_path):
    """the whole branch through RendererPathTracing::render(): RenderInfo.fileType = PNG, exposure != 0, writeAllFiles"""
    eng = capi.HostEngine()
    assert eng.backend_ok(), eng.last_error()
    eng.build_scene("MeshLight")
    eng.set_render_info(width=96, height=64, samples=32, batch_size=8)
    rad, alb, nrm = (x.copy() for x in eng.render_to_memory())
    eng.set_output("png", exposure=1.25, write_all_files=True)
    out = str(tmp_path / "frame")
    eng.render(out)
    got = read_png(out + ".png")
    want = reference_png_bytes(rad, 1.25)  # the render is deterministic: same image as render_to_memory
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1 and np.mean(d > 0) < 2e-3
    assert got[..., :3].max() == 254 and 20 < got[..., :3].mean() < 235  # exposure pushed the lit floor into the clamp, not the whole image
    # AOV files stay HDR and unexposed (VulkanRendererPathTracing.cpp:966-969)
    from imgmetrics import rgbe_roundtrip
    assert np.allclose(capi.read_hdr(out + "_albedo.hdr")[..., :3], rgbe_roundtrip(alb), atol=1e-6)
    assert np.allclose(capi.read_hdr(out + "_normal.hdr")[..., :3], rgbe_roundtrip(nrm), atol=1e-6)
    assert not os.path.exists(out + ".hdr")
    eng.close()
