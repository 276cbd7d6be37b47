"""Parity tests proper: the CUDA product against the CPU oracle and against the reference's golden images.
All of them call through the C-ABI (include/ptc.h) or through the C++ RendererPathTracing plugin.  Need a B200."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from imgmetrics import firefly_mask, mean_lum_ratio, mse, mse_masked, p99_rel_err, rgbe_roundtrip
from test_oracle_units import HIERARCHIES, check_lbvh, check_wide_bvh, look_down_params, make_quad_scene
from oracle import loader as oracle_loader

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "tests", "golden", "reference_images")

GOLDEN_SCENES = ["FurnacePBR", "FurnaceLambert", "EnvironmentMap", "EnvironmentMapPBR00", "EnvironmentMapPBR01", "EnvironmentMapPBR10",
                 "EnvironmentMapPBR11", "EnvironmentMapLambert", "Volume0", "Volume1", "Volume2", "Volume3", "Volume4", "Volume5", "Volume6",
                 "Volume7", "Volume8", "Volume9", "PointLight", "DirectionalLight", "MeshLight", "Transparency", "NormalMap", "GLTF", "Hierarchy",
                 "DepthOfField", "SharedComponents"]


@pytest.fixture(scope="module")
def engine(capi):
    eng = capi.HostEngine()  # default backend: the CUDA library
    assert eng.backend_ok(), eng.last_error()
    yield eng
    eng.close()


def both(capi, desc, hierarchy=None):
    out = {}
    for label, lib in (("cuda", capi.load_cuda()), ("oracle", oracle_loader.load_oracle())):
        ctx = capi.Context(lib)
        ctx.upload_scene(desc)
        ctx.build_accel(hierarchy)
        out[label] = ctx
    return out["cuda"], out["oracle"]


# ---------------------------------------------------------------- (1) LBVH build: bit-exact against the CPU reference build
@pytest.mark.parametrize("scene,kw", [("Volume5", {}), ("Cornell", {}), ("MeshLight", {}), ("SharedComponents", {}), ("Hierarchy", {}),
                                      ("Atrium", dict(texture_size=4, scale=0.3)), ("Atrium", dict(texture_size=4, scale=1.0))])
@pytest.mark.parametrize("hierarchy", HIERARCHIES)
def test_lbvh_bit_exact(capi, engine, scene, kw, hierarchy):
    """Morton codes, sorted order, hierarchy (Karras radix tree or PLOC) and node boxes equal the CPU reference build."""
    engine.build_scene(scene, **kw)
    cu, orc = both(capi, engine.scene_desc(), hierarchy)
    a, b = cu.get_lbvh(), orc.get_lbvh()
    assert a["n"] == b["n"] > 0
    for k in ("morton", "order", "parent", "left", "right", "aabb"):
        assert np.array_equal(a[k], b[k]), "LBVH field %s differs (%d entries)" % (k, int(np.sum(a[k] != b[k])))
    if a["n"] < 50000:
        check_lbvh(a)
    cu.close()
    orc.close()


@pytest.mark.parametrize("scene,kw", [("Volume5", {}), ("Cornell", {}), ("MeshLight", {}), ("SharedComponents", {}), ("Hierarchy", {}),
                                      ("Atrium", dict(texture_size=4, scale=0.3)), ("Atrium", dict(texture_size=4, scale=1.0))])
@pytest.mark.parametrize("hierarchy", HIERARCHIES)
def test_wide_bvh_bit_exact(capi, engine, scene, kw, hierarchy):
    """The on-device collapse of the LBVH into the 8-wide compressed BVH equals the CPU reference collapse byte for byte:
    same children per node, same octant slots, same exponents and quantised boxes, same breadth-first numbering, same
    triangle order."""
    engine.build_scene(scene, **kw)
    cu, orc = both(capi, engine.scene_desc(), hierarchy)
    a, b = cu.get_wide_bvh(), orc.get_wide_bvh()
    assert a["n_tris"] == b["n_tris"] > 0 and a["n_nodes"] == b["n_nodes"] > 0
    assert np.array_equal(a["tri_order"], b["tri_order"])
    diff = np.any(a["words"] != b["words"], axis=1)
    assert not diff.any(), "wide nodes differ: %d of %d, first %d" % (int(diff.sum()), len(diff), int(np.argmax(diff)))
    if a["n_tris"] < 50000:
        check_wide_bvh(a, cu.get_lbvh())
    assert cu.stats()["n_bvh_nodes"] == a["n_nodes"]
    cu.close()
    orc.close()


@pytest.mark.parametrize("bits", [10, 16, 21])
def test_every_morton_tier_bit_exact(capi, engine, bits):
    """30-, 48- and 63-bit Morton codes (4, 6 and 8 passes of the hand-written radix sort): the tier is a function of the triangle
    count, PTC_MORTON_BITS forces it on both sides so that a scene of modest size reaches the 63-bit path"""
    engine.build_scene("Atrium", texture_size=4, scale=0.2)
    os.environ["PTC_MORTON_BITS"] = str(bits)
    try:
        cu, orc = both(capi, engine.scene_desc())
        a, b = cu.get_lbvh(), orc.get_lbvh()
        wa, wb = cu.get_wide_bvh(), orc.get_wide_bvh()
    finally:
        del os.environ["PTC_MORTON_BITS"]
    assert a["n"] == b["n"] > 20000
    assert int(a["morton"].max()).bit_length() in range(3 * bits - 5, 3 * bits + 1)
    for k in ("morton", "order", "parent", "left", "right", "aabb"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(wa["words"], wb["words"]) and np.array_equal(wa["tri_order"], wb["tri_order"])
    cu.close()
    orc.close()


def test_lbvh_single_triangle_and_empty(capi):
    d, keep = make_quad_scene(capi)
    keep[2][0].tri_count = 1
    keep[3][0].num_triangles = 1
    cu, orc = both(capi, C.byref(d))
    a, b = cu.get_lbvh(), orc.get_lbvh()
    assert a["n"] == b["n"] == 1 and np.array_equal(a["aabb"], b["aabb"]) and np.array_equal(a["morton"], b["morton"])
    wa, wb = cu.get_wide_bvh(), orc.get_wide_bvh()
    assert wa["n_nodes"] == wb["n_nodes"] == 1 and np.array_equal(wa["words"], wb["words"]) and np.array_equal(wa["tri_order"], wb["tri_order"])
    rays = np.array([[-0.5, 2, 0.5, 1e-3, 0, -1, 0, 1e4], [0.9, 2, -0.9, 1e-3, 0, -1, 0, 1e4]], np.float32)
    ra, rb = cu.trace_closest(rays), orc.trace_closest(rays)
    assert np.array_equal(ra[0], rb[0]) and np.allclose(ra[2], rb[2])
    cu.close()
    orc.close()
    e = capi.ptc_scene_desc()
    cu, orc = both(capi, C.byref(e))
    assert cu.get_lbvh()["n"] == 0 and cu.get_wide_bvh()["n_nodes"] == 0
    ra = cu.trace_closest(rays)
    assert list(ra[0]) == [-1, -1]
    ia, _, _ = cu.render(look_down_params(capi, bg=(0.25, 0.5, 0.75)))
    ib, _, _ = orc.render(look_down_params(capi, bg=(0.25, 0.5, 0.75)))
    assert np.allclose(ia, ib, atol=1e-6) and np.allclose(ia[..., :3], (0.25, 0.5, 0.75), atol=1e-6)
    cu.close()
    orc.close()


# ---------------------------------------------------------------- (2) ray sets: ids identical, |dt| <= 1e-4 * max(1, t)
def ray_set(rng, n, lo, hi, aim=None):
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    if aim is None:
        d = rng.normal(size=(n, 3)).astype(np.float32)
    else:
        d = (rng.uniform(aim[0], aim[1], (n, 3)) - o).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, np.full((n, 1), 1e-3, np.float32), d, np.full((n, 1), 1e4, np.float32)], axis=1)


@pytest.mark.parametrize("scene,kw,box,coplanar", [("EnvironmentMap", {}, 4.0, False), ("Hierarchy", {}, 12.0, False),
                                                   ("SharedComponents", {}, 60.0, False), ("Cornell", {}, 1.0, True),
                                                   ("Atrium", dict(texture_size=4, scale=0.25), 8.0, False)])
@pytest.mark.parametrize("hierarchy", HIERARCHIES)
def test_ray_set_parity(capi, engine, scene, kw, box, coplanar, hierarchy):
    engine.build_scene(scene, **kw)
    cu, orc = both(capi, engine.scene_desc(), hierarchy)
    rng = np.random.default_rng(11)
    rays = np.concatenate([ray_set(rng, 30000, -box, box), ray_set(rng, 30000, -box, box, aim=(-box / 4, box / 4))])
    ia, pa, ta, ua, va = cu.trace_closest(rays)
    ib, pb, tb, ub, vb = orc.trace_closest(rays)
    assert (ib >= 0).mean() > 0.2
    same = (ia == ib) & (pa == pb)
    # a differing id is only acceptable on a shared edge / coplanar overlap where both report the same t
    # (the Cornell boxes stand ON the floor: their bottom faces are coplanar with it and every ray through them is a tie)
    bad = ~same
    assert bad.mean() <= (0.03 if coplanar else 2e-4), "id mismatches: %d" % int(bad.sum())
    if bad.any():
        assert np.all(np.abs(ta[bad] - tb[bad]) <= 1e-4 * np.maximum(1.0, tb[bad]))
    hit = same & (ib >= 0)
    assert np.all(np.abs(ta[hit] - tb[hit]) <= 1e-4 * np.maximum(1.0, tb[hit]))
    # barycentrics: 1e-4 for 99.9 % of the hits; rays grazing a sliver triangle (tiny determinant) amplify the FMA-contraction
    # difference between nvcc and gcc, so the worst case is bounded at 1e-2 instead
    db = np.maximum(np.abs(ua[hit] - ub[hit]), np.abs(va[hit] - vb[hit]))
    assert np.percentile(db, 99.9) <= 1e-4 and db.max() <= 1e-2, (np.percentile(db, 99.9), db.max())
    cu.close()
    orc.close()


# ---------------------------------------------------------------- (3) BSDF: 1e-5 relative on 1e5 random configurations
def test_bsdf_parity(capi):
    cu, orc = capi.Context(capi.load_cuda()), capi.Context(oracle_loader.load_oracle())
    rng = np.random.default_rng(5)
    n = 100000
    def dirs():
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d[:, 1] = np.abs(d[:, 1])
        return d
    wi, wo = dirs(), dirs()
    wi[: n // 20, 1] *= -1  # some below the horizon
    params = np.stack([rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0.035, 1, n)],
                      axis=1).astype(np.float32)
    fa, pa = cu.bsdf_eval(params, wi, wo)
    fb, pb = orc.bsdf_eval(params, wi, wo)
    tol = 1e-5
    assert np.all(np.abs(fa - fb) <= tol * np.maximum(np.abs(fb), 1e-3))
    assert np.all(np.abs(pa - pb) <= tol * np.maximum(np.abs(pb), 1e-3))
    u = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    wa, fa, pa = cu.bsdf_sample(params, wo, u)
    wb, fb, pb = orc.bsdf_sample(params, wo, u)
    # sampling goes through sincos whose last bits differ between libdevice and glibc, and a glossy lobe amplifies that in
    # f and pdf: directions within 2e-4; f/pdf must equal the device's OWN eval at its direction (1e-5) and the oracle's to 1e-3
    # for 99.9 % of the samples
    ok = pb >= 1e-6
    assert np.all(np.abs(wa - wb)[ok] <= 2e-4)
    fe, pe = cu.bsdf_eval(params, wa, wo)
    oka = pa >= 1e-6
    assert np.all(np.abs(pa - pe)[oka] <= tol * np.maximum(np.abs(pe[oka]), 1e-3))
    assert np.all(np.abs(fa - fe)[oka] <= tol * np.maximum(np.abs(fe[oka]), 1e-3))
    relp = np.abs(pa - pb)[ok] / np.maximum(np.abs(pb[ok]), 1e-3)
    assert np.percentile(relp, 99) <= 1e-3 and np.percentile(relp, 99.9) <= 5e-2
    assert np.mean((pa < 1e-6) != (pb < 1e-6)) < 1e-4
    cu.close()
    orc.close()


# ---------------------------------------------------------------- (4) environment: equirect -> cubemap kernel + cubemap lookup
def test_env_lookup_parity(capi, engine):
    engine.build_scene("EnvironmentMapLambert")
    cu, orc = both(capi, engine.scene_desc())
    rng = np.random.default_rng(9)
    d = rng.normal(size=(50000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    a, b = cu.env_lookup(d), orc.env_lookup(d)
    # hardware bilinear weights have 8 fractional bits; the resampling itself went through the same equirect taps
    err = np.abs(a - b).max(axis=1) / np.maximum(b.max(axis=1), 0.05)
    assert np.percentile(err, 99) < 0.03 and np.median(err) < 2e-3
    assert abs(a.mean() / b.mean() - 1) < 1e-3
    cu.close()
    orc.close()


def test_srgb_table_is_the_hardware_s(capi):
    """sRGB texels are decoded by the texture unit with a fixed table (hardware-defined) that differs from the analytic
    curve by up to 5e-3 relative; the oracle carries a copy (oracle/srgb_table.h, read through this hook): it must be what the
    device does, bit for bit, and stay a plausible sRGB curve"""
    cu, orc = capi.Context(capi.load_cuda()), capi.Context(oracle_loader.load_oracle())
    a, b = cu.srgb_table(), orc.srgb_table()
    assert np.array_equal(a, b)
    c = np.arange(256) / 255.0
    analytic = np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    assert a[0] == 0.0 and a[255] == 1.0 and np.all(np.diff(a) > 0)
    assert np.max(np.abs(a - analytic) / np.maximum(analytic, 1e-3)) < 1e-2
    cu.close()
    orc.close()


# ---------------------------------------------------------------- (5) renders: CUDA vs oracle at matched samples (same RNG streams)
def test_sampler_points_bit_exact(capi):
    """Both samplers (default xorshift stream, shuffled Owen-scrambled Sobol) are integer arithmetic: identical on both sides."""
    cu, orc = capi.Context(capi.load_cuda()), capi.Context(oracle_loader.load_oracle())
    for flags in (0, capi.PTC_FLAG_SAMPLER_SOBOL):
        for (px, py, w, first, count, dim) in [(0, 0, 256, 0, 1024, 0), (1919, 1079, 1920, 4000, 96, 7), (5, 9, 64, 1 << 20, 33, 40)]:
            a = cu.sampler_points(px, py, w, first, count, dim, flags)
            b = orc.sampler_points(px, py, w, first, count, dim, flags)
            assert np.array_equal(a, b), (flags, px, py, dim)
    cu.close()
    orc.close()


SOBOL_SCENES = ["Cornell", "Volume5", "DepthOfField", "EnvironmentMapPBR01", "MeshLight"]


@pytest.mark.parametrize("scene,sobol", [(s, False) for s in GOLDEN_SCENES + ["Denoise", "Cornell"]] + [(s, True) for s in SOBOL_SCENES])
def test_render_matches_oracle(capi, engine, scene, sobol):
    engine.build_scene(scene)
    engine.set_render_info(width=128, height=128, samples=8, batch_size=4)
    desc, rp = engine.scene_desc(), engine.render_params()
    if sobol:
        rp.flags |= capi.PTC_FLAG_SAMPLER_SOBOL
    cu, orc = both(capi, desc)
    ra, aa, na = cu.render(rp)
    rb, ab, nb = orc.render(rp)
    sa, sb = cu.stats(), orc.stats()
    # both consume the same random streams, so only float rounding (and the rare path it flips) can differ
    d = np.abs(ra[..., :3] - rb[..., :3]).max(axis=-1)
    # magnified image textures: the texture unit's filter arithmetic is hardware defined; 2 % instead of 1 %
    limit = 0.02 if scene in ("NormalMap", "Transparency", "GLTF") else 0.01
    assert np.mean(d > 1e-3 * np.maximum(1.0, rb[..., :3].max(axis=-1))) < limit, "radiance differs in %.3f %% of pixels" % (100 * np.mean(d > 1e-3))
    assert abs(ra[..., :3].mean() / max(rb[..., :3].mean(), 1e-9) - 1) < 2e-3
    assert np.mean(np.abs(aa - ab).max(axis=-1) > 1e-3) < limit and np.mean(np.abs(na - nb).max(axis=-1) > 1e-3) < 0.01
    # filter-weight rounding only, never a different texel (silhouette pixels where rounding moves one sample to another surface aside)
    assert np.mean(np.abs(aa - ab).max(axis=-1) > 2e-2) < 2e-3
    assert np.all(ra[..., 3] == 1.0)
    assert abs(sa["segments"] - sb["segments"]) <= 1e-3 * sb["segments"]
    # the device does not trace probe rays that miss the world boxes of all emitters (they can only return black)
    assert sa["probe_rays"] <= sb["probe_rays"] + 1e-3 * max(sb["probe_rays"], 1000)
    cu.close()
    orc.close()


# ---------------------------------------------------------------- (6) renders: CUDA vs the reference's golden images
GOLDEN_LIMITS = {"mse": 2e-4, "lum": 0.01, "p99": 0.05}
GOLDEN_SPP_FACTOR = 8  # "render at >= 8x the golden's spp"
# low-variance scenes (delta or small lights) must do much better
TIGHT = {"PointLight": 1e-5, "DirectionalLight": 1e-5, "FurnaceLambert": 5e-6, "SharedComponents": 1e-5}
# Goldens that are themselves far from converged at their 2048 spp: a point light / an emissive plane INSIDE a scattering medium
# (1 / d^2 of lightSampling.glsl:20-31 next to the light).  profiles/r2_golden_fireflies.json: at 8x the samples 69 % (Volume4) and
# 89 % (Volume8) of the whole squared error sits in 4-5 isolated spike pixels OF THE GOLDEN, and what remains equals the noise a
# 2048-spp render of this estimator carries (our own 1x render against our own 8x render).  So these four are held to the SAME
# 2e-4 / 5 % limits after (a) dropping the golden's spike pixels (imgmetrics.firefly_mask) and (b) never asking for less than 1.5x
# that measured noise floor; the mean luminance must still agree within 1 % (it did not need loosening).
NOISY_GOLDEN = {"Volume4", "Volume5", "Volume8", "Volume9"}


@pytest.mark.parametrize("scene", GOLDEN_SCENES)
def test_render_matches_reference_golden(capi, engine, scene, tmp_path):
    """Same recipe as the reference's TEST_F (RenderTests.cpp), rendered through RendererPathTracing::render() at 8x the golden's
    samples, compared with assets/unittests/<scene>_ref.hdr after the same RGBE quantisation."""
    engine.build_scene(scene)
    ri = engine.render_info()
    engine.set_render_info(samples=GOLDEN_SPP_FACTOR * ri["samples"])
    out = str(tmp_path / (scene + "_test"))
    engine.render(out)  # writes <out>.hdr like the reference
    img = capi.read_hdr(out + ".hdr")
    ref = capi.read_hdr(os.path.join(REF_DIR, scene + "_ref.hdr"))
    # the RGBE writer turns a NaN into a number: the images handed back in memory must be finite too
    assert all(np.isfinite(x).all() for x in engine.render_to_memory()), "non-finite pixel"
    lum, p99 = mean_lum_ratio(img, ref), p99_rel_err(img, ref)
    assert abs(lum - 1) <= GOLDEN_LIMITS["lum"], "mean luminance ratio %.4f" % lum
    if scene in NOISY_GOLDEN:
        engine.set_render_info(samples=ri["samples"])
        own = rgbe_roundtrip(engine.render_to_memory()[0])  # this estimator at the golden's own sample count
        floor_mse, floor_p99 = mse(own, img), p99_rel_err(own, img)
        spikes = firefly_mask(ref)
        assert spikes.sum() <= 16, "the golden's spikes are a handful of pixels, not a region (%d)" % int(spikes.sum())
        m = mse_masked(img, ref, spikes)
        assert m <= max(GOLDEN_LIMITS["mse"], 1.5 * floor_mse), "MSE %.3e (noise floor of a %d-spp render: %.3e)" % (m, ri["samples"], floor_mse)
        assert p99 <= max(GOLDEN_LIMITS["p99"], 1.5 * floor_p99), "p99 rel err %.4f (noise floor %.4f)" % (p99, floor_p99)
        return
    m = mse(img, ref)
    limit = TIGHT.get(scene, GOLDEN_LIMITS["mse"])
    assert m <= limit, "MSE %.3e > %.1e" % (m, limit)
    assert p99 <= GOLDEN_LIMITS["p99"] * (2 if scene in ("Transparency", "NormalMap", "GLTF", "DepthOfField", "EnvironmentMap") else 1), "p99 rel err %.4f" % p99


def test_denoise_aovs_match_reference(capi, engine, tmp_path):
    """Denoise_ref_{radiance,albedo,normal}.hdr pin the AOV conventions (first-hit albedo, n*0.5+0.5, background albedo = env)."""
    engine.build_scene("Denoise")
    ri = engine.render_info()
    engine.set_render_info(samples=GOLDEN_SPP_FACTOR * ri["samples"])
    out = str(tmp_path / "Denoise_test")
    engine.render(out)
    for suffix, lim in (("_radiance", 2e-4), ("_albedo", 2e-4), ("_normal", 5e-5)):
        img = capi.read_hdr(out + suffix + ".hdr")
        ref = capi.read_hdr(os.path.join(REF_DIR, "Denoise_ref" + suffix + ".hdr"))
        assert mse(img, ref) <= lim, (suffix, mse(img, ref))
        assert abs(mean_lum_ratio(img, ref) - 1) <= 0.01


# ---------------------------------------------------------------- (7) edge cases and partitions
def test_samples_dropped_and_alpha(capi):
    d, keep = make_quad_scene(capi)
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(C.byref(d))
    cu.build_accel()
    rad, alb, nrm = cu.render(look_down_params(capi, spp=70, batch=16))
    assert cu.stats()["segments"] == 2 * 16 * 16 * 64  # trap T7
    assert np.allclose(rad[..., :3], 0.5, atol=2e-6) and np.all(rad[..., 3] == 1.0)
    assert np.allclose(alb[..., :3], 0.5, atol=1e-6)
    cu.close()


# ---------------------------------------------------------------- environment importance sampling (PTC_FLAG_ENV_IMPORTANCE)
def test_env_sampler_parity(capi, engine):
    """same tables (built on the host from the same input, bit-identical) and same inversion: directions and densities agree to
    float rounding of the device's sin / cos / atan2"""
    engine.build_scene("EnvironmentMapLambert")
    cu, orc = both(capi, engine.scene_desc())
    u = np.random.default_rng(5).uniform(size=(100000, 2)).astype(np.float32)
    da, pa = cu.env_sample(u)
    db, pb = orc.env_sample(u)
    assert np.abs(da - db).max() < 2e-6
    assert np.max(np.abs(pa / pb - 1)) < 1e-4
    d = np.random.default_rng(6).normal(size=(100000, 3)).astype(np.float32)
    qa, qb = cu.env_pdf(d), orc.env_pdf(d)
    # a direction within rounding of a bin edge may fall into the neighbouring bin on one side
    assert np.mean(np.abs(qa / qb - 1) > 1e-4) < 1e-3
    cu.close()
    orc.close()


@pytest.mark.parametrize("scene,env_type", [("EnvironmentMapLambert", 1), ("EnvironmentMapPBR00", 1), ("EnvironmentMapPBR11", 2), ("Volume0", None),
                                            ("NormalMap", None), ("MeshLight", None), ("DepthOfField", None)])
def test_env_importance_matches_oracle(capi, engine, scene, env_type):
    engine.build_scene(scene)
    engine.set_render_info(width=128, height=128, samples=8, batch_size=4)
    desc, rp = engine.scene_desc(), engine.render_params()
    rp.flags |= capi.PTC_FLAG_ENV_IMPORTANCE
    if env_type is not None:
        rp.scene.background[3] = float(env_type)
    cu, orc = both(capi, desc)
    ra, aa, na = cu.render(rp)
    rb, ab, nb = orc.render(rp)
    sa, sb = cu.stats(), orc.stats()
    d = np.abs(ra[..., :3] - rb[..., :3]).max(axis=-1)
    assert np.mean(d > 1e-3 * np.maximum(1.0, rb[..., :3].max(axis=-1))) < 0.02, "radiance differs in %.3f %% of pixels" % (100 * np.mean(d > 1e-3))
    assert abs(ra[..., :3].mean() / max(rb[..., :3].mean(), 1e-9) - 1) < 3e-3
    assert abs(sa["segments"] - sb["segments"]) <= 1e-3 * sb["segments"]
    # (shadow-ray counts are not comparable: the device drops requests whose BSDF value is black before tracing them)
    if rp.scene.background[3] != 0.0:
        rp.flags &= ~capi.PTC_FLAG_ENV_IMPORTANCE
        cu.render(rp, want_aovs=False)
        assert cu.stats()["shadow_rays"] < sa["shadow_rays"]  # the flag really adds the environment to the light pick
    cu.close()
    orc.close()


@pytest.mark.parametrize("env", [{"PTC_OVERLAP": "0", "PTC_MAX_SLOTS": str(96 * 96 * 3)}, {"PTC_OVERLAP": "5,2"}, {}, {"PTC_OVERLAP": "4,3", "PTC_MAX_SLOTS": str(96 * 96)}])
def test_chunked_and_overlapped_wavefronts_equal_the_plain_render(capi, engine, env):
    """a batch cut into chunks (what 4K frames need) and the two-stream overlap of wavefronts (PTC_OVERLAP, read when the context is
    created; {} = the default, two uncapped wavefronts) only change the order of float additions into the accumulators, compared with
    one wavefront at a time (PTC_OVERLAP=0)"""
    engine.build_scene("Cornell")
    engine.set_render_info(width=96, height=96, samples=24, batch_size=8)
    desc, rp = engine.scene_desc(), engine.render_params()

    def render():
        ctx = capi.Context(capi.load_cuda())
        ctx.upload_scene(desc)
        ctx.build_accel()
        out = ctx.render(rp)
        st = ctx.stats()
        ctx.close()
        return out, st

    def with_env(e):
        saved = {k: os.environ.get(k) for k in e}
        os.environ.update(e)
        try:
            return render()
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    plain, sp = with_env({"PTC_OVERLAP": "0"})
    other, so = with_env(env)
    assert so["segments"] == sp["segments"] and so["shadow_rays"] == sp["shadow_rays"] and so["probe_rays"] == sp["probe_rays"]
    assert so["kernel_launches"] > sp["kernel_launches"]  # really ran in more pieces
    for a, b in zip(plain, other):
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, float(np.abs(a).max()))


@pytest.mark.parametrize("scene,kw", [("Cornell", {}), ("MeshLight", {}), ("Progressive", {"scale": 0.05})])
def test_probe_folded_into_next_segment_equals_traced_probe(capi, engine, scene, kw):
    """opaque, media-free scenes: a surviving path's BSDF probe is evaluated on its next segment's closest hit instead of being traced
    (PTC_NO_PROBE_FUSION=1 traces it).  Same terms in the same order: images agree to rounding, far fewer probe rays."""
    engine.build_scene(scene, **kw)
    engine.set_render_info(width=128, height=96, samples=16, batch_size=8)
    desc, rp = engine.scene_desc(), engine.render_params()
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(desc)
    cu.build_accel()
    fused = cu.render(rp)[0]
    sf = cu.stats()
    os.environ["PTC_NO_PROBE_FUSION"] = "1"
    try:
        traced = cu.render(rp)[0]
        st = cu.stats()
    finally:
        del os.environ["PTC_NO_PROBE_FUSION"]
    cu.close()
    assert np.isfinite(fused).all() and np.isfinite(traced).all()
    assert st["probe_rays"] > 4 * max(sf["probe_rays"], 1) and sf["segments"] == st["segments"]
    d = np.abs(fused[..., :3] - traced[..., :3]).max(axis=-1)
    # a surface inside the first millimetre of the ray is seen by the traced probe only: a handful of pixels at most
    assert np.mean(d > 1e-4 * np.maximum(1.0, traced[..., :3].max(axis=-1))) < 2e-3
    assert abs(fused[..., :3].mean() / traced[..., :3].mean() - 1) < 1e-4


def test_shadow_rays_through_alpha_tested_foliage(capi, engine):
    """media-free scene with transparent materials and a light: the device looks at every triangle a shadow ray crosses in one
    traversal (order-independent product of 1 - alpha), the oracle walks the candidates nearest first (raySecondary.rahit.glsl)"""
    engine.build_scene("Instanced", texture_size=64, scale=0.004)
    engine.set_render_info(width=160, height=90, samples=8, batch_size=4)
    desc, rp = engine.scene_desc(), engine.render_params()
    cu, orc = both(capi, desc)
    ra = cu.render(rp)[0]
    rb = orc.render(rp)[0]
    sa, sb = cu.stats(), orc.stats()
    assert sb["shadow_rays"] > 10000 and sa["shadow_rays"] > 10000
    d = np.abs(ra[..., :3] - rb[..., :3]).max(axis=-1)
    assert np.mean(d > 1e-3 * np.maximum(1.0, rb[..., :3].max(axis=-1))) < 0.03, "radiance differs in %.3f %% of pixels" % (100 * np.mean(d > 1e-3))
    assert abs(ra[..., :3].mean() / rb[..., :3].mean() - 1) < 5e-3
    assert abs(sa["segments"] - sb["segments"]) <= 2e-3 * sb["segments"]
    cu.close()
    orc.close()


def test_render_progress_polled_from_another_thread(capi):
    """renderProgress() is read by the UI thread while render() runs on a worker (MainWindow.cpp:874-896): monotone, within [0, 1],
    1 when the call returns; with many short batches intermediate values are seen"""
    import threading
    eng = capi.HostEngine()
    eng.build_scene("Cornell")
    eng.set_render_info(width=256, height=256, samples=4 * 96, batch_size=4)
    assert eng.render_progress() in (0.0, 1.0)
    seen = []
    stop = threading.Event()

    def poll():
        while not stop.is_set():
            seen.append(eng.render_progress())

    t = threading.Thread(target=poll)
    t.start()
    try:
        eng.render_to_memory()  # ctypes releases the GIL during the call
    finally:
        stop.set()
        t.join()
    assert eng.render_progress() == 1.0
    vals = np.array(seen, np.float64)
    assert len(vals) > 10 and vals.min() >= 0.0 and vals.max() <= 1.0
    rising = vals[np.argmax(vals < 0.999):] if np.any(vals < 0.999) else vals  # from the first poll inside the render on
    assert np.all(np.diff(rising) >= -1e-7), "progress went backwards"
    assert np.any((vals > 0.02) & (vals < 0.98)), "no intermediate progress value observed"
    eng.close()


def test_texture_identity_cache(capi):
    """ptc_texture.uid: content with a non-zero uid is immutable by contract and keeps its device copy across uploads
    (the reference uploads textures once, at import); uid 0 or a new uid uploads again."""
    d, keep = make_quad_scene(capi)
    T, tex_data = keep[5], keep[6]
    albedo = tex_data[1]  # sRGB albedo texture of the quad's material (tex1[0] = 1)
    albedo[0], albedo[1], albedo[2] = 255, 255, 254  # not the all-white fast path
    for i in range(3):
        T[i].uid = 100 + i
    cu = capi.Context(capi.load_cuda())

    def albedo_aov():
        cu.upload_scene(C.byref(d))
        cu.build_accel()
        return cu.render(look_down_params(capi, spp=16, batch=16))[1][..., :3].mean(axis=(0, 1))

    a0 = albedo_aov()
    assert np.allclose(a0[:2], 0.5, atol=1e-3)
    albedo[0] = 0  # same uid: the caller broke the contract, the device copy is (legitimately) still the old one
    assert np.allclose(albedo_aov(), a0, atol=1e-7)
    T[1].uid = 200  # new identity: uploaded again
    a1 = albedo_aov()
    assert a1[0] < 1e-3 and np.allclose(a1[1], 0.5, atol=1e-3)
    albedo[0] = 255
    T[1].uid = 0  # anonymous content is always uploaded
    assert np.allclose(albedo_aov(), a0, atol=1e-7)
    albedo[1] = 0
    a2 = albedo_aov()
    assert a2[1] < 1e-3
    cu.close()


@pytest.mark.parametrize("mode,w,h,tile", [("tile", 160, 96, 32), ("sample", 160, 96, 32),
                                             ("tile", 150, 90, 32),   # ragged tiles at the right and bottom edges
                                             ("tile", 37, 29, 32),    # two tiles for three ranks: one rank owns no pixel
                                             ("tile", 33, 17, 0)])    # tile_size 0 = the default edge of 32
def test_partition_sums_to_full_render(capi, engine, mode, w, h, tile):
    import partition_util as parallel
    engine.build_scene("MeshLight")
    engine.set_render_info(width=w, height=h, samples=16, batch_size=4)
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(engine.scene_desc())
    cu.build_accel()
    full = np.stack(cu.render(engine.render_params()))
    seg_full = cu.stats()["segments"]
    world = 3
    total = np.zeros_like(full)
    seg = 0
    for r in range(world):
        rp = parallel.partition(engine.render_params(), r, world, mode, tile_size=tile)
        part = np.stack(cu.render(rp))
        total += part
        seg += cu.stats()["segments"]
    assert seg == seg_full
    assert np.allclose(total[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)
    assert np.all(total[..., 3] == 1.0) and np.all(full[..., 3] == 1.0)  # alpha is written by rank 0 alone: the sum keeps it
    cu.close()


def test_errors_are_reported_not_thrown(capi):
    cu = capi.Context(capi.load_cuda())
    with pytest.raises(RuntimeError, match="ptc_upload_scene|ptc_build_accel"):
        cu.build_accel()
    d, keep = make_quad_scene(capi)
    keep[3][0].material_index = 7
    with pytest.raises(RuntimeError, match="material index"):
        cu.upload_scene(C.byref(d))
    cu.close()


# ---------------------------------------------------------------- (8) BASELINE-size properties (1920x1080 atrium)
def test_full_size_determinism_and_linearity(capi, engine):
    """At the benchmark's size the oracle is too slow, so check size-independent properties: two runs are bit-identical,
    and radiance is linear in the environment intensity (same random streams -> exactly 2x up to rounding)."""
    engine.build_scene("Atrium", texture_size=64)
    engine.set_render_info(samples=4, batch_size=4)
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(engine.scene_desc())
    cu.build_accel()
    rp = engine.render_params()
    assert (rp.width, rp.height) == (1920, 1080)
    a = cu.render(rp, want_aovs=False)
    sa = cu.stats()
    b = cu.render(rp, want_aovs=False)
    assert np.array_equal(a, b) and sa["segments"] == cu.stats()["segments"]
    rp.scene.exposure[1] = 2.0
    c = cu.render(rp, want_aovs=False)
    assert np.allclose(c[..., :3], 2.0 * a[..., :3], rtol=1e-5, atol=1e-7)
    assert 3.0 < sa["segments"] / (1920 * 1080 * 4) < 5.0
    cu.close()
