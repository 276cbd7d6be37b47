"""The reference's optional PMJ02BN sampler: include/rng/rng_pmj.glsl:20-107 with the reference's own tables
(math/PMJSequences.cpp, math/BlueNoise.cpp -> assets/tables, tools/extract_sampler_tables.py).

* the committed tables equal a fresh extraction from /root/reference (when that exists) and their recorded checksums;
* a pure-Python restatement of the GLSL written here, independently of the oracle's C++, pins the oracle integer-exactly;
* net properties of what a pixel receives; * (GPU) the product equals the oracle bit for bit, and renders with it."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import loader as oracle_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLES = os.path.join(ROOT, "assets", "tables")
M32, M64 = 0xffffffff, 0xffffffffffffffff
PMJ_SEED = 2873468793
ONEMINUSEPSILON = np.float32(0.999999)


# ---------------------------------------------------------------- pure-Python restatement of rng_pmj.glsl (small cases only)
def mix_bits(v):  # rng_pmj.glsl:30-37
    v ^= v >> 31
    v = (v * 9202493588570546565) & M64
    v ^= v >> 27
    v = (v * 9357036318526133325) & M64
    v ^= v >> 33
    return v


def permutation_element(i, l, p):  # rng_pmj.glsl:39-69
    w = l - 1
    for sh in (1, 2, 4, 8, 16):
        w |= w >> sh
    while True:
        i ^= p
        i = (i * 0xe170893d) & M32
        i ^= p >> 16
        i ^= (i & w) >> 4
        i ^= p >> 8
        i = (i * 0x0929eb3f) & M32
        i ^= p >> 23
        i ^= (i & w) >> 1
        i = (i * (1 | p >> 27)) & M32
        i = (i * 0x6935fa69) & M32
        i ^= (i & w) >> 11
        i = (i * 0x74dcb303) & M32
        i ^= (i & w) >> 2
        i = (i * 0x9e501cc3) & M32
        i ^= (i & w) >> 2
        i = (i * 0xc860a3df) & M32
        i &= w
        i ^= i >> 5
        if i < l:
            break
    return ((i + p) & M32) % l


def pmj_hash(px, py, dim):  # rng_pmj.glsl:73-74
    return mix_bits(((px << 48) ^ (py << 32) ^ (dim << 16) ^ PMJ_SEED) & M64) & M32


def py_rand2d(tables, px, py, dim, sample, spp):  # rng_pmj.glsl:85-107
    pmj, _ = tables
    idx, inst = sample, dim // 2
    if inst >= 16:
        idx = permutation_element(sample, spp, pmj_hash(px, py, dim))
    u = pmj[inst % 16, idx % 16384]
    return np.minimum(u, ONEMINUSEPSILON)


def py_rand1d(tables, px, py, dim, sample, spp):  # rng_pmj.glsl:71-83 + bluenoise.glsl:1-8
    _, blue = tables
    idx = permutation_element(sample, spp, pmj_hash(px, py, dim))
    delta = blue[dim % 48, px % 128, py % 128]
    return min((np.float32(idx) + delta) / np.float32(spp), ONEMINUSEPSILON)


@pytest.fixture(scope="module")
def tables(capi):
    return capi.load_sampler_tables()


@pytest.fixture(scope="module")
def oracle_ctx(capi, tables):
    ctx = capi.Context(oracle_loader.load_oracle())
    ctx.set_sampler_tables(tables)
    yield ctx
    ctx.close()


def test_tables_are_the_references(capi, tables):
    meta = json.load(open(os.path.join(TABLES, "tables.json")))
    pmj, blue = tables
    assert pmj.shape == (16, 16384, 2) and blue.shape == (48, 128, 128)
    assert hashlib.sha256(pmj.astype("<f4").tobytes()).hexdigest() == meta["pmj02bn"]["sha256_f32"]
    assert hashlib.sha256(blue.astype("<f4").tobytes()).hexdigest() == meta["bluenoise"]["sha256_f32"]
    # first literals of the reference's arrays (math/PMJSequences.cpp:13-14, math/BlueNoise.cpp:12)
    assert pmj[0, 0, 0] == np.float32(0.1804527938365936) and pmj[0, 0, 1] == np.float32(0.7133938074111938)
    assert pmj[0, 1, 0] == np.float32(0.6353681683540344)
    assert blue[0, 0, 0] == np.float32(0.500244140625) and blue[0, 0, 2] == np.float32(0.8553466796875)
    assert 0 <= pmj.min() and pmj.max() < 1 and 0 <= blue.min() and blue.max() < 1
    ref_math = "/root/reference/src/lib/vengine/math/PMJSequences.cpp"
    if os.path.exists(ref_math):  # this container only: a fresh extraction reproduces the committed files bit for bit
        before = {f: hashlib.sha256(open(os.path.join(TABLES, f), "rb").read()).hexdigest() for f in ("pmj02bn.f32", "bluenoise.u16")}
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extract_sampler_tables.py")], check=True, stdout=subprocess.DEVNULL)
        after = {f: hashlib.sha256(open(os.path.join(TABLES, f), "rb").read()).hexdigest() for f in before}
        assert before == after


def test_permutation_element_is_a_permutation():
    for l, p in ((16, 12345), (64, 0xdeadbeef), (100, 7), (1024, 0x9e3779b9)):
        assert sorted(permutation_element(i, l, p) for i in range(l)) == list(range(l))


def test_oracle_equals_python_restatement(capi, tables, oracle_ctx):
    """integer arithmetic and table lookups: exact equality"""
    for (px, py, w, spp, dim) in [(0, 0, 256, 64, 0), (17, 0, 256, 16, 2), (5, 9, 64, 32, 0), (1919, 1079, 1920, 8, 5), (130, 200, 256, 100, 31)]:
        start = py * w + py  # raygen.rgen.glsl:59
        got = oracle_ctx.sampler_points(px, py, w, 0, spp, dim, capi.PTC_FLAG_SAMPLER_PMJ)
        want = np.array([py_rand2d(tables, px, py, start + dim, i, spp) for i in range(spp)], np.float32)
        assert np.array_equal(got, want), (px, py, dim)
        got1 = oracle_ctx.sampler_points(px, py, w, 0, spp, dim, capi.PTC_FLAG_SAMPLER_PMJ | capi.PTC_SAMPLER_HOOK_1D)
        want1 = np.array([[py_rand1d(tables, px, py, start + dim, i, spp), py_rand1d(tables, px, py, start + dim + 1, i, spp)] for i in range(spp)], np.float32)
        assert np.array_equal(got1, want1), (px, py, dim)


def is_02_net(points):
    """every elementary interval of area 1/N in base 2 holds exactly one of the N = 2^m points"""
    n = len(points)
    m = n.bit_length() - 1
    assert 1 << m == n
    for a in range(m + 1):
        nx, ny = 1 << a, 1 << (m - a)
        cells = (np.floor(points[:, 0] * nx).astype(int) * ny + np.floor(points[:, 1] * ny).astype(int))
        if len(np.unique(cells)) != n:
            return False
    return True


def test_pixels_receive_02_nets(capi, tables, oracle_ctx):
    """row 0 starts at dimension 0 -> sequence 0 in table order: every power-of-two prefix is a (0, m, 2)-net (progressive multi-jittered
    (0,2) sequence); other rows start at dimension y * width + y >= 32, where the sample index is permuted first: the samplesPerPixel
    points of the pixel are then a permutation of a table prefix, i.e. the same net"""
    for n in (16, 64, 256, 1024):
        pts = oracle_ctx.sampler_points(3, 0, 256, 0, n, 0, capi.PTC_FLAG_SAMPLER_PMJ)
        assert np.array_equal(pts, np.minimum(tables[0][0, :n], ONEMINUSEPSILON)) and is_02_net(pts)
    for (px, py, w, n) in ((5, 9, 64, 256), (100, 37, 256, 64)):
        pts = oracle_ctx.sampler_points(px, py, w, 0, n, 0, capi.PTC_FLAG_SAMPLER_PMJ)
        seq = ((py * w + py) // 2) % 16
        assert is_02_net(pts)
        assert sorted(map(tuple, pts)) == sorted(map(tuple, np.minimum(tables[0][seq, :n], ONEMINUSEPSILON)))
    # rand1D: one sample per stratum [k / spp, (k + 1) / spp), jittered by the pixel's blue-noise value
    v = oracle_ctx.sampler_points(40, 11, 256, 0, 128, 4, capi.PTC_FLAG_SAMPLER_PMJ | capi.PTC_SAMPLER_HOOK_1D)[:, 0]
    assert sorted(np.floor(v * 128).astype(int)) == list(range(128))
    assert np.allclose((v * 128) % 1.0, tables[1][(11 * 256 + 11 + 4) % 48, 40 % 128, 11 % 128], atol=1e-4)


def test_pmj_needs_tables(capi):
    ctx = capi.Context(oracle_loader.load_oracle())
    with pytest.raises(RuntimeError, match="ptc_set_sampler_tables"):
        ctx.sampler_points(0, 0, 16, 0, 4, 0, capi.PTC_FLAG_SAMPLER_PMJ)
    ctx.close()


def test_pmj_lowers_the_error_of_a_small_render(capi, tables):
    """the point of the sampler: at 16 spp the stratified points beat the default random stream against a converged image"""
    eng = capi.HostEngine()
    eng.build_scene("FurnaceLambert")
    eng.set_render_info(width=48, height=48, samples=1024, batch_size=64)
    ctx = oracle_loader.oracle_context(eng)
    ctx.set_sampler_tables(tables)
    ref = ctx.render(eng.render_params(), want_aovs=False)
    eng.set_render_info(samples=16, batch_size=16)
    rp = eng.render_params()
    plain = ctx.render(rp, want_aovs=False)
    rp.flags |= capi.PTC_FLAG_SAMPLER_PMJ
    pmj = ctx.render(rp, want_aovs=False)
    e_plain, e_pmj = float(np.mean((plain - ref) ** 2)), float(np.mean((pmj - ref) ** 2))
    assert abs(pmj[..., :3].mean() / ref[..., :3].mean() - 1) < 0.01  # same expectation
    assert e_pmj < 0.8 * e_plain, (e_pmj, e_plain)
    ctx.close()
    eng.close()


# ---------------------------------------------------------------- the product
@pytest.mark.gpu
def test_cuda_points_equal_oracle_points(capi, tables, oracle_ctx):
    cu = capi.Context(capi.load_cuda())
    with pytest.raises(RuntimeError, match="ptc_set_sampler_tables"):
        cu.sampler_points(0, 0, 16, 0, 4, 0, capi.PTC_FLAG_SAMPLER_PMJ)
    cu.set_sampler_tables(tables)
    for hook in (0, capi.PTC_SAMPLER_HOOK_1D):
        for (px, py, w, first, count, dim) in [(0, 0, 256, 0, 1024, 0), (1919, 1079, 1920, 0, 96, 7), (5, 9, 64, 0, 333, 40), (77, 3, 128, 0, 4096, 1)]:
            a = cu.sampler_points(px, py, w, first, count, dim, capi.PTC_FLAG_SAMPLER_PMJ | hook)
            b = oracle_ctx.sampler_points(px, py, w, first, count, dim, capi.PTC_FLAG_SAMPLER_PMJ | hook)
            assert np.array_equal(a, b), (hook, px, py, dim)
    cu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["Cornell", "Volume5", "DepthOfField", "EnvironmentMapPBR01"])
def test_render_with_pmj_matches_oracle(capi, tables, scene):
    eng = capi.HostEngine()
    assert eng.backend_ok()
    eng.build_scene(scene)
    eng.set_render_info(width=96, height=96, samples=16, batch_size=8)
    desc, rp = eng.scene_desc(), eng.render_params()
    rp.flags |= capi.PTC_FLAG_SAMPLER_PMJ
    out = {}
    for label, lib in (("cuda", capi.load_cuda()), ("oracle", oracle_loader.load_oracle())):
        ctx = capi.Context(lib)
        ctx.set_sampler_tables(tables)
        ctx.upload_scene(desc)
        ctx.build_accel()
        out[label] = (ctx.render(rp), ctx.stats())
        ctx.close()
    (ra, aa, na), sa = out["cuda"]
    (rb, ab, nb), sb = out["oracle"]
    d = np.abs(ra[..., :3] - rb[..., :3]).max(axis=-1)
    assert np.mean(d > 1e-3 * np.maximum(1.0, rb[..., :3].max(axis=-1))) < 0.01
    assert abs(ra[..., :3].mean() / rb[..., :3].mean() - 1) < 2e-3
    assert abs(sa["segments"] - sb["segments"]) <= 1e-3 * sb["segments"]
    # and through the plugin: RenderInfo selects the sampler, the engine loads assets/tables itself
    eng.set_render_options(sampler="pmj")
    img = eng.render_to_memory()[0]
    assert np.allclose(img, ra, atol=1e-6)
    eng.close()
