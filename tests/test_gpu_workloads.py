"""BASELINE.json's workload configurations as CUDA-vs-oracle IMAGE tests at reduced scale, the callers either
side of the path (scene.json import, the BallOnPlane frame sequence) rendered through the CUDA plugin, and the multi-GPU
product path (one context over several GPUs / one process per GPU, NCCL inside the core).  Need a B200 (two for the last group)."""
import json
import os
import sys

import numpy as np
import pytest

from imgmetrics import rgbe_roundtrip
from oracle import loader as oracle_loader

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def engine(capi):
    eng = capi.HostEngine()
    assert eng.backend_ok(), eng.last_error()
    yield eng
    eng.close()


def compare_with_oracle(capi, eng, rp=None, pixel_limit=0.01, mean_limit=2e-3, aov_limit=0.01):
    """the metrics of test_gpu_parity.test_render_matches_oracle: both sides consume the same random streams, so only float
    rounding (and the rare path it flips) can differ"""
    desc = eng.scene_desc()
    rp = rp or eng.render_params()
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(desc)
    cu.build_accel()
    ra, aa, na = cu.render(rp)
    sa = cu.stats()
    cu.close()
    orc = capi.Context(oracle_loader.load_oracle())
    orc.upload_scene(desc)
    orc.build_accel()
    rb, ab, nb = orc.render(rp)
    sb = orc.stats()
    orc.close()
    assert np.isfinite(ra).all() and np.isfinite(rb).all()
    d = np.abs(ra[..., :3] - rb[..., :3]).max(axis=-1)
    frac = float(np.mean(d > 1e-3 * np.maximum(1.0, rb[..., :3].max(axis=-1))))
    assert frac < pixel_limit, "radiance differs in %.3f %% of pixels" % (100 * frac)
    assert abs(ra[..., :3].mean() / max(rb[..., :3].mean(), 1e-9) - 1) < mean_limit
    assert np.mean(np.abs(aa - ab).max(axis=-1) > 1e-3) < aov_limit and np.mean(np.abs(na - nb).max(axis=-1) > 1e-3) < aov_limit
    assert np.all(ra[..., 3] == 1.0)
    assert abs(sa["segments"] - sb["segments"]) <= 2e-3 * sb["segments"]
    return ra, rb, sa, sb


# scene, recipe options, width, height, samples, batch | what the config exercises
WORKLOADS = [
    ("Atrium", dict(scale=0.1, texture_size=64), 128, 72, 8, 4),       # C2: 24 textured PBR classes (layered textures), HDR env, depth 9
    ("Fog", dict(scale=0.05, texture_size=32), 128, 72, 8, 4),         # C3: camera medium, point + mesh lights, depth 32, RR, DOF -> k_shade<1,1>, tmaxNext cut
    ("Progressive", dict(scale=0.05, camera=0), 128, 72, 16, 8),       # C5 perspective camera, many emitters (fused probes)
    ("Progressive", dict(scale=0.05, camera=1), 128, 72, 16, 8),       # C5 ORTHOGRAPHIC camera (true parallel rays, trap T10)
    ("Instanced", dict(scale=0.004, texture_size=64), 160, 90, 8, 4),  # C4: alpha-tested foliage + directional light
    ("Cornell", {}, 96, 96, 16, 8),                                    # C1
]
# Scenes whose paths bounce off glossy, finely textured surfaces many times.  Product and oracle draw the same random numbers, but a
# sampled direction already differs in its last bits (sincos of libdevice vs glibc, test_bsdf_parity: 2e-4), a glossy lobe amplifies
# that, and after two or three such bounces the two paths land on different texels / triangles: from then on they are two independent
# samples of the same estimator (profiles/r2_atrium_parity_diag.log: 0.05 % of the pixels differ at depth 2, 7.5 % at depth 4, 17 % at
# depth 9; the first-hit AOVs agree to 5e-5).  These configs are therefore held to the per-pixel limits at depth 2, where the paths
# still coincide, and to STATISTICAL limits at their full depth with 16 times the samples.
CHAOTIC = ("Atrium", "Fog")


@pytest.mark.parametrize("scene,kw,w,h,spp,batch", WORKLOADS, ids=["C2-Atrium", "C3-Fog", "C5-persp", "C5-ortho", "C4-Instanced", "C1-Cornell"])
def test_workload_configs_match_oracle(capi, engine, scene, kw, w, h, spp, batch):
    engine.build_scene(scene, **kw)
    full = engine.render_info()
    engine.set_render_info(width=w, height=h, samples=spp, batch_size=batch)  # depth stays the config's (9 / 32 / 6 / ...)
    rp = engine.render_params()
    if scene == "Fog":
        assert full["depth"] == 32 and rp.scene.volumes[0] >= 0 and rp.scene.exposure[2] > 0  # medium + depth of field are really on
    if kw.get("camera") == 1:
        assert rp.camera_type == capi.PTC_CAMERA_ORTHOGRAPHIC and rp.ortho_width > 0
    if scene in CHAOTIC:
        rp.depth = 2
        compare_with_oracle(capi, engine, rp)  # the strict limits of every other scene
        return
    loose = scene == "Instanced"  # alpha-tested silhouettes of magnified 64 x 64 leaf cards
    ra, rb, sa, sb = compare_with_oracle(capi, engine, rp, pixel_limit=0.03 if loose else 0.01, mean_limit=5e-3 if loose else 2e-3,
                                         aov_limit=0.03 if loose else 0.01)
    assert rb[..., :3].mean() > 1e-4


@pytest.mark.parametrize("scene,kw", [("Atrium", dict(scale=0.1, texture_size=64)), ("Fog", dict(scale=0.05, texture_size=32))], ids=["C2-Atrium", "C3-Fog"])
def test_deep_glossy_configs_match_oracle_statistically(capi, engine, scene, kw):
    """full depth (9 / 32 with roulette), 128 spp at 64 x 36: means, block means and segment counts of two converging estimates"""
    engine.build_scene(scene, **kw)
    engine.set_render_info(width=64, height=36, samples=128, batch_size=16)
    desc, rp = engine.scene_desc(), engine.render_params()
    res = {}
    for label, lib in (("cuda", capi.load_cuda()), ("oracle", oracle_loader.load_oracle())):
        ctx = capi.Context(lib)
        ctx.upload_scene(desc)
        ctx.build_accel()
        res[label] = (ctx.render(rp), ctx.stats())
        ctx.close()
    (ra, aa, na), sa = res["cuda"]
    (rb, ab, nb), sb = res["oracle"]
    assert np.isfinite(ra).all() and np.all(ra[..., 3] == 1.0)
    # first-hit AOVs do not depend on the later bounces: per pixel
    assert np.mean(np.abs(aa - ab).max(axis=-1) > 1e-3) < 0.01 and np.mean(np.abs(na - nb).max(axis=-1) > 1e-3) < 0.01
    assert abs(sa["segments"] / sb["segments"] - 1) < 2e-3       # path lengths (roulette, misses) have the same distribution
    assert abs(ra[..., :3].mean() / rb[..., :3].mean() - 1) < 0.01  # 295 k paths each: the means of two independent estimates agree to ~0.3 %
    blocks = lambda x: x[..., :3].reshape(9, 4, 16, 4, 3).mean(axis=(1, 3, 4))  # noqa: E731  4 x 4 pixel block luminance
    ba, bb = blocks(ra), blocks(rb)
    rel = np.abs(ba - bb) / np.maximum(bb, 0.05 * bb.mean())
    assert np.median(rel) < 0.05 and np.percentile(rel, 95) < 0.2, (float(np.median(rel)), float(np.percentile(rel, 95)))


@pytest.mark.parametrize("w,h", [(64, 40), (1056, 1000), (2048, 1031)])
def test_host_readback_equals_device_targets(capi, engine, w, h):
    """getRenderTargetData x3 (...PathTracing.cpp:890-893): the host images go through pinned staging chunks of 16 MB - one
    chunk, one chunk plus a short tail, several chunks with a ragged end - and must be the device targets byte for byte"""
    import torch
    engine.build_scene("Cornell")
    engine.set_render_info(width=w, height=h, samples=2, batch_size=1)
    desc, rp = engine.scene_desc(), engine.render_params()
    ctx = capi.Context(capi.load_cuda(), device=0)
    ctx.upload_scene(desc)
    ctx.build_accel()
    host = ctx.render(rp)
    dev = [torch.full((h, w, 4), -1.0, dtype=torch.float32, device="cuda:0") for _ in range(3)]
    ctx.render_device(rp, *[t.data_ptr() for t in dev])
    torch.cuda.synchronize()
    ctx.close()
    for a, b in zip(host, dev):
        assert np.array_equal(np.asarray(a).reshape(h, w, 4), b.cpu().numpy())
    assert np.all(np.asarray(host[0]).reshape(h, w, 4)[..., 3] == 1.0)


def test_orthographic_rays_are_parallel(capi, engine):
    """trap T10 (parity unpinned in the reference): the build renders a TRUE orthographic view.  Property test: over a flat floor
    seen head-on, the first-hit normal AOV is constant and the albedo AOV of a checker of emissive spheres does not depend on
    the distance to the camera - moving the orthographic camera back along its axis leaves the AOVs unchanged."""
    engine.build_scene("Progressive", scale=0.05, camera=1)
    engine.set_render_info(width=96, height=54, samples=4, batch_size=4)
    desc, rp = engine.scene_desc(), engine.render_params()
    assert rp.camera_type == capi.PTC_CAMERA_ORTHOGRAPHIC
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(desc)
    cu.build_accel()
    _, a0, n0 = cu.render(rp)
    # translate the camera 5 units backwards along its own viewing axis (view_inverse column 2 = camera +z in world space)
    vi = np.array(rp.scene.view_inverse, np.float32).reshape(4, 4).T.copy()
    vi[:3, 3] += 5.0 * vi[:3, 2]
    v = np.linalg.inv(vi)
    for k, val in enumerate(vi.T.reshape(-1)):
        rp.scene.view_inverse[k] = float(val)
    for k, val in enumerate(v.T.reshape(-1)):
        rp.scene.view[k] = float(val)
    _, a1, n1 = cu.render(rp)
    cu.close()
    # same pixels see the same surfaces (jitter is the same stream); a perspective camera would zoom out
    assert np.mean(np.abs(a0 - a1).max(axis=-1) > 1e-3) < 0.01 and np.mean(np.abs(n0 - n1).max(axis=-1) > 1e-3) < 0.01


def test_scene_json_renders_on_cuda(capi, engine, tmp_path):
    """SURVEY 8f rank 2 through the product: a scene written by Scene::exportScene, read back by importScene into a second engine
    and rendered by RendererPathTracing::render() on the CUDA core equals the render of the original recipe, and the oracle's"""
    engine.build_scene("MeshLight")
    engine.set_render_info(width=96, height=72, samples=16, batch_size=8)
    img_a = engine.render_to_memory()[0].copy()
    seg_a = engine.stats()["segments"]
    engine.export_scene(str(tmp_path))
    other = capi.HostEngine()
    assert other.backend_ok()
    other.import_scene(str(tmp_path / "scene.json"))
    other.set_render_info(width=96, height=72, samples=16, batch_size=8, depth=engine.render_info()["depth"])
    img_b = other.render_to_memory()[0].copy()
    assert other.stats()["segments"] > 0 and abs(other.stats()["segments"] - seg_a) <= 2e-3 * seg_a
    # transforms travel through the file as decimal text / Euler angles: same scene up to rounding
    d = np.abs(img_a[..., :3] - img_b[..., :3]).max(axis=-1)
    assert np.mean(d > 1e-3 * np.maximum(1.0, img_a[..., :3].max(axis=-1))) < 0.02
    assert abs(img_a[..., :3].mean() / img_b[..., :3].mean() - 1) < 2e-3
    compare_with_oracle(capi, other)
    other.close()


def test_ball_on_plane_frame_sequence(capi):
    """The reference's demo sequence (src/bin/offlinerender/PtSceneBallOnPlane.cpp:8-55): the camera orbits, every frame is one
    render() call.  Each frame must equal the oracle's render of that frame; across frames the environment cubemap and the textures
    stay resident on the device (ptc_env.uid / ptc_texture.uid) - only geometry and records are uploaded again."""
    eng = capi.HostEngine()
    assert eng.backend_ok()
    eng.build_scene("BallOnPlane")
    ri = eng.render_info()
    assert (ri["samples"], ri["batch_size"]) == (64, 64)  # PtSceneBallOnPlane.cpp:38-39
    eng.set_render_info(width=96, height=96, samples=16, batch_size=16)
    frames, uploads = [], []
    for f in range(3):
        eng.set_sequence_frame(f)
        img = eng.render_to_memory()[0].copy()
        st = eng.stats()
        frames.append(img)
        uploads.append(st["upload_bytes"])
        ref = oracle_loader.oracle_render(eng)[0]
        d = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        assert np.mean(d > 1e-3 * np.maximum(1.0, ref[..., :3].max(axis=-1))) < 0.01, "frame %d" % f
        assert abs(img[..., :3].mean() / ref[..., :3].mean() - 1) < 2e-3
    assert np.abs(frames[0] - frames[1]).mean() > 1e-3 and np.abs(frames[1] - frames[2]).mean() > 1e-3  # the camera really moved
    # frame 0 uploads the 3072 x 1536 RGBA32F environment (75 MB); later frames only geometry + records (< 2 MB)
    assert uploads[0] > 50e6 and uploads[1] < 5e6 and uploads[2] == uploads[1], uploads
    eng.close()


def test_offlinerender_frame_sequence_and_png(capi, tmp_path):
    """the product binary on the CUDA core: BallOnPlane render sequence to PNG files (the demo's file type)"""
    import subprocess
    from PIL import Image
    exe = os.path.join(capi.LIB_DIR, "offlinerender")
    r = subprocess.run([exe, "--scene", "BallOnPlane", "--frames", "2", "--width", "64", "--height", "48", "--spp", "8", "--batch", "8",
                        "--out", str(tmp_path / "f")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    st = json.loads(r.stdout.strip().splitlines()[-1])
    assert st["backend"] == "cuda-sm_100a" and st["segments"] > 0
    a, b = np.asarray(Image.open(str(tmp_path / "f0.png"))), np.asarray(Image.open(str(tmp_path / "f1.png")))
    assert a.shape == b.shape == (48, 64, 4) and np.all(a[..., 3] == 254) and np.abs(a.astype(int) - b.astype(int)).mean() > 0.5


# ---------------------------------------------------------------- multi-GPU inside the product
def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["tile", "sample", "none"])
def test_multi_device_context_equals_single_device(capi, engine, mode):
    """ptc_create with several devices: the context replicates the scene, partitions the render, reduces with NCCL onto device 0.
    Same samples as the one-GPU render - only the float summation order of the accumulators differs."""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    engine.build_scene("MeshLight")
    engine.set_render_info(width=160, height=96, samples=32, batch_size=4)
    desc, rp = engine.scene_desc(), engine.render_params()
    one = capi.Context(capi.load_cuda(), device=0)
    one.upload_scene(desc)
    one.build_accel()
    full = np.stack(one.render(rp))
    seg_full = one.stats()["segments"]
    one.close()
    devs = list(range(min(n, 4)))
    multi = capi.Context(capi.load_cuda(), device=devs)
    assert multi.device_count() == len(devs)
    multi.upload_scene(desc)
    multi.build_accel()
    rp.split_mode = {"tile": capi.PTC_SPLIT_TILE, "sample": capi.PTC_SPLIT_SAMPLE, "none": capi.PTC_SPLIT_NONE}[mode]
    rp.tile_size = 16
    part = np.stack(multi.render(rp))
    st = multi.stats()
    multi.close()
    assert st["segments"] == seg_full
    assert np.allclose(part[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)
    assert np.all(part[..., 3] == 1.0)  # alpha after the multi-rank sum
    assert st["reduce_ms"] > 0.0


def test_geometry_reaches_the_other_devices_over_the_communicator(capi, engine, monkeypatch):
    """several devices in one process: the vertex / index pools are copied to the first device only and broadcast from there
    (ncclBroadcast); the render is bit-identical to the one with a host copy per device"""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    engine.build_scene("MeshLight")
    engine.set_render_info(width=160, height=96, samples=16, batch_size=4)
    desc, rp = engine.scene_desc(), engine.render_params()
    rp.split_mode, rp.tile_size = capi.PTC_SPLIT_TILE, 16
    images = {}
    for label, min_bytes in (("broadcast", "0"), ("host copies", "-1")):
        monkeypatch.setenv("PTC_SCENE_BROADCAST_MIN_BYTES", min_bytes)
        multi = capi.Context(capi.load_cuda(), device=list(range(min(n, 4))))
        multi.upload_scene(desc)
        multi.build_accel()
        images[label] = np.stack(multi.render(rp))
        multi.close()
    assert np.array_equal(images["broadcast"], images["host copies"])
    assert np.isfinite(images["broadcast"]).all() and images["broadcast"][0, ..., :3].mean() > 0.01


def test_engine_with_several_devices_renders_through_the_plugin(capi):
    """RendererPathTracing::render() itself uses more than one B200 (VERDICT r1 'missing' item 1)"""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    single = capi.HostEngine()
    single.build_scene("Cornell")
    single.set_render_info(width=128, height=128, samples=32, batch_size=8)
    a = single.render_to_memory()[0].copy()
    seg = single.stats()["segments"]
    single.close()
    eng = capi.HostEngine(devices=list(range(min(n, 8))))
    assert eng.device_count() == min(n, 8)
    eng.build_scene("Cornell")
    eng.set_render_info(width=128, height=128, samples=32, batch_size=8)
    for split in ("tile", "sample"):
        eng.set_render_options(split=split)
        b = eng.render_to_memory()[0]
        assert eng.stats()["segments"] == seg
        assert np.allclose(a[..., :3], b[..., :3], rtol=1e-5, atol=1e-6) and np.all(b[..., 3] == 1.0)
    eng.close()


def _rank_worker(rank, world, id_path, out_dir):
    sys.path.insert(0, ROOT)
    import time
    import torch
    from vviewer_b200 import capi
    torch.cuda.set_device(rank)
    eng = capi.HostEngine()
    # one rank makes the id, the launcher (here: a file) hands it to all (ptc_comm_unique_id / ptc_comm_init_rank)
    if rank == 0:
        open(id_path + ".tmp", "wb").write(eng.comm_unique_id())
        os.rename(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.01)
    eng.comm_init_rank(open(id_path, "rb").read(), rank, world)
    eng.build_scene("Cornell")
    eng.set_render_info(width=128, height=128, samples=32, batch_size=8)
    eng.set_render_options(split="tile" if rank >= 0 else None)
    img = eng.render_to_memory()[0]
    st = eng.stats()
    if rank == 0:
        np.save(os.path.join(out_dir, "img.npy"), img)
    np.save(os.path.join(out_dir, "seg%d.npy" % rank), np.array([st["segments"]]))
    eng.close()


def test_one_process_per_gpu_communicator(capi, tmp_path):
    """the torchrun shape: every process owns one GPU and one engine, the engines join one NCCL communicator, render() is
    collective and rank 0 receives the image"""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_rank_worker, args=(world, str(tmp_path / "nccl_id"), str(tmp_path)), nprocs=world, join=True)
    img = np.load(str(tmp_path / "img.npy"))
    seg = sum(int(np.load(str(tmp_path / ("seg%d.npy" % r)))[0]) for r in range(world))
    single = capi.HostEngine()
    single.build_scene("Cornell")
    single.set_render_info(width=128, height=128, samples=32, batch_size=8)
    a = single.render_to_memory()[0]
    assert single.stats()["segments"] in (seg, seg // world)  # every rank reports the group's statistics or its own share
    assert np.allclose(a[..., :3], img[..., :3], rtol=1e-5, atol=1e-6) and np.all(img[..., 3] == 1.0)
    single.close()


# ---------------------------------------------------------------- two-level acceleration structure (VulkanScene.cpp:306-381)
def _instanced(engine, **kw):
    engine.build_scene("Instanced", **kw)
    return engine.scene_desc()


def test_two_level_build_bit_exact(capi, engine):
    """every tree of the two-level structure - one per mesh over object-space triangles, one over the instances' world boxes - equals the
    oracle's restatement byte for byte: node words, primitive order, bounds"""
    desc = _instanced(engine, texture_size=16, scale=0.01)
    d = desc.contents
    cu, orc = capi.Context(capi.load_cuda()), capi.Context(oracle_loader.load_oracle())
    for ctx in (cu, orc):
        ctx.upload_scene(desc)
        ctx.set_accel_mode(capi.PTC_ACCEL_TWO_LEVEL)
        ctx.build_accel()
    assert cu.stats()["accel_levels"] == 2
    seen_tris = 0
    for level in [-1] + list(range(d.n_meshes)):
        a, b = cu.get_accel_level(level), orc.get_accel_level(level)
        assert a["n_prims"] == b["n_prims"] and a["n_nodes"] == b["n_nodes"], level
        if a["n_prims"] == 0:
            continue
        assert np.array_equal(a["order"], b["order"]), level
        assert np.array_equal(a["words"], b["words"]), "level %d: %d nodes differ" % (level, int(np.any(a["words"] != b["words"], axis=1).sum()))
        assert np.array_equal(a["box"], b["box"]), level
        assert sorted(a["order"]) == list(range(a["n_prims"]))
        if level >= 0:
            seen_tris += a["n_prims"]
    assert cu.get_accel_level(-1)["n_prims"] == d.n_instances
    used = {d.instances[i].mesh_index for i in range(d.n_instances)}
    assert seen_tris == sum(d.meshes[m].tri_count for m in used)
    with pytest.raises(RuntimeError, match="two levels"):
        cu.get_wide_bvh()
    cu.close()
    orc.close()


@pytest.mark.parametrize("scene,kw,box", [("Instanced", dict(texture_size=16, scale=0.01), 60.0), ("SharedComponents", {}, 60.0), ("Hierarchy", {}, 12.0)])
def test_two_level_ray_sets_match_oracle(capi, engine, scene, kw, box):
    """closest hits through instance transforms: ids identical to the oracle's world-space queries, |dt| <= 1e-4 max(1, t)"""
    from test_gpu_parity import ray_set
    engine.build_scene(scene, **kw)
    desc = engine.scene_desc()
    cu, orc = capi.Context(capi.load_cuda()), capi.Context(oracle_loader.load_oracle())
    for ctx in (cu, orc):
        ctx.upload_scene(desc)
        ctx.set_accel_mode(capi.PTC_ACCEL_TWO_LEVEL)
        ctx.build_accel()
    rng = np.random.default_rng(21)
    rays = np.concatenate([ray_set(rng, 30000, -box, box), ray_set(rng, 30000, -box, box, aim=(-box / 4, box / 4))])
    ia, pa, ta, ua, va = cu.trace_closest(rays)
    ib, pb, tb, ub, vb = orc.trace_closest(rays)
    assert (ib >= 0).mean() > 0.1
    bad = ~((ia == ib) & (pa == pb))
    # object-space vs world-space intersection: a different rounding can flip the winner only on shared edges / coplanar overlaps
    assert bad.mean() <= 2e-3, "id mismatches: %d" % int(bad.sum())
    if bad.any():
        assert np.all(np.abs(ta[bad] - tb[bad]) <= 1e-3 * np.maximum(1.0, tb[bad]))
    hit = ~bad & (ib >= 0)
    assert np.all(np.abs(ta[hit] - tb[hit]) <= 1e-4 * np.maximum(1.0, tb[hit]))
    db = np.maximum(np.abs(ua[hit] - ub[hit]), np.abs(va[hit] - vb[hit]))
    assert np.percentile(db, 99.9) <= 1e-3
    # and the flat structure of the same scene answers the same
    flat = capi.Context(capi.load_cuda())
    flat.upload_scene(desc)
    flat.set_accel_mode(capi.PTC_ACCEL_FLAT)
    flat.build_accel()
    assert flat.stats()["accel_levels"] == 1
    ic, pc, tc, _, _ = flat.trace_closest(rays)
    assert np.mean((ia != ic) | (pa != pc)) <= 2e-3
    for c in (cu, orc, flat):
        c.close()


@pytest.mark.parametrize("scene,kw", [("Instanced", dict(texture_size=64, scale=0.004)), ("SharedComponents", {}), ("MeshLight", {}), ("Volume5", {}),
                                      ("Transparency", {})])
def test_two_level_render_matches_oracle(capi, engine, scene, kw):
    """the whole pipeline (path rays, shadow chains incl. the all-hits foliage mode, probes, media) over the two-level structure"""
    engine.build_scene(scene, **kw)
    engine.set_render_info(width=128, height=96, samples=8, batch_size=4)
    desc, rp = engine.scene_desc(), engine.render_params()
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(desc)
    cu.set_accel_mode(capi.PTC_ACCEL_TWO_LEVEL)
    cu.build_accel()
    ra, aa, na = cu.render(rp)
    sa = cu.stats()
    assert sa["accel_levels"] == 2
    cu.close()
    orc = oracle_loader.oracle_context(engine)
    rb, ab, nb = orc.render(rp)
    sb = orc.stats()
    orc.close()
    loose = scene in ("Instanced", "Transparency")
    d = np.abs(ra[..., :3] - rb[..., :3]).max(axis=-1)
    assert np.mean(d > 1e-3 * np.maximum(1.0, rb[..., :3].max(axis=-1))) < (0.03 if loose else 0.01)
    assert abs(ra[..., :3].mean() / rb[..., :3].mean() - 1) < (5e-3 if loose else 2e-3)
    assert np.mean(np.abs(aa - ab).max(axis=-1) > 1e-3) < (0.03 if loose else 0.01) and np.mean(np.abs(na - nb).max(axis=-1) > 1e-3) < (0.03 if loose else 0.01)
    assert abs(sa["segments"] - sb["segments"]) <= 2e-3 * sb["segments"] and np.all(ra[..., 3] == 1.0)


def test_accel_mode_auto(capi, engine):
    """AUTO flattens (measured faster, profiles/r2_c4_flat_vs_two_level.log) until the flattened data would not fit a quarter of the
    device memory; the threshold is overridable, which is how this test reaches the other branch with a small scene"""
    desc = _instanced(engine, texture_size=16, scale=0.02)
    cu = capi.Context(capi.load_cuda())
    cu.upload_scene(desc)
    cu.build_accel()
    flat = cu.stats()
    assert flat["accel_levels"] == 1
    cu.close()
    os.environ["PTC_TWO_LEVEL_MIN_BYTES"] = "1000000"
    try:
        cu = capi.Context(capi.load_cuda())
    finally:
        del os.environ["PTC_TWO_LEVEL_MIN_BYTES"]
    cu.upload_scene(desc)
    cu.build_accel()
    two = cu.stats()
    assert two["accel_levels"] == 2 and two["n_triangles"] == flat["n_triangles"] and two["traversal_bytes"] < 0.5 * flat["traversal_bytes"]
    engine.build_scene("Atrium", texture_size=4, scale=0.3)  # hardly any instancing: stays flat whatever the size
    cu.upload_scene(engine.scene_desc())
    cu.build_accel()
    assert cu.stats()["accel_levels"] == 1
    cu.close()


def test_packed_normal_roughness_tap_is_bit_identical(capi, engine):
    """materials whose normal and roughness maps have the same size get ONE tap on the device (roughness in the alpha channel of a copy
    of the normal map): the filter treats the four channels of a linear RGBA8 texture alike, so the image does not change by a bit"""
    for scene, kw in (("Atrium", dict(scale=0.1, texture_size=64)), ("GLTF", {})):
        engine.build_scene(scene, **kw)
        engine.set_render_info(width=128, height=72, samples=8, batch_size=4)
        desc, rp = engine.scene_desc(), engine.render_params()
        out = {}
        for packing in (True, False):
            if not packing:
                os.environ["PTC_NO_TEXTURE_PACKING"] = "1"
            try:
                cu = capi.Context(capi.load_cuda())
                cu.upload_scene(desc)
            finally:
                os.environ.pop("PTC_NO_TEXTURE_PACKING", None)
            cu.build_accel()
            out[packing] = np.stack(cu.render(rp))
            cu.close()
        assert np.array_equal(out[True], out[False]), scene
        assert out[True][0, ..., :3].mean() > 1e-3
