"""Host-side logic (CPU): image I/O, camera matrices, scene flattening rules of the reference."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from imgmetrics import rgbe_roundtrip
from oracle import loader as oracle_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def engine(capi):
    eng = capi.HostEngine()
    yield eng
    eng.close()


def test_hdr_roundtrip(capi, tmp_path):
    rng = np.random.default_rng(0)
    img = np.zeros((37, 53, 4), np.float32)
    img[..., :3] = rng.uniform(0, 4, (37, 53, 3)) ** 3
    img[5:9, 3:40, :3] = 0.25  # a run for the RLE encoder
    img[..., 3] = 1
    p = str(tmp_path / "t.hdr")
    capi.write_hdr(p, img)
    back = capi.read_hdr(p)
    assert back.shape == img.shape
    # the file holds exactly the truncating RGBE quantisation of stbi_write_hdr
    assert np.allclose(back[..., :3], rgbe_roundtrip(img), rtol=0, atol=0)
    rel = np.abs(back[..., :3] - img[..., :3]).max(axis=-1) / img[..., :3].max(axis=-1)
    assert rel.max() < 1.0 / 128.0


def test_hdr_reader_matches_opencv_on_golden(capi):
    cv2 = pytest.importorskip("cv2")
    p = os.path.join(ROOT, "tests", "golden", "reference_images", "FurnacePBR_ref.hdr")
    ours = capi.read_hdr(p)[..., :3]
    theirs = cv2.imread(p, cv2.IMREAD_UNCHANGED)[..., ::-1]
    assert ours.shape == (256, 256, 3)
    # OpenCV decodes with (m + 0.5) * 2^(e-136); stb (the reference) with m * 2^(e-136)
    assert np.allclose(ours, theirs, rtol=0.01, atol=1e-6)


def test_scene_list_and_unknown_scene(engine):
    names = engine.scene_list()
    for n in ("FurnacePBR", "Volume9", "MeshLight", "SharedComponents", "Cornell", "Atrium"):
        assert n in names
    with pytest.raises(RuntimeError):
        engine.build_scene("NoSuchScene")


def test_render_info_defaults_follow_recipe(engine):
    engine.build_scene("DepthOfField")
    assert engine.render_info() == dict(width=256, height=256, samples=2048, batch_size=4, depth=9)
    engine.build_scene("Cornell")
    assert engine.render_info() == dict(width=512, height=512, samples=64, batch_size=8, depth=8)


def test_camera_matrices(engine):
    """glm::lookAt / glm::perspective (depth 0..1, flipped y) as in VulkanRendererPathTracing.cpp:149-157."""
    engine.build_scene("FurnacePBR")  # camera at (0,0,2) looking down -z, fov 60, 256x256
    rp = engine.render_params()
    view = np.array(rp.scene.view, np.float32).reshape(4, 4).T
    vinv = np.array(rp.scene.view_inverse, np.float32).reshape(4, 4).T
    proj = np.array(rp.scene.projection, np.float32).reshape(4, 4).T
    pinv = np.array(rp.scene.projection_inverse, np.float32).reshape(4, 4).T
    assert np.allclose(view @ vinv, np.eye(4), atol=1e-6)
    assert np.allclose(proj @ pinv, np.eye(4), atol=1e-5)
    assert np.allclose(vinv[:3, 3], [0, 0, 2])
    t = math.tan(math.radians(60) / 2)
    assert np.isclose(proj[0, 0], 1 / t, rtol=1e-6) and np.isclose(proj[1, 1], -1 / t, rtol=1e-6)
    zn, zf = 0.5, 50.0
    assert np.isclose(proj[2, 2], zf / (zn - zf)) and np.isclose(proj[2, 3], -(zf * zn) / (zf - zn)) and proj[3, 2] == -1
    assert rp.scene.volumes[0] == -1 and rp.scene.volumes[1] == zn and rp.scene.volumes[2] == zf
    assert rp.scene.background[3] == 0.0 and tuple(rp.scene.background[:3]) == (1.0, 1.0, 1.0)


def test_flatten_mesh_light_and_order(engine):
    """Instances.cpp:100-181 + VulkanInstances.cpp:66-109: ComponentLight objects first, then mesh lights; mesh light rows
    are the rows of the model matrix; emissive mesh is also an ordinary instance."""
    engine.build_scene("MeshLight")
    d = engine.scene_desc().contents
    assert d.n_instances == 4 and d.n_light_instances == 1
    li = d.light_instances[0]
    assert li.info[3] == 2
    inst = d.instances[li.info[1]]
    model = np.array(inst.model, np.float32).reshape(4, 4).T
    rows = np.array([list(li.position), list(li.position1), list(li.position2)], np.float32)
    assert np.allclose(rows, model[:3, :])
    assert inst.num_triangles == 2
    mat = d.materials[inst.material_index]
    assert tuple(mat.emissive) == pytest.approx((0.3, 0.3, 0.7, 15.0))
    # plane scaled 0.4, rotated 90 deg about x, at (0,2,0)
    assert np.allclose(model[:3, 3], [0, 2, 0]) and np.isclose(np.linalg.norm(model[:3, 0]), 0.4, rtol=1e-5)


def test_flatten_keeps_geometry_pools_while_the_meshes_stay(capi):
    """Meshes are immutable after import (the reference uploads them once, VulkanMesh): flattening the same scene again - here
    after a camera move of the render sequence - hands out the SAME vertex / index pools; another scene gets new ones with
    its own content."""
    eng = capi.HostEngine()
    try:
        eng.build_scene("BallOnPlane")
        d = eng.scene_desc().contents
        first = (C.addressof(d.vertices.contents), C.addressof(d.indices.contents), int(d.n_vertices), int(d.n_indices), d.n_meshes)
        v0 = np.array(d.vertices[0].position, np.float32).copy()
        view0 = np.array(eng.render_params().scene.view, np.float32).copy()
        eng.set_sequence_frame(7)  # invalidates the flattened scene: only the camera differs
        d = eng.scene_desc().contents
        again = (C.addressof(d.vertices.contents), C.addressof(d.indices.contents), int(d.n_vertices), int(d.n_indices), d.n_meshes)
        assert again == first
        assert not np.array_equal(np.array(eng.render_params().scene.view, np.float32), view0)
        assert np.array_equal(np.array(d.vertices[0].position, np.float32), v0)
        eng.build_scene("Cornell")  # other meshes
        d = eng.scene_desc().contents
        assert (int(d.n_vertices), int(d.n_indices)) != first[2:4]
        tri = sum(d.meshes[m].tri_count for m in range(d.n_meshes))
        assert tri * 3 == d.n_indices
        for m in range(d.n_meshes):
            me = d.meshes[m]
            assert me.first_vertex + me.vertex_count <= d.n_vertices and me.first_index + 3 * me.tri_count <= d.n_indices
    finally:
        eng.close()


def test_flatten_directional_light_unnormalised(engine):
    """Trap T4: directional light direction = modelMatrix * (0,0,1,0), not normalised under scaled parents."""
    engine.build_scene("Hierarchy")
    d = engine.scene_desc().contents
    assert d.n_light_instances == 1
    li = d.light_instances[0]
    assert li.info[3] == 1
    n = np.linalg.norm(list(li.position)[:3])
    assert np.isclose(n, 1.1 * 0.5, rtol=1e-5)  # root2 scale 1.1 x l1_1 scale 0.5


def test_flatten_volumes_and_transparency(engine):
    engine.build_scene("Volume6")
    d = engine.scene_desc().contents
    rp = engine.render_params()
    cube = [d.instances[i] for i in range(d.n_instances) if d.instances[i].num_triangles == 12][0]
    assert cube.id[1] != -1 and cube.id[2] != -1 and cube.id[1] != cube.id[2]
    assert rp.scene.volumes[0] == cube.id[1]  # camera volume = the front-facing (fog) volume
    assert rp.scene.volumes[2] == 10.0
    mat = d.materials[cube.material_index]
    assert mat.metallic_roughness_ao[3] == 1.0 and mat.albedo[3] == pytest.approx(0.2)
    vol = d.materials[int(cube.id[2])]
    assert tuple(vol.albedo)[:3] == pytest.approx((0.2, 0.4, 0.0)) and tuple(vol.metallic_roughness_ao)[:3] == pytest.approx((0.8, 0.4, 0.2))
    assert int(vol.uv_tiling[2]) == 3


def test_shared_components_instancing(engine):
    engine.build_scene("SharedComponents")
    d = engine.scene_desc().contents
    assert d.n_instances == 100 * 100 and d.n_meshes == 1 and d.meshes[0].tri_count == 12
    assert d.n_light_instances == 1 and d.light_instances[0].info[3] == 1


def test_base_textures_and_png_loading(engine):
    """Texture slots 0,1,2 = white / whiteColor / normalmapdefault (VulkanTextures.cpp:71-76); PNGs are stored bottom row first."""
    engine.build_scene("Transparency")
    d = engine.scene_desc().contents
    assert d.n_textures >= 4
    t0, t1, t2 = d.textures[0], d.textures[1], d.textures[2]
    assert (t0.width, t0.height, t0.srgb) == (1, 1, 0) and (t1.srgb, t2.srgb) == (1, 0)
    assert [t2.data[i] for i in range(4)] == [0x80, 0x80, 0xFF, 0xFF]
    cv2 = pytest.importorskip("cv2")
    ref = cv2.imread(os.path.join(ROOT, "assets", "textures", "checkerboard.png"), cv2.IMREAD_UNCHANGED)
    ref = cv2.cvtColor(ref, cv2.COLOR_BGRA2RGBA)[::-1]  # flipped vertically (trap T11)
    tex = [d.textures[i] for i in range(d.n_textures) if d.textures[i].width == ref.shape[1] and d.textures[i].height == ref.shape[0]][0]
    ours = np.ctypeslib.as_array(tex.data, shape=(tex.height, tex.width, tex.channels))
    assert np.array_equal(ours, ref)


def test_obj_import_conventions(engine):
    """plane.obj: 4 unique vertices, 2 triangles, uv kept as in the file, unit normals, tangent along +u = +x."""
    engine.build_scene("NormalMap")
    d = engine.scene_desc().contents
    m = d.meshes[0]
    assert (m.tri_count, m.vertex_count) == (2, 4)
    verts = [d.vertices[m.first_vertex + i] for i in range(4)]
    for v in verts:
        assert tuple(v.normal) == pytest.approx((0, 1, 0))
        assert tuple(v.tangent) == pytest.approx((1, 0, 0), abs=1e-5)
        # u grows with x, v (file convention) grows with -z
        assert v.uv[0] == pytest.approx(0.5 + 0.4999 * v.position[0], abs=1e-4)
        assert v.uv[1] == pytest.approx(0.5 - 0.4999 * v.position[2], abs=1e-4)


def test_atrium_is_sponza_class(capi):
    eng = capi.HostEngine()
    eng.build_scene("Atrium", texture_size=8)
    d = eng.scene_desc().contents
    tris = sum(d.instances[i].num_triangles for i in range(d.n_instances))
    assert 285000 <= tris <= 315000
    used = {d.instances[i].material_index for i in range(d.n_instances)}
    assert len(used) >= 20 and d.n_light_instances == 0
    assert eng.render_info() == dict(width=1920, height=1080, samples=1024, batch_size=16, depth=9)
    eng.close()


@pytest.mark.parametrize("scene,kw", [("Progressive", {"scale": 0.05}), ("Cornell", {}), ("Fog", {"scale": 0.02, "texture_size": 16}),
                                      ("Instanced", {"scale": 0.004, "texture_size": 16})])
def test_workload_scenes_render_finite_images(capi, scene, kw):
    """emissive meshes must not contain zero-area triangles: the light sampler divides by the triangle area (lightSampling.glsl:44-100)
    and the power heuristic of an infinite pdf is NaN - the sphere of the progressive workload once had collapsed pole triangles"""
    eng = capi.HostEngine()
    eng.build_scene(scene, **kw)
    eng.set_render_info(width=96, height=64, samples=8, batch_size=8)
    rad, alb, nrm = oracle_loader.oracle_render(eng)
    assert np.isfinite(rad).all() and np.isfinite(alb).all() and np.isfinite(nrm).all()
    assert rad[..., :3].mean() > 0
    eng.close()
