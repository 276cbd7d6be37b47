"""ctypes bindings for the two C interfaces of this repository.

* ``include/ptc.h``           — the drop-in boundary (libptc_cuda.so; the oracle exports the same surface)
* ``include/vengine_host.h``  — C view of the C++ host library (scene model + RendererPathTracing)

Nothing here computes anything: it only declares structures / prototypes and loads shared libraries.
The product library is ``vviewer_b200/_lib/libptc_cuda.so``; loading it fails loudly when it is missing
(there is no CPU fallback, and nothing in this package knows where the test oracle lives).
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "vviewer_b200", "_lib")
CUDA_LIB = os.path.join(LIB_DIR, "libptc_cuda.so")
HOST_LIB = os.path.join(LIB_DIR, "libvengine_host.so")

f32 = C.c_float
u32 = C.c_uint32
u64 = C.c_uint64


class ptc_vertex(C.Structure):
    _fields_ = [("position", f32 * 3), ("uv", f32 * 2), ("normal", f32 * 3), ("color", f32 * 3), ("tangent", f32 * 3),
                ("bitangent", f32 * 3)]


class ptc_mesh(C.Structure):
    _fields_ = [("first_index", u32), ("tri_count", u32), ("first_vertex", u32), ("vertex_count", u32)]


class ptc_instance(C.Structure):
    _fields_ = [("model", f32 * 16), ("id", f32 * 4), ("material_index", u32), ("mesh_index", u32), ("num_triangles", u32),
                ("pad", u32 * 9)]


class ptc_material(C.Structure):
    _fields_ = [("albedo", f32 * 4), ("metallic_roughness_ao", f32 * 4), ("emissive", f32 * 4), ("tex1", u32 * 4),
                ("tex2", u32 * 4), ("uv_tiling", f32 * 4), ("pad1", u32 * 4), ("pad2", u32 * 4)]


class ptc_light_data(C.Structure):
    _fields_ = [("color", f32 * 4), ("type", u32 * 4), ("pad1", u32 * 4), ("pad2", u32 * 4)]


class ptc_light_instance(C.Structure):
    _fields_ = [("info", u32 * 4), ("position", f32 * 4), ("position1", f32 * 4), ("position2", f32 * 4)]


class ptc_texture(C.Structure):
    _fields_ = [("width", u32), ("height", u32), ("channels", u32), ("srgb", u32), ("data", C.POINTER(C.c_uint8)), ("uid", u64)]


class ptc_env(C.Structure):
    _fields_ = [("equirect_rgba", C.POINTER(f32)), ("width", u32), ("height", u32), ("uid", u64)]


class ptc_scene_desc(C.Structure):
    _fields_ = [("vertices", C.POINTER(ptc_vertex)), ("n_vertices", u64), ("indices", C.POINTER(u32)), ("n_indices", u64),
                ("meshes", C.POINTER(ptc_mesh)), ("n_meshes", u32), ("instances", C.POINTER(ptc_instance)), ("n_instances", u32),
                ("materials", C.POINTER(ptc_material)), ("n_materials", u32), ("light_data", C.POINTER(ptc_light_data)),
                ("n_light_data", u32), ("light_instances", C.POINTER(ptc_light_instance)), ("n_light_instances", u32),
                ("textures", C.POINTER(ptc_texture)), ("n_textures", u32), ("env", ptc_env)]


class ptc_scene_data(C.Structure):
    _fields_ = [("view", f32 * 16), ("view_inverse", f32 * 16), ("projection", f32 * 16), ("projection_inverse", f32 * 16),
                ("exposure", f32 * 4), ("background", f32 * 4), ("volumes", f32 * 4)]


class ptc_render_params(C.Structure):
    _fields_ = [("scene", ptc_scene_data), ("samples", u32), ("batch_size", u32), ("depth", u32), ("width", u32), ("height", u32),
                ("camera_type", u32), ("ortho_width", f32), ("ortho_height", f32), ("split_mode", u32), ("rank", u32), ("world", u32),
                ("tile_size", u32), ("flags", u32), ("reserved", u32 * 7)]


class ptc_stats(C.Structure):
    _fields_ = [("segments", u64), ("path_rays", u64), ("shadow_rays", u64), ("shadow_hops", u64), ("probe_rays", u64),
                ("probe_hops", u64), ("render_ms", C.c_double), ("trace_ms", C.c_double), ("shade_ms", C.c_double),
                ("shadow_ms", C.c_double), ("build_ms", C.c_double), ("trace_launches", u64), ("kernel_launches", u64),
                ("n_triangles", u64), ("n_bvh_nodes", u64), ("scene_bytes", u64), ("reserved", u64 * 4),
                ("upload_bytes", u64), ("reduce_ms", C.c_double), ("reserved_ms", C.c_double), ("accel_levels", u64), ("traversal_bytes", u64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["reserved"] = list(self.reserved)
        return d


PTC_SPLIT_NONE, PTC_SPLIT_TILE, PTC_SPLIT_SAMPLE = 0, 1, 2
PTC_CAMERA_PERSPECTIVE, PTC_CAMERA_ORTHOGRAPHIC = 0, 1
PTC_FLAG_WORLD_ORIGIN_PROBE_PDF = 1
PTC_FLAG_TIME_KERNELS = 2
PTC_FLAG_SAMPLER_SOBOL = 4
PTC_FLAG_ENV_IMPORTANCE = 8
PTC_FLAG_SAMPLER_PMJ = 16
PTC_SAMPLER_HOOK_1D = 0x80000000
PTC_HIERARCHY_LBVH, PTC_HIERARCHY_PLOC = 0, 1
PTC_ACCEL_AUTO, PTC_ACCEL_FLAT, PTC_ACCEL_TWO_LEVEL = 0, 1, 2

# every symbol include/ptc.h declares
PTC_SYMBOLS = ["ptc_set_accel_mode", "ptc_get_accel_level", "ptc_srgb_table", "ptc_set_sampler_tables", "ptc_create", "ptc_destroy", "ptc_device_count", "ptc_comm_unique_id", "ptc_comm_init_rank", "ptc_last_error", "ptc_backend_name", "ptc_upload_scene", "ptc_set_build_options", "ptc_build_accel", "ptc_render",
               "ptc_render_device", "ptc_progress", "ptc_get_stats", "ptc_trace_closest", "ptc_get_lbvh", "ptc_get_wide_bvh", "ptc_bsdf_eval",
               "ptc_bsdf_sample", "ptc_sampler_points", "ptc_env_lookup", "ptc_env_sample", "ptc_env_pdf"]
VH_SYMBOLS = ["vh_set_sequence_frame", "vh_set_output", "vh_write_image", "vh_set_devices", "vh_device_count", "vh_comm_unique_id", "vh_comm_init_rank", "vh_set_render_options", "vh_engine_create", "vh_engine_destroy", "vh_backend_ok", "vh_last_error", "vh_scene_list", "vh_build_scene",
              "vh_set_render_info", "vh_get_render_info", "vh_scene_desc", "vh_render_params", "vh_render_to_memory", "vh_render",
              "vh_get_stats", "vh_read_hdr", "vh_write_hdr", "vh_import_model", "vh_add_model", "vh_import_scene", "vh_export_scene",
              "vh_describe", "vh_decode_image", "vh_render_progress"]

_fp = C.POINTER(f32)
_ip = C.POINTER(C.c_int)


def _declare_ptc(lib):
    vp = C.c_void_p
    lib.ptc_create.argtypes = [C.POINTER(vp), _ip, C.c_int]
    lib.ptc_create.restype = C.c_int
    lib.ptc_destroy.argtypes = [vp]
    lib.ptc_destroy.restype = None
    lib.ptc_last_error.argtypes = [vp]
    lib.ptc_last_error.restype = C.c_char_p
    lib.ptc_set_accel_mode.argtypes = [vp, u32]
    lib.ptc_set_accel_mode.restype = C.c_int
    lib.ptc_get_accel_level.argtypes = [vp, C.c_int32, C.POINTER(u64), C.POINTER(u64), vp, vp, vp]
    lib.ptc_get_accel_level.restype = C.c_int
    lib.ptc_srgb_table.argtypes = [vp, vp]
    lib.ptc_srgb_table.restype = C.c_int
    lib.ptc_set_sampler_tables.argtypes = [vp, vp, u32, u32, vp, u32, u32]
    lib.ptc_set_sampler_tables.restype = C.c_int
    lib.ptc_device_count.argtypes = [vp]
    lib.ptc_device_count.restype = C.c_int
    lib.ptc_comm_unique_id.argtypes = [vp]
    lib.ptc_comm_unique_id.restype = C.c_int
    lib.ptc_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.ptc_comm_init_rank.restype = C.c_int
    lib.ptc_backend_name.argtypes = []
    lib.ptc_backend_name.restype = C.c_char_p
    lib.ptc_upload_scene.argtypes = [vp, C.POINTER(ptc_scene_desc)]
    lib.ptc_upload_scene.restype = C.c_int
    lib.ptc_set_build_options.argtypes = [vp, u32, u32]
    lib.ptc_set_build_options.restype = C.c_int
    lib.ptc_build_accel.argtypes = [vp]
    lib.ptc_build_accel.restype = C.c_int
    lib.ptc_render.argtypes = [vp, C.POINTER(ptc_render_params), vp, vp, vp]
    lib.ptc_render.restype = C.c_int
    lib.ptc_render_device.argtypes = [vp, C.POINTER(ptc_render_params), vp, vp, vp]
    lib.ptc_render_device.restype = C.c_int
    lib.ptc_progress.argtypes = [vp]
    lib.ptc_progress.restype = f32
    lib.ptc_get_stats.argtypes = [vp, C.POINTER(ptc_stats)]
    lib.ptc_get_stats.restype = C.c_int
    lib.ptc_trace_closest.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp]
    lib.ptc_trace_closest.restype = C.c_int
    lib.ptc_get_lbvh.argtypes = [vp, C.POINTER(u64), vp, vp, vp, vp, vp, vp]
    lib.ptc_get_lbvh.restype = C.c_int
    lib.ptc_get_wide_bvh.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), vp, vp]
    lib.ptc_get_wide_bvh.restype = C.c_int
    lib.ptc_bsdf_eval.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    lib.ptc_bsdf_eval.restype = C.c_int
    lib.ptc_bsdf_sample.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.ptc_bsdf_sample.restype = C.c_int
    lib.ptc_sampler_points.argtypes = [vp, u32, u32, u32, u32, u32, u32, u32, vp]
    lib.ptc_sampler_points.restype = C.c_int
    lib.ptc_env_lookup.argtypes = [vp, C.c_int, vp, vp]
    lib.ptc_env_lookup.restype = C.c_int
    lib.ptc_env_sample.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.ptc_env_sample.restype = C.c_int
    lib.ptc_env_pdf.argtypes = [vp, C.c_int, vp, vp]
    lib.ptc_env_pdf.restype = C.c_int
    return lib


def _declare_vh(lib):
    vp = C.c_void_p
    lib.vh_set_devices.argtypes = [vp, _ip, C.c_int]
    lib.vh_set_devices.restype = C.c_int
    lib.vh_device_count.argtypes = [vp]
    lib.vh_device_count.restype = C.c_int
    lib.vh_comm_unique_id.argtypes = [vp, vp]
    lib.vh_comm_unique_id.restype = C.c_int
    lib.vh_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.vh_comm_init_rank.restype = C.c_int
    lib.vh_set_render_options.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.vh_set_render_options.restype = None
    lib.vh_set_sequence_frame.argtypes = [vp, C.c_int]
    lib.vh_set_sequence_frame.restype = C.c_int
    lib.vh_set_output.argtypes = [vp, C.c_int, f32, C.c_int, C.c_int]
    lib.vh_set_output.restype = None
    lib.vh_write_image.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, vp, C.c_int, f32]
    lib.vh_write_image.restype = C.c_int
    lib.vh_engine_create.argtypes = [C.c_char_p]
    lib.vh_engine_create.restype = vp
    lib.vh_engine_destroy.argtypes = [vp]
    lib.vh_engine_destroy.restype = None
    lib.vh_backend_ok.argtypes = [vp]
    lib.vh_backend_ok.restype = C.c_int
    lib.vh_last_error.argtypes = [vp]
    lib.vh_last_error.restype = C.c_char_p
    lib.vh_scene_list.argtypes = []
    lib.vh_scene_list.restype = C.c_char_p
    lib.vh_build_scene.argtypes = [vp, C.c_char_p, C.c_int, f32, C.c_int]
    lib.vh_build_scene.restype = C.c_int
    lib.vh_set_render_info.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.vh_set_render_info.restype = None
    lib.vh_get_render_info.argtypes = [vp, _ip, _ip, _ip, _ip, _ip]
    lib.vh_get_render_info.restype = None
    lib.vh_scene_desc.argtypes = [vp]
    lib.vh_scene_desc.restype = C.POINTER(ptc_scene_desc)
    lib.vh_render_params.argtypes = [vp, C.POINTER(ptc_render_params)]
    lib.vh_render_params.restype = C.c_int
    lib.vh_render_to_memory.argtypes = [vp, vp, vp, vp]
    lib.vh_render_to_memory.restype = C.c_int
    lib.vh_render.argtypes = [vp, C.c_char_p]
    lib.vh_render.restype = C.c_int
    lib.vh_get_stats.argtypes = [vp, C.POINTER(ptc_stats)]
    lib.vh_get_stats.restype = C.c_int
    lib.vh_render_progress.argtypes = [vp]
    lib.vh_render_progress.restype = f32
    lib.vh_read_hdr.argtypes = [C.c_char_p, _ip, _ip, vp]
    lib.vh_read_hdr.restype = C.c_int
    lib.vh_write_hdr.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, vp]
    lib.vh_write_hdr.restype = C.c_int
    lib.vh_import_model.argtypes = [vp, C.c_char_p, C.c_int]
    lib.vh_import_model.restype = C.c_int
    lib.vh_add_model.argtypes = [vp, C.c_char_p]
    lib.vh_add_model.restype = C.c_int
    lib.vh_import_scene.argtypes = [vp, C.c_char_p]
    lib.vh_import_scene.restype = C.c_int
    lib.vh_export_scene.argtypes = [vp, C.c_char_p]
    lib.vh_export_scene.restype = C.c_int
    lib.vh_describe.argtypes = [vp]
    lib.vh_describe.restype = C.c_char_p
    lib.vh_decode_image.argtypes = [vp, C.c_uint64, C.c_int, _ip, _ip, _ip, _ip, vp, C.c_uint64]
    lib.vh_decode_image.restype = C.c_int
    return lib


def load_ptc(path):
    if not os.path.exists(path):
        raise RuntimeError("path-tracing backend library not found: %s (run `make` / __graft_entry__.build())" % path)
    return _declare_ptc(C.CDLL(path, mode=C.RTLD_LOCAL))


_cuda = None
_host = None


def load_cuda():
    """The product: hand-written sm_100a CUDA behind include/ptc.h. Raises if the library is absent."""
    global _cuda
    if _cuda is None:
        _cuda = load_ptc(CUDA_LIB)
    return _cuda


def load_host():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError("host library not found: %s (run `make`)" % HOST_LIB)
        _host = _declare_vh(C.CDLL(HOST_LIB, mode=C.RTLD_LOCAL))
    return _host


def load_sampler_tables(root=None):
    """assets/tables/pmj02bn.f32 + bluenoise.u16 -> (float32 [16, 16384, 2], float32 [48, 128, 128])"""
    d = os.path.join(root or ROOT, "assets", "tables")
    pmj = np.fromfile(os.path.join(d, "pmj02bn.f32"), "<f4").reshape(16, 16384, 2).copy()
    blue = (np.fromfile(os.path.join(d, "bluenoise.u16"), "<u2").astype(np.float32) / np.float32(65536.0)).reshape(48, 128, 128).copy()
    return pmj, blue


def np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """RAII wrapper of a ptc_ctx of one backend library."""

    def __init__(self, lib, device=None):
        """device: None (current device), an index, or a list of indices (one context driving several GPUs over NCCL)"""
        self.lib = lib
        self.ctx = C.c_void_p()
        if device is None:
            rc = lib.ptc_create(C.byref(self.ctx), None, 0)
        else:
            ids = [int(device)] if isinstance(device, int) else [int(x) for x in device]
            d = (C.c_int * len(ids))(*ids)
            rc = lib.ptc_create(C.byref(self.ctx), d, len(ids))
        if rc != 0 or not self.ctx:
            msg = lib.ptc_last_error(self.ctx).decode() if self.ctx else "no context"
            raise RuntimeError("ptc_create failed (%s): %s" % (lib.ptc_backend_name().decode(), msg))

    def close(self):
        if self.ctx:
            self.lib.ptc_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, self.lib.ptc_last_error(self.ctx).decode()))

    def device_count(self):
        return int(self.lib.ptc_device_count(self.ctx))

    def unique_id(self):
        """128 opaque bytes (ncclGetUniqueId) to hand to every rank's comm_init_rank"""
        buf = (C.c_uint8 * 128)()
        if self.lib.ptc_comm_unique_id(buf) != 0:
            raise RuntimeError("ptc_comm_unique_id failed (no NCCL library?)")
        return bytes(buf)

    def comm_init_rank(self, id_bytes, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(id_bytes)
        self._check(self.lib.ptc_comm_init_rank(self.ctx, buf, int(rank), int(world)), "ptc_comm_init_rank")

    def set_sampler_tables(self, tables=None):
        """the PMJ02BN / blue-noise tables (default: assets/tables, extracted from the reference by tools/extract_sampler_tables.py)"""
        pmj, blue = tables or load_sampler_tables()
        self._tables = (pmj, blue)
        self._check(self.lib.ptc_set_sampler_tables(self.ctx, np_ptr(pmj), 16, 16384, np_ptr(blue), 48, 128), "ptc_set_sampler_tables")

    def srgb_table(self):
        out = np.zeros(256, np.float32)
        self._check(self.lib.ptc_srgb_table(self.ctx, np_ptr(out)), "ptc_srgb_table")
        return out

    def upload_scene(self, desc_ptr):
        self._check(self.lib.ptc_upload_scene(self.ctx, desc_ptr), "ptc_upload_scene")

    def set_build_options(self, hierarchy, ploc_radius=0):
        self._check(self.lib.ptc_set_build_options(self.ctx, hierarchy, ploc_radius), "ptc_set_build_options")

    def set_accel_mode(self, mode):
        self._check(self.lib.ptc_set_accel_mode(self.ctx, int(mode)), "ptc_set_accel_mode")

    def get_accel_level(self, level):
        """two-level structure: level -1 = instance tree, m >= 0 = tree of mesh m"""
        nn, npr = u64(), u64()
        box = np.zeros(6, np.float32)
        self._check(self.lib.ptc_get_accel_level(self.ctx, int(level), C.byref(nn), C.byref(npr), None, None, np_ptr(box)), "ptc_get_accel_level")
        words = np.zeros((max(nn.value, 1), 20), np.uint32)
        order = np.zeros(max(npr.value, 1), np.uint32)
        self._check(self.lib.ptc_get_accel_level(self.ctx, int(level), C.byref(nn), C.byref(npr), np_ptr(words), np_ptr(order), None), "ptc_get_accel_level")
        return {"n_nodes": nn.value, "n_prims": npr.value, "words": words[:nn.value], "order": order[:npr.value], "box": box}

    def build_accel(self, hierarchy=None, ploc_radius=0):
        if hierarchy is not None:
            self.set_build_options(hierarchy, ploc_radius)
        self._check(self.lib.ptc_build_accel(self.ctx), "ptc_build_accel")

    def render(self, params, want_aovs=True):
        n = params.width * params.height * 4
        rad = np.zeros(n, np.float32)
        alb = np.zeros(n, np.float32) if want_aovs else None
        nrm = np.zeros(n, np.float32) if want_aovs else None
        self._check(self.lib.ptc_render(self.ctx, C.byref(params), np_ptr(rad), np_ptr(alb) if want_aovs else None,
                                        np_ptr(nrm) if want_aovs else None), "ptc_render")
        shape = (params.height, params.width, 4)
        if want_aovs:
            return rad.reshape(shape), alb.reshape(shape), nrm.reshape(shape)
        return rad.reshape(shape)

    def render_device(self, params, d_rad, d_alb, d_nrm):
        self._check(self.lib.ptc_render_device(self.ctx, C.byref(params), C.c_void_p(d_rad), C.c_void_p(d_alb) if d_alb else None,
                                               C.c_void_p(d_nrm) if d_nrm else None), "ptc_render_device")

    def stats(self):
        s = ptc_stats()
        self._check(self.lib.ptc_get_stats(self.ctx, C.byref(s)), "ptc_get_stats")
        return s.as_dict()

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        inst = np.zeros(n, np.int32)
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        u = np.zeros(n, np.float32)
        v = np.zeros(n, np.float32)
        self._check(self.lib.ptc_trace_closest(self.ctx, np_ptr(rays), n, np_ptr(inst), np_ptr(prim), np_ptr(t), np_ptr(u), np_ptr(v)),
                    "ptc_trace_closest")
        return inst, prim, t, u, v

    def get_lbvh(self):
        n = u64(0)
        self._check(self.lib.ptc_get_lbvh(self.ctx, C.byref(n), None, None, None, None, None, None), "ptc_get_lbvh")
        n = int(n.value)
        nn = max(2 * n - 1, 0)
        out = dict(n=n, morton=np.zeros(n, np.uint64), order=np.zeros(n, np.uint32), parent=np.zeros(nn, np.int32),
                   left=np.zeros(nn, np.int32), right=np.zeros(nn, np.int32), aabb=np.zeros((nn, 6), np.float32))
        if n:
            k = u64(0)
            self._check(self.lib.ptc_get_lbvh(self.ctx, C.byref(k), np_ptr(out["morton"]), np_ptr(out["order"]), np_ptr(out["parent"]),
                                              np_ptr(out["left"]), np_ptr(out["right"]), np_ptr(out["aabb"])), "ptc_get_lbvh")
        return out

    def get_wide_bvh(self):
        """8-wide compressed BVH: words (n_nodes, 20) uint32 and tri_order (n_tris,) uint32."""
        nn, nt = u64(0), u64(0)
        self._check(self.lib.ptc_get_wide_bvh(self.ctx, C.byref(nn), C.byref(nt), None, None), "ptc_get_wide_bvh")
        out = dict(n_nodes=int(nn.value), n_tris=int(nt.value), words=np.zeros((int(nn.value), 20), np.uint32),
                   tri_order=np.zeros(int(nt.value), np.uint32))
        if out["n_nodes"]:
            self._check(self.lib.ptc_get_wide_bvh(self.ctx, C.byref(nn), C.byref(nt), np_ptr(out["words"]), np_ptr(out["tri_order"])),
                        "ptc_get_wide_bvh")
        return out

    def sampler_points(self, px, py, width, first_index, count, dimension, flags):
        out = np.zeros((count, 2), np.float32)
        self._check(self.lib.ptc_sampler_points(self.ctx, px, py, width, first_index, count, dimension, flags, np_ptr(out)), "ptc_sampler_points")
        return out

    def bsdf_eval(self, params, wi, wo):
        params = np.ascontiguousarray(params, np.float32)
        wi = np.ascontiguousarray(wi, np.float32)
        wo = np.ascontiguousarray(wo, np.float32)
        n = params.shape[0]
        f = np.zeros((n, 3), np.float32)
        pdf = np.zeros(n, np.float32)
        self._check(self.lib.ptc_bsdf_eval(self.ctx, n, np_ptr(params), np_ptr(wi), np_ptr(wo), np_ptr(f), np_ptr(pdf)), "ptc_bsdf_eval")
        return f, pdf

    def bsdf_sample(self, params, wo, u):
        params = np.ascontiguousarray(params, np.float32)
        wo = np.ascontiguousarray(wo, np.float32)
        u = np.ascontiguousarray(u, np.float32)
        n = params.shape[0]
        wi = np.zeros((n, 3), np.float32)
        f = np.zeros((n, 3), np.float32)
        pdf = np.zeros(n, np.float32)
        self._check(self.lib.ptc_bsdf_sample(self.ctx, n, np_ptr(params), np_ptr(wo), np_ptr(u), np_ptr(wi), np_ptr(f), np_ptr(pdf)),
                    "ptc_bsdf_sample")
        return wi, f, pdf

    def env_lookup(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        out = np.zeros_like(dirs)
        self._check(self.lib.ptc_env_lookup(self.ctx, dirs.shape[0], np_ptr(dirs), np_ptr(out)), "ptc_env_lookup")
        return out

    def env_sample(self, u01):
        u01 = np.ascontiguousarray(u01, np.float32).reshape(-1, 2)
        dirs, pdf = np.zeros((u01.shape[0], 3), np.float32), np.zeros(u01.shape[0], np.float32)
        self._check(self.lib.ptc_env_sample(self.ctx, u01.shape[0], np_ptr(u01), np_ptr(dirs), np_ptr(pdf)), "ptc_env_sample")
        return dirs, pdf

    def env_pdf(self, dirs):
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        pdf = np.zeros(dirs.shape[0], np.float32)
        self._check(self.lib.ptc_env_pdf(self.ctx, dirs.shape[0], np_ptr(dirs), np_ptr(pdf)), "ptc_env_pdf")
        return pdf


class HostEngine:
    """The C++ vengine host library: scene recipes, flattening, RendererPathTracing."""

    def __init__(self, asset_root=None, devices=None):
        """devices: None = the current device; a list of indices = one engine driving several GPUs (NCCL inside the core)"""
        self.lib = load_host()
        a = (asset_root or ROOT).encode()
        self.h = C.c_void_p(self.lib.vh_engine_create(a))
        if devices is not None and len(devices) > 1:
            self.set_devices(devices)

    def set_devices(self, devices):
        ids = (C.c_int * len(devices))(*[int(d) for d in devices])
        if self.lib.vh_set_devices(self.h, ids, len(devices)) != 0:
            raise RuntimeError("cannot use devices %s: %s" % (list(devices), self.last_error()))

    def device_count(self):
        return int(self.lib.vh_device_count(self.h))

    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        if self.lib.vh_comm_unique_id(self.h, buf) != 0:
            raise RuntimeError("vh_comm_unique_id failed")
        return bytes(buf)

    def comm_init_rank(self, id_bytes, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(id_bytes)
        if self.lib.vh_comm_init_rank(self.h, buf, int(rank), int(world)) != 0:
            raise RuntimeError("vh_comm_init_rank failed: %s" % self.last_error())

    def set_render_options(self, split=None, sampler=None, env_importance=None):
        """split: None | "default" | "tile" | "sample"; sampler: None | "default" | "sobol" | "pmj" """
        sp = -1 if split is None else {"default": 0, "tile": 1, "sample": 2}[split]
        sm = -1 if sampler is None else {"default": 0, "sobol": 1, "pmj": 2}[sampler]
        ei = -1 if env_importance is None else int(bool(env_importance))
        self.lib.vh_set_render_options(self.h, sp, sm, ei)

    def close(self):
        if self.h:
            self.lib.vh_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def backend_ok(self):
        return bool(self.lib.vh_backend_ok(self.h))

    def last_error(self):
        return self.lib.vh_last_error(self.h).decode()

    def scene_list(self):
        return self.lib.vh_scene_list().decode().split(",")

    def build_scene(self, name, texture_size=0, scale=0.0, camera=0):
        rc = self.lib.vh_build_scene(self.h, name.encode(), int(texture_size), float(scale), int(camera))
        if rc == 3:
            raise RuntimeError("scene recipe %r failed: %s" % (name, self.last_error()))
        if rc != 0:
            raise RuntimeError("unknown scene recipe %r" % name)

    def import_model(self, path, import_materials=True):
        """Engine::importModel: .obj (+ .mtl), .gltf, .glb"""
        if self.lib.vh_import_model(self.h, path.encode(), int(bool(import_materials))) != 0:
            raise RuntimeError(self.last_error())

    def add_model(self, model_name):
        """addModel3D under the scene root + Scene::update"""
        if self.lib.vh_add_model(self.h, model_name.encode()) != 0:
            raise RuntimeError(self.last_error())

    def import_scene(self, scene_json):
        if self.lib.vh_import_scene(self.h, scene_json.encode()) != 0:
            raise RuntimeError(self.last_error())

    def export_scene(self, directory):
        if self.lib.vh_export_scene(self.h, directory.encode()) != 0:
            raise RuntimeError(self.last_error())

    def describe(self):
        import json
        return json.loads(self.lib.vh_describe(self.h).decode())

    def set_render_info(self, width=0, height=0, samples=0, batch_size=0, depth=0):
        self.lib.vh_set_render_info(self.h, width, height, samples, batch_size, depth)

    def render_info(self):
        v = [C.c_int() for _ in range(5)]
        self.lib.vh_get_render_info(self.h, *[C.byref(x) for x in v])
        return dict(zip(["width", "height", "samples", "batch_size", "depth"], [x.value for x in v]))

    def scene_desc(self):
        return self.lib.vh_scene_desc(self.h)

    def render_params(self):
        p = ptc_render_params()
        self.lib.vh_render_params(self.h, C.byref(p))
        return p

    def render_to_memory(self, out=None):
        """RendererPathTracing::render() into host arrays; `out` = three reusable float32 arrays of width*height*4"""
        ri = self.render_info()
        n = ri["width"] * ri["height"] * 4
        if out is not None:
            rad, alb, nrm = (o.reshape(-1) for o in out)
            assert all(o.size == n and o.dtype == np.float32 and o.flags.c_contiguous for o in (rad, alb, nrm))
        else:
            rad, alb, nrm = (np.empty(n, np.float32) for _ in range(3))
        rc = self.lib.vh_render_to_memory(self.h, np_ptr(rad), np_ptr(alb), np_ptr(nrm))
        if rc != 0:
            raise RuntimeError("RendererPathTracing::render failed: %s" % self.last_error())
        shape = (ri["height"], ri["width"], 4)
        return rad.reshape(shape), alb.reshape(shape), nrm.reshape(shape)

    def set_sequence_frame(self, frame):
        if self.lib.vh_set_sequence_frame(self.h, int(frame)) != 0:
            raise RuntimeError("no scene / camera")

    def set_output(self, file_type=None, exposure=0.0, write_all_files=None, denoise=None):
        """file_type: None | "hdr" | "png" """
        ft = -1 if file_type is None else {"hdr": 0, "png": 1}[file_type]
        self.lib.vh_set_output(self.h, ft, float(exposure), -1 if write_all_files is None else int(write_all_files),
                               -1 if denoise is None else int(denoise))

    def render_progress(self):
        """RendererPathTracing::renderProgress(); safe to call from another thread while a render runs"""
        return float(self.lib.vh_render_progress(self.h))

    def render(self, filename):
        rc = self.lib.vh_render(self.h, filename.encode())
        if rc != 0:
            raise RuntimeError("RendererPathTracing::render failed: %s" % self.last_error())

    def stats(self):
        s = ptc_stats()
        self.lib.vh_get_stats(self.h, C.byref(s))
        return s.as_dict()


def decode_image(data, flip=False):
    """PNG / JPEG bytes -> (uint8 array H x W x C, channel count of the file) through the host library's own decoders"""
    lib = load_host()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    w, h, c, sc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    if lib.vh_decode_image(buf, len(data), int(flip), C.byref(w), C.byref(h), C.byref(c), C.byref(sc), None, 0) != 0:
        raise RuntimeError("cannot decode image")
    out = np.zeros((h.value, w.value, c.value), np.uint8)
    if lib.vh_decode_image(buf, len(data), int(flip), None, None, None, None, np_ptr(out), out.size) != 0:
        raise RuntimeError("cannot decode image")
    return out, sc.value


def read_hdr(path):
    lib = load_host()
    w, h = C.c_int(), C.c_int()
    if lib.vh_read_hdr(path.encode(), C.byref(w), C.byref(h), None) != 0:
        raise RuntimeError("cannot read %s" % path)
    out = np.zeros((h.value, w.value, 4), np.float32)
    lib.vh_read_hdr(path.encode(), C.byref(w), C.byref(h), np_ptr(out))
    return out


def write_image(path_no_ext, img, file_type="png", exposure=0.0):
    """writeToDisk of the host library (+ applyExposure for PNG): <path>.png / <path>.hdr"""
    lib = load_host()
    img = np.ascontiguousarray(img, np.float32)
    h, w, c = img.shape
    if lib.vh_write_image(path_no_ext.encode(), w, h, c, np_ptr(img), {"hdr": 0, "png": 1}[file_type], float(exposure)) != 0:
        raise RuntimeError("cannot write %s" % path_no_ext)


def write_hdr(path, img):
    lib = load_host()
    img = np.ascontiguousarray(img, np.float32)
    h, w, c = img.shape
    if lib.vh_write_hdr(path.encode(), w, h, c, np_ptr(img)) != 0:
        raise RuntimeError("cannot write %s" % path)
