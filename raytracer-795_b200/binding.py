"""ctypes binding of include/rt795.h (the C ABI).

This is the thin test/bench driver over the shared library; the product is the C ABI itself.  Struct layouts
here must match include/rt795.h field for field (tests/test_abi.py checks sizes against the compiled library).
There is no CPU fallback: `load_core()` raises if librt795.so is missing, and rendering raises if CUDA is unusable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CORE_SO = os.path.join(HERE, "lib", "librt795.so")

RT_OK = 0
(RT_ERR_INVALID, RT_ERR_CUDA, RT_ERR_OOM, RT_ERR_IO, RT_ERR_PARSE, RT_ERR_OVERFLOW) = (-1, -2, -3, -4, -5, -6)
RT_INTEGRATOR_REFERENCE, RT_INTEGRATOR_PATH = 0, 1
RT_BVH_AUTO, RT_BVH_HOST_SAH, RT_BVH_DEVICE_PLOC, RT_BVH_DEVICE_SAH = 0, 1, 2, 3
RT_DEBUG_BVH2, RT_DEBUG_BVH4, RT_DEBUG_BVH4Q, RT_DEBUG_TRIS, RT_DEBUG_OBJ_ROOT4 = 0, 1, 2, 3, 4
RT_OBJ_SPHERE, RT_OBJ_MESH = 0, 1
RT_IMG_RGBA8, RT_IMG_RGBA32F = 0, 1

f3 = C.c_float * 3


class RtMaterial(C.Structure):
    _fields_ = [("type", C.c_int32), ("brdf", C.c_int32), ("phong_exp", C.c_int32), ("is_rough", C.c_int32),
                ("ambient", f3), ("diffuse", f3), ("specular", f3), ("mirror", f3),
                ("refraction_index", C.c_float), ("absorption_index", C.c_float), ("roughness", C.c_float),
                ("absorption_coeff", f3), ("emission", f3)]


class RtTexture(C.Structure):
    _fields_ = [("type", C.c_int32), ("decal_mode", C.c_int32), ("interpolation", C.c_int32),
                ("noise_conversion", C.c_int32), ("normalizer", C.c_int32), ("noise_scale", C.c_float),
                ("bump_factor", C.c_float), ("image", C.c_int32)]


class RtImage(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("format", C.c_int32), ("pixels", C.c_void_p)]


class RtMesh(C.Structure):
    _fields_ = [("first_tri", C.c_int32), ("tri_count", C.c_int32), ("smooth", C.c_int32)]


class RtSphere(C.Structure):
    _fields_ = [("center", f3), ("radius", C.c_float)]


class RtObject(C.Structure):
    _fields_ = [("kind", C.c_int32), ("geom", C.c_int32), ("mat_index", C.c_int32), ("tex", C.c_int32 * 2),
                ("tex_offset", C.c_int32), ("inv", C.c_float * 12), ("inv_t", C.c_float * 9), ("blur", f3), ("radiance", f3)]


class RtLight(C.Structure):
    _fields_ = [("type", C.c_int32), ("position", f3), ("direction", f3), ("intensity", f3), ("size", C.c_float),
                ("coverage_deg", C.c_float), ("falloff_deg", C.c_float), ("image", C.c_int32), ("object", C.c_int32)]


class RtSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("max_recursion_depth", C.c_int32), ("shadow_ray_eps", C.c_float),
                ("int_test_eps", C.c_float), ("background_color", f3), ("ambient_light", f3),
                ("env_light_index", C.c_int32), ("background_texture", C.c_int32),
                ("num_vertices", C.c_int32), ("vertices", C.POINTER(C.c_float)), ("vertex_normals", C.POINTER(C.c_float)),
                ("num_texcoords", C.c_int32), ("texcoords", C.POINTER(C.c_float)),
                ("num_triangles", C.c_int32), ("tri_indices", C.POINTER(C.c_int32)),
                ("num_meshes", C.c_int32), ("meshes", C.POINTER(RtMesh)),
                ("num_spheres", C.c_int32), ("spheres", C.POINTER(RtSphere)),
                ("num_objects", C.c_int32), ("objects", C.POINTER(RtObject)),
                ("num_materials", C.c_int32), ("materials", C.POINTER(RtMaterial)),
                ("num_textures", C.c_int32), ("textures", C.POINTER(RtTexture)),
                ("num_images", C.c_int32), ("images", C.POINTER(RtImage)),
                ("num_lights", C.c_int32), ("lights", C.POINTER(RtLight)), ("bvh_builder", C.c_int32)]


class RtCameraDesc(C.Structure):
    _fields_ = [("position", f3), ("gaze", f3), ("up", f3), ("left", C.c_float), ("right", C.c_float),
                ("bottom", C.c_float), ("top", C.c_float), ("distance", C.c_float), ("nx", C.c_int32), ("ny", C.c_int32),
                ("num_samples", C.c_int32), ("focus_distance", C.c_float), ("aperture_size", C.c_float),
                ("is_dof", C.c_int32), ("left_handed", C.c_int32)]


class RtRenderOpts(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("seed", C.c_uint64), ("tile_rank", C.c_int32), ("tile_count", C.c_int32),
                ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("pt_max_depth", C.c_int32), ("pt_nee", C.c_int32),
                ("pt_russian_roulette", C.c_int32), ("pt_importance", C.c_int32), ("wave_rays", C.c_int32),
                ("queue_capacity", C.c_int32), ("collect_counters", C.c_int32), ("resolve_owned_only", C.c_int32), ("parity_exact", C.c_int32)]


class RtTonemap(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("key", C.c_float), ("burn_percent", C.c_float), ("saturation", C.c_float), ("gamma", C.c_float)]


class RtHit(C.Structure):
    _fields_ = [("full", C.c_int32), ("object", C.c_int32), ("primitive", C.c_int32), ("t", C.c_float),
                ("point", f3), ("normal", f3)]


class RtStats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_secondary", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("launches", C.c_uint64), ("waves", C.c_uint64), ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64),
                ("node_visits_shadow", C.c_uint64), ("prim_tests_shadow", C.c_uint64), ("diag", C.c_uint64 * 8),
                ("ms_total", C.c_double), ("ms_extend", C.c_double), ("ms_shade", C.c_double), ("ms_shadow", C.c_double),
                ("ms_raygen", C.c_double), ("ms_resolve", C.c_double),
                ("launches_extend", C.c_uint64), ("launches_shade", C.c_uint64), ("launches_shadow", C.c_uint64),
                ("bvh_nodes", C.c_uint64), ("bvh_bytes", C.c_uint64), ("tri_bytes", C.c_uint64),
                ("ms_bvh_build", C.c_double), ("ms_upload", C.c_double), ("ms_bvh_device", C.c_double),
                ("bvh_device_meshes", C.c_int32), ("bvh_ploc_iterations", C.c_int32), ("bvh2_nodes", C.c_uint64)]


HIT_DTYPE = np.dtype([("full", "<i4"), ("object", "<i4"), ("primitive", "<i4"), ("t", "<f4"),
                      ("point", "<f4", (3,)), ("normal", "<f4", (3,))])

# every symbol include/rt795.h declares (tests/test_abi.py asserts the library exports all of them)
CORE_SYMBOLS = ["rt_scene_create", "rt_scene_destroy", "rt_render", "rt_render_device", "rt_trace_primary", "rt_tonemap",
                "rt_get_stats", "rt_group_create", "rt_group_destroy", "rt_group_render", "rt_group_num_devices", "rt_group_scene",
                "rt_group_uses_peer_memory", "rt_host_alloc_pinned", "rt_host_free_pinned", "rt_debug_read_bvh", "rt_last_error", "rt_abi_version", "rt_device_count", "rt_host_load_scene",
                "rt_host_save_scene_blob", "rt_host_load_scene_blob", "rt_host_load_scene_cached",
                "rt_host_camera_render_opts", "rt_host_camera_tonemap",
                "rt_host_free_scene", "rt_host_scene_desc", "rt_host_num_cameras", "rt_host_camera",
                "rt_host_camera_image_name", "rt_host_save_image", "rt_host_save_f32"]


class RtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rt795 error {code}: {msg}")
        self.code = code


_core = None


def load_core() -> C.CDLL:
    """Loads librt795.so.  Fails loudly if the CUDA extension has not been built (no fallback)."""
    global _core
    if _core is None:
        if not os.path.exists(CORE_SO):
            raise RuntimeError(f"{CORE_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(make -C raytracer-795_b200). There is no CPU fallback.")
        lib = C.CDLL(os.environ.get("RT795_CORE_SO", CORE_SO))      # RT795_CORE_SO: A/B builds of the same library (profiles/ab_*.sh)
        P = C.POINTER
        lib.rt_scene_create.argtypes = [P(RtSceneDesc), C.c_int, P(C.c_void_p)]
        lib.rt_scene_destroy.argtypes = [C.c_void_p]
        lib.rt_scene_destroy.restype = None
        lib.rt_render.argtypes = [C.c_void_p, P(RtCameraDesc), P(RtRenderOpts), C.c_void_p]
        lib.rt_render_device.argtypes = [C.c_void_p, P(RtCameraDesc), P(RtRenderOpts), C.c_void_p, C.c_void_p]
        lib.rt_trace_primary.argtypes = [C.c_void_p, P(RtCameraDesc), C.c_void_p]
        lib.rt_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        lib.rt_get_stats.argtypes = [C.c_void_p, P(RtStats)]
        lib.rt_group_create.argtypes = [P(RtSceneDesc), C.c_int, P(C.c_int), P(C.c_void_p)]
        lib.rt_group_destroy.argtypes = [C.c_void_p]
        lib.rt_group_destroy.restype = None
        lib.rt_group_render.argtypes = [C.c_void_p, P(RtCameraDesc), P(RtRenderOpts), C.c_void_p]
        lib.rt_group_num_devices.argtypes = [C.c_void_p]
        lib.rt_group_scene.argtypes = [C.c_void_p, C.c_int]
        lib.rt_group_scene.restype = C.c_void_p
        lib.rt_group_uses_peer_memory.argtypes = [C.c_void_p, C.c_int]
        lib.rt_debug_read_bvh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, P(C.c_uint64), P(C.c_int32)]
        lib.rt_last_error.restype = C.c_char_p
        lib.rt_host_load_scene.argtypes = [C.c_char_p, P(C.c_void_p)]
        lib.rt_host_save_scene_blob.argtypes = [C.c_void_p, C.c_char_p]
        lib.rt_host_load_scene_blob.argtypes = [C.c_char_p, C.c_int, P(C.c_void_p)]
        lib.rt_host_load_scene_cached.argtypes = [C.c_char_p, C.c_char_p, P(C.c_void_p), P(C.c_int)]
        lib.rt_host_free_scene.argtypes = [C.c_void_p]
        lib.rt_host_free_scene.restype = None
        lib.rt_host_scene_desc.argtypes = [C.c_void_p]
        lib.rt_host_scene_desc.restype = P(RtSceneDesc)
        lib.rt_host_num_cameras.argtypes = [C.c_void_p]
        lib.rt_host_camera.argtypes = [C.c_void_p, C.c_int]
        lib.rt_host_camera.restype = P(RtCameraDesc)
        lib.rt_host_camera_render_opts.argtypes = [C.c_void_p, C.c_int, P(RtRenderOpts)]
        lib.rt_host_camera_tonemap.argtypes = [C.c_void_p, C.c_int, P(RtTonemap)]
        lib.rt_host_camera_image_name.argtypes = [C.c_void_p, C.c_int]
        lib.rt_host_camera_image_name.restype = C.c_char_p
        lib.rt_host_save_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        lib.rt_host_save_f32.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        _core = lib
    return _core


def _check(code: int) -> None:
    if code != RT_OK:
        raise RtError(code, load_core().rt_last_error().decode(errors="replace"))


def default_opts(**kw) -> RtRenderOpts:
    o = RtRenderOpts()
    o.integrator = RT_INTEGRATOR_REFERENCE
    o.seed = 795
    o.tile_rank, o.tile_count = 0, 1
    o.pt_nee, o.pt_russian_roulette, o.pt_importance = 1, 1, 1
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class HostScene:
    """rt_host_load_scene: XML (+PLY/PNG/EXR) -> RtSceneDesc, the reference's Scene ctor + renderScene prologue."""

    def __init__(self, xml_path: Optional[str] = None, *, blob: Optional[str] = None, check_sources: bool = True, cached: bool = False):
        """HostScene(xml): parse.  HostScene(blob=path): rt_host_load_scene_blob.  HostScene(xml, cached=True[, blob=path]):
        rt_host_load_scene_cached (self.cache_hit tells whether the blob was used)."""
        self.lib = load_core()
        self.h = C.c_void_p()
        self.cache_hit = False
        if cached:
            hit = C.c_int(0)
            _check(self.lib.rt_host_load_scene_cached(os.fsencode(xml_path), os.fsencode(blob) if blob else None, C.byref(self.h), C.byref(hit)))
            self.cache_hit = bool(hit.value)
        elif blob is not None:
            _check(self.lib.rt_host_load_scene_blob(os.fsencode(blob), 1 if check_sources else 0, C.byref(self.h)))
        else:
            _check(self.lib.rt_host_load_scene(os.fsencode(xml_path), C.byref(self.h)))
        self.desc = self.lib.rt_host_scene_desc(self.h)

    def save_blob(self, path: str) -> None:
        _check(self.lib.rt_host_save_scene_blob(self.h, os.fsencode(path)))

    @property
    def num_cameras(self) -> int:
        return self.lib.rt_host_num_cameras(self.h)

    def camera(self, i: int = 0) -> RtCameraDesc:
        return self.lib.rt_host_camera(self.h, i).contents

    def render_opts(self, i: int = 0) -> RtRenderOpts:
        o = RtRenderOpts()
        _check(self.lib.rt_host_camera_render_opts(self.h, i, C.byref(o)))
        return o

    def tonemap(self, i: int = 0) -> "RtTonemap":
        t = RtTonemap()
        _check(self.lib.rt_host_camera_tonemap(self.h, i, C.byref(t)))
        return t

    def image_name(self, i: int = 0) -> str:
        return self.lib.rt_host_camera_image_name(self.h, i).decode()

    def close(self) -> None:
        if self.h:
            self.lib.rt_host_free_scene(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """rt_scene_create / rt_render / rt_trace_primary on one CUDA device."""

    def __init__(self, desc, device: int = 0, bvh_builder: Optional[int] = None):
        self.lib = load_core()
        self.h = C.c_void_p()
        if bvh_builder is not None:                               # same scene, another RT_BVH_* builder: shallow copy of the descriptor
            src = desc.contents if hasattr(desc, "contents") else desc
            d = RtSceneDesc()
            C.memmove(C.byref(d), C.byref(src), C.sizeof(RtSceneDesc))
            d.bvh_builder = bvh_builder
            self._desc = d
            desc = C.byref(d)
        _check(self.lib.rt_scene_create(desc, device, C.byref(self.h)))

    def render(self, cam: RtCameraDesc, opts: Optional[RtRenderOpts] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        opts = opts or default_opts()
        if out is None:
            out = np.empty((cam.ny, cam.nx, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == cam.nx * cam.ny * 3
        _check(self.lib.rt_render(self.h, C.byref(cam), C.byref(opts), out.ctypes.data))
        return out

    def render_device(self, cam: RtCameraDesc, opts: RtRenderOpts, d_ptr: int, stream: Optional[int] = None) -> None:
        """stream: a cudaStream_t handle; None = the library's own stream.  0 is the CUDA legacy default stream (what
        torch.cuda.current_stream().cuda_stream is unless a side stream is current): it is passed as cudaStreamLegacy so that the
        frame really is ordered with the caller's work on that stream (barriers, copies) - a NULL handle means "library stream"
        to rt_render_device."""
        CUDA_STREAM_LEGACY = 1
        h = 0 if stream is None else (stream if stream else CUDA_STREAM_LEGACY)
        _check(self.lib.rt_render_device(self.h, C.byref(cam), C.byref(opts), C.c_void_p(d_ptr), C.c_void_p(h)))

    def trace_primary(self, cam: RtCameraDesc) -> np.ndarray:
        out = np.zeros((cam.ny, cam.nx), dtype=HIT_DTYPE)
        _check(self.lib.rt_trace_primary(self.h, C.byref(cam), out.ctypes.data))
        return out

    def tonemap(self, rgb: np.ndarray, key: float = 0.18, burn: float = 1.0, saturation: float = 1.0, gamma: float = 2.2) -> np.ndarray:
        ny, nx, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        out = np.empty((ny, nx, 3), dtype=np.uint8)
        _check(self.lib.rt_tonemap(self.h, rgb.ctypes.data, nx, ny, key, burn, saturation, gamma, out.ctypes.data))
        return out

    def read_bvh(self, what: int):
        """(array, tlas_root): one acceleration-structure array as float32 (nodes, triangles) / uint32 (quantised nodes) / int32 rows."""
        size, root = C.c_uint64(0), C.c_int32(0)
        _check(self.lib.rt_debug_read_bvh(self.h, what, None, 0, C.byref(size), C.byref(root)))
        buf = np.empty(size.value // 4, dtype=np.uint32)
        _check(self.lib.rt_debug_read_bvh(self.h, what, buf.ctypes.data, buf.nbytes, C.byref(size), C.byref(root)))
        row = {RT_DEBUG_BVH2: 16, RT_DEBUG_BVH4: 32, RT_DEBUG_BVH4Q: 16, RT_DEBUG_TRIS: 12, RT_DEBUG_OBJ_ROOT4: 1}[what]
        return buf.reshape(-1, row), int(root.value)

    def stats(self) -> RtStats:
        s = RtStats()
        _check(self.lib.rt_get_stats(self.h, C.byref(s)))
        return s

    def close(self) -> None:
        if self.h:
            self.lib.rt_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SceneGroup:
    """rt_group_create / rt_group_render: the scene replicated on several CUDA devices of this process, one frame split into
    interleaved tiles, gathered on devices[0] over peer memory."""

    def __init__(self, desc, devices):
        self.lib = load_core()
        self.h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        _check(self.lib.rt_group_create(desc, len(devices), devs, C.byref(self.h)))
        self.devices = list(devices)

    def render(self, cam: RtCameraDesc, opts: Optional[RtRenderOpts] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        opts = opts or default_opts()
        if out is None:
            out = np.empty((cam.ny, cam.nx, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == cam.nx * cam.ny * 3
        _check(self.lib.rt_group_render(self.h, C.byref(cam), C.byref(opts), out.ctypes.data))
        return out

    def stats(self, index: int = 0) -> RtStats:
        s = RtStats()
        _check(self.lib.rt_get_stats(self.lib.rt_group_scene(self.h, index), C.byref(s)))
        return s

    def uses_peer_memory(self, index: int) -> bool:
        return bool(self.lib.rt_group_uses_peer_memory(self.h, index))

    def close(self) -> None:
        if self.h:
            self.lib.rt_group_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_f32(path: str) -> np.ndarray:
    """Reads a raw float32 frame written by rt_host_save_f32 (int32 nx, ny, then ny*nx*3 floats)."""
    raw = np.fromfile(path, dtype=np.uint8)
    nx, ny = np.frombuffer(raw[:8], dtype="<i4")
    return np.frombuffer(raw[8:], dtype="<f4").reshape(ny, nx, 3).copy()
