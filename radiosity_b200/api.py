"""ctypes bindings of include/rad_cuda.h (librad_cuda.so) and of the host library's C view
(libradiosity_host.so).  Thin by design: numpy arrays in, numpy arrays out, every call goes straight
through the C ABI — this is the path the `-m gpu` parity tests and bench.py's e2e leg measure."""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SELECT_REFERENCE, SELECT_TOPK = 0, 1
FLAG_KEEP_ITEMBUFFER = 1
FACE_NAMES = ("UP", "DOWN", "LEFT", "RIGHT", "FRONT")      # atlas order, Main.h:210-211
# Camera::PatchLook values (Camera.h:18-24) of the atlas faces above
FACE_TO_LOOK = (1, 2, 3, 4, 0)


class RadError(RuntimeError):
    pass


class RadConfig(ctypes.Structure):
    _fields_ = [("hemicube_side", ctypes.c_uint32), ("hemicubes", ctypes.c_uint32), ("max_patches", ctypes.c_uint32),
                ("device", ctypes.c_int32), ("select_mode", ctypes.c_uint32), ("reflectivity", ctypes.c_float),
                ("projection", ctypes.c_float * 16), ("flags", ctypes.c_uint32)]


class RadStats(ctypes.Structure):
    _fields_ = [("batches_done", ctypes.c_uint32), ("shots_done", ctypes.c_uint32), ("stopped", ctypes.c_uint32),
                ("last_energy_len", ctypes.c_float), ("gpu_ms", ctypes.c_float), ("kernel_launches", ctypes.c_uint32),
                ("big_triangles", ctypes.c_uint32), ("queue_overflow", ctypes.c_uint32)]


_vp, _u32, _i32, _f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32, ctypes.c_float
_P = ctypes.POINTER

# name -> (restype, argtypes); the list mirrors include/rad_cuda.h one to one (tests/test_abi.py checks it)
CUDA_SIGNATURES = {
    "rad_create": (ctypes.c_int, [_P(_vp), _P(RadConfig)]),
    "rad_destroy": (ctypes.c_int, [_vp]),
    "rad_last_error": (ctypes.c_char_p, [_vp]),
    "rad_set_formfactors": (ctypes.c_int, [_vp, _vp, _u32]),
    "rad_upload_scene": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _u32]),
    "rad_upload_state": (ctypes.c_int, [_vp, _vp, _vp]),
    "rad_download_state": (ctypes.c_int, [_vp, _vp, _vp]),
    "rad_select": (ctypes.c_int, [_vp, _vp, _vp]),
    "rad_set_emitters": (ctypes.c_int, [_vp, _vp, _u32]),
    "rad_render_hemicubes": (ctypes.c_int, [_vp]),
    "rad_process_hemicubes": (ctypes.c_int, [_vp]),
    "rad_apply": (ctypes.c_int, [_vp, _P(_f32)]),
    "rad_shoot": (ctypes.c_int, [_vp, _u32, ctypes.c_int, _P(RadStats)]),
    "rad_save_state": (ctypes.c_int, [_vp]),
    "rad_restore_state": (ctypes.c_int, [_vp]),
    "rad_upload_neighbours": (ctypes.c_int, [_vp, _vp, _u32]),
    "rad_shade_vertices": (ctypes.c_int, [_vp, _vp, _P(_f32)]),
    "rad_read_itembuffer": (ctypes.c_int, [_vp, _u32, _vp]),
    "rad_read_depthbuffer": (ctypes.c_int, [_vp, _u32, _vp]),
    "rad_read_formfactors": (ctypes.c_int, [_vp, _u32, _vp]),
    "rad_read_mvp": (ctypes.c_int, [_vp, _u32, _u32, _vp]),
    "rad_write_itembuffer": (ctypes.c_int, [_vp, _u32, _vp]),
    "rad_bench_process": (ctypes.c_int, [_vp, _u32, _P(_f32)]),
    "rad_bench_atomics": (ctypes.c_int, [_vp, _u32, ctypes.c_uint64, _P(_f32)]),
    "rad_profile_batch": (ctypes.c_int, [_vp, _vp]),
    "rad_nccl_unique_id": (ctypes.c_int, [_vp]),
    "rad_comm_init": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "rad_set_partition": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "rad_peer_handle": (ctypes.c_int, [_vp, _vp]),
    "rad_peer_init": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "rad_batch_partial": (ctypes.c_int, [_vp]),
    "rad_read_delta": (ctypes.c_int, [_vp, _vp]),
    "rad_write_delta": (ctypes.c_int, [_vp, _vp]),
    "rad_batch_finish": (ctypes.c_int, [_vp, _P(_f32)]),
    "rad_patch_count": (_u32, [_vp]),
    "rad_atlas_width": (_u32, [_vp]),
    "rad_atlas_height": (_u32, [_vp]),
    "rad_version": (ctypes.c_char_p, []),
}

_cuda = None
_host = None


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no Python/CPU fallback for the compute path)")
    return ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)


def cuda_lib():
    """librad_cuda.so with typed signatures."""
    global _cuda
    if _cuda is None:
        lib = _load(os.environ.get("RAD_CUDA_LIB", "librad_cuda.so"))   # (RAD_CUDA_LIB: A/B builds of the same library, scripts/ only)
        for name, (res, args) in CUDA_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _cuda = lib
    return _cuda


def host_lib():
    """libradiosity_host.so (C view of the C++ host API)."""
    global _host
    if _host is None:
        cuda_lib()
        lib = _load("libradiosity_host.so")
        lib.radhost_scene_new.restype = _vp
        for f in ("radhost_scene_free", "radhost_scene_load_cornell"):
            getattr(lib, f).argtypes = [_vp]
        lib.radhost_scene_load_obj.argtypes = [_vp, ctypes.c_char_p]
        lib.radhost_scene_set_area.argtypes = [_vp, ctypes.c_double]
        lib.radhost_scene_patch_count.argtypes = [_vp]; lib.radhost_scene_patch_count.restype = _u32
        lib.radhost_scene_get.argtypes = [_vp] * 6
        lib.radhost_scene_set_state.argtypes = [_vp] * 3
        lib.radhost_scene_neighbours.argtypes = [_vp, _vp]
        lib.radhost_scene_select.argtypes = [_vp, _u32, _vp, _vp]
        lib.radhost_scene_select_single.argtypes = [_vp]; lib.radhost_scene_select_single.restype = _u32
        lib.radhost_patch_geom.argtypes = [_vp, _u32, _vp, _vp, _vp]
        lib.radhost_mvp.argtypes = [_vp, _u32, ctypes.c_int, _vp]
        lib.radhost_projection.argtypes = [_vp]
        lib.radhost_config.argtypes = [_u32, _u32, _u32, ctypes.c_double, _vp]
        lib.radhost_formfactors.argtypes = [_u32, _u32, _vp]
        lib.radhost_solver_new.argtypes = [_vp, ctypes.c_int, _u32, _u32, ctypes.c_char_p, _u32]; lib.radhost_solver_new.restype = _vp
        lib.radhost_solver_free.argtypes = [_vp]
        lib.radhost_solver_shoot.argtypes = [_vp, _u32, ctypes.c_int, _P(RadStats)]
        lib.radhost_solver_sync_to_scene.argtypes = [_vp]
        lib.radhost_solver_sync_from_scene.argtypes = [_vp]
        lib.radhost_solver_ctx.argtypes = [_vp]; lib.radhost_solver_ctx.restype = _vp
        lib.radhost_solver_error.argtypes = [_vp]; lib.radhost_solver_error.restype = ctypes.c_char_p
        lib.radhost_solver_pass_counter.argtypes = [_vp]; lib.radhost_solver_pass_counter.restype = _u32
        lib.radhost_solver_running.argtypes = [_vp]
        lib.radhost_scene_smooth_shade.argtypes = [_vp, _vp]
        lib.radhost_solver_shade.argtypes = [_vp, _vp]
        lib.radhost_scene_save.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int]
        lib.radhost_scene_load.argtypes = [_vp, ctypes.c_char_p]
        lib.radhost_scene_export_ply.argtypes = [_vp, _vp, ctypes.c_char_p, ctypes.c_float]
        lib.radhost_sizeof_patch.restype = _u32
        _host = lib
    return _host


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Scene:
    """ModelContainer of the host library (reference API: ModelContainer.h)."""

    def __init__(self, area=0.5, obj=None, static_mesh=None, scale=0.01, flip=False, emissive_material=-1):
        self.lib = host_lib()
        self.h = self.lib.radhost_scene_new()
        if static_mesh is not None:                              # TestModel.h-style export (StaticMeshModel.h)
            self.lib.radhost_scene_load_static_mesh.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_float, ctypes.c_int, ctypes.c_int]
            if not self.lib.radhost_scene_load_static_mesh(self.h, os.fsencode(static_mesh), float(scale), 1 if flip else 0, int(emissive_material)):
                raise RadError(f"cannot load static mesh {static_mesh}")
        elif obj is None:
            self.lib.radhost_scene_load_cornell(self.h)          # ModelContainer::load()
        elif not self.lib.radhost_scene_load_obj(self.h, os.fsencode(obj)):
            raise RadError(f"cannot load OBJ scene {obj}")
        self.lib.radhost_scene_set_area(self.h, float(area))     # scene.maxPatchArea (Main.cpp:843)
        self.P = int(self.lib.radhost_scene_patch_count(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.radhost_scene_free(self.h)
            self.h = None

    def arrays(self):
        """(verts[P,12], indices[P,6], color[P,3], radiosity[P,3], illumination[P,3])"""
        P = self.P
        v = np.zeros((P, 12), np.float32); ix = np.zeros((P, 6), np.int32)
        c = np.zeros((P, 3), np.float32); r = np.zeros((P, 3), np.float32); i = np.zeros((P, 3), np.float32)
        self.lib.radhost_scene_get(self.h, _ptr(v), _ptr(ix), _ptr(c), _ptr(r), _ptr(i))
        return v, ix, c, r, i

    def set_state(self, rad=None, illum=None):
        rad = _f32c(rad) if rad is not None else None
        illum = _f32c(illum) if illum is not None else None
        self.lib.radhost_scene_set_state(self.h, _ptr(rad), _ptr(illum))

    def neighbours(self):
        out = np.zeros((self.P, 8), np.int32)
        self.lib.radhost_scene_neighbours(self.h, _ptr(out))
        return out

    def save(self, path, fmt=0):
        """SaveToFile (.rr checkpoint): fmt 0 portable, 1 reference Win32 layout, 2 reference LP64 layout"""
        if not self.lib.radhost_scene_save(self.h, os.fsencode(path), fmt):
            raise RadError(f"cannot write {path}")

    def load(self, path):
        """LoadFromFile: replaces the scene by the checkpoint's patches (LoadingModel)"""
        if not self.lib.radhost_scene_load(self.h, os.fsencode(path)):
            raise RadError(f"cannot read {path}")
        self.P = int(self.lib.radhost_scene_patch_count(self.h))

    def smooth_shade(self):
        """Colors::smoothShadePatch for every patch on the host (reference API) -> float[P,12]"""
        out = np.zeros((self.P, 12), np.float32)
        self.lib.radhost_scene_smooth_shade(self.h, _ptr(out))
        return out

    def export_ply(self, path, colors12, exposure=1.0):
        """ExportPly (MeshExport.h): the scene's quads with the display stage's vertex colours as a binary PLY"""
        c = _f32c(colors12)
        if c.size != self.P * 12:
            raise RadError("export_ply: colours must be float[P, 12]")
        if not self.lib.radhost_scene_export_ply(self.h, _ptr(c), os.fsencode(path), float(exposure)):
            raise RadError(f"cannot write {path}")

    def select(self, count):
        """ModelContainer::getHighestRadiosityPatchesId -> (ids, is_null)"""
        ids = np.zeros(count, np.uint32); nul = np.zeros(count, np.int32)
        self.lib.radhost_scene_select(self.h, count, _ptr(ids), _ptr(nul))
        return ids, nul

    def mvp(self, patch, look):
        out = np.zeros(16, np.float32)
        self.lib.radhost_mvp(self.h, patch, look, _ptr(out))
        return out


def projection():
    out = np.zeros(16, np.float32)
    host_lib().radhost_projection(_ptr(out))
    return out


def formfactors(side, hemicubes=1):
    """precomputeHemicubeFormFactors() for Config(hemicube=side, hemicubes=hemicubes)."""
    out = np.zeros(3 * side * side * hemicubes, np.float32)
    host_lib().radhost_formfactors(side, hemicubes, _ptr(out))
    return out


class Context:
    """rad_ctx of include/rad_cuda.h."""

    def __init__(self, side, hemicubes, max_patches, device=0, select_mode=SELECT_REFERENCE, flags=0, reflectivity=0.3):
        self.lib = cuda_lib()
        cfg = RadConfig()
        cfg.hemicube_side, cfg.hemicubes, cfg.max_patches = side, hemicubes, max_patches
        cfg.device, cfg.select_mode, cfg.reflectivity, cfg.flags = device, select_mode, reflectivity, flags
        proj = projection()
        for i in range(16):
            cfg.projection[i] = float(proj[i])
        self.h = _vp()
        rc = self.lib.rad_create(ctypes.byref(self.h), ctypes.byref(cfg))
        if rc != 0:
            self.h = None
            raise RadError(f"rad_create failed ({rc}): {self.lib.rad_last_error(None).decode()}")
        self.N, self.k = side, hemicubes
        self.W, self.H = 2 * side, side + side // 2
        self.RES = self.W * self.H
        self.P = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.rad_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise RadError(f"{what} failed ({rc}): {self.lib.rad_last_error(self.h).decode()}")

    def set_formfactors(self, ff):
        ff = _f32c(ff)
        self._ck(self.lib.rad_set_formfactors(self.h, _ptr(ff), ff.size), "rad_set_formfactors")

    def upload_scene(self, verts, color, rad, illum):
        verts, color, rad, illum = _f32c(verts), _f32c(color), _f32c(rad), _f32c(illum)
        P = verts.size // 12
        self._ck(self.lib.rad_upload_scene(self.h, _ptr(verts), _ptr(color), _ptr(rad), _ptr(illum), P), "rad_upload_scene")
        self.P = P

    def upload_state(self, rad, illum):
        rad, illum = _f32c(rad), _f32c(illum)
        self._ck(self.lib.rad_upload_state(self.h, _ptr(rad), _ptr(illum)), "rad_upload_state")

    def download_state(self, out=None):
        """(radiosity[P,3], illumination[P,3]); `out` = a pair of preallocated float32 arrays (page-locked ones are
        written by the copy engine directly)"""
        if out is not None:
            rad, illum = out
            assert rad.dtype == np.float32 and illum.dtype == np.float32 and rad.size == 3 * self.P and illum.size == 3 * self.P
            assert rad.flags["C_CONTIGUOUS"] and illum.flags["C_CONTIGUOUS"]
        else:
            rad = np.zeros((self.P, 3), np.float32); illum = np.zeros((self.P, 3), np.float32)
        self._ck(self.lib.rad_download_state(self.h, _ptr(rad), _ptr(illum)), "rad_download_state")
        return rad, illum

    def select(self):
        ids = np.zeros(self.k, np.uint32); valid = np.zeros(self.k, np.uint32)
        self._ck(self.lib.rad_select(self.h, _ptr(ids), _ptr(valid)), "rad_select")
        return ids, valid

    def set_emitters(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._ck(self.lib.rad_set_emitters(self.h, _ptr(ids), ids.size), "rad_set_emitters")

    def render(self):
        self._ck(self.lib.rad_render_hemicubes(self.h), "rad_render_hemicubes")

    def process(self):
        self._ck(self.lib.rad_process_hemicubes(self.h), "rad_process_hemicubes")

    def apply(self):
        last = _f32()
        self._ck(self.lib.rad_apply(self.h, ctypes.byref(last)), "rad_apply")
        return last.value

    def shoot(self, n_batches, stop_test=False):
        st = RadStats()
        self._ck(self.lib.rad_shoot(self.h, n_batches, 1 if stop_test else 0, ctypes.byref(st)), "rad_shoot")
        return st

    def save_state(self):
        self._ck(self.lib.rad_save_state(self.h), "rad_save_state")

    def restore_state(self):
        self._ck(self.lib.rad_restore_state(self.h), "rad_restore_state")

    def read_itembuffer(self, hi=0):
        out = np.zeros((self.H, self.W), np.uint32)
        self._ck(self.lib.rad_read_itembuffer(self.h, hi, _ptr(out)), "rad_read_itembuffer")
        return out

    def write_itembuffer(self, hi, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        assert ids.size == self.RES
        self._ck(self.lib.rad_write_itembuffer(self.h, hi, _ptr(ids)), "rad_write_itembuffer")

    def read_depthbuffer(self, hi=0):
        out = np.zeros((self.H, self.W), np.uint32)
        self._ck(self.lib.rad_read_depthbuffer(self.h, hi, _ptr(out)), "rad_read_depthbuffer")
        return out

    def read_formfactors(self, hi=0):
        out = np.zeros(self.P, np.float32)
        self._ck(self.lib.rad_read_formfactors(self.h, hi, _ptr(out)), "rad_read_formfactors")
        return out

    def read_mvp(self, hi, face):
        out = np.zeros(16, np.float32)
        self._ck(self.lib.rad_read_mvp(self.h, hi, face, _ptr(out)), "rad_read_mvp")
        return out

    def upload_neighbours(self, nb8):
        nb8 = np.ascontiguousarray(nb8, dtype=np.int32)
        self._ck(self.lib.rad_upload_neighbours(self.h, _ptr(nb8), nb8.size // 8), "rad_upload_neighbours")

    def shade_vertices(self):
        """display stage on the device -> (float[P,12] vertex colours, gpu ms)"""
        out = np.zeros((self.P, 12), np.float32); ms = _f32()
        self._ck(self.lib.rad_shade_vertices(self.h, _ptr(out), ctypes.byref(ms)), "rad_shade_vertices")
        return out, ms.value

    def bench_process(self, repeat=20):
        ms = _f32()
        self._ck(self.lib.rad_bench_process(self.h, repeat, ctypes.byref(ms)), "rad_bench_process")
        return ms.value

    def bench_atomics(self, pattern=0, count=1 << 26):
        """measured RED.MIN.64 rate (1e9 atomics/s) for pattern 0 raster-like, 1 coalesced, 2 scattered"""
        g = _f32()
        self._ck(self.lib.rad_bench_atomics(self.h, pattern, count, ctypes.byref(g)), "rad_bench_atomics")
        return g.value

    def profile_batch(self):
        ms = np.zeros(6, np.float32)
        self._ck(self.lib.rad_profile_batch(self.h, _ptr(ms)), "rad_profile_batch")
        return ms

    # ---- multi-GPU ----
    @staticmethod
    def nccl_unique_id():
        buf = (ctypes.c_char * 128)()
        rc = cuda_lib().rad_nccl_unique_id(buf)
        if rc != 0:
            raise RadError(f"rad_nccl_unique_id failed ({rc}): {cuda_lib().rad_last_error(None).decode()}")
        return bytes(buf)

    def comm_init(self, rank, world, uid):
        self._ck(self.lib.rad_comm_init(self.h, rank, world, ctypes.c_char_p(uid)), "rad_comm_init")

    def peer_handle(self):
        """64-byte CUDA IPC handle of this rank's exchange buffer (fused dB exchange over peer memory)"""
        buf = ctypes.create_string_buffer(64)
        self._ck(self.lib.rad_peer_handle(self.h, buf), "rad_peer_handle")
        return buf.raw

    def peer_init(self, rank, world, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._ck(self.lib.rad_peer_init(self.h, rank, world, ctypes.c_char_p(blob)), "rad_peer_init")

    def set_partition(self, rank, world):
        self._ck(self.lib.rad_set_partition(self.h, rank, world), "rad_set_partition")

    def batch_partial(self):
        self._ck(self.lib.rad_batch_partial(self.h), "rad_batch_partial")

    def read_delta(self):
        out = np.zeros((self.P, 3), np.float32)
        self._ck(self.lib.rad_read_delta(self.h, _ptr(out)), "rad_read_delta")
        return out

    def write_delta(self, dB):
        dB = _f32c(dB)
        self._ck(self.lib.rad_write_delta(self.h, _ptr(dB)), "rad_write_delta")

    def batch_finish(self):
        last = _f32()
        self._ck(self.lib.rad_batch_finish(self.h, ctypes.byref(last)), "rad_batch_finish")
        return last.value


def context_for_scene(scene, side, hemicubes=1, **kw):
    """Context with the scene and its dFF table uploaded (what RadiositySolver::init does in C++)."""
    ctx = Context(side, hemicubes, scene.P, **kw)
    ctx.set_formfactors(formfactors(side, 1))
    v, _, c, r, i = scene.arrays()
    ctx.upload_scene(v, c, r, i)
    return ctx
