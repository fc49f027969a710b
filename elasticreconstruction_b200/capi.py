"""ctypes binding of libelastic_b200.so (the C ABI declared in include/elastic_b200.h).

This is the Python face of the product used by tests/ and bench.py; the drop-in executables under
apps/ bind the same entry points from C++.  There is no fallback path: if the shared library has not
been built (python -m elasticreconstruction_b200.build) importing a symbol raises, and creating a
handle on a machine without a CUDA device raises with the library's error text.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ER_B200_LIB") or os.path.join(HERE, "libelastic_b200.so")   # override: experiments only

_lib = None

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_u16p = C.POINTER(C.c_uint16)
c_ip = C.POINTER(C.c_int)
c_i32p = C.POINTER(C.c_int32)
c_u64p = C.POINTER(C.c_uint64)
c_szp = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); kept in one table so tests can check it against include/elastic_b200.h
SIGNATURES = {
    "er_last_error": (C.c_char_p, []),
    "er_version": (C.c_int, []),
    "er_build_arch": (C.c_char_p, []),
    "er_mat4_mul": (None, [c_dp, c_dp, c_dp]),
    "er_mat4_inverse": (None, [c_dp, c_dp]),
    "er_host_alloc": (C.c_void_p, [C.c_size_t]),
    "er_host_free": (None, [C.c_void_p]),
    "er_tsdf_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_size_t]),
    "er_tsdf_destroy": (None, [C.c_void_p]),
    "er_tsdf_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "er_tsdf_set_camera": (C.c_int, [C.c_void_p] + [C.c_float] * 5),
    "er_tsdf_set_control_grids": (C.c_int, [C.c_void_p, c_fp, C.c_int, C.c_int, C.c_float]),
    "er_tsdf_set_partition": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "er_tsdf_set_stats": (C.c_int, [C.c_void_p, C.c_int]),
    "er_tsdf_get_stats": (C.c_int, [C.c_void_p, c_u64p, c_u64p, c_u64p]),
    "er_tsdf_get_class_stats": (C.c_int, [C.c_void_p, c_u64p, c_u64p, c_u64p, c_u64p]),
    "er_tsdf_reproject": (C.c_int, [C.c_void_p, c_u16p, c_dp, c_dp, c_dp, c_dp, C.c_int]),
    "er_tsdf_integrate": (C.c_int, [C.c_void_p, c_u16p, c_dp]),
    "er_tsdf_integrate_ctr": (C.c_int, [C.c_void_p, c_u16p, c_dp, c_dp, c_dp, c_dp, C.c_int]),
    "er_tsdf_integrate_ctr_batch": (C.c_int, [C.c_void_p, c_u16p, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int]),
    "er_tsdf_integrate_ctr_device": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int]),
    "er_tsdf_profile_reproject_device": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int, c_fp]),
    "er_tsdf_inputs_consumed": (C.c_int, [C.c_void_p]),
    "er_tsdf_integrate_batch": (C.c_int, [C.c_void_p, c_u16p, c_dp, C.c_int]),
    "er_tsdf_integrate_device": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, C.c_int]),
    "er_tsdf_profile_device": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, C.c_int, c_fp, c_fp, c_fp]),
    "er_tsdf_set_fusion": (C.c_int, [C.c_void_p, C.c_int]),
    "er_tsdf_sync": (C.c_int, [C.c_void_p]),
    "er_tsdf_num_blocks": (C.c_int, [C.c_void_p, c_ip]),
    "er_tsdf_block_keys": (C.c_int, [C.c_void_p, c_ip, C.c_int]),
    "er_tsdf_read_block": (C.c_int, [C.c_void_p, C.c_int, c_fp, c_fp]),
    "er_tsdf_extract": (C.c_int, [C.c_void_p, c_fp, C.c_size_t, c_szp]),
    "er_tsdf_extract_begin": (C.c_int, [C.c_void_p, c_fp, C.c_size_t]),
    "er_tsdf_extract_wait": (C.c_int, [C.c_void_p, c_szp]),
    "er_tsdf_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "er_tsdf_load": (C.c_int, [C.c_void_p, C.c_char_p]),
    "er_tsdf_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "er_tsdf_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_fp]),
    "er_tsdf_launch_count": (C.c_uint64, [C.c_void_p]),
    "er_cloud_upload": (C.c_void_p, [C.c_int, c_fp, c_fp, C.c_size_t, C.c_double]),
    "er_cloud_free": (None, [C.c_void_p]),
    "er_cloud_size": (C.c_size_t, [C.c_void_p]),
    "er_icp_count_inliers": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, C.c_double, c_ip]),
    "er_cloud_get_meta": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "er_cloud_from_blob": (C.c_void_p, [C.c_int, C.c_void_p, C.c_void_p]),
    "er_icp_count_inliers_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_ip, c_ip, C.c_int, c_dp, C.c_double, c_ip]),
    "er_icp_correspondences_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_ip, c_ip, C.c_int, c_dp, C.c_double,
                                               C.c_double, c_i32p, C.c_size_t, c_szp, c_dp]),
    "er_icp_release": (C.c_int, [C.c_int]),
    "er_icp_align": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, C.c_double, C.c_int, C.c_double, C.c_int, c_dp, c_ip, c_dp]),
    "er_icp_align_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_ip, c_ip, C.c_int, c_dp, C.c_double, C.c_int,
                                     C.c_double, C.c_int, c_dp, c_ip, c_dp]),
    "er_icp_correspondences": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, C.c_double, C.c_double, c_i32p, C.c_size_t,
                                         c_szp, c_dp]),
    "er_icp_event_record": (C.c_int, [C.c_int, C.c_int]),
    "er_icp_event_elapsed_ms": (C.c_int, [C.c_int, C.c_int, C.c_int, c_fp]),
    "er_icp_launch_count": (C.c_uint64, [C.c_int]),
    "er_icp_sync": (C.c_int, [C.c_int]),
    "er_fragopt_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_double]),
    "er_fragopt_destroy": (None, [C.c_void_p]),
    "er_fragopt_set_cloud": (C.c_int, [C.c_void_p, C.c_int, c_fp, C.c_size_t, c_szp]),
    "er_fragopt_set_pairs": (C.c_int, [C.c_void_p, C.c_int, c_ip, c_ip, c_szp, c_i32p]),
    "er_fragopt_update_pose": (C.c_int, [C.c_void_p, c_dp]),
    "er_fragopt_update_normals": (C.c_int, [C.c_void_p, c_dp]),
    "er_fragopt_update_points_pn": (C.c_int, [C.c_void_p, c_dp]),
    "er_fragopt_get_points": (C.c_int, [C.c_void_p, C.c_int, c_fp, C.c_size_t, c_dp]),
    "er_fragopt_slac_assemble": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
    "er_fragopt_slac_solve": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
    "er_fragopt_slac_get_jj": (C.c_int, [C.c_void_p, c_dp]),
    "er_fragopt_slac_accumulate": (C.c_int, [C.c_void_p, C.c_void_p]),
    "er_fragopt_rigid_assemble": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
    "er_fragopt_nonrigid_assemble": (C.c_int, [C.c_void_p, C.c_double]),
    "er_fragopt_nonrigid_factor": (C.c_int, [C.c_void_p, C.c_size_t, c_i32p, c_i32p, c_dp]),
    "er_fragopt_nonrigid_solve": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "er_fragopt_nonrigid_get_aa": (C.c_int, [C.c_void_p, c_dp]),
    "er_fragopt_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "er_fragopt_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_fp]),
    "er_fragopt_launch_count": (C.c_uint64, [C.c_void_p]),
    "er_fragopt_sync": (C.c_int, [C.c_void_p]),
    "er_fragopt_last_assemble_stats": (C.c_int, [C.c_void_p, c_fp, c_u64p, c_u64p]),
}


class ElasticError(RuntimeError):
    pass


def lib():
    """Load libelastic_b200.so; raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ElasticError(LIB_PATH + " is missing: run `python -m elasticreconstruction_b200.build` "
                               "(the CUDA extension is the product; there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise ElasticError("libelastic_b200: rc=%d: %s" % (rc, lib().er_last_error().decode("utf-8", "replace")))


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _fp(a):
    return a.ctypes.data_as(c_fp)


def _u16p(a):
    return a.ctypes.data_as(c_u16p)


def _mat(T):
    return np.ascontiguousarray(T, dtype=np.float64).reshape(16)


def mat4_mul(a, b) -> np.ndarray:
    out = np.empty(16, dtype=np.float64)
    lib().er_mat4_mul(_dp(_mat(a)), _dp(_mat(b)), _dp(out))
    return out.reshape(4, 4)


def mat4_inverse(m) -> np.ndarray:
    out = np.empty(16, dtype=np.float64)
    lib().er_mat4_inverse(_dp(_mat(m)), _dp(out))
    return out.reshape(4, 4)


class HostBuffer:
    """Page-locked host memory from er_host_alloc (frame buffers / result buffers of the *_batch and extract calls)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().er_host_alloc(self.nbytes)
        if not self.ptr:
            raise ElasticError("er_host_alloc failed: " + lib().er_last_error().decode("utf-8", "replace"))

    def array(self, dtype, count=None):
        dt = np.dtype(dtype)
        n = self.nbytes // dt.itemsize if count is None else count
        return np.ctypeslib.as_array(C.cast(C.c_void_p(self.ptr), C.POINTER(C.c_uint8)), shape=(self.nbytes,)).view(dt)[:n]

    def close(self):
        if getattr(self, "ptr", None):
            lib().er_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TSDFVolume:
    """Device-resident counterpart of the reference's TSDFVolume (Integrate/TSDFVolume.h:14-69)."""

    def __init__(self, cols: int = 640, rows: int = 480, device: int = 0, max_blocks: int = 1024,
                 voxel: float = 0.0, trunc: float = 0.0):
        self.cols, self.rows = cols, rows
        h = lib().er_tsdf_create(device, cols, rows, voxel, 64, trunc, max_blocks)
        if not h:
            raise ElasticError("er_tsdf_create failed: " + lib().er_last_error().decode("utf-8", "replace"))
        self._h = C.c_void_p(h)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            lib().er_tsdf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def set_stream(self, cuda_stream_ptr: int):
        _check(lib().er_tsdf_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_camera(self, fx, fy, cx, cy, integration_trunc):
        _check(lib().er_tsdf_set_camera(self._h, fx, fy, cx, cy, integration_trunc))

    def set_control_grids(self, grids, res: int, length: float):
        g = np.ascontiguousarray(grids, dtype=np.float32)
        num = g.shape[0]
        assert g.size == num * (res + 1) ** 3 * 3
        _check(lib().er_tsdf_set_control_grids(self._h, _fp(g), num, res, length))
        self._grids_key = (id(grids), res, length)

    def set_partition(self, rank: int, world: int):
        _check(lib().er_tsdf_set_partition(self._h, rank, world))

    def set_stats(self, enable: bool):
        _check(lib().er_tsdf_set_stats(self._h, int(enable)))

    def set_fusion(self, frames_per_launch: int):
        _check(lib().er_tsdf_set_fusion(self._h, frames_per_launch))

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().er_tsdf_get_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        o, fl, idle, allone = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().er_tsdf_get_class_stats(self._h, C.byref(o), C.byref(fl), C.byref(idle), C.byref(allone)))
        return dict(frames=a.value, touched_blocks=b.value, updated_voxels=c.value, one_pairs=o.value, full_pairs=fl.value,
                    full_idle_pairs=idle.value, full_all_one_pairs=allone.value)

    # -- the hot path
    def integrate(self, depth, T):
        """ScaleDepth + Integrate (IntegrateApp.cpp:224-225) for one host frame; asynchronous."""
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        assert d.size == self.cols * self.rows
        _check(lib().er_tsdf_integrate(self._h, _u16p(d), _dp(_mat(T))))

    def integrate_batch(self, depth, T):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        Ts = np.ascontiguousarray(T, dtype=np.float64).reshape(-1, 16)
        n = Ts.shape[0]
        assert d.size == n * self.cols * self.rows
        _check(lib().er_tsdf_integrate_batch(self._h, _u16p(d), _dp(Ts), n))

    def integrate_ptr(self, host_ptr: int, T, n: int):
        """Frames in (pinned) host memory given by raw address (e.g. a torch pinned tensor's data_ptr)."""
        Ts = np.ascontiguousarray(T, dtype=np.float64).reshape(-1, 16)
        _check(lib().er_tsdf_integrate_batch(self._h, C.cast(C.c_void_p(host_ptr), c_u16p), _dp(Ts), n))

    def integrate_device(self, device_ptr: int, T, n: int):
        """Frames already resident in HBM (device pointer)."""
        Ts = np.ascontiguousarray(T, dtype=np.float64).reshape(-1, 16)
        assert Ts.shape[0] >= n
        _check(lib().er_tsdf_integrate_device(self._h, C.c_void_p(device_ptr), _dp(Ts), n))

    def profile_device(self, device_ptr: int, T, n: int):
        """Per-kernel device time over n resident frames: (touch, classify, integrate) ms, event-bracketed."""
        Ts = np.ascontiguousarray(T, dtype=np.float64).reshape(-1, 16)
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(lib().er_tsdf_profile_device(self._h, C.c_void_p(device_ptr), _dp(Ts), n, C.byref(a), C.byref(b),
                                            C.byref(c)))
        return a.value, b.value, c.value

    def reproject(self, depth, Ti, T0, A0, Ai, grid_or_chunk, res=None, length=None):
        """CIntegrateApp::Reproject on one frame; returns the new depth image.

        `grid_or_chunk` is either the chunk index into the grids given to set_control_grids(), or (for
        signature compatibility with the oracle) one lattice array, which is then uploaded as chunk 0."""
        d = np.array(depth, dtype=np.uint16, order="C", copy=True)
        chunk = grid_or_chunk
        if not isinstance(grid_or_chunk, (int, np.integer)):
            g = np.ascontiguousarray(grid_or_chunk, dtype=np.float32)
            self.set_control_grids(g.reshape(1, -1, 3), res, length)
            chunk = 0
        _check(lib().er_tsdf_reproject(self._h, _u16p(d), _dp(_mat(Ti)), _dp(_mat(T0)), _dp(_mat(A0)),
                                       _dp(_mat(Ai)), int(chunk)))
        return d

    def integrate_ctr(self, depth, Ti, T0, A0, Ai, chunk: int):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        _check(lib().er_tsdf_integrate_ctr(self._h, _u16p(d), _dp(_mat(Ti)), _dp(_mat(T0)), _dp(_mat(A0)),
                                           _dp(_mat(Ai)), int(chunk)))

    def integrate_ctr_batch(self, depth, Ti, T0, A0, Ai, chunks):
        """Reproject + ScaleDepth + Integrate for n frames in batches (IntegrateApp.cpp:217-225 per frame)."""
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        Ti = np.ascontiguousarray(Ti, dtype=np.float64).reshape(-1, 16)
        Ai = np.ascontiguousarray(Ai, dtype=np.float64).reshape(-1, 16)
        ch = np.ascontiguousarray(chunks, dtype=np.int32)
        n = len(ch)
        assert Ti.shape[0] == n and Ai.shape[0] == n and d.size == n * self.cols * self.rows
        _check(lib().er_tsdf_integrate_ctr_batch(self._h, _u16p(d), _dp(Ti), _dp(_mat(T0)), _dp(_mat(A0)), _dp(Ai),
                                                 ch.ctypes.data_as(c_ip), n))

    def _ctr_args(self, Ti, T0, A0, Ai, chunks):
        Ti = np.ascontiguousarray(Ti, dtype=np.float64).reshape(-1, 16)
        Ai = np.ascontiguousarray(Ai, dtype=np.float64).reshape(-1, 16)
        ch = np.ascontiguousarray(chunks, dtype=np.int32)
        assert Ti.shape[0] >= len(ch) and Ai.shape[0] >= len(ch)
        return Ti, _mat(T0), _mat(A0), Ai, ch

    def integrate_ctr_device(self, device_ptr: int, Ti, T0, A0, Ai, chunks):
        """ControlGrid mode with the frames resident in HBM."""
        Ti, T0, A0, Ai, ch = self._ctr_args(Ti, T0, A0, Ai, chunks)
        _check(lib().er_tsdf_integrate_ctr_device(self._h, C.c_void_p(device_ptr), _dp(Ti), _dp(T0), _dp(A0), _dp(Ai),
                                                  ch.ctypes.data_as(c_ip), len(ch)))

    def profile_reproject_device(self, device_ptr: int, Ti, T0, A0, Ai, chunks) -> float:
        """Device milliseconds of the re-projection alone over the resident frames."""
        Ti, T0, A0, Ai, ch = self._ctr_args(Ti, T0, A0, Ai, chunks)
        ms = C.c_float()
        _check(lib().er_tsdf_profile_reproject_device(self._h, C.c_void_p(device_ptr), _dp(Ti), _dp(T0), _dp(A0), _dp(Ai),
                                                      ch.ctypes.data_as(c_ip), len(ch), C.byref(ms)))
        return ms.value

    def inputs_consumed(self):
        _check(lib().er_tsdf_inputs_consumed(self._h))

    def sync(self):
        _check(lib().er_tsdf_sync(self._h))

    # -- inspection / output
    def num_blocks(self) -> int:
        n = C.c_int()
        _check(lib().er_tsdf_num_blocks(self._h, C.byref(n)))
        return n.value

    def block_keys(self) -> np.ndarray:
        n = self.num_blocks()
        keys = np.empty(n, dtype=np.int32)
        if n:
            _check(lib().er_tsdf_block_keys(self._h, keys.ctypes.data_as(c_ip), n))
        return keys

    def read_block(self, key: int):
        sdf = np.empty(64 ** 3, dtype=np.float32)
        w = np.empty(64 ** 3, dtype=np.float32)
        _check(lib().er_tsdf_read_block(self._h, int(key), _fp(sdf), _fp(w)))
        return sdf, w

    def extract(self) -> np.ndarray:
        """SaveWorld predicate (TSDFVolume.cpp:104-132): float32 [n,4] = x,y,z (voxel units), tsdf."""
        n = C.c_size_t()
        _check(lib().er_tsdf_extract(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, 4), dtype=np.float32)
        if n.value:
            m = C.c_size_t()
            _check(lib().er_tsdf_extract(self._h, _fp(out), n.value, C.byref(m)))
            assert m.value == n.value
        return out

    def extract_count(self) -> int:
        n = C.c_size_t()
        _check(lib().er_tsdf_extract(self._h, None, 0, C.byref(n)))
        return n.value

    def extract_into(self, hostbuf: "HostBuffer", cap_rows: int) -> int:
        """SaveWorld's voxel list into a (pinned) host buffer of cap_rows x 16 bytes; returns the number of rows."""
        assert hostbuf.nbytes >= cap_rows * 16
        n = C.c_size_t()
        _check(lib().er_tsdf_extract(self._h, C.cast(C.c_void_p(hostbuf.ptr), c_fp), cap_rows, C.byref(n)))
        return n.value

    def extract_begin(self, hostbuf: "HostBuffer", cap_rows: int):
        """Start SaveWorld's extraction into a page-locked buffer without blocking (er_tsdf_extract_begin)."""
        assert hostbuf.nbytes >= cap_rows * 16
        _check(lib().er_tsdf_extract_begin(self._h, C.cast(C.c_void_p(hostbuf.ptr), c_fp), cap_rows))

    def extract_wait(self) -> int:
        n = C.c_size_t()
        _check(lib().er_tsdf_extract_wait(self._h, C.byref(n)))
        return n.value

    def save(self, path: str):
        """Raw volume checkpoint (er_tsdf_save)."""
        _check(lib().er_tsdf_save(self._h, path.encode()))

    def load(self, path: str):
        """Add the blocks of a checkpoint to this volume (er_tsdf_load)."""
        _check(lib().er_tsdf_load(self._h, path.encode()))

    def save_world(self, path: str) -> int:
        """Write world.pcd exactly as pcl::io::savePCDFile(binary) does for PointXYZI (TSDFVolume.cpp:130)."""
        pts = self.extract()
        n = len(pts)
        hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\n"
               "SIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\n"
               "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n))
        with open(path, "wb") as f:
            f.write(hdr.encode("ascii"))
            f.write(pts.astype("<f4").tobytes())
        return n

    # -- device timing
    def event_record(self, slot: int):
        _check(lib().er_tsdf_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        _check(lib().er_tsdf_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(lib().er_tsdf_launch_count(self._h))


# ================================================================================================== ICP
class Cloud:
    """One fragment resident on the device: the reference's pointclouds_[i] (CorresApp.cpp:82-98) plus the
    uniform-grid search structure standing in for pcl::KdTreeFLANN.  cell_size >= every radius used later."""

    def __init__(self, xyz, normals, device: int = 0, cell_size: float = 0.03):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        self.normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        assert self.xyz.shape == self.normals.shape
        self.device = device
        h = lib().er_cloud_upload(device, _fp(self.xyz), _fp(self.normals), len(self.xyz), float(cell_size))
        if not h:
            raise ElasticError("er_cloud_upload failed: " + lib().er_last_error().decode("utf-8", "replace"))
        self._h = C.c_void_p(h)

    META_BYTES = 8 * 7 + 8 + 4 * 5 + 4 * 4          # sizeof(er_cloud_meta) = 100, padded to 104

    def __len__(self):
        return int(lib().er_cloud_size(self._h))

    def blob(self):
        """(meta bytes, device pointer, blob bytes) of the built cloud (er_cloud_get_meta)."""
        meta = (C.c_uint8 * 104)()
        ptr = C.c_void_p()
        _check(lib().er_cloud_get_meta(self._h, meta, C.byref(ptr)))
        nbytes = int(np.frombuffer(bytes(meta), dtype=np.uint64, count=2)[1])
        return bytes(meta), ptr.value, nbytes

    @classmethod
    def from_blob(cls, meta: bytes, device_ptr: int, device: int = 0):
        """A cloud built elsewhere (er_cloud_from_blob copies the blob at device_ptr into its own allocation)."""
        self = cls.__new__(cls)
        self.xyz = self.normals = None
        self.device = device
        buf = (C.c_uint8 * 104).from_buffer_copy(meta)
        h = lib().er_cloud_from_blob(device, buf, C.c_void_p(device_ptr))
        if not h:
            raise ElasticError("er_cloud_from_blob failed: " + lib().er_last_error().decode("utf-8", "replace"))
        self._h = C.c_void_p(h)
        return self

    def close(self):
        if getattr(self, "_h", None):
            lib().er_cloud_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def icp_count_inliers(tgt: Cloud, src: Cloud, T, dist: float) -> int:
    """Gate of CCorresApp::Registration (CorresApp.cpp:250-264)."""
    n = C.c_int()
    _check(lib().er_icp_count_inliers(tgt._h, src._h, _dp(_mat(T)), float(dist), C.byref(n)))
    return n.value


def icp_align(tgt: Cloud, src: Cloud, Tinit, max_dist=0.03, max_iter=20, eps=1e-6, fixed_iters=False):
    """pcl::IterativeClosestPoint as configured at CorresApp.cpp:295-306.  Returns (T, iterations, fitness)."""
    out = np.empty(16, dtype=np.float64)
    it = C.c_int()
    fit = C.c_double()
    _check(lib().er_icp_align(tgt._h, src._h, _dp(_mat(Tinit)), float(max_dist), int(max_iter), float(eps),
                              int(bool(fixed_iters)), _dp(out), C.byref(it), C.byref(fit)))
    return out.reshape(4, 4), it.value, fit.value


def icp_align_batch(clouds, tgt_idx, src_idx, Tinit, max_dist=0.03, max_iter=20, eps=1e-6, fixed_iters=False):
    """Many pairs in flight on one device (pair loop of CCorresApp::Registration, CorresApp.cpp:220)."""
    n = len(tgt_idx)
    arr = (C.c_void_p * len(clouds))(*[c._h for c in clouds])
    ti = np.ascontiguousarray(tgt_idx, dtype=np.int32)
    si = np.ascontiguousarray(src_idx, dtype=np.int32)
    Ti = np.ascontiguousarray(Tinit, dtype=np.float64).reshape(n, 16)
    out = np.empty((n, 16), dtype=np.float64)
    its = np.empty(n, dtype=np.int32)
    fit = np.empty(n, dtype=np.float64)
    _check(lib().er_icp_align_batch(arr, len(clouds), ti.ctypes.data_as(c_ip), si.ctypes.data_as(c_ip), n, _dp(Ti),
                                    float(max_dist), int(max_iter), float(eps), int(bool(fixed_iters)), _dp(out),
                                    its.ctypes.data_as(c_ip), _dp(fit)))
    return out.reshape(n, 4, 4), its, fit


def icp_correspondences(tgt: Cloud, src: Cloud, T, dist=0.015, normal_cos=0.8660, want_info=True):
    """FindCorrespondence inner loop (CorresApp.cpp:145-161,186-208): (pairs [m,2] = (target, source), info 6x6)."""
    cap = len(src)
    pairs = np.empty((max(cap, 1), 2), dtype=np.int32)
    n = C.c_size_t()
    info = np.zeros(36, dtype=np.float64) if want_info else None
    _check(lib().er_icp_correspondences(tgt._h, src._h, _dp(_mat(T)), float(dist), float(normal_cos),
                                        pairs.ctypes.data_as(c_i32p), cap, C.byref(n),
                                        _dp(info) if want_info else None))
    return pairs[: n.value].copy(), (info.reshape(6, 6) if want_info else None)


def icp_count_inliers_batch(clouds, tgt_idx, src_idx, T, dist: float) -> np.ndarray:
    """The gate for a whole candidate list in one launch (pair loop of CorresApp.cpp:220-264)."""
    n = len(tgt_idx)
    arr = (C.c_void_p * len(clouds))(*[c._h for c in clouds])
    ti = np.ascontiguousarray(tgt_idx, dtype=np.int32)
    si = np.ascontiguousarray(src_idx, dtype=np.int32)
    Ts = np.ascontiguousarray(T, dtype=np.float64).reshape(n, 16)
    cnt = np.zeros(n, dtype=np.int32)
    _check(lib().er_icp_count_inliers_batch(arr, len(clouds), ti.ctypes.data_as(c_ip), si.ctypes.data_as(c_ip), n, _dp(Ts),
                                            float(dist), cnt.ctypes.data_as(c_ip)))
    return cnt


def icp_correspondences_batch(clouds, tgt_idx, src_idx, T, dist=0.015, normal_cos=0.8660, want_info=True):
    """FindCorrespondence for a list of pairs (CorresApp.cpp:121-208): ([pairs_p [m_p,2]], info [n,6,6] or None)."""
    n = len(tgt_idx)
    arr = (C.c_void_p * len(clouds))(*[c._h for c in clouds])
    ti = np.ascontiguousarray(tgt_idx, dtype=np.int32)
    si = np.ascontiguousarray(src_idx, dtype=np.int32)
    Ts = np.ascontiguousarray(T, dtype=np.float64).reshape(n, 16)
    cap = int(sum(len(clouds[j]) for j in src_idx))
    rows = np.empty((max(cap, 1), 2), dtype=np.int32)
    start = np.zeros(n + 1, dtype=np.uintp)
    info = np.zeros((n, 36), dtype=np.float64) if want_info else None
    _check(lib().er_icp_correspondences_batch(arr, len(clouds), ti.ctypes.data_as(c_ip), si.ctypes.data_as(c_ip), n, _dp(Ts),
                                              float(dist), float(normal_cos), rows.ctypes.data_as(c_i32p), cap,
                                              start.ctypes.data_as(c_szp), _dp(info) if want_info else None))
    out = [rows[int(start[p]):int(start[p + 1])].copy() for p in range(n)]
    return out, (info.reshape(n, 6, 6) if want_info else None)


def icp_release(device: int = 0):
    _check(lib().er_icp_release(device))


def icp_event_record(device: int, slot: int):
    _check(lib().er_icp_event_record(device, slot))


def icp_event_elapsed_ms(device: int, a: int, b: int) -> float:
    ms = C.c_float()
    _check(lib().er_icp_event_elapsed_ms(device, a, b, C.byref(ms)))
    return ms.value


def icp_launch_count(device: int) -> int:
    return int(lib().er_icp_launch_count(device))


# ================================================================================================== FragmentOptimizer
class FragOpt:
    """COptApp's device-side state (include/elastic_b200.h, er_fragopt_*): the fragments' points with their lattice
    coordinates, the correspondences of all pairs, and the assembly of the SLAC / rigid normal equations.
    `library` = an alternative ctypes library exporting the same entry points (the CPU emulation of the tests)."""

    def __init__(self, num: int, resolution: int = 8, length: float = 3.0, device: int = 0, library=None):
        self._L = library if library is not None else lib()
        if library is not None:
            for name, (res, args) in SIGNATURES.items():
                if name.startswith("er_fragopt_") or name == "er_last_error":
                    fn = getattr(library, name)
                    fn.restype, fn.argtypes = res, args
        self.num, self.res, self.length = int(num), int(resolution), float(length)
        self.nper = (self.res + 1) ** 3 * 3
        self.M = 6 * self.num + self.nper
        self.sizes = [0] * self.num
        h = self._L.er_fragopt_create(int(device), self.num, self.res, self.length)
        if not h:
            raise ElasticError("er_fragopt_create failed: " + self._L.er_last_error().decode("utf-8", "replace"))
        self._h = C.c_void_p(h)

    def _check(self, rc):
        if rc != 0:
            raise ElasticError(self._L.er_last_error().decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._L.er_fragopt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_cloud(self, l: int, xyzn) -> int:
        a = np.ascontiguousarray(xyzn, dtype=np.float32).reshape(-1, 6)
        kept = C.c_size_t()
        self._check(self._L.er_fragopt_set_cloud(self._h, int(l), _fp(a), len(a), C.byref(kept)))
        self.sizes[l] = kept.value
        return kept.value

    def set_pairs(self, pairs, corres):
        """pairs = [(i, j)], corres = list of [m,2] int arrays (index into cloud i, index into cloud j)."""
        n = len(pairs)
        i0 = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int32)
        i1 = np.ascontiguousarray([p[1] for p in pairs], dtype=np.int32)
        start = np.zeros(n + 1, dtype=np.uintp)
        for k, c in enumerate(corres):
            start[k + 1] = start[k] + len(c)
        rows = (np.concatenate([np.asarray(c, dtype=np.int32).reshape(-1, 2) for c in corres]) if n and start[n]
                else np.zeros((1, 2), dtype=np.int32))
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        self._check(self._L.er_fragopt_set_pairs(self._h, n, i0.ctypes.data_as(c_ip), i1.ctypes.data_as(c_ip),
                                                 start.ctypes.data_as(c_szp), rows.ctypes.data_as(c_i32p)))

    def update_pose(self, T):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(self.num, 16)
        self._check(self._L.er_fragopt_update_pose(self._h, _dp(T)))

    def update_normals(self, ctr):
        c = np.ascontiguousarray(ctr, dtype=np.float64).reshape(self.num * self.nper)
        self._check(self._L.er_fragopt_update_normals(self._h, _dp(c)))

    def update_points_pn(self, ctr):
        c = np.ascontiguousarray(ctr, dtype=np.float64).reshape(self.num * self.nper)
        self._check(self._L.er_fragopt_update_points_pn(self._h, _dp(c)))

    def points(self, l: int, ctr=None) -> np.ndarray:
        out = np.empty((max(self.sizes[l], 1), 6), dtype=np.float32)
        c = None if ctr is None else np.ascontiguousarray(ctr, dtype=np.float64).reshape(self.num * self.nper)
        self._check(self._L.er_fragopt_get_points(self._h, int(l), _fp(out), self.sizes[l], None if c is None else _dp(c)))
        return out[: self.sizes[l]]

    def slac_assemble(self, rot_t):
        """Returns (Jb [M], score); the matrix stays on the device (slac_jj() copies it out)."""
        r = np.ascontiguousarray(rot_t, dtype=np.float64).reshape(self.num * 9)
        Jb = np.empty(self.M, dtype=np.float64)
        score = C.c_double()
        self._check(self._L.er_fragopt_slac_assemble(self._h, _dp(r), _dp(Jb), C.byref(score)))
        return Jb, score.value

    def slac_accumulate(self, other: "FragOpt"):
        """self's assembled matrix += other's (several devices, er_fragopt_slac_accumulate)."""
        self._check(self._L.er_fragopt_slac_accumulate(self._h, other._h))

    def slac_jj(self) -> np.ndarray:
        jj = np.empty((self.M, self.M), dtype=np.float64)
        self._check(self._L.er_fragopt_slac_get_jj(self._h, _dp(jj)))
        return jj

    def slac_solve(self, base, rhs) -> np.ndarray:
        b = None if base is None else np.ascontiguousarray(base, dtype=np.float64).reshape(self.M * self.M)
        r = np.ascontiguousarray(rhs, dtype=np.float64).reshape(self.M)
        x = np.empty(self.M, dtype=np.float64)
        self._check(self._L.er_fragopt_slac_solve(self._h, None if b is None else _dp(b), _dp(r), _dp(x)))
        return x

    def rigid_assemble(self):
        n6 = 6 * self.num
        JJ = np.empty((n6, n6), dtype=np.float64)
        Jb = np.empty(n6, dtype=np.float64)
        score = C.c_double()
        self._check(self._L.er_fragopt_rigid_assemble(self._h, _dp(JJ), _dp(Jb), C.byref(score)))
        return JJ, Jb, score.value

    def nonrigid_assemble(self, weight: float = 1.0):
        self._check(self._L.er_fragopt_nonrigid_assemble(self._h, float(weight)))

    def nonrigid_aa(self) -> np.ndarray:
        n = self.num * self.nper
        aa = np.empty((n, n), dtype=np.float64)
        self._check(self._L.er_fragopt_nonrigid_get_aa(self._h, _dp(aa)))
        return aa

    def nonrigid_factor(self, rows, cols, vals):
        r = np.ascontiguousarray(rows, dtype=np.int32)
        c = np.ascontiguousarray(cols, dtype=np.int32)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        self._check(self._L.er_fragopt_nonrigid_factor(self._h, len(v), r.ctypes.data_as(c_i32p), c.ctypes.data_as(c_i32p), _dp(v)))

    def nonrigid_solve(self, rhs) -> np.ndarray:
        n = self.num * self.nper
        r = np.ascontiguousarray(rhs, dtype=np.float64).reshape(n)
        x = np.empty(n, dtype=np.float64)
        self._check(self._L.er_fragopt_nonrigid_solve(self._h, _dp(r), _dp(x)))
        return x

    def assemble_stats(self):
        ms = C.c_float()
        rows = C.c_uint64()
        fl = C.c_uint64()
        self._check(self._L.er_fragopt_last_assemble_stats(self._h, C.byref(ms), C.byref(rows), C.byref(fl)))
        return ms.value, rows.value, fl.value

    def launch_count(self) -> int:
        return int(self._L.er_fragopt_launch_count(self._h))
