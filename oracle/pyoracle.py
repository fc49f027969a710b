"""ctypes front-end of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does (tests/test_layout.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
REF_INTEGRATE = os.path.join(_HERE, "_ref", "Integrate_ref")

_lib = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and, where /root/reference exists, the verbatim reference)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(LIB_PATH)
            for s in ("tsdf_oracle.c", "icp_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/liboracle.so"])
    if os.path.isdir("/root/reference/Integrate") and (force or not os.path.exists(REF_INTEGRATE)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        fp = C.POINTER(C.c_float)
        u16p = C.POINTER(C.c_uint16)
        ip = C.POINTER(C.c_int)
        L.oracle_tsdf_create.restype = C.c_void_p
        L.oracle_tsdf_create.argtypes = [C.c_int, C.c_int]
        L.oracle_tsdf_destroy.argtypes = [C.c_void_p]
        L.oracle_tsdf_set_camera.argtypes = [C.c_void_p] + [C.c_float] * 5
        L.oracle_tsdf_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.oracle_tsdf_set_voxel.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.oracle_scale_depth.argtypes = [C.c_void_p, u16p, fp]
        L.oracle_tsdf_integrate.restype = C.c_int
        L.oracle_tsdf_integrate.argtypes = [C.c_void_p, u16p, dp]
        L.oracle_tsdf_last_updated_voxels.restype = C.c_longlong
        L.oracle_tsdf_last_updated_voxels.argtypes = [C.c_void_p]
        L.oracle_tsdf_num_blocks.restype = C.c_int
        L.oracle_tsdf_num_blocks.argtypes = [C.c_void_p]
        L.oracle_tsdf_block_keys.argtypes = [C.c_void_p, ip]
        L.oracle_tsdf_read_block.restype = C.c_int
        L.oracle_tsdf_read_block.argtypes = [C.c_void_p, C.c_int, fp, fp]
        L.oracle_tsdf_extract.restype = C.c_size_t
        L.oracle_tsdf_extract.argtypes = [C.c_void_p, fp, C.c_size_t]
        L.oracle_reproject.argtypes = [C.c_void_p, u16p, dp, dp, dp, dp, fp, C.c_int, C.c_float]
        L.oracle_mat4_inverse.argtypes = [dp, dp]
        L.oracle_mat4_mul.argtypes = [dp, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u16p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


def mat4_inverse(T) -> np.ndarray:
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    out = np.empty(16, dtype=np.float64)
    lib().oracle_mat4_inverse(_dp(T), _dp(out))
    return out.reshape(4, 4)


def mat4_mul(a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(16)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(16)
    out = np.empty(16, dtype=np.float64)
    lib().oracle_mat4_mul(_dp(a), _dp(b), _dp(out))
    return out.reshape(4, 4)


class OracleTSDF:
    """CPU restatement of TSDFVolume (+ CIntegrateApp::Reproject)."""

    def __init__(self, cols: int = 640, rows: int = 480, threads: int = 8):
        self.cols, self.rows = cols, rows
        self._h = C.c_void_p(lib().oracle_tsdf_create(cols, rows))
        lib().oracle_tsdf_set_threads(self._h, threads)

    def close(self):
        if self._h:
            lib().oracle_tsdf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_camera(self, fx, fy, cx, cy, integration_trunc):
        lib().oracle_tsdf_set_camera(self._h, fx, fy, cx, cy, integration_trunc)

    def set_voxel(self, unit_length, tsdf_trunc):
        lib().oracle_tsdf_set_voxel(self._h, unit_length, tsdf_trunc)

    def scale_depth(self, depth) -> np.ndarray:
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        out = np.empty((self.rows, self.cols), dtype=np.float32)
        lib().oracle_scale_depth(self._h, _u16p(depth), _fp(out))
        return out

    def integrate(self, depth, T) -> int:
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        return lib().oracle_tsdf_integrate(self._h, _u16p(depth), _dp(T))

    def last_updated_voxels(self) -> int:
        return int(lib().oracle_tsdf_last_updated_voxels(self._h))

    def reproject(self, depth, Ti, T0, A0, Ai, ctr, res, length) -> np.ndarray:
        depth = np.array(depth, dtype=np.uint16, order="C", copy=True)
        m = [np.ascontiguousarray(x, dtype=np.float64).reshape(16) for x in (Ti, T0, A0, Ai)]
        ctr = np.ascontiguousarray(ctr, dtype=np.float32)
        lib().oracle_reproject(self._h, _u16p(depth), _dp(m[0]), _dp(m[1]), _dp(m[2]), _dp(m[3]), _fp(ctr),
                               int(res), C.c_float(length))
        return depth

    def num_blocks(self) -> int:
        return lib().oracle_tsdf_num_blocks(self._h)

    def block_keys(self) -> np.ndarray:
        keys = np.empty(self.num_blocks(), dtype=np.int32)
        if len(keys):
            lib().oracle_tsdf_block_keys(self._h, keys.ctypes.data_as(C.POINTER(C.c_int)))
        return keys

    def read_block(self, key: int):
        sdf = np.empty(64 ** 3, dtype=np.float32)
        w = np.empty(64 ** 3, dtype=np.float32)
        if lib().oracle_tsdf_read_block(self._h, int(key), _fp(sdf), _fp(w)) != 0:
            raise KeyError(key)
        return sdf, w

    def extract(self) -> np.ndarray:
        n = lib().oracle_tsdf_extract(self._h, None, 0)
        out = np.empty((n, 4), dtype=np.float32)
        if n:
            lib().oracle_tsdf_extract(self._h, _fp(out), n)
        return out


def sort_world(xyzi: np.ndarray) -> np.ndarray:
    """Canonical order for order-free comparison of SaveWorld output (keyed by integer voxel coords)."""
    xyzi = np.asarray(xyzi, dtype=np.float32).reshape(-1, 4)
    k = xyzi[:, :3].astype(np.int64)
    order = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
    return xyzi[order]


def world_digest(xyzi: np.ndarray) -> str:
    import hashlib
    return hashlib.sha256(sort_world(xyzi).tobytes()).hexdigest()


def run_reference_integrate(args, cwd=None, timing: bool = False):
    """Run the verbatim reference binary (oracle/_ref/Integrate_ref).  Returns (stdout+stderr, ms_per_frame)."""
    env = dict(os.environ)
    if timing:
        env["ER_SHIM_TIMING"] = "1"
    p = subprocess.run([REF_INTEGRATE] + list(args), cwd=cwd, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, check=True)
    ms = None
    for line in p.stdout.splitlines():
        if line.startswith("[shim]") and "ms_per_frame=" in line:
            ms = float(line.split("ms_per_frame=")[1])
    return p.stdout, ms
