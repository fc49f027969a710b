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
REF_INTEGRATE_12MM = os.path.join(_HERE, "_ref", "Integrate_ref_12mm")     # unit_length_ patched to 0.012 (Makefile: ref12)
REF_FRAGOPT = os.path.join(_HERE, "_ref", "FragmentOptimizer_ref")       # verbatim FragmentOptimizer (Makefile: fragopt)
EIGEN_CHECK = os.path.join(_HERE, "_ref", "libeigen_check.so")              # the reference's vendored Eigen 3.1.2 (Makefile: eigen)

_lib = None


def build(force: bool = False) -> None:
    """Compile the C restatement (and, where /root/reference exists, the verbatim reference)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(LIB_PATH)
            for s in ("tsdf_oracle.c", "icp_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/liboracle.so"])
    if os.path.isdir("/root/reference/Integrate") and (force or not os.path.exists(REF_INTEGRATE)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    if os.path.isdir("/root/reference/Integrate") and (force or not os.path.exists(REF_INTEGRATE_12MM)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref12"])
    if os.path.isdir("/root/reference/FragmentOptimizer/external/Eigen") and (force or not os.path.exists(EIGEN_CHECK)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "eigen"])
    if os.path.isdir("/root/reference/FragmentOptimizer") and (force or not os.path.exists(REF_FRAGOPT)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "fragopt"])


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
        L.oracle_tsdf_last_touched.restype = C.c_int
        L.oracle_tsdf_last_touched.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
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
        i32p = C.POINTER(C.c_int32)
        L.oracle_cloud_create.restype = C.c_void_p
        L.oracle_cloud_create.argtypes = [fp, fp, C.c_int, C.c_double]
        L.oracle_cloud_free.argtypes = [C.c_void_p]
        L.oracle_nn_brute.restype = C.c_int
        L.oracle_nn_brute.argtypes = [C.c_void_p, fp, fp]
        L.oracle_nn_grid.restype = C.c_int
        L.oracle_nn_grid.argtypes = [C.c_void_p, fp, fp]
        L.oracle_icp_count_inliers.restype = C.c_int
        L.oracle_icp_count_inliers.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_double, C.c_int]
        L.oracle_solve6.restype = C.c_int
        L.oracle_solve6.argtypes = [dp, dp, dp]
        L.oracle_icp_align.restype = C.c_int
        L.oracle_icp_align.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_double, C.c_int, C.c_double, C.c_int, dp, ip, dp,
                                       C.c_int]
        L.oracle_icp_align_batch.argtypes = [C.POINTER(C.c_void_p), ip, ip, C.c_int, dp, C.c_double, C.c_int, C.c_double,
                                             C.c_int, dp, ip, dp, C.c_int]
        L.oracle_icp_correspondences.restype = C.c_int
        L.oracle_icp_correspondences.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_double, C.c_double, i32p, C.c_int, dp,
                                                 C.c_int]
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

    def last_touched(self) -> np.ndarray:
        """Keys of the blocks the last integrate() touched (== its IntegrateVolumeUnit calls), first-touch order."""
        n = lib().oracle_tsdf_last_touched(self._h, None)
        keys = np.zeros(max(n, 1), dtype=np.int32)
        lib().oracle_tsdf_last_touched(self._h, keys.ctypes.data_as(C.POINTER(C.c_int)))
        return keys[:n]

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


class OracleCloud:
    """One fragment for the ICP oracle (icp_oracle.c): xyz/normals float32 [n,3]; cell >= any search radius."""

    def __init__(self, xyz, nrm, cell: float = 0.03):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        self.nrm = np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        self._h = C.c_void_p(lib().oracle_cloud_create(_fp(self.xyz), _fp(self.nrm), len(self.xyz), float(cell)))

    def __len__(self):
        return len(self.xyz)

    def close(self):
        if self._h:
            lib().oracle_cloud_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nn(self, q, brute=False):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(3)
        d2 = C.c_float()
        fn = lib().oracle_nn_brute if brute else lib().oracle_nn_grid
        return fn(self._h, _fp(q), C.byref(d2)), d2.value


def icp_count_inliers(tgt: OracleCloud, src: OracleCloud, T, dist: float, threads: int = 8) -> int:
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    return lib().oracle_icp_count_inliers(tgt._h, src._h, _dp(T), float(dist), threads)


def icp_align(tgt: OracleCloud, src: OracleCloud, Tinit, max_dist=0.03, max_iter=20, eps=1e-6, fixed_iters=False,
              threads: int = 8):
    """Returns (T [4,4] float64, iterations, mean NN^2 of the last iteration, converged)."""
    Tinit = np.ascontiguousarray(Tinit, dtype=np.float64).reshape(16)
    out = np.empty(16, dtype=np.float64)
    it = C.c_int()
    fit = C.c_double()
    conv = lib().oracle_icp_align(tgt._h, src._h, _dp(Tinit), float(max_dist), int(max_iter), float(eps),
                                  int(bool(fixed_iters)), _dp(out), C.byref(it), C.byref(fit), threads)
    return out.reshape(4, 4), it.value, fit.value, bool(conv)


def icp_align_batch(clouds, tgt_idx, src_idx, Tinit, max_dist=0.03, max_iter=20, eps=1e-6, fixed_iters=False,
                    threads: int = 8):
    """Pair loop of CCorresApp::Registration (omp over pairs, `threads` threads, each pair single-threaded)."""
    n = len(tgt_idx)
    arr = (C.c_void_p * len(clouds))(*[c._h for c in clouds])
    ti = np.ascontiguousarray(tgt_idx, dtype=np.int32)
    si = np.ascontiguousarray(src_idx, dtype=np.int32)
    Ti = np.ascontiguousarray(Tinit, dtype=np.float64).reshape(n, 16)
    out = np.empty((n, 16), dtype=np.float64)
    its = np.empty(n, dtype=np.int32)
    fit = np.empty(n, dtype=np.float64)
    lib().oracle_icp_align_batch(arr, ti.ctypes.data_as(C.POINTER(C.c_int)), si.ctypes.data_as(C.POINTER(C.c_int)), n,
                                 _dp(Ti), float(max_dist), int(max_iter), float(eps), int(bool(fixed_iters)), _dp(out),
                                 its.ctypes.data_as(C.POINTER(C.c_int)), _dp(fit), threads)
    return out.reshape(n, 4, 4), its, fit


def icp_correspondences(tgt: OracleCloud, src: OracleCloud, T, dist=0.015, normal_cos=0.8660, want_info=True,
                        threads: int = 8):
    """Returns (pairs int32 [m,2] = (target idx, source idx), info [6,6] or None)."""
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    pairs = np.empty((len(src), 2), dtype=np.int32)
    info = np.zeros(36, dtype=np.float64) if want_info else None
    m = lib().oracle_icp_correspondences(tgt._h, src._h, _dp(T), float(dist), float(normal_cos),
                                         pairs.ctypes.data_as(C.POINTER(C.c_int32)), len(src),
                                         _dp(info) if want_info else None, threads)
    return pairs[:m].copy(), (info.reshape(6, 6) if want_info else None)


def solve6(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(36)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(6)
    x = np.empty(6, dtype=np.float64)
    rc = lib().oracle_solve6(_dp(A), _dp(b), _dp(x))
    return x if rc == 0 else None


def sort_world(xyzi: np.ndarray) -> np.ndarray:
    """Canonical order for order-free comparison of SaveWorld output (keyed by integer voxel coords)."""
    xyzi = np.asarray(xyzi, dtype=np.float32).reshape(-1, 4)
    k = xyzi[:, :3].astype(np.int64)
    order = np.lexsort((k[:, 2], k[:, 1], k[:, 0]))
    return xyzi[order]


def world_digest(xyzi: np.ndarray) -> str:
    import hashlib
    return hashlib.sha256(sort_world(xyzi).tobytes()).hexdigest()


def run_reference_integrate(args, cwd=None, timing: bool = False, binary: str = None):
    """Run the verbatim reference binary (oracle/_ref/Integrate_ref).  Returns (stdout+stderr, ms_per_frame)."""
    env = dict(os.environ)
    if timing:
        env["ER_SHIM_TIMING"] = "1"
    p = subprocess.run([binary or REF_INTEGRATE] + list(args), cwd=cwd, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, check=True)
    ms = None
    for line in p.stdout.splitlines():
        if line.startswith("[shim]") and "ms_per_frame=" in line:
            ms = float(line.split("ms_per_frame=")[1])
    return p.stdout, ms


def run_reference_fragopt(args, cwd=None, threads: int = None, timeout: float = 3600.0):
    """Run the verbatim reference FragmentOptimizer (oracle/_ref/FragmentOptimizer_ref) in `cwd` (it writes pose.log,
    output.ctr ... there).  Its pair loop sums per-pair matrices under an OpenMP critical section in arrival order
    (OptApp.cpp:562-569), so results vary at rounding level from run to run; threads=1 (OMP_THREAD_LIMIT) fixes the
    order.  Returns stdout+stderr."""
    env = dict(os.environ)
    if threads is not None:
        env["OMP_THREAD_LIMIT"] = str(threads)
    p = subprocess.run([REF_FRAGOPT] + list(args), cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, check=True, timeout=timeout)
    return p.stdout
