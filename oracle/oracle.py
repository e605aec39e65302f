"""ctypes bindings of the CPU oracle (oracle/_build/liblocus_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by locus_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblocus_oracle.so")


def build(force=False):
    srcs = ["kdtree.c", "bfgs_oracle.c", "gicp_oracle.c", "voxel_oracle.c", "normals_oracle.c", "ndt_oracle.c", "lb_oracle.h"]
    if not force and os.path.exists(_SO):
        mt = os.path.getmtime(_SO)
        if all(os.path.getmtime(os.path.join(_HERE, s)) <= mt for s in srcs):
            return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class GicpParams(C.Structure):
    _fields_ = [
        ("k_correspondences", C.c_int), ("gicp_epsilon", C.c_double),
        ("rotation_epsilon", C.c_double), ("transformation_epsilon", C.c_double),
        ("corr_dist_threshold", C.c_double), ("max_iterations", C.c_int),
        ("max_inner_iterations", C.c_int), ("num_threads", C.c_int),
        ("source_cov_from_normals", C.c_int), ("target_cov_from_normals", C.c_int),
        ("optimizer", C.c_int),
    ]


class GicpResult(C.Structure):
    _fields_ = [
        ("final_transformation", C.c_float * 16), ("nr_iterations", C.c_int),
        ("converged", C.c_int), ("n_correspondences", C.c_int), ("delta", C.c_double),
        ("n_fdf_evals", C.c_long), ("n_inner_iterations", C.c_long),
        ("t_covariances_s", C.c_double), ("t_iterations_s", C.c_double), ("t_total_s", C.c_double),
        ("t_lookups_s", C.c_double), ("t_optimization_s", C.c_double), ("status", C.c_int),
    ]


class VoxelParams(C.Structure):
    _fields_ = [
        ("leaf", C.c_float * 3), ("filter_field_offset", C.c_int),
        ("filter_limit_min", C.c_double), ("filter_limit_max", C.c_double),
        ("filter_limit_negative", C.c_int), ("min_points_per_voxel", C.c_int),
        ("downsample_all_data", C.c_int),
        ("body_enabled", C.c_int), ("body_min", C.c_float * 3), ("body_max", C.c_float * 3), ("body_rotation", C.c_float),
    ]


class NdtParams(C.Structure):
    _fields_ = [
        ("resolution", C.c_float), ("step_size", C.c_double), ("outlier_ratio", C.c_double),
        ("transformation_epsilon", C.c_double), ("max_iterations", C.c_int), ("min_points_per_voxel", C.c_int),
        ("min_covar_eigvalue_mult", C.c_double), ("search_method", C.c_int), ("num_threads", C.c_int),
    ]


class NdtResult(C.Structure):
    _fields_ = [
        ("final_transformation", C.c_float * 16), ("converged", C.c_int), ("nr_iterations", C.c_int),
        ("trans_probability", C.c_double), ("n_evaluations", C.c_long), ("pose", C.c_double * 6), ("status", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.og_gicp_default_params.argtypes = [C.POINTER(GicpParams)]
        L.og_gicp_align.restype = C.c_int
        L.og_gicp_align.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
            C.POINTER(GicpParams), C.c_void_p, C.POINTER(GicpResult), C.c_void_p, C.c_void_p, C.c_void_p]
        L.og_gicp_target_prepare.restype = C.c_void_p
        L.og_gicp_target_prepare.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(GicpParams)]
        L.og_gicp_target_free.argtypes = [C.c_void_p]
        L.og_gicp_target_set_covariances.argtypes = [C.c_void_p, C.c_void_p]
        L.og_gicp_align_prepared.restype = C.c_int
        L.og_gicp_align_prepared.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(GicpParams),
                                             C.c_void_p, C.POINTER(GicpResult)]
        L.og_gicp_covariances.restype = C.c_int
        L.og_gicp_covariances.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.og_gicp_fitness.restype = C.c_double
        L.og_gicp_fitness.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_double, C.c_int]
        L.og_gicp_fdf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.og_gicp_apply_state.argtypes = [C.c_void_p, C.c_void_p]
        L.og_kdtree_build.restype = C.c_void_p
        L.og_kdtree_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.og_kdtree_free.argtypes = [C.c_void_p]
        L.og_kdtree_knn.restype = C.c_int
        L.og_kdtree_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.og_kdtree_radius.restype = C.c_int
        L.og_kdtree_radius.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        L.og_kdtree_nn_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.og_bfgs_minimize_quadratic.restype = C.c_int
        L.og_bfgs_minimize_quadratic.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double]
        L.og_voxel_filter.restype = C.c_int
        L.og_voxel_filter.argtypes = [
            C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
            C.c_void_p, C.c_int, C.POINTER(VoxelParams), C.c_void_p, C.POINTER(C.c_size_t),
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.og_normalize_pcloud.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.og_compute_ap.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.og_ndt_default_params.argtypes = [C.POINTER(NdtParams)]
        L.og_ndt_target_build.restype = C.c_void_p
        L.og_ndt_target_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(NdtParams)]
        L.og_ndt_target_free.argtypes = [C.c_void_p]
        L.og_ndt_target_info.restype = C.c_int
        L.og_ndt_target_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.og_ndt_target_leaves.argtypes = [C.c_void_p] * 6
        L.og_ndt_derivatives.restype = C.c_int
        L.og_ndt_derivatives.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.og_ndt_hessian.restype = C.c_int
        L.og_ndt_hessian.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.og_ndt_align.restype = C.c_int
        L.og_ndt_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(NdtResult)]
        L.og_ndt_pose_to_matrix.argtypes = [C.c_void_p, C.c_void_p]
        L.og_ndt_euler_xyz.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_params(**kw):
    p = GicpParams()
    lib().og_gicp_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _as_cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


def gicp_align(src, tgt, params=None, guess=None, src_normal_off=-1, tgt_normal_off=-1,
               want_cov=False, want_aligned=False):
    """src/tgt: (n, stride) float32 arrays with xyz in columns 0..2."""
    src = _as_cloud(src); tgt = _as_cloud(tgt)
    params = params or default_params()
    res = GicpResult()
    g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
    sc = np.zeros((src.shape[0], 9)) if want_cov else None
    tc = np.zeros((tgt.shape[0], 9)) if want_cov else None
    al = np.zeros((src.shape[0], 3), dtype=np.float32) if want_aligned else None
    rc = lib().og_gicp_align(_p(src), src.shape[0], src.shape[1], src_normal_off,
                             _p(tgt), tgt.shape[0], tgt.shape[1], tgt_normal_off,
                             C.byref(params), _p(g), C.byref(res), _p(sc), _p(tc), _p(al))
    out = {
        "status": rc,
        "T": np.array(res.final_transformation, dtype=np.float32).reshape(4, 4),
        "iterations": res.nr_iterations, "converged": bool(res.converged),
        "n_corr": res.n_correspondences, "delta": res.delta,
        "n_evals": res.n_fdf_evals, "n_inner": res.n_inner_iterations,
        "t_cov": res.t_covariances_s, "t_iter": res.t_iterations_s, "t_total": res.t_total_s,
        "t_lookups": res.t_lookups_s, "t_opt": res.t_optimization_s,
    }
    if want_cov:
        out["src_cov"] = sc.reshape(-1, 3, 3); out["tgt_cov"] = tc.reshape(-1, 3, 3)
    if want_aligned:
        out["aligned"] = al
    return out


def _result_dict(rc, res):
    return {
        "status": rc,
        "T": np.array(res.final_transformation, dtype=np.float32).reshape(4, 4),
        "iterations": res.nr_iterations, "converged": bool(res.converged),
        "n_corr": res.n_correspondences, "delta": res.delta,
        "n_evals": res.n_fdf_evals, "n_inner": res.n_inner_iterations,
        "t_cov": res.t_covariances_s, "t_iter": res.t_iterations_s, "t_total": res.t_total_s,
        "t_lookups": res.t_lookups_s, "t_opt": res.t_optimization_s,
    }


class PreparedTarget:
    """setInputTarget once, align many sources against it (the reference keeps the target's kd-tree and covariances
    until the next setInputTarget): the unchanged-submap case of BASELINE configs[2]."""

    def __init__(self, tgt, params, tgt_normal_off=-1, cov=None):
        """cov (n, 3, 3): use these target covariances instead of computing them from the target as it is now"""
        self.tgt = _as_cloud(tgt)
        self.h = lib().og_gicp_target_prepare(_p(self.tgt), self.tgt.shape[0], self.tgt.shape[1], tgt_normal_off,
                                              C.byref(params))
        if not self.h:
            raise ValueError("og_gicp_target_prepare: empty target or fewer points than k_correspondences")
        if cov is not None:
            c = np.ascontiguousarray(cov, dtype=np.float64).reshape(-1, 9)
            assert c.shape[0] == self.tgt.shape[0]
            lib().og_gicp_target_set_covariances(self.h, _p(c))

    def align(self, src, params, guess=None, src_normal_off=-1):
        src = _as_cloud(src)
        res = GicpResult()
        g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
        rc = lib().og_gicp_align_prepared(_p(src), src.shape[0], src.shape[1], src_normal_off, self.h, C.byref(params),
                                          _p(g), C.byref(res))
        return _result_dict(rc, res)

    def close(self):
        if getattr(self, "h", None):
            lib().og_gicp_target_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def set_sum_chunk(c):
    """summation-order probe (og_set_sum_chunk): 0 = the reference's serial order"""
    lib().og_set_sum_chunk.argtypes = [C.c_int]
    lib().og_set_sum_chunk(int(c))


def covariances(pts, k=20, eps=1e-3, num_threads=1):
    pts = _as_cloud(pts)
    out = np.zeros((pts.shape[0], 9))
    rc = lib().og_gicp_covariances(_p(pts), pts.shape[0], pts.shape[1], k, eps, num_threads, _p(out))
    if rc:
        raise RuntimeError("og_gicp_covariances rc=%d" % rc)
    return out.reshape(-1, 3, 3)


def fitness(src, tgt, T, max_range=float(np.finfo(np.float64).max), num_threads=1):
    src = _as_cloud(src); tgt = _as_cloud(tgt)
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
    return lib().og_gicp_fitness(_p(src), src.shape[0], src.shape[1], _p(tgt), tgt.shape[0], tgt.shape[1],
                                 _p(T), max_range, num_threads)


def fdf(src4, tgt4, M, x):
    src4 = np.ascontiguousarray(src4, dtype=np.float32); tgt4 = np.ascontiguousarray(tgt4, dtype=np.float32)
    M = np.ascontiguousarray(M, dtype=np.float64).reshape(-1, 9)
    x = np.ascontiguousarray(x, dtype=np.float64)
    f = C.c_double(); g = np.zeros(6)
    lib().og_gicp_fdf(_p(src4), _p(tgt4), _p(M), src4.shape[0], _p(x), C.byref(f), _p(g))
    return f.value, g


def apply_state(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    T = np.zeros(16, dtype=np.float32)
    lib().og_gicp_apply_state(_p(x), _p(T))
    return T.reshape(4, 4)


class KdTree:
    def __init__(self, pts):
        self.pts = _as_cloud(pts)
        self.h = lib().og_kdtree_build(_p(self.pts), self.pts.shape[0], self.pts.shape[1])

    def __del__(self):
        if getattr(self, "h", None):
            lib().og_kdtree_free(self.h); self.h = None

    def knn(self, q, k):
        q = np.ascontiguousarray(q, dtype=np.float32)
        idx = np.zeros(k, dtype=np.int32); d2 = np.zeros(k, dtype=np.float32)
        c = lib().og_kdtree_knn(self.h, _p(q), k, _p(idx), _p(d2))
        return idx[:c], d2[:c]

    def radius(self, q, radius2, cap=64):
        """every point with d2 < radius2 (strict), ascending (d2, index): (indices, d2)"""
        q = np.ascontiguousarray(q, dtype=np.float32)
        idx = np.zeros(cap, np.int32); d2 = np.zeros(cap, np.float32)
        n = lib().og_kdtree_radius(self.h, _p(q), np.float32(radius2), _p(idx), _p(d2), cap)
        n = min(n, cap)
        return idx[:n].copy(), d2[:n].copy()

    def nn_batch(self, q, num_threads=1):
        q = _as_cloud(q)
        idx = np.zeros(q.shape[0], dtype=np.int32); d2 = np.zeros(q.shape[0], dtype=np.float32)
        lib().og_kdtree_nn_batch(self.h, _p(q), q.shape[0], q.shape[1], _p(idx), _p(d2), num_threads)
        return idx, d2


def bfgs_quadratic(A, b, x0, max_iters=100, grad_tol=1e-8):
    A = np.ascontiguousarray(A, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    it = lib().og_bfgs_minimize_quadratic(_p(A), _p(b), A.shape[0], _p(x), max_iters, grad_tol)
    return x, it


def voxel_filter(blob, point_step, leaf, x_off=0, y_off=4, z_off=8, float_fields=None,
                 filter_field_offset=-1, limit_min=-3.4028234663852886e38, limit_max=3.4028234663852886e38,
                 negative=False, min_points_per_voxel=0, downsample_all_data=True, body=None):
    """blob: uint8 array of n*point_step bytes.  Returns dict(out, voxel_idx, first_pt, count, min_b, div_b, rc)."""
    blob = np.ascontiguousarray(blob, dtype=np.uint8).reshape(-1)
    n = blob.size // point_step
    if float_fields is None:
        float_fields = [x_off, y_off, z_off]
    ffo = np.array(float_fields, dtype=np.uint32)
    P = VoxelParams()
    if np.isscalar(leaf):
        leaf = (leaf, leaf, leaf)
    P.leaf[0], P.leaf[1], P.leaf[2] = [np.float32(v) for v in leaf]
    P.filter_field_offset = filter_field_offset
    P.filter_limit_min = limit_min; P.filter_limit_max = limit_max
    P.filter_limit_negative = int(negative); P.min_points_per_voxel = min_points_per_voxel
    P.downsample_all_data = int(downsample_all_data)
    if body is not None:          # (min3, max3, rotation_z): the BodyFilter nodelet ahead of the voxel grid
        P.body_enabled = 1
        for i in range(3):
            P.body_min[i] = np.float32(body[0][i]); P.body_max[i] = np.float32(body[1][i])
        P.body_rotation = np.float32(body[2])
    out = np.zeros(max(n, 1) * point_step, dtype=np.uint8)
    vidx = np.zeros(max(n, 1), dtype=np.int32); first = np.zeros(max(n, 1), dtype=np.int32)
    cnt = np.zeros(max(n, 1), dtype=np.int32)
    min_b = np.zeros(3, dtype=np.int32); div_b = np.zeros(3, dtype=np.int32)
    n_out = C.c_size_t(0)
    rc = lib().og_voxel_filter(_p(blob), n, point_step, x_off, y_off, z_off, _p(ffo), len(ffo), C.byref(P),
                               _p(out), C.byref(n_out), _p(vidx), _p(first), _p(cnt), _p(min_b), _p(div_b))
    m = n_out.value
    return {"rc": rc, "out": out[: m * point_step].reshape(m, point_step), "voxel_idx": vidx[:m],
            "first_pt": first[:m], "count": cnt[:m], "min_b": min_b, "div_b": div_b}


def normalize_pcloud(xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.zeros_like(xyz)
    lib().og_normalize_pcloud(_p(xyz), xyz.shape[0], _p(out))
    return out


def compute_ap(query_xyz, ref_normals, correspondences):
    q = np.ascontiguousarray(query_xyz, dtype=np.float32).reshape(-1, 3)
    nr = np.ascontiguousarray(ref_normals, dtype=np.float32).reshape(-1, 3)
    co = np.ascontiguousarray(correspondences, dtype=np.int64)
    Ap = np.zeros(36)
    lib().og_compute_ap(_p(q), q.shape[0], _p(nr), _p(co), _p(Ap))
    return Ap.reshape(6, 6)


def normals_knn(pts, k=20, viewpoint=(0.0, 0.0, 0.0), num_threads=8):
    """point_cloud_filter::NormalComputation, k-NN mode (normals_oracle.c).  pts: (n, >=3) float32.
    Returns (n, 4) float32: nx, ny, nz, curvature."""
    p = _as_cloud(pts)
    vp = np.ascontiguousarray(viewpoint, dtype=np.float32)
    out = np.zeros((p.shape[0], 4), dtype=np.float32)
    lib().og_normals_knn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    rc = lib().og_normals_knn(_p(p), p.shape[0], p.shape[1], int(k), _p(vp), _p(out), int(num_threads))
    if rc != 0:
        raise ValueError("og_normals_knn: need 3 <= k <= n")
    return out


def normals_radius(pts, radius, viewpoint=(0.0, 0.0, 0.0), num_threads=8):
    """point_cloud_filter::NormalComputation, radius mode + NaN-normal removal (normals_oracle.c).
    Returns (out4 (n, 4) float32 with NaN rows where a point has fewer than 3 neighbours, valid_idx int32)."""
    p = _as_cloud(pts)
    vp = np.ascontiguousarray(viewpoint, dtype=np.float32)
    out = np.zeros((p.shape[0], 4), dtype=np.float32)
    vi = np.zeros(max(p.shape[0], 1), dtype=np.int32)
    lib().og_normals_radius.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib().og_normals_radius.restype = C.c_int
    m = lib().og_normals_radius(_p(p), p.shape[0], p.shape[1], float(radius), _p(vp), _p(out), _p(vi), int(num_threads))
    if m < 0:
        raise ValueError("og_normals_radius failed")
    return out, vi[:m]


# ---------------------------------------------------------------- NDT (row f4; ndt_oracle.c -- parity unpinned)
def ndt_params(**kw):
    p = NdtParams()
    lib().og_ndt_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class NdtTarget:
    """setInputTarget of pclomp::NormalDistributionsTransform: the target's voxel Gaussians + centroid kd-tree."""

    def __init__(self, tgt, params=None):
        self.tgt = _as_cloud(tgt)
        self.params = params or ndt_params()
        self.h = lib().og_ndt_target_build(_p(self.tgt), self.tgt.shape[0], self.tgt.shape[1], C.byref(self.params))
        nv = C.c_int(); na = C.c_int()
        self.min_b = np.zeros(3, np.int32); self.div_b = np.zeros(3, np.int32)
        self.status = lib().og_ndt_target_info(self.h, C.byref(nv), C.byref(na), _p(self.min_b), _p(self.div_b))
        self.n_valid, self.n_all = nv.value, na.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().og_ndt_target_free(self.h); self.h = None

    def leaves(self):
        n = self.n_valid
        out = {"leaf_idx": np.zeros(n, np.int32), "nr_points": np.zeros(n, np.int32), "mean": np.zeros((n, 3)),
               "icov": np.zeros((n, 9)), "centroid": np.zeros((n, 3), np.float32)}
        lib().og_ndt_target_leaves(self.h, _p(out["leaf_idx"]), _p(out["nr_points"]), _p(out["mean"]), _p(out["icov"]),
                                   _p(out["centroid"]))
        return out

    def derivatives(self, src, T, pose, compute_hessian=True):
        src = _as_cloud(src)
        T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        score = C.c_double(); g = np.zeros(6); H = np.zeros((6, 6))
        lib().og_ndt_derivatives(self.h, _p(src), src.shape[0], src.shape[1], _p(T), _p(pose), int(compute_hessian),
                                 C.byref(score), _p(g), _p(H))
        return score.value, g, H

    def hessian(self, src, T, pose):
        src = _as_cloud(src)
        T = np.ascontiguousarray(T, dtype=np.float32).reshape(16)
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        H = np.zeros((6, 6))
        lib().og_ndt_hessian(self.h, _p(src), src.shape[0], src.shape[1], _p(T), _p(pose), _p(H))
        return H

    def align(self, src, guess=None):
        src = _as_cloud(src)
        g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
        r = NdtResult()
        rc = lib().og_ndt_align(self.h, _p(src), src.shape[0], src.shape[1], _p(g), C.byref(r))
        return {"status": rc, "T": np.array(r.final_transformation, dtype=np.float32).reshape(4, 4),
                "converged": bool(r.converged), "iterations": r.nr_iterations, "trans_probability": r.trans_probability,
                "evaluations": r.n_evaluations, "pose": np.array(r.pose)}


def ndt_pose_to_matrix(pose):
    pose = np.ascontiguousarray(pose, dtype=np.float64); T = np.zeros(16, np.float32)
    lib().og_ndt_pose_to_matrix(_p(pose), _p(T))
    return T.reshape(4, 4)


def ndt_euler_xyz(T):
    T = np.ascontiguousarray(T, dtype=np.float32).reshape(16); e = np.zeros(3, np.float32)
    lib().og_ndt_euler_xyz(_p(T), _p(e))
    return e
