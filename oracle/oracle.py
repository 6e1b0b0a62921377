"""ctypes binding of oracle/libea_oracle.so (CPU fp64 restatement of the reference).

TEST INFRASTRUCTURE ONLY.  Mirrors, on the CPU, what the reference computes:
standalone/utils.cpp:38-83,201-281 (preprocessing), standalone/utils.h:38-99
(EAResidue) and standalone/standalone_edge_align.cpp:256-301 (Ceres solve).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libea_oracle.so")

LOSS_TRIVIAL, LOSS_CAUCHY, LOSS_HUBER = 0, 1, 2
STRATEGY_LM, STRATEGY_DOGLEG = 0, 1
TERMINATION = {1: "CONVERGENCE_GRADIENT", 2: "CONVERGENCE_FUNCTION", 3: "CONVERGENCE_PARAMETER",
               4: "CONVERGENCE_MIN_RADIUS", 5: "NO_CONVERGENCE", 6: "FAILURE_EVAL_X0", 7: "FAILURE_INVALID_STEPS"}


class Options(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("jacobi_scaling", C.c_int),
                ("max_consecutive_invalid", C.c_int), ("loss_type", C.c_int), ("loss_scale", C.c_double),
                ("strategy", C.c_int), ("pad", C.c_int)]


class Summary(C.Structure):
    _fields_ = [("termination", C.c_int), ("iterations", C.c_int), ("accepted", C.c_int), ("rejected", C.c_int),
                ("n_residuals", C.c_int), ("jac_evals", C.c_int), ("res_evals", C.c_int), ("pad", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double)]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "pad"}
        d["termination_name"] = TERMINATION.get(self.termination, "?")
        return d


class PairCfg(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("n_levels", C.c_int), ("stride", C.c_int), ("thresh", C.c_int),
                ("use_median", C.c_int), ("norm_mode", C.c_int), ("pad", C.c_int),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("zscale", C.c_double)]


def build(force=False):
    """Compile libea_oracle.so with the committed Makefile (g++)."""
    src = os.path.join(_HERE, "ea_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.eo_align_batch.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def default_options(**kw):
    o = Options()
    lib().eo_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(type(o), k):
            raise TypeError('default_options: unknown field %r' % k)
        setattr(o, k, v)
    return o


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _stage(name, src, out_shape, out_dtype=np.uint8):
    src = _u8(src)
    h, w = src.shape[:2]
    out = np.empty(out_shape, out_dtype)
    getattr(lib(), name)(_p(src, C.c_uint8), w, h, out.ctypes.data_as(C.c_void_p))
    return out


def gaussian3(bgr): return _stage("eo_gaussian3_u8c3", bgr, bgr.shape)
def box3(bgr): return _stage("eo_box3_u8c3", bgr, bgr.shape)
def rgb2gray(bgr): return _stage("eo_rgb2gray", bgr, bgr.shape[:2])
def laplacian3_abs(gray): return _stage("eo_laplacian3_abs", gray, gray.shape)
def median3(mask): return _stage("eo_median3", mask, mask.shape)
def chamfer3_dt(mask): return _stage("eo_chamfer3_dt", mask, mask.shape, np.float32)
def half_linear(bgr): return _stage("eo_half_linear_u8c3", bgr, (bgr.shape[0] // 2, bgr.shape[1] // 2, 3))


def canny(img, t1, t2, l2=False):
    """cv::Canny(img, t1, t2, 3, l2) on 8UC1 or 8UC3."""
    img = _u8(img)
    h, w = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((h, w), np.uint8)
    lib().eo_canny(_p(img, C.c_uint8), w, h, cn, C.c_double(t1), C.c_double(t2), 1 if l2 else 0, _p(out, C.c_uint8))
    return out


def exact_edt(mask): return _stage("eo_exact_edt", mask, mask.shape, np.float32)


def half_nearest(depth):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = depth.shape
    out = np.empty((h // 2, w // 2), np.uint16)
    lib().eo_half_nearest_u16(_p(depth, C.c_uint16), w, h, _p(out, C.c_uint16))
    return out


def normalize_minmax(dt, alpha, beta):
    out = np.ascontiguousarray(dt, dtype=np.float32).copy()
    lib().eo_normalize_minmax(_p(out, C.c_float), out.size, C.c_double(alpha), C.c_double(beta))
    return out


def get_aX(bgr, depth, K, zscale=5000.0, thresh=35):
    """standalone/utils.cpp:201-281.  Returns (xyz [N,3] f64, uvd [N,3] i32) in row-major pixel order."""
    bgr = _u8(bgr); depth = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = depth.shape
    cap = w * h
    xyz = np.empty((cap, 3), np.float64); uvd = np.empty((cap, 3), np.int32)
    n = lib().eo_get_aX(_p(bgr, C.c_uint8), _p(depth, C.c_uint16), w, h, C.c_double(K[0]), C.c_double(K[1]),
                        C.c_double(K[2]), C.c_double(K[3]), C.c_double(zscale), thresh,
                        _p(xyz, C.c_double), _p(uvd, C.c_int), cap)
    return xyz[:n].copy(), uvd[:n].copy()


def get_distance_transform(bgr, thresh=35, use_median=1, norm_mode=1):
    """standalone/utils.cpp:38-83.  Returns (dt f32 [h,w], mask u8 [h,w] with 0 == edge)."""
    bgr = _u8(bgr)
    h, w = bgr.shape[:2]
    dt = np.empty((h, w), np.float32); mask = np.empty((h, w), np.uint8)
    lib().eo_get_distance_transform(_p(bgr, C.c_uint8), w, h, thresh, use_median, norm_mode, _p(dt, C.c_float),
                                    _p(mask, C.c_uint8))
    return dt, mask


def evaluate(xyz, dt, K, pose7, stride=1, options=None):
    """One Ceres-style evaluation (Jet autodiff + local parameterisation + loss corrector).

    Returns dict(residuals, raw, J, cost, b, H, ok)."""
    xyz = np.ascontiguousarray(xyz, np.float64); dt = np.ascontiguousarray(dt, np.float32)
    pose7 = np.ascontiguousarray(pose7, np.float64)
    h, w = dt.shape
    n = (len(xyz) + stride - 1) // stride
    r = np.empty(n); raw = np.empty(n); J = np.empty((n, 6)); sums = np.empty(28)
    o = options or default_options()
    rc = lib().eo_eval(_p(xyz, C.c_double), len(xyz), stride, _p(dt, C.c_float), w, h, C.c_double(K[0]),
                       C.c_double(K[1]), C.c_double(K[2]), C.c_double(K[3]), _p(pose7, C.c_double), C.byref(o),
                       _p(r, C.c_double), _p(raw, C.c_double), _p(J, C.c_double), _p(sums, C.c_double))
    H = np.zeros((6, 6)); k = 7
    for a in range(6):
        for c in range(a, 6):
            H[a, c] = H[c, a] = sums[k]; k += 1
    return dict(residuals=r, raw=raw, J=J, cost=sums[0], b=sums[1:7].copy(), H=H, sums=sums, ok=(rc == 0))


def evaluate_raw(xyz, dt, K, pose7, stride=1):
    xyz = np.ascontiguousarray(xyz, np.float64); dt = np.ascontiguousarray(dt, np.float32)
    pose7 = np.ascontiguousarray(pose7, np.float64)
    h, w = dt.shape
    n = (len(xyz) + stride - 1) // stride
    raw = np.empty(n); uv = np.empty((n, 2))
    rc = lib().eo_eval_raw(_p(xyz, C.c_double), len(xyz), stride, _p(dt, C.c_float), w, h, C.c_double(K[0]),
                           C.c_double(K[1]), C.c_double(K[2]), C.c_double(K[3]), _p(pose7, C.c_double),
                           _p(raw, C.c_double), _p(uv, C.c_double))
    return raw, uv, rc == 0


def solve(xyz, dt, K, pose7, stride=30, options=None, trace_cap=256):
    """ceres::Solve restatement (SEA:256-301).  Returns (pose7, summary dict, trace [k,7])."""
    xyz = np.ascontiguousarray(xyz, np.float64); dt = np.ascontiguousarray(dt, np.float32)
    pose = np.array(pose7, np.float64)
    h, w = dt.shape
    o = options or default_options()
    s = Summary(); tr = np.zeros((trace_cap, 7)); tn = C.c_int(0)
    lib().eo_solve(_p(xyz, C.c_double), len(xyz), stride, _p(dt, C.c_float), w, h, C.c_double(K[0]), C.c_double(K[1]),
                   C.c_double(K[2]), C.c_double(K[3]), _p(pose, C.c_double), C.byref(o), C.byref(s),
                   _p(tr, C.c_double), trace_cap, C.byref(tn))
    return pose, s.asdict(), tr[:tn.value].copy()


def pair_cfg(w, h, K, n_levels=1, stride=30, thresh=35, use_median=1, norm_mode=1, zscale=5000.0):
    return PairCfg(w, h, n_levels, stride, thresh, use_median, norm_mode, 0, K[0], K[1], K[2], K[3], zscale)


def align_pair(ref_bgr, ref_depth, now_bgr, cfg, pose7, options=None):
    ref_bgr = _u8(ref_bgr); now_bgr = _u8(now_bgr); ref_depth = np.ascontiguousarray(ref_depth, np.uint16)
    pose = np.array(pose7, np.float64)
    o = options or default_options()
    S = (Summary * cfg.n_levels)()
    lib().eo_align_pair(_p(ref_bgr, C.c_uint8), _p(ref_depth, C.c_uint16), _p(now_bgr, C.c_uint8), C.byref(cfg),
                        C.byref(o), _p(pose, C.c_double), S)
    return pose, [s.asdict() for s in S]


def align_batch(bgr, depth, ref_idx, now_idx, cfg, poses, options=None, n_threads=1, include_preprocess=True):
    """Multi-threaded batch (one pair per thread).  Returns (poses, summaries, seconds)."""
    bgr = _u8(bgr); depth = np.ascontiguousarray(depth, np.uint16)
    ref_idx = np.ascontiguousarray(ref_idx, np.int32); now_idx = np.ascontiguousarray(now_idx, np.int32)
    poses = np.array(poses, np.float64).reshape(-1, 7).copy()
    n = len(ref_idx)
    o = options or default_options()
    S = (Summary * (n * cfg.n_levels))()
    sec = lib().eo_align_batch(_p(bgr, C.c_uint8), _p(depth, C.c_uint16), n, _p(ref_idx, C.c_int), _p(now_idx, C.c_int),
                               C.byref(cfg), C.byref(o), _p(poses, C.c_double), S, n_threads,
                               1 if include_preprocess else 0)
    return poses, [s.asdict() for s in S], sec


def hardware_threads():
    return lib().eo_hardware_threads()


def quat_to_matrix(pose7):
    pose7 = np.ascontiguousarray(pose7, np.float64); T = np.empty(16)
    lib().eo_quat_to_matrix(_p(pose7, C.c_double), _p(T, C.c_double))
    return T.reshape(4, 4)


def quat_plus(x7, d6):
    x7 = np.ascontiguousarray(x7, np.float64); d6 = np.ascontiguousarray(d6, np.float64); out = np.empty(7)
    lib().eo_quat_plus(_p(x7, C.c_double), _p(d6, C.c_double), _p(out, C.c_double))
    return out


# ---- multi-camera problems / residual variants (standalone/utils.h:101-421) ------------------------------------
class ViewC(C.Structure):
    _fields_ = [("pts", C.POINTER(C.c_double)), ("n_total", C.c_int), ("stride", C.c_int), ("dt", C.POINTER(C.c_float)),
                ("w", C.c_int), ("h", C.c_int), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("use_dist", C.c_int), ("dist", C.c_double * 5), ("use_rig", C.c_int), ("T21", C.c_double * 16),
                ("T12", C.c_double * 16)]


def make_views(views):
    """views: list of dict(xyz, dt, K, stride=1, dist=None (k1,k2,p1,p2,k3), T21=None (4x4), T12=None)."""
    arr = (ViewC * len(views))()
    keep = []
    for i, v in enumerate(views):
        xyz = np.ascontiguousarray(v["xyz"], np.float64); dt = np.ascontiguousarray(v["dt"], np.float32)
        keep += [xyz, dt]
        a = arr[i]
        a.pts = _p(xyz, C.c_double); a.n_total = len(xyz); a.stride = int(v.get("stride", 1)); a.dt = _p(dt, C.c_float)
        a.h, a.w = dt.shape
        a.fx, a.fy, a.cx, a.cy = [float(k) for k in v["K"]]
        d = v.get("dist")
        a.use_dist = 0 if d is None else 1
        for k in range(5):
            a.dist[k] = 0.0 if d is None else float(d[k])
        T21 = v.get("T21")
        a.use_rig = 0 if T21 is None else 1
        t21 = np.eye(4) if T21 is None else np.asarray(T21, np.float64)
        t12 = np.linalg.inv(t21) if v.get("T12") is None else np.asarray(v["T12"], np.float64)
        for k in range(16):
            a.T21[k] = float(t21.flat[k]); a.T12[k] = float(t12.flat[k])
    return arr, keep


def evaluate_views(views, pose7, options=None):
    arr, keep = make_views(views)
    n = sum((len(v["xyz"]) + int(v.get("stride", 1)) - 1) // int(v.get("stride", 1)) for v in views)
    pose7 = np.ascontiguousarray(pose7, np.float64)
    r = np.empty(n); raw = np.empty(n); J = np.empty((n, 6)); sums = np.empty(28)
    o = options or default_options()
    rc = lib().eo_eval_views(arr, len(views), _p(pose7, C.c_double), C.byref(o), _p(r, C.c_double), _p(raw, C.c_double),
                             _p(J, C.c_double), _p(sums, C.c_double))
    H = np.zeros((6, 6)); k = 7
    for a in range(6):
        for c in range(a, 6):
            H[a, c] = H[c, a] = sums[k]; k += 1
    return dict(residuals=r, raw=raw, J=J, cost=sums[0], b=sums[1:7].copy(), H=H, sums=sums, ok=(rc == 0))


def solve_views(views, pose7, options=None):
    arr, keep = make_views(views)
    pose = np.array(pose7, np.float64)
    o = options or default_options()
    s = Summary()
    lib().eo_solve_views(arr, len(views), _p(pose, C.c_double), C.byref(o), C.byref(s))
    return pose, s.asdict()
