"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/phovo_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under the product package does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libphovo_oracle.so")
MAXL = 10


class Config(C.Structure):
    """POD mirror of `phovo_config` (include/phovo_b200.h)."""
    _fields_ = [
        ("mode", C.c_int32), ("num_levels", C.c_int32),
        ("blur_filter_size", C.c_int32 * MAXL), ("max_num_iterations", C.c_int32 * MAXL),
        ("grad_scale", C.c_double * MAXL), ("lambda_step", C.c_double * MAXL),
        ("min_gradient_norm", C.c_double * MAXL),
        ("min_depth", C.c_double), ("max_depth", C.c_double),
        ("visualize_iterations", C.c_int32),
        ("function_tolerance", C.c_double * MAXL), ("gradient_tolerance", C.c_double * MAXL),
        ("parameter_tolerance", C.c_double * MAXL),
        ("initial_trust_region_radius", C.c_double * MAXL),
        ("max_trust_region_radius", C.c_double * MAXL),
        ("min_trust_region_radius", C.c_double * MAXL),
        ("min_relative_decrease", C.c_double * MAXL),
        ("num_threads", C.c_int32), ("num_linear_solver_threads", C.c_int32),
        ("minimizer_progress_to_stdout", C.c_int32), ("reserved", C.c_int32),
    ]


class IterStats(C.Structure):
    """POD mirror of `phovo_iter_stats`."""
    _fields_ = [
        ("level", C.c_int32), ("iteration", C.c_int32), ("num_valid", C.c_int32), ("accepted", C.c_int32),
        ("H", C.c_double * 21), ("g", C.c_double * 6), ("grad_norm", C.c_double), ("cost", C.c_double),
        ("radius", C.c_double), ("state_in", C.c_double * 6), ("state_out", C.c_double * 6),
    ]

    def as_dict(self):
        return dict(level=self.level, iteration=self.iteration, num_valid=self.num_valid,
                    accepted=self.accepted, H=np.array(self.H[:]), g=np.array(self.g[:]),
                    grad_norm=self.grad_norm, cost=self.cost, radius=self.radius,
                    state_in=np.array(self.state_in[:]), state_out=np.array(self.state_out[:]))


def build(force=False):
    src = os.path.join(_HERE, "phovo_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.pho_create.restype = vp
        L.pho_destroy.argtypes = [vp]
        L.pho_set_config.argtypes = [vp, C.POINTER(Config)]
        L.pho_set_options.argtypes = [vp, C.c_int, C.c_int]
        L.pho_set_intrinsics.argtypes = [vp, dp]
        L.pho_set_source.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_int, C.c_int]
        L.pho_set_target.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int]
        L.pho_set_initial_state.argtypes = [vp, dp]
        L.pho_optimize.argtypes = [vp]
        L.pho_get_state.argtypes = [vp, dp]
        L.pho_get_rt.argtypes = [vp, dp]
        L.pho_num_iter_stats.argtypes = [vp]
        L.pho_get_iter_stats.argtypes = [vp, C.c_int, C.POINTER(IterStats)]
        L.pho_level_image.argtypes = [vp, C.c_int, C.c_int, ip, ip]
        L.pho_level_image.restype = dp
        L.pho_eval.argtypes = [vp, C.c_int, dp, C.POINTER(IterStats), vp, vp]
        L.pho_winner_map.argtypes = [vp, C.c_int, dp, vp]
        L.pho_convert_u8.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, vp]
        L.pho_level_size.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
        L.pho_resize_level.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.pho_gaussian_blur.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double]
        L.pho_scharr.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, vp]
        L.pho_align_batch.argtypes = [C.POINTER(Config), dp, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                      C.c_int, C.c_int, vp, vp, dp]
        L.pho_align_batch.restype = C.c_double
        L.pho_state_to_rt.argtypes = [dp, dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_config(mode=0, num_levels=5, max_iters=(0, 0, 5, 20, 50), grad_scale=0.0625, lam=1.0,
                min_grad_norm=300.0, blur=0, min_depth=0.3, max_depth=5.0, **ceres):
    """Config with reference constructor defaults (AN:430-443); scalars broadcast over levels."""
    cfg = Config()
    cfg.mode, cfg.num_levels = mode, num_levels
    cfg.min_depth, cfg.max_depth = min_depth, max_depth

    def fill(arr, v, cast):
        vals = list(v) if hasattr(v, "__len__") else [v] * MAXL
        for i in range(MAXL):
            arr[i] = cast(vals[i] if i < len(vals) else vals[-1])
    fill(cfg.max_num_iterations, list(max_iters) + [0] * (MAXL - len(max_iters)), int)
    fill(cfg.grad_scale, grad_scale, float)
    fill(cfg.lambda_step, lam, float)
    fill(cfg.min_gradient_norm, min_grad_norm, float)
    fill(cfg.blur_filter_size, blur, int)
    defaults = dict(function_tolerance=1e-4, gradient_tolerance=1e-3, parameter_tolerance=1e-6,
                    initial_trust_region_radius=1e4, max_trust_region_radius=1e8,
                    min_trust_region_radius=1e-32, min_relative_decrease=1e-3)
    defaults.update(ceres)
    for k, v in defaults.items():
        fill(getattr(cfg, k), v, float)
    return cfg


class Oracle:
    """Mirrors the reference call sequence (ReadConfigurationFile/SetIntrinsicMatrix/SetSourceFrame/
    SetTargetFrame/SetInitialStateVector/Optimize/GetOptimalStateVector)."""

    def __init__(self, cfg, K, storage_f32=False, lean=False):
        self.L = lib()
        self.h = self.L.pho_create()
        self.cfg = cfg
        self.L.pho_set_config(self.h, C.byref(cfg))
        self.L.pho_set_options(self.h, int(storage_f32), int(lean))   # storage_f32: 0 | 1 (all fp32) | 2 (fp32, depth fp64)
        self.K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        self.L.pho_set_intrinsics(self.h, _d(self.K))

    def __del__(self):
        try:
            self.L.pho_destroy(self.h)
        except Exception:
            pass

    def set_source(self, gray, depth):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float64)
        r, c = gray.shape
        self.shape = (r, c)
        self.L.pho_set_source(self.h, gray.ctypes.data, gray.strides[0], depth.ctypes.data, depth.strides[0], r, c)

    def set_target(self, gray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        r, c = gray.shape
        self.L.pho_set_target(self.h, gray.ctypes.data, gray.strides[0], r, c)

    def set_initial_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        self.L.pho_set_initial_state(self.h, _d(s))

    def optimize(self):
        self.L.pho_optimize(self.h)

    def state(self):
        s = np.zeros(6)
        self.L.pho_get_state(self.h, _d(s))
        return s

    def rt(self):
        m = np.zeros(16)
        self.L.pho_get_rt(self.h, _d(m))
        return m.reshape(4, 4)

    def iter_stats(self):
        out = []
        for i in range(self.L.pho_num_iter_stats(self.h)):
            s = IterStats()
            self.L.pho_get_iter_stats(self.h, i, C.byref(s))
            out.append(s.as_dict())
        return out

    def level_image(self, which, level):
        r, c = C.c_int32(), C.c_int32()
        p = self.L.pho_level_image(self.h, which, level, C.byref(r), C.byref(c))
        return np.ctypeslib.as_array(p, shape=(r.value, c.value)).copy()

    def eval(self, level, state, want_residuals=False, want_jacobian=False):
        state = np.ascontiguousarray(state, dtype=np.float64)
        r, c = self.level_image(0, level).shape
        res = np.zeros(r * c) if want_residuals else None
        jac = np.zeros((r * c, 6)) if want_jacobian else None
        s = IterStats()
        self.L.pho_eval(self.h, level, _d(state), C.byref(s),
                        res.ctypes.data if res is not None else None,
                        jac.ctypes.data if jac is not None else None)
        d = s.as_dict()
        d["residuals"], d["jacobian"] = res, jac
        return d

    def winner_map(self, level, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        r, c = self.level_image(0, level).shape
        w = np.zeros(r * c, dtype=np.int32)
        self.L.pho_winner_map(self.h, level, _d(state), w.ctypes.data)
        return w


def align_batch(cfg, K, gray0, depth0, gray1, num_threads=1, lean=False):
    """CPU baseline helper: returns (states[P,6], iterations[P,MAXL], wall_seconds, optimize_seconds)."""
    L = lib()
    gray0 = np.ascontiguousarray(gray0, dtype=np.uint8)
    gray1 = np.ascontiguousarray(gray1, dtype=np.uint8)
    depth0 = np.ascontiguousarray(depth0, dtype=np.float64)
    P, r, c = gray0.shape
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    states = np.zeros((P, 6))
    iters = np.zeros((P, MAXL), dtype=np.int32)
    opt = C.c_double(0)
    wall = L.pho_align_batch(C.byref(cfg), _d(K), P, r, c, gray0.ctypes.data, depth0.ctypes.data,
                             gray1.ctypes.data, num_threads, int(lean), states.ctypes.data,
                             iters.ctypes.data, C.byref(opt))
    return states, iters, wall, opt.value


# thin wrappers over the stand-alone image ops (used by the cv2 pin tests)
def convert_u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros(img.shape)
    lib().pho_convert_u8(img.ctypes.data, img.strides[0], img.shape[0], img.shape[1], out.ctypes.data)
    return out


def level_size(rows, cols, level):
    r, c = C.c_int32(), C.c_int32()
    lib().pho_level_size(rows, cols, level, C.byref(r), C.byref(c))
    return r.value, c.value


def resize_level(img, level):
    img = np.ascontiguousarray(img, dtype=np.float64)
    r, c = level_size(img.shape[0], img.shape[1], level)
    out = np.zeros((r, c))
    lib().pho_resize_level(img.ctypes.data, img.shape[0], img.shape[1], level, out.ctypes.data)
    return out


def gaussian_blur(img, ksize, sigma=3.0):
    out = np.ascontiguousarray(img, dtype=np.float64).copy()
    lib().pho_gaussian_blur(out.ctypes.data, out.shape[0], out.shape[1], ksize, sigma)
    return out


def scharr(img, dx, dy, scale):
    img = np.ascontiguousarray(img, dtype=np.float64)
    out = np.zeros(img.shape)
    lib().pho_scharr(img.ctypes.data, img.shape[0], img.shape[1], dx, dy, scale, out.ctypes.data)
    return out
