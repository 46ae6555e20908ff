"""The reference's named configurations as data (values of config_files/*.yml in the reference
tree, cited per entry), plus a writer for the OpenCV-FileStorage YAML dialect its
ReadConfigurationFile expects.  /root/reference does not exist on the GPU box, so tests and
bench.py take their configurations from here; tests/test_config.py checks this table against the
reference's own files whenever the reference tree is present.
"""
import os

# key order and spelling as read by CPhotoconsistencyOdometryAnalytic.h:586-606 / Ceres.h:531-575
K_LEVELS = "numOptimizationLevels"
K_BLUR = "blurFilterSize (at each level)"
K_GRAD = "imageGradientsScalingFactor (at each level)"
K_LAMBDA = "lambda_optimization_step (at each level)"
K_ITERS = "max_num_iterations (at each level)"
K_MINGRAD = "min_gradient_norm (at each level)"
K_FTOL = "function_tolerance (at each level)"
K_GTOL = "gradient_tolerance (at each level)"
K_PTOL = "parameter_tolerance (at each level)"
K_R0 = "initial_trust_region_radius (at each level)"
K_RMAX = "max_trust_region_radius (at each level)"
K_RMIN = "min_trust_region_radius (at each level)"
K_ETA = "min_relative_decrease (at each level)"


def _analytic(levels, iters, mingrad, blur=None, viz=0):
    n = len(iters)
    return {K_LEVELS: levels, K_BLUR: blur or [0] * n, K_GRAD: [0.0625] * n, K_LAMBDA: [1] * n,
            K_ITERS: iters, K_MINGRAD: mingrad, "visualizeIterations": viz}


def _ceres(levels, blur, grad, iters, ftol, gtol, ptol, r0, rmax, rmin, eta, nt=2, nlt=2, progress=0):
    return {K_LEVELS: levels, K_BLUR: blur, K_GRAD: grad, K_ITERS: iters, K_FTOL: ftol, K_GTOL: gtol,
            K_PTOL: ptol, K_R0: r0, K_RMAX: rmax, K_RMIN: rmin, K_ETA: eta, "num_threads": nt,
            "num_linear_solver_threads": nlt, "minimizer_progress_to_stdout": progress, "visualizeIterations": 0}


REFERENCE_CONFIGS = {
    # config_files/config_4_level_optimization_analytic.yml:2-8  (BASELINE configs 1 and 4)
    "config_4_level_optimization_analytic": _analytic(4, [0, 0, 20, 50], [300] * 4),
    # config_files/config_5_level_optimization_analytic.yml:2-8  (BASELINE config 2)
    "config_5_level_optimization_analytic": _analytic(5, [0, 0, 5, 20, 50], [300] * 5),
    # config_files/config_6_level_optimization_analytic.yml:2-8  (BASELINE config 5)
    "config_6_level_optimization_analytic": _analytic(6, [0, 0, 5, 20, 50, 50], [100, 100, 100, 100, 100, 10]),
    # config_files/config_only_level_0_analytic.yml:2-8 (more entries than levels; visualisation on)
    "config_only_level_0_analytic": _analytic(1, [5000, 0, 0, 0], [300] * 4, viz=1),
    # config_files/config_5_level_optimization_ceres.yml:2-16  (BASELINE config 3; 4 radii for 5 levels)
    "config_5_level_optimization_ceres": _ceres(
        5, [0] * 5, [0.0625] * 5, [2, 2, 5, 10, 50], [1e-4] * 5, [1e-3] * 5, [1e-4, 1e-4, 1e-6, 1e-6, 1e-6],
        [1e8, 1e4, 1e4, 1e4, 1e4], [1e8] * 5, [1e-32] * 4, [1e-3] * 5, progress=1),
    # config_files/config_4_level_optimization_ceres.yml:2-16
    "config_4_level_optimization_ceres": _ceres(
        4, [0] * 4, [0.0625] * 4, [2, 4, 5, 50], [1e-4] * 4, [1e-3] * 4, [1e-4, 1e-4, 1e-6, 1e-6],
        [1e8, 1e4, 1e4, 1e4], [1e8] * 4, [1e-32] * 3, [1e-3] * 4, progress=1),
    # config_files/config_3_level_optimization_ceres.yml:2-16 (blur 3 at level 2)
    "config_3_level_optimization_ceres": _ceres(
        3, [0, 0, 3], [1, 2, 0.005], [50, 50, 50], [1e-4, 1e-5, 1e-6], [1e-3] * 3, [1e-4, 1e-4, 1e-10],
        [1e8, 1e8, 1e4], [1e20, 1e20, 1e8], [1e-32] * 3, [1e-1, 1e-2, 1e-3]),
    # config_files/config_only_level_0_ceres.yml:2-16
    "config_only_level_0_ceres": _ceres(
        1, [0] * 4, [0.006, 0.01, 0.01, 0.01], [50, 0, 0, 0], [1e-4] * 4, [1e-3] * 4, [1e-4, 1e-4, 1e-6, 1e-6],
        [1e2, 1e4, 1e4, 1e4], [1e8] * 4, [1e-32] * 3, [1e-4, 1e-3, 1e-3, 1e-3], nlt=1),
    # config_files/config_only_level_1_ceres.yml:2-16 (blur 5 at level 1)
    "config_only_level_1_ceres": _ceres(
        2, [0, 5, 3], [0.5, 0.5, 0.0625], [0, 40, 0], [1e-4] * 3, [1e-3] * 3, [1e-4, 1e-4, 1e-6],
        [1e8, 1e4, 1e4], [1e8] * 3, [1e-32] * 3, [1e-1, 1e-1, 1e-3]),
    # config_files/config_only_level_2_ceres.yml:2-16
    "config_only_level_2_ceres": _ceres(
        3, [0, 5, 3], [0.5, 0.5, 0.0625], [0, 0, 50], [1e-4] * 3, [1e-3] * 3, [1e-4, 1e-4, 1e-6],
        [1e8, 1e4, 1e4], [1e8] * 3, [1e-32] * 3, [1e-1, 1e-1, 1e-3]),
    # config_files/config_only_level_3_ceres.yml:2-16
    "config_only_level_3_ceres": _ceres(
        4, [0] * 4, [0.5, 0.5, 1, 0.01], [0, 0, 0, 50], [1e-4] * 4, [1e-3] * 4, [1e-4, 1e-4, 1e-6, 1e-6],
        [1e8, 1e4, 1e4, 1e4], [1e8] * 4, [1e-32] * 3, [1e-1, 1e-1, 1e-2, 1e-3], nlt=1),
    # config_files/config_only_level_4_ceres.yml:2-16
    "config_only_level_4_ceres": _ceres(
        5, [0] * 5, [0.0625] * 5, [0, 0, 0, 0, 50], [1e-4] * 5, [1e-3] * 5, [1e-4, 1e-4, 1e-6, 1e-6, 1e-6],
        [1e8, 1e4, 1e4, 1e4, 1e4], [1e8] * 5, [1e-32] * 4, [1e-3] * 5),
}


# NOT from the reference tree: extra configurations used by the parity tests (every level active,
# lambda != 1, low thresholds so that levels run to their iteration caps).
EXTRA_CONFIGS = {
    "test_3_level_all_active": _analytic(3, [3, 6, 10], [1., 1., 1.]),
}
EXTRA_CONFIGS["test_3_level_all_active"][K_LAMBDA] = [0.8, 1, 0.9]


def _lookup(name):
    return REFERENCE_CONFIGS[name] if name in REFERENCE_CONFIGS else EXTRA_CONFIGS[name]


def to_yaml(values):
    """OpenCV FileStorage YAML 1.0 text for a config dict."""
    lines = ["%YAML:1.0"]
    for k, v in values.items():
        if isinstance(v, (list, tuple)):
            lines.append("%s: [%s]" % (k, ", ".join(repr(x) if isinstance(x, float) else str(x) for x in v)))
        else:
            lines.append("%s: %s" % (k, v))
    return "\n".join(lines) + "\n"


def write_yaml(name, directory):
    path = os.path.join(directory, name + ".yml")
    with open(path, "w") as f:
        f.write(to_yaml(_lookup(name)))
    return path


def to_config(name, capi, mode=None):
    """capi.Config for a named reference configuration without going through a file."""
    v = _lookup(name)
    cfg = capi.default_config() if hasattr(capi, "default_config") else capi.Config()
    n = capi.MAXL

    def fill(arr, vals, cast):
        if vals is None:
            return
        for i in range(n):
            arr[i] = cast(vals[i] if i < len(vals) else vals[-1])
    cfg.num_levels = v[K_LEVELS]
    fill(cfg.blur_filter_size, v.get(K_BLUR), int)
    fill(cfg.grad_scale, v.get(K_GRAD), float)
    fill(cfg.lambda_step, v.get(K_LAMBDA), float)
    fill(cfg.max_num_iterations, v.get(K_ITERS), int)
    fill(cfg.min_gradient_norm, v.get(K_MINGRAD), float)
    fill(cfg.function_tolerance, v.get(K_FTOL), float)
    fill(cfg.gradient_tolerance, v.get(K_GTOL), float)
    fill(cfg.parameter_tolerance, v.get(K_PTOL), float)
    fill(cfg.initial_trust_region_radius, v.get(K_R0), float)
    fill(cfg.max_trust_region_radius, v.get(K_RMAX), float)
    fill(cfg.min_trust_region_radius, v.get(K_RMIN), float)
    fill(cfg.min_relative_decrease, v.get(K_ETA), float)
    for l in range(cfg.num_levels, n):
        cfg.max_num_iterations[l] = 0
    is_ceres = K_FTOL in v
    cfg.mode = mode if mode is not None else (2 if is_ceres else 0)
    cfg.min_depth, cfg.max_depth = 0.3, 5.0
    return cfg
