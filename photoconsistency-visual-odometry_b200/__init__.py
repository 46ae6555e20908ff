"""phovo-b200: B200-native (sm_100a) photoconsistency alignment hot path.

The directory name follows the reference repository's name; because it contains hyphens import it
with importlib:

    import importlib
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    odo = phovo.CPhotoconsistencyOdometryCuda()

Layout: csrc/ (CUDA kernels + the C ABI of include/phovo_b200.h), capi.py (ctypes binding),
odometry.py (host-side mirror of the reference's CPhotoconsistencyOdometry interface),
synth.py (deterministic synthetic RGB-D scene), dataset.py (recorded-sequence reader + the VO app's
loop), build.py (in-tree nvcc build).
"""
from . import capi, configs, dataset, sharded, synth  # noqa: F401
from .capi import (Config, IterStats, PhovoError, MODE_ANALYTIC_REF, MODE_ANALYTIC_FIXED,  # noqa: F401
                   MODE_CERES, MODE_BIOBJECTIVE, DEPTH_F64, DEPTH_F32, DEPTH_U16, MAXL, default_config,
                   parse_config_yaml, state_to_rt)
from .odometry import CPhotoconsistencyOdometryCuda  # noqa: F401
from .build import build  # noqa: F401
