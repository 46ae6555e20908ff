"""The third reference solver, CPhotoconsistencyOdometryBiObjective (photometric + depth rows with its
row aliasing, BiObjective.h:242-452), on the GPU (PHOVO_MODE_BIOBJECTIVE) against outputs of the
REFERENCE'S OWN header: tests/golden/ref_bi_*.npz (minted from oracle/_ref) and, when the prebuilt
library is present, live at 640x480."""
import glob
import os

import numpy as np
import pytest

from helpers import GOLDEN, assert_pose_close, g_rel_err, h_rel_err
from test_reference_source_pins import pack

pytestmark = pytest.mark.gpu
FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_bi_*.npz")))


def run(phovo, cfg_name, K, g0, d0, g1, d1, graph=None):
    cfg = phovo.configs.to_config(cfg_name, phovo.capi, mode=phovo.MODE_BIOBJECTIVE)
    odo = phovo.CPhotoconsistencyOdometryCuda()
    odo.SetConfig(cfg)
    odo.SetIntrinsicMatrix(K)
    if graph is not None:
        odo.SetUseGraph(graph)
    odo.SetSourceFrame(g0, d0)
    odo.SetTargetFrame(g1, d1)
    odo.SetInitialStateVector(np.zeros(6))
    odo.Optimize()
    return odo, odo.GetOptimalStateVector(), odo.IterationStats()


@pytest.mark.parametrize("name", FIXTURES)
def test_biobjective_matches_reference_golden(phovo, name):
    assert len(FIXTURES) >= 2
    gd = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    odo, s, log = run(phovo, str(gd["config"]), gd["K"], gd["gray0"], gd["depth0"], gd["gray1"], gd["depth1"])
    assert len(log) == len(gd["n"])                     # executed iterations, every level
    for e, n, H, g in zip(log, gd["n"], gd["H"], gd["g"]):
        assert h_rel_err(e["H"], pack(H)) < 1e-5 and g_rel_err(e["g"], g) < 1e-5      # north-star bar
        assert h_rel_err(e["H"], pack(H)) < 1e-9 and g_rel_err(e["g"], g) < 1e-8
    assert_pose_close(s, gd["state"], name)
    assert np.max(np.abs(s - gd["state"])) < 1e-9
    assert odo.LastPath() == 2          # default driver: persistent cooperative kernel per level
    # graph and plain-stream drivers: bitwise the same as each other, equal to rounding with the default
    _, s1, log1 = run(phovo, str(gd["config"]), gd["K"], gd["gray0"], gd["depth0"], gd["gray1"], gd["depth1"], graph=True)
    _, s2, log2 = run(phovo, str(gd["config"]), gd["K"], gd["gray0"], gd["depth0"], gd["gray1"], gd["depth1"], graph=False)
    assert np.array_equal(s1, s2) and len(log1) == len(log2) == len(log)
    assert np.max(np.abs(s1 - s)) < 1e-11
    # the batch entry (waves of per-pair slots, one CTA per pair through every level) on the same pair, twice in a batch
    odo.BatchSetDebugFlags(8)
    st, it = odo.BatchAlign(np.stack([gd["gray0"]] * 2), np.stack([gd["depth0"]] * 2), np.stack([gd["gray1"]] * 2), depth1=np.stack([gd["depth1"]] * 2))
    assert odo.BatchLastPath() == 3
    for p in range(2):
        assert int(it[p].sum()) == len(gd["n"])
        assert_pose_close(st[p], gd["state"], name + " (batch)")
        assert np.max(np.abs(st[p] - gd["state"])) < 1e-9


def test_biobjective_matches_reference_source_live_640x480(phovo, tmp_path):
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref library not present")
    K = phovo.synth.K_FRAME_ALIGNMENT
    name = "config_4_level_optimization_analytic"
    g0, d0, g1, d1 = phovo.synth.make_pair(480, 640, K=K, seed=39)
    yml = phovo.configs.write_yaml(name, str(tmp_path))
    sref, iters = ref_py.ReferenceBiObjective(yml, K).align(g0, d0, g1, d1)
    odo, s, log = run(phovo, name, K, g0, d0, g1, d1)
    assert len(log) == len(iters) and len(log) > 0
    for e, it in zip(log, iters):
        assert h_rel_err(e["H"], pack(it["H"])) < 1e-9 and g_rel_err(e["g"], it["g"]) < 1e-8
    assert np.max(np.abs(s - sref)) < 1e-9
    # the batch entry at 640x480 against the reference's own solver (wave path)
    odo.BatchSetDebugFlags(8)
    st, itb = odo.BatchAlign(g0[None], d0[None], g1[None], depth1=d1[None])
    assert odo.BatchLastPath() == 3 and int(itb.sum()) == len(iters)
    assert np.max(np.abs(st[0] - sref)) < 1e-9


def test_biobjective_needs_target_depth_also_in_a_batch(phovo):
    K = phovo.synth.K_FRAME_ALIGNMENT
    g0, d0, g1, d1 = phovo.synth.make_pair(120, 160, K=K, seed=1)
    cfg = phovo.configs.to_config("test_3_level_all_active", phovo.capi, mode=phovo.MODE_BIOBJECTIVE)
    odo = phovo.CPhotoconsistencyOdometryCuda()
    odo.SetConfig(cfg)
    odo.SetIntrinsicMatrix(K)
    odo.SetSourceFrame(g0, d0)
    odo.SetTargetFrame(g1)                       # no depth
    odo.SetInitialStateVector(np.zeros(6))
    with pytest.raises(phovo.PhovoError) as e:
        odo.Optimize()
    assert e.value.code == phovo.capi.E_INVALID
    with pytest.raises(phovo.PhovoError) as e:
        odo.BatchAlign(g0[None], d0[None], g1[None])           # the batch entry needs the target depth too
    assert e.value.code == phovo.capi.E_INVALID
    # with it, the batch runs in waves of per-pair slots (one CTA per pair): same iterations, states to the last bits;
    # the pool of per-pair contexts (debug flag 4) is bitwise the per-pair API
    odo.SetTargetFrame(g1, d1)
    odo.Optimize()
    for flags, path in ((8, 3), (4, 2), (0, 2)):
        odo.BatchSetDebugFlags(flags)
        st, it = odo.BatchAlign(np.stack([g0, g0, g0]), np.stack([d0, d0, d0]), np.stack([g1, g1, g1]), depth1=np.stack([d1, d1, d1]))
        assert odo.BatchLastPath() == path
        for p in range(3):
            if path == 2: assert np.array_equal(st[p], odo.GetOptimalStateVector())
            else: assert np.max(np.abs(st[p] - odo.GetOptimalStateVector())) < 1e-10
            assert int(it[p].sum()) == len(odo.IterationStats()) > 0
    odo.BatchSetDebugFlags(0)
