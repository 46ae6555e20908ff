import sys, importlib, os
ROOT='/root/repo'; sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/oracle'); sys.path.insert(0,ROOT+'/tests')
import numpy as np
import oracle_py as op
from helpers import *
phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
def conv(cfg): return op.Config.from_buffer_copy(bytes(cfg))
# golden
gd = load_golden("pair_96x128_ref")
cfg = phovo.default_config(); cfg.mode=0; cfg.num_levels=3
for l in range(10):
    cfg.max_num_iterations[l] = int(gd["iters"][l]) if l<3 else 0
    cfg.min_gradient_norm[l] = float(gd["min_grad"])
odo = phovo.CPhotoconsistencyOdometryCuda(); odo.SetConfig(cfg); odo.SetIntrinsicMatrix(gd["K"])
odo.SetSourceFrame(gd["gray0"], gd["depth0"]); odo.SetTargetFrame(gd["gray1"]); odo.SetInitialStateVector(np.zeros(6)); odo.Optimize()
for i in range(3):
    lvl, st = int(gd["eval%d_level"%i]), gd["eval%d_state"%i]
    if cfg.max_num_iterations[lvl]==0: continue
    e = odo.EvalNormalEquations(lvl, st)
    res,_ = odo.EvalResiduals(lvl, st, odo.LevelImage(0,lvl).shape)
    gr = gd["eval%d_res"%i]
    print("eval",i,"lvl",lvl,"valid",e["num_valid"],int(gd["eval%d_count"%i]),"maxdiff",np.abs(res-gr).max())
    bad = np.flatnonzero((np.abs(res)>1e-6)!=(np.abs(gr)>1e-6))
    print(" bad idx", bad[:10], res[bad[:10]], gr[bad[:10]], gd["eval%d_winner"%i][bad[:10]])
# ceres LM
K = phovo.synth.K_FRAME_ALIGNMENT
g0,d0,g1,_ = phovo.synth.make_pair(240,320,seed=12)
cfg = phovo.configs.to_config("config_4_level_optimization_ceres", phovo.capi)
odo = phovo.CPhotoconsistencyOdometryCuda(); odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
odo.SetSourceFrame(g0,d0); odo.SetTargetFrame(g1); odo.SetInitialStateVector(np.zeros(6)); odo.Optimize()
log = odo.IterationStats()
for f32 in (False, True):
    o = op.Oracle(conv(cfg), K, storage_f32=f32); o.set_source(g0,d0); o.set_target(g1); o.set_initial_state(np.zeros(6)); o.optimize()
    print("oracle f32 storage", f32, "final diff", np.abs(o.state()-odo.GetOptimalStateVector()).max())
    for a,b in zip(log,o.iter_stats()):
        print(a["level"],a["iteration"],a["accepted"],b["accepted"],a["num_valid"],b["num_valid"],"cost %.9f %.9f"%(a["cost"],b["cost"]),"rad %g %g"%(a["radius"],b["radius"]), "dstate %.2e"%np.abs(a["state_in"]-b["state_in"]).max())
