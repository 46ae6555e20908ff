#!/usr/bin/env python
"""Parity at the BASELINE size over many pairs: N random 640x480 pairs (the bench's generator) through the
batch kernels and through the CPU oracle on all host threads; counts iteration-count mismatches and poses over
the north-star bar.  One JSON line."""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2048)
    ap.add_argument("--config", default="config_4_level_optimization_analytic")
    ap.add_argument("--depth", default="u16", choices=["u16", "f32"])
    ap.add_argument("--reference-sample", type=int, default=128, help="pairs also run through the reference's own header (oracle/_ref)")
    args = ap.parse_args()
    phovo = importlib.import_module("photoconsistency-visual-odometry_b200")
    phovo.build()
    import oracle_py, torch
    oracle_py.build()
    K = phovo.synth.K_FRAME_ALIGNMENT if "4_level" in args.config else phovo.synth.K_VISUAL_ODOMETRY
    cfg = phovo.configs.to_config(args.config, phovo.capi)
    g0, d0, g1, _ = phovo.synth.render_batch_torch(args.pairs, 480, 640, K, torch.device("cuda", 0), seed0=5000)
    odo = phovo.CPhotoconsistencyOdometryCuda()
    odo.SetConfig(cfg); odo.SetIntrinsicMatrix(K)
    if args.depth == "u16":
        raw = torch.clamp(torch.round(d0.to(torch.float64) * 5000.), 0, 32767).to(torch.int16)
        st, it = odo.BatchAlign(g0, raw, g1, depth_scale=1. / 5000.)
        d_host = raw.cpu().numpy().astype(np.float64) * (1. / 5000.)
    else:
        st, it = odo.BatchAlign(g0, d0, g1)
        d_host = d0.cpu().numpy().astype(np.float64)
    t0 = time.time()
    ost, oit, _, _ = oracle_py.align_batch(oracle_py.Config.from_buffer_copy(bytes(cfg)), K, g0.cpu().numpy(), d_host, g1.cpu().numpy(),
                                           num_threads=os.cpu_count(), lean=True)
    dt = np.max(np.abs(st[:, :3] - ost[:, :3]), axis=1); dr = np.max(np.abs(st[:, 3:] - ost[:, 3:]), axis=1)
    # a sample also against the REFERENCE'S OWN header (oracle/_ref, where the prebuilt library travelled)
    ref_stats = {}
    try:
        import ref_py, tempfile
        if ref_py.available() and cfg.mode == 0:
            ref = ref_py.Reference(phovo.configs.write_yaml(args.config, tempfile.mkdtemp()), K)
            gh0, gh1 = g0.cpu().numpy(), g1.cpu().numpy()
            worst, mism, n = 0., 0, 0
            for p in range(0, args.pairs, max(1, args.pairs // args.reference_sample)):
                rs, _, riters = ref.align(gh0[p], d_host[p], gh1[p])
                n += 1
                if len(riters) != int(it[p].sum()): mism += 1
                else: worst = max(worst, float(np.max(np.abs(st[p] - rs))))
            ref_stats = {"reference_header_pairs": n, "reference_header_iteration_mismatches": mism, "reference_header_worst_state_diff": worst}
    except Exception as e:  # pragma: no cover
        ref_stats = {"reference_header_error": repr(e)}
    print(json.dumps({**ref_stats, "pairs": args.pairs, "config": args.config, "depth": args.depth, "iteration_count_mismatches": int((it != oit).any(axis=1).sum()),
                      "poses_over_bar": int(((dt >= 1e-4) | (dr >= 1e-5)).sum()), "worst_translation_diff": float(dt.max()), "worst_rotation_diff": float(dr.max()),
                      "mean_iterations": {str(l): float(it[:, l].mean()) for l in range(cfg.num_levels) if cfg.max_num_iterations[l] > 0},
                      "pairs_at_iteration_cap": int((it[:, cfg.num_levels - 1] == cfg.max_num_iterations[cfg.num_levels - 1]).sum()),
                      "cpu_oracle_seconds": time.time() - t0}))


if __name__ == "__main__":
    main()
