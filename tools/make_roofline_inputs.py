#!/usr/bin/env python
"""profiles/roofline_inputs.json: the per-unit figures bench.py's `roofline` object needs that only a
profiler can give -- DRAM bytes per pair (ncu dram__bytes_read+write of the k_batch_level launches of ONE step /
its pairs), fp64 arithmetic thread-instructions per executed pixel-iteration (ncu sass op counters /
pixel-iterations of that launch), and the measured FP64 issue peak (tools/fp64_peak.cu).

usage: make_roofline_inputs.py <prof.ncu-rep> <log holding the bench JSON line of the profiled run> <fp64_peak.jsonl> <tag>"""
import csv, io, json, subprocess, sys


def l2_bytes(L, u):
    """L2 traffic of a launch: lts__t_bytes when the capture has it, else 32-byte sectors (ncu --set full has lts__t_sectors.sum)."""
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if "lts__t_bytes.sum" in L:
        return float(L["lts__t_bytes.sum"]) * scale[u["lts__t_bytes.sum"]]
    mult = {"sector": 1.0, "Ksector": 1e3, "Msector": 1e6, "Gsector": 1e9}.get(u["lts__t_sectors.sum"], 1.0)
    return float(L["lts__t_sectors.sum"]) * mult * 32.0


def main(rep, log, peak, tag):
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    u = dict(zip(raw[0], raw[1]))
    every = [dict(zip(raw[0], r)) for r in raw[2:]]
    launches = [L for L in every if "k_batch_level" in L.get("Kernel Name", "")]       # one k_batch_level launch per active pyramid level
    pyramid = [L for L in every if "k_batch_pyramid" in L.get("Kernel Name", "")]     # present when captured with -k regex:k_batch
    d = launches[0]
    line = json.loads([l for l in open(log) if l.startswith("{")][-1])
    pairs = line["config"]["pairs_per_gpu"]
    it = line.get("mean_iterations_per_pair") or line["config"]["mean_iterations_per_pair"]
    rows, cols = line["config"]["rows"], line["config"]["cols"]
    px_iters = pairs * sum(v * round(rows * 0.5 ** int(l)) * round(cols * 0.5 ** int(l)) for l, v in it.items())
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    fp64 = dram = dur = l2 = 0.0
    warp_inst = sched_cycles = 0.0
    for L in launches:
        warp_inst += float(L["smsp__inst_executed.sum"])
        sched_cycles += float(L["sm__cycles_elapsed.avg"]) * 4 * float(L.get("launch__sm_count", 148))   # 4 schedulers per SM
        l2 += l2_bytes(L, u)
        cycles = float(L["sm__cycles_elapsed.avg"])
        fp64 += sum(float(L["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k]) for k in ("dfma", "dmul", "dadd")) * cycles
        dram += sum(float(L["dram__bytes_%s.sum" % k]) * scale[u["dram__bytes_%s.sum" % k]] for k in ("read", "write"))
        dur += float(L["gpu__time_duration.sum"]) * tscale[u["gpu__time_duration.sum"]]
    best = max(json.loads(l)["dfma_per_s"] for l in open(peak) if l.startswith("{"))
    out = {"tag": tag, "kernel": "k_batch_level (one launch per active level, summed)", "profiled_launches": len(launches),
           "profiled_pairs": pairs, "profiled_duration_ms": dur,
           "dram_bytes_per_pair": dram / pairs, "fp64_thread_inst_per_px_iter": fp64 / px_iters,
           "fp64_pipe_active_pct_of_active_per_launch": [float(L["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]) for L in launches],
           "fp64_peak_thread_inst_per_s": best, "fp64_peak_source": "tools/fp64_peak.cu on this pool's B200 (profiles/r01_fp64_peak.jsonl), DFMA issue rate",
           # issue-slot accounting (tools/issue_probe.cu: nothing issues in the shadow of a DFMA, so an FP64-pipe instruction
           # costs the scheduler two issue cycles and every other instruction one): per 32 pixel-iterations of one warp
           "warp_inst_per_32_px_iter": warp_inst / (px_iters / 32.0),
           "fp64_warp_inst_per_32_px_iter": fp64 / px_iters,
           "issue_cycles_floor_per_32_px_iter": (warp_inst + fp64 / 32.0) / (px_iters / 32.0),
           "scheduler_cycles_per_32_px_iter": sched_cycles / (px_iters / 32.0),
           "issue_floor_frac": (warp_inst + fp64 / 32.0) / sched_cycles,
           "l2_bytes_per_pair": l2 / pairs, "l2_gbs": l2 / (dur * 1e-3) / 1e9, "dram_gbs": dram / (dur * 1e-3) / 1e9,
           "registers_per_thread": int(d["launch__registers_per_thread"])}
    if pyramid:
        Lp = pyramid[0]
        pd = sum(float(Lp["dram__bytes_%s.sum" % k]) * scale[u["dram__bytes_%s.sum" % k]] for k in ("read", "write"))
        pt = float(Lp["gpu__time_duration.sum"]) * tscale[u["gpu__time_duration.sum"]]
        pl2 = l2_bytes(Lp, u)
        out.update({"pyramid_dram_bytes_per_pair": pd / pairs, "pyramid_profiled_duration_ms": pt,
                    "pyramid_dram_gbs": pd / (pt * 1e-3) / 1e9, "pyramid_l2_gbs": pl2 / (pt * 1e-3) / 1e9})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
