#!/usr/bin/env python
"""profiles/roofline_inputs.json: the per-unit figures bench.py's `roofline` object needs that only a
profiler can give -- DRAM bytes per pair (ncu dram__bytes_read+write of ONE k_batch_align launch /
its pairs), fp64 arithmetic thread-instructions per executed pixel-iteration (ncu sass op counters /
pixel-iterations of that launch), and the measured FP64 issue peak (tools/fp64_peak.cu).

usage: make_roofline_inputs.py <prof.ncu-rep> <log holding the bench JSON line of the profiled run> <fp64_peak.jsonl> <tag>"""
import csv, io, json, subprocess, sys


def main(rep, log, peak, tag):
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    d = dict(zip(raw[0], raw[2])); u = dict(zip(raw[0], raw[1]))
    line = json.loads([l for l in open(log) if l.startswith("{")][-1])
    pairs = line["config"]["pairs_per_gpu"]
    it = line["config"]["mean_iterations_per_pair"]
    rows, cols = line["config"]["rows"], line["config"]["cols"]
    px_iters = pairs * sum(v * round(rows * 0.5 ** int(l)) * round(cols * 0.5 ** int(l)) for l, v in it.items())
    cycles = float(d["sm__cycles_elapsed.avg"])
    fp64 = sum(float(d["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k]) for k in ("dfma", "dmul", "dadd")) * cycles
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum(float(d["dram__bytes_%s.sum" % k]) * scale[u["dram__bytes_%s.sum" % k]] for k in ("read", "write"))
    best = max(json.loads(l)["dfma_per_s"] for l in open(peak) if l.startswith("{"))
    out = {"tag": tag, "kernel": "k_batch_align", "profiled_pairs": pairs, "profiled_duration_ms": float(d["gpu__time_duration.sum"]),
           "dram_bytes_per_pair": dram / pairs, "fp64_thread_inst_per_px_iter": fp64 / px_iters,
           "fp64_pipe_active_pct_of_active": float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]),
           "fp64_peak_thread_inst_per_s": best, "fp64_peak_source": "tools/fp64_peak.cu on this pool's B200 (profiles/r01_fp64_peak.jsonl), DFMA issue rate",
           "registers_per_thread": int(d["launch__registers_per_thread"])}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
