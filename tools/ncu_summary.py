#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py cite:
duration, DRAM traffic, pipe utilisation, stall reasons, and the share of warp samples per phase
(phases are delimited by BAR.SYNC in the SASS).   usage: ncu_summary.py prof.ncu-rep > profiles/x.txt"""
import collections, csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep):
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for li, vals in enumerate(raw[2:]):
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print("== launch %d: %s" % (li, d.get("Kernel Name", "?")[:110]))
        for k in KEYS:
            if k in d:
                print("  %-72s %s %s" % (k, d[k], u[k]))
        try:   # achieved DRAM and L2 bandwidth of the launch (north star: both against the ~8 TB/s HBM peak / the L2 roofline)
            sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ts = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
            t = float(d["gpu__time_duration.sum"]) * ts[u["gpu__time_duration.sum"]]
            dram = sum(float(d["dram__bytes_%s.sum" % k]) * sc[u["dram__bytes_%s.sum" % k]] for k in ("read", "write"))
            sect = float(d["lts__t_sectors.sum"]) * {"sector": 1.0, "Ksector": 1e3, "Msector": 1e6, "Gsector": 1e9}.get(u["lts__t_sectors.sum"], 1.0)
            print("  %-72s %.1f GB/s DRAM, %.1f GB/s L2 (32 B sectors)" % ("achieved bandwidth", dram / t / 1e9, sect * 32 / t / 1e9))
        except Exception:
            pass
        st = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
        print("  stall reasons (warps per issue-active cycle): " + ", ".join(
            "%s %.2f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    # one table per kernel: header row starts with "Address"
    i = 0
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            name = src[i][1]; hdr = src[i + 1]; i += 2
            iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
            segs = [[0, 0, collections.Counter()]]
            while i < len(src) and src[i] and src[i][0] != "Kernel Name":
                r = src[i]; s = r[isrc].strip()
                m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
                op = m.group(2).split(".")[0] if m else s
                segs[-1][0] += int(r[iS]); segs[-1][1] += int(r[iI]); segs[-1][2][op] += int(r[iI])
                if "BAR.SYNC" in s:
                    segs.append([0, 0, collections.Counter()])
                i += 1
            ts = sum(x[0] for x in segs) or 1; ti = sum(x[1] for x in segs) or 1
            print("== phases of %s (SASS split at BAR.SYNC): %% of warp samples / %% of warp instructions / top opcodes" % name[:80])
            for k, (s_, i_, ops) in enumerate(segs):
                if s_ * 200 < ts and i_ * 200 < ti:
                    continue
                print("  seg %2d  %5.1f%%  %5.1f%%  %s" % (k, 100. * s_ / ts, 100. * i_ / ti, " ".join("%s:%.0f%%" % (o, 100. * c / max(i_, 1)) for o, c in ops.most_common(8))))
        else:
            i += 1


if __name__ == "__main__":
    main(sys.argv[1])
