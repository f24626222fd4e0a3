#!/usr/bin/env python3
"""Per-phase summary of an `ncu --set full --import-source on` capture of the standard solve kernels.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; ncu -i prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_phases.py src.csv [raw.csv]

Phases are cut at SASS landmarks (cp.async wait, first STTM, first pair shuffle, first LDTM, the
reciprocal that opens the recovery loop, UBLKCP) and every phase reports its share of the warp
samples, executed instructions per warp, FP64 instructions, shared-memory wavefronts (ideal / excess)
and its top stall reasons."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    return {h: i for i, h in enumerate(hdr)}, rows[2:], rows[0][1]


def main():
    ix, data, name = load(sys.argv[1])
    src = [r[ix["Source"]].strip() for r in data]
    n_warps = max(int(r[ix["Instructions Executed"]]) for r in data[:50]) or 1

    def first(pred, start=0):
        for i in range(start, len(src)):
            if pred(src[i]):
                return i
        return len(src)

    i_in = first(lambda s: "DEPBAR.LE SB0" in s or "LDGDEPBAR" in s) + 1
    i_sttm = first(lambda s: "STTM" in s)
    i_mid = first(lambda s: "SHFL.BFLY" in s, i_in)
    i_ldtm = first(lambda s: "LDTM" in s)
    i_rec = first(lambda s: "MUFU.RCP64H" in s, i_ldtm) if i_ldtm < len(src) else len(src)
    i_ub = first(lambda s: "UBLKCP" in s)
    cuts = [("inputs", 0, i_in), ("forward", i_in, max(i_mid - 60, i_in)), ("middle", max(i_mid - 60, i_in), max(i_ldtm - 40, i_mid)),
            ("backsub", max(i_ldtm - 40, i_mid), max(i_rec - 40, i_ldtm)), ("recovery", max(i_rec - 40, i_ldtm), min(i_ub + 60, len(src))),
            ("tail", min(i_ub + 60, len(src)), len(src))]
    stalls = [h for h in ix if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    print(name)
    print("rows %d, samples %d, warps %d; landmarks: inputs<%d sttm %d shfl %d ldtm %d rcp %d ublkcp %d" %
          (len(data), tot, n_warps, i_in, i_sttm, i_mid, i_ldtm, i_rec, i_ub))
    for label, a, b in cuts:
        rs = data[a:b]
        if not rs:
            continue
        s = sum(int(r[ix["# Samples"]]) for r in rs)
        inst = sum(int(r[ix["Instructions Executed"]]) for r in rs) / n_warps
        fp64 = sum(int(r[ix["Instructions Executed"]]) for r in rs
                   if re.search(r"\b(DFMA|DMUL|DADD|DSETP)\b", r[ix["Source"]])) / n_warps
        wf = sum(int(r[ix["L1 Wavefronts Shared"]]) for r in rs) / n_warps
        ex = sum(int(r[ix["L1 Wavefronts Shared Excessive"]]) for r in rs) / n_warps
        st = {h: sum(int(r[ix[h]]) for r in rs) for h in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        print("%-9s samples %5.1f%%  inst/warp %6.0f  fp64 %5.0f  smem wavefronts/warp %5.0f (excess %4.0f)  stalls %s" %
              (label, 100.0 * s / tot, inst, fp64, wf, ex, " ".join("%s:%d" % (k[6:], v) for k, v in top)))
    ops = collections.Counter()
    for r in data:
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
        if m:
            ops[m.group(2)] += int(r[ix["Instructions Executed"]]) / n_warps
    print("executed per warp: " + " ".join("%s:%.0f" % kv for kv in ops.most_common(24)))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        want = ("gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
                "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
                "smsp__thread_inst_executed_per_inst_executed.ratio")
        for h, u, v in zip(rows[0], rows[1], rows[2]):
            if h in want:
                print("%-70s %-14s %s" % (h, u, v))


if __name__ == "__main__":
    main()
