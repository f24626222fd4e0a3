#!/usr/bin/env python3
"""Turn what tools/capture_profiles.sh brought back in gpurun_out/ into the tracked summaries under profiles/:
r2_bench_n1.json, r2_launches.csv (+ _summary.txt), r2_solve_standard_tm_ncu_summary.txt,
r2_solve_standard_tm_sass_histogram.txt, solve_standard_traffic.json.   (run from the repo root, ncu on PATH)"""
import collections
import csv
import json
import re
import shutil
import subprocess

G, P = "gpurun_out/", "profiles/"
KERNEL = "_ZN7minsnap2tm24solve_standard_tm_kernelILi3ELb0ELb0ELb0EEEvNS_4fast10FastParamsEii14CUtensorMap_sti"

for page, out in (("source", "final_tm_src.csv"), ("raw", "final_tm_raw.csv")):
    with open(G + out, "w") as fh:
        subprocess.run(["ncu", "-i", G + "final_tm.ncu-rep", "--page", page, "--csv"], stdout=fh, stderr=subprocess.DEVNULL)
shutil.copy(G + "final_bench_n1.json", P + "r2_bench_n1.json")
shutil.copy(G + "final_launches.csv", P + "r2_launches.csv")
bench = json.load(open(G + "final_bench_n1.json"))

raw = list(csv.reader(open(G + "final_tm_raw.csv")))
h, u, d = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct"]
out = ["ncu --set full --import-source on --clock-control none -k regex:solve_standard_tm -s 8 -c 1 python tools/quick_bench.py --steps 6",
       "(65,536 x K=10, D=3, no cost, no status: the bench's timed step; taken after the plain run of the same command by",
       " tools/capture_profiles.sh.  Per-launch times under ncu are cold-cache and serialised -- no programmatic dependent launch",
       " overlap --; the bench's %.1f us is the warm back-to-back figure of the same build)" % (bench["ms_per_step"] * 1e3), ""]
for w in want:
    if w in h:
        i = h.index(w)
        out.append("%-72s %-16s %s" % (w, u[i], d[i]))
out += ["", "== tools/ncu_hotspots.py (warp samples per 200 SASS rows; instructions with >= 30 samples) =="]
out.append(subprocess.run(["python", "tools/ncu_hotspots.py", G + "final_tm_src.csv", "200", "30"], capture_output=True, text=True).stdout)
rows = list(csv.reader(open(G + "final_tm_src.csv")))
ix = {k: i for i, k in enumerate(rows[1])}
data = rows[2:]
nw = max(int(r[ix["Instructions Executed"]]) for r in data[:50])
cnt = collections.Counter()
for r in data:
    t = r[ix["Source"]].strip().split()
    if t:
        cnt[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += int(r[ix["Instructions Executed"]])
out.append("== executed instructions per warp (%d warps; a warp runs 1.73 batches of 16 trajectories on average) ==" % nw)
out.append(" ".join("%s:%d" % (k, v // nw) for k, v in cnt.most_common(28)))
open(P + "r2_solve_standard_tm_ncu_summary.txt", "w").write("\n".join(out) + "\n")


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


rd = to_bytes(d[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_read.sum")])
wr = to_bytes(d[h.index("dram__bytes_write.sum")], u[h.index("dram__bytes_write.sum")])
json.dump({"kernel": "solve_standard_tm_kernel<3,false,false>", "launch": "65,536 x K=10, D=3", "dram_bytes_read": rd,
           "dram_bytes_write": wr, "traffic": rd + wr, "algorithmic_bytes": 179830784,
           "source": "ncu --set full --clock-control none, profiles/r2_solve_standard_tm_ncu_summary.txt",
           "note": "below the algorithmic bytes because the tail of the output is still in L2 when the kernel ends",
           "dram_bytes_per_launch": rd + wr}, open(P + "solve_standard_traffic.json", "w"), indent=1)

sass = subprocess.run(["cuobjdump", "-sass", "-fun", KERNEL, "mav_trajectory_generation_cmake_b200/lib/obj/minsnap_standard.o"],
                      capture_output=True, text=True).stdout
ops = collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m:
        t = m.group(1).split()
        ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
names = {"STTM": "tcgen05.st", "LDTM": "tcgen05.ld", "UTMASTG": "cp.async.bulk.tensor store", "UBLKPF": "cp.async.bulk.prefetch.L2",
         "UTCATOMSWS": "tcgen05.alloc/dealloc", "LDGSTS": "cp.async", "UTMACCTL": "prefetch.tensormap", "UTMACMDFLUSH": "bulk commit",
         "ACQBULK": "griddepcontrol.wait"}
open(P + "r2_solve_standard_tm_sass_histogram.txt", "w").write(
    "SASS opcode histogram (static), solve_standard_tm_kernel<3,false,false>, sm_100a, %d instructions (end of round 2)\n" % sum(ops.values())
    + "Blackwell-specific: " + ", ".join("%s %d (%s)" % (k, ops.get(k, 0), v) for k, v in names.items()) + "\n"
    + " ".join("%s:%d" % kv for kv in ops.most_common()) + "\n")

lrows = [r for r in csv.reader(open(G + "final_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in lrows:
    name = re.sub(r"\(.*", "", r[4])
    agg[name][0] += 1
    agg[name][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 1  (first 600 launches;",
         "cold-cache, serialised times: shares, not absolutes)", "%-80s %6s %12s %6s" % ("kernel", "count", "total us", "share")]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    lines.append("%-80s %6d %12.1f %5.1f%%" % (k[:80], v[0], v[1] / 1e3, 100 * v[1] / tot))
open(P + "r2_launches_summary.txt", "w").write("\n".join(lines) + "\n")
print("profiles refreshed: %.2f us per step, frac %.4f" % (bench["ms_per_step"] * 1e3, bench["roofline"]["frac"]))
