#!/usr/bin/env python3
"""Where a kernel's warp samples go, from an `ncu --set full --import-source on` capture.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_hotspots.py src.csv [window] [min_samples]

Prints the totals per stall reason, the samples and executed instructions per window of SASS rows, and every
instruction with at least `min_samples` samples together with its top stall reasons."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    win = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    thr = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    S = lambda i: int(data[i][ix["# Samples"]])
    E = lambda i: int(data[i][ix["Instructions Executed"]])
    n_warps = max(E(i) for i in range(min(50, len(data)))) or 1
    total = sum(S(i) for i in range(len(data)))
    print(rows[0][1][:100])
    print("rows %d  samples %d  warps %d  instructions per warp %.0f" %
          (len(data), total, n_warps, sum(E(i) for i in range(len(data))) / n_warps))
    tot = collections.Counter()
    for r in data:
        for c in stall:
            tot[c[6:]] += int(r[ix[c]] or 0)
    print("stalls: " + "  ".join("%s %.1f%%" % (c, 100.0 * n / total) for c, n in tot.most_common(10)))

    def top(r):
        d = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall), reverse=True)
        return " ".join("%s:%d" % (c, n) for n, c in d[:3] if n > 0)

    print("-- windows of %d rows: first row, samples (%%), instructions per warp, first instruction" % win)
    for a in range(0, len(data), win):
        b = min(len(data), a + win)
        n = sum(S(i) for i in range(a, b))
        if n:
            print("%5d %5d %5.1f%% %7.0f  %s" % (a, n, 100.0 * n / total, sum(E(i) for i in range(a, b)) / n_warps,
                                               data[a][ix["Source"]].strip()[:50]))
    print("-- instructions with >= %d samples" % thr)
    for i, r in enumerate(data):
        if S(i) >= thr:
            print("%5d %4d x%-6d %-60s | %s" % (i, S(i), E(i), r[ix["Source"]].strip()[:60], top(r)))


if __name__ == "__main__":
    main()
