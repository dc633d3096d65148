import csv, sys, glob, statistics
for f in sorted(glob.glob(sys.argv[1] + "*")):
    rows = list(csv.DictReader(open(f)))
    by = {}
    for r in rows:
        by[(int(r["iteration"]), r["kernel"])] = tuple(int(r[k]) for k in ("entry_ns", "start_ns", "end_ns"))
    its = sorted({i for i, _ in by})
    names = ["spmv_dot", "update_xr", "update_p"]
    dur = {k: [] for k in names}; gap = {k: [] for k in names}; early = {k: [] for k in names}
    period = []
    for i in its:
        for j, k in enumerate(names):
            if (i, k) not in by: continue
            e, s, t = by[(i, k)]
            dur[k].append((t - s) / 1e3)
            early[k].append((s - e) / 1e3)
            prev = by.get((i, names[j - 1])) if j else by.get((i - 1, names[2]))
            if prev: gap[k].append((s - prev[2]) / 1e3)
        if (i, "spmv_dot") in by and (i + 1, "spmv_dot") in by:
            period.append((by[(i + 1, "spmv_dot")][1] - by[(i, "spmv_dot")][1]) / 1e3)
    print(f)
    for k in names:
        if dur[k]:
            print("  %-10s run %.1f us (min %.1f max %.1f)  gap-before %.1f us  resident-before-wait %.1f us" % (
                k, statistics.mean(dur[k]), min(dur[k]), max(dur[k]), statistics.mean(gap[k]) if gap[k] else -1, statistics.mean(early[k])))
    if period: print("  iteration period %.1f us" % statistics.mean(period))
