"""Key counters of an `ncu --set full` report (raw page, CSV on stdin) as markdown for profiles/.

    ncu -i gpurun_out/<tag>.ncu-rep --page raw --csv | python profiles/summarize_ncu.py \
        [--traffic-json profiles/roofline_traffic.json --workload "C2 ..." --source profiles/<tag>_ncu.md]

One table row per captured launch, then a per-launch detail block (L1 wavefronts, L2 sectors by eviction class - the
matrix stream is loaded evict-first, x / y / row pointers evict-normal, so the evict-normal hit rate IS the L2 hit rate
on x for the gather kernels - and the top stall reasons).  With --traffic-json the DRAM bytes per launch of every
captured kernel instantiation are written (merged) into that file, keyed by the instantiation's name, which is where
bench.py reads `roofline.traffic` from: no hand-copied numbers.
"""
import argparse
import csv
import json
import os
import re
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak")]
DETAIL = [
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU data-pipe wavefronts, % of peak"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global load sectors"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "L1 global load tag wavefronts"),
    ("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "L1 global load sector hit %"),
    ("lts__t_sectors_srcunit_tex.sum", "L2 sectors from the SMs"),
    ("lts__t_sectors_srcunit_tex_evict_first_lookup_hit.sum", "  evict-first (matrix stream) hit"),
    ("lts__t_sectors_srcunit_tex_evict_first_lookup_miss.sum", "  evict-first (matrix stream) miss"),
    ("lts__t_sectors_srcunit_tex_evict_normal_lookup_hit.sum", "  evict-normal (x, y, row pointers) hit"),
    ("lts__t_sectors_srcunit_tex_evict_normal_lookup_miss.sum", "  evict-normal (x, y, row pointers) miss"),
    ("lts__t_sectors_srcunit_tex_evict_last_lookup_hit.sum", "  evict-last (kept vectors) hit"),
    ("lts__t_sectors_srcunit_tex_evict_last_lookup_miss.sum", "  evict-last (kept vectors) miss"),
    ("lts__t_sectors_srcunit_ltcfabric.sum", "L2 sectors crossing the die-to-die fabric"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall: membar / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def kernel_name(raw):
    """`void caskb200::<unnamed>::spmv_ell_persistent_kernel<2, 0, 0>(...)` -> spmv_ell_persistent_kernel<2,false,0>"""
    name = raw.split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("<unnamed>::", "")
    name = re.sub(r"^(\w+::)+", "", name).replace(" ", "")
    m = re.match(r"(spmv_ell_persistent_kernel)<(\d+),(\d+|true|false),(\d+)>", name)
    if m:
        b = {"0": "false", "1": "true"}.get(m.group(3), m.group(3))
        name = "%s<%s,%s,%s>" % (m.group(1), m.group(2), b, m.group(4))
    return name


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traffic-json")
    ap.add_argument("--workload", default="")
    ap.add_argument("--source", default="")
    args = ap.parse_args()
    r = list(csv.reader(sys.stdin))
    if len(r) < 3:
        print("(no launches captured)")
        return
    h, units = r[0], r[1]
    col = {k: i for i, k in enumerate(h)}
    print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|" + "---:|" * len(WANT))
    traffic = {}
    for row in r[2:]:
        name = kernel_name(row[col["Kernel Name"]])
        cells = []
        for m, _ in WANT:
            if m in col:
                v = num(row[col[m]])
                cells.append("%s %s" % ("%.4g" % v if v is not None else row[col[m]], units[col[m]]))
            else:
                cells.append("-")
        print("| `%s` | " % name + " | ".join(cells) + " |")
        try:
            rd = num(row[col["dram__bytes_read.sum"]]) * SCALE[units[col["dram__bytes_read.sum"]]]
            wr = num(row[col["dram__bytes_write.sum"]]) * SCALE[units[col["dram__bytes_write.sum"]]]
            du = num(row[col["gpu__time_duration.sum"]])
            du_u = units[col["gpu__time_duration.sum"]]
            du_us = du * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(du_u, 1)
            t = traffic.setdefault(name, {"launches": 0, "read": 0.0, "write": 0.0, "us": 0.0, "nh": 0.0, "nm": 0.0})
            t["launches"] += 1; t["read"] += rd; t["write"] += wr; t["us"] += du_us
            kh, km = "lts__t_sectors_srcunit_tex_evict_normal_lookup_hit.sum", "lts__t_sectors_srcunit_tex_evict_normal_lookup_miss.sum"
            if kh in col and km in col:
                t["nh"] += num(row[col[kh]]) or 0.0
                t["nm"] += num(row[col[km]]) or 0.0
        except (KeyError, TypeError):
            pass
    for row in r[2:]:
        name = kernel_name(row[col["Kernel Name"]])
        print("\n`%s` (launch id %s):\n" % (name, row[col["ID"]] if "ID" in col else "?"))
        vals = {}
        for m, label in DETAIL:
            if m in col and row[col[m]] != "":
                v = num(row[col[m]])
                vals[m] = v
                print("* %s: %s %s" % (label.strip(), "%.6g" % v if v is not None else row[col[m]], units[col[m]]))
        nh = vals.get("lts__t_sectors_srcunit_tex_evict_normal_lookup_hit.sum")
        nm = vals.get("lts__t_sectors_srcunit_tex_evict_normal_lookup_miss.sum")
        if nh is not None and nm is not None and nh + nm > 0:
            print("* => L2 hit rate of the evict-normal class (x gathers, y, row pointers): %.1f %%; its misses = %.3g GB of DRAM reads"
                  % (100.0 * nh / (nh + nm), nm * 32 / 1e9))
        fh = vals.get("lts__t_sectors_srcunit_tex_evict_first_lookup_hit.sum")
        fm = vals.get("lts__t_sectors_srcunit_tex_evict_first_lookup_miss.sum")
        if fh is not None and fm is not None and fh + fm > 0:
            print("* => matrix stream (evict-first): %.3g GB requested at L2, %.3g GB missed to DRAM" % ((fh + fm) * 32 / 1e9, fm * 32 / 1e9))
    if args.traffic_json:
        cur = {}
        if os.path.exists(args.traffic_json):
            with open(args.traffic_json) as f:
                cur = json.load(f)
        kern = cur.setdefault("kernels", {})
        for name, t in traffic.items():
            n = t["launches"]
            kern[name + (" @ " + args.workload if args.workload else "")] = {
                "kernel": name, "workload": args.workload,
                "dram_bytes_per_launch": int(round((t["read"] + t["write"]) / n)),
                "dram_read_bytes": int(round(t["read"] / n)), "dram_write_bytes": int(round(t["write"] / n)),
                "duration_us_under_ncu": round(t["us"] / n, 2), "launches_averaged": n, "source": args.source,
                # evict-normal class = x gathers / x windows, y, row pointers (the matrix stream is evict-first)
                "l2_hit_rate_on_x_pct": round(100.0 * t["nh"] / (t["nh"] + t["nm"]), 2) if t["nh"] + t["nm"] > 0 else None}
        cur.pop("spmv_ell_persistent_kernel_bytes_per_launch", None)
        cur.pop("source", None)
        with open(args.traffic_json, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
