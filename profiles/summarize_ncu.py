"""Key counters of an `ncu --set full` report (raw page, CSV on stdin) as a markdown table for profiles/."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak")]
r = list(csv.reader(sys.stdin))
h, units = r[0], r[1]
print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
print("|---|" + "---:|" * len(WANT))
for row in r[2:]:
    name = row[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    cells = []
    for m, _ in WANT:
        if m in h:
            i = h.index(m)
            v = row[i]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            cells.append("%s %s" % (v, units[i]))
        else:
            cells.append("-")
    print("| `%s` | " % name + " | ".join(cells) + " |")
