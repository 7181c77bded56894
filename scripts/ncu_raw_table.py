"""Markdown table + per-class DRAM traffic from an `ncu -i X.ncu-rep --page raw --csv` export of the hot kernels.
usage: ncu_raw_table.py raw.csv [traffic.json [workload key]]   (launches under 50 us are left out of the table)"""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__t_sectors.sum", "L2 sectors"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "regs")]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def klass(name):
    if "k_lcc_scan<1" in name or "k_lcc_first" in name:
        return "scan_first"
    if "k_lcc_scan<0, 0, 1" in name or "k_lcc_xlate8" in name:
        return "scan_xlate"
    if "k_lcc_scan" in name:
        return "scan_later"
    for k in ("k_nem1_close_cycle", "k_nem1_expand", "k_init_flags", "k_init_assign", "k_lcc_commit"):
        if k in name:
            return k
    return "other"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {c: hdr.index(c) for c, _ in COLS if c in hdr}
    print("| kernel | class | " + " | ".join(t for c, t in COLS if c in ix) + " |")
    print("|---|---|" + "---:|" * len(ix))
    agg = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        vals = {}
        for c, _ in COLS:
            if c in ix:
                vals[c] = float(r[ix[c]].replace(",", "")) * SCALE.get(units[ix[c]], 1)
        k = klass(name)
        if vals.get("gpu__time_duration.sum", 0) < 0.05:
            continue
        cells = []
        for c, _ in COLS:
            if c not in ix:
                continue
            v = vals[c]
            if c == "gpu__time_duration.sum":
                cells.append("%.1f us" % (v * 1e3))
            elif c.startswith("dram__bytes"):
                cells.append("%.1f MB" % (v / 1e6))
            elif c in ("lts__t_sectors.sum", "smsp__inst_executed.sum"):
                cells.append("%.1f M" % (v / 1e6))
            else:
                cells.append("%.1f" % v)
        print("| `%s` | %s | %s |" % (name.split("(")[0].replace("void ", ""), k, " | ".join(cells)))
        a = agg.setdefault(k, {"bytes": 0.0, "ms": 0.0, "n": 0})
        a["bytes"] += vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
        a["ms"] += vals["gpu__time_duration.sum"]
        a["n"] += 1
    out = {k: {"dram_bytes_per_launch": a["bytes"] / a["n"], "launches": a["n"], "avg_us": a["ms"] / a["n"] * 1e3,
               "dram_gbs": a["bytes"] / (a["ms"] * 1e-3) / 1e9} for k, a in agg.items()}
    print("\n```json\n" + json.dumps(out, indent=1) + "\n```")
    if len(sys.argv) > 2:
        key = sys.argv[3] if len(sys.argv) > 3 else "rmat_s26_cyclic_n1"
        try:
            doc = json.load(open(sys.argv[2]))
        except Exception:
            doc = {}
        doc[key] = out
        json.dump(doc, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
