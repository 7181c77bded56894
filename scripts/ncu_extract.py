"""Extracts the judged metrics from an .ncu-rep (read here with `ncu -i ... --page raw --csv`)."""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__pcsamp_sample_count", "smsp__pcsamp_warps_issue_stalled_long_scoreboard"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = (r[i], units[i])
        res.append(d)
    return res


if __name__ == "__main__":
    summary = {}
    for key, path in (a.split("=") for a in sys.argv[1:]):
        rs = rows_of(path)
        tot_b, tot_us = 0.0, 0.0
        print("## %s (%s)" % (key, path))
        print("| kernel | us | DRAM read MB | DRAM write MB | DRAM % of peak | warps active % | issue active % | L2 hit % | regs | long-scoreboard share |")
        print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
        for d in rs:
            rd = to_bytes(*d["dram__bytes_read.sum"])
            wr = to_bytes(*d["dram__bytes_write.sum"])
            us = to_us(*d["gpu__time_duration.sum"])
            tot_b += rd + wr
            tot_us += us
            ls = float(d["smsp__pcsamp_warps_issue_stalled_long_scoreboard"][0].replace(",", "")) / max(1.0, float(d["smsp__pcsamp_sample_count"][0].replace(",", "")))
            print("| %s | %.1f | %.1f | %.1f | %s | %s | %s | %s | %s | %.2f |" % (
                d["kernel"][:34], us, rd / 1e6, wr / 1e6, d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0][:5],
                d["sm__warps_active.avg.pct_of_peak_sustained_active"][0][:5],
                d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0][:5], d["lts__t_sector_hit_rate.pct"][0][:5],
                d["launch__registers_per_thread"][0], ls))
        summary[key] = {"dram_bytes_per_launch": tot_b / len(rs), "launches": len(rs), "avg_us": tot_us / len(rs),
                        "dram_gbs": tot_b / tot_us / 1e3}
        print()
    print("```json\n" + json.dumps(summary, indent=1) + "\n```")
    json.dump(summary, open("profiles/traffic.json", "w"), indent=1)
