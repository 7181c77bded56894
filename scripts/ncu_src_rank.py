"""Ranks CUDA source lines of an `ncu --page source --csv --print-source cuda,sass` export by executed
instructions and stall samples, per kernel."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Function Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            iS, iE, iL2 = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L2 Theoretical Sectors Global")
            agg, te, ts, tl = {}, 0, 0, 0
            j = i + 2
            while j < len(rows) and not (rows[j] and rows[j][0] in ("Function Name", "File Path")):
                r = rows[j]
                j += 1
                if len(r) <= iL2 or r[2] != "-":
                    continue
                try:
                    e, sm, l2 = int(r[iE]), int(r[iS]), int(r[iL2] or 0)
                except ValueError:
                    continue
                k = (int(r[0]), r[1].strip()[:100])
                a = agg.get(k, (0, 0, 0))
                agg[k] = (a[0] + e, a[1] + sm, a[2] + l2)
                te, ts, tl = te + e, ts + sm, tl + l2
            print("== %s\n   warp instructions %d, samples %d, L2 sectors %d" % (name[:110], te, ts, tl))
            for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
                print("%5.1f%% instr %5.1f%% samp %5.1f%% l2  L%d: %s" % (100 * a[0] / max(te, 1), 100 * a[1] / max(ts, 1), 100 * a[2] / max(tl, 1), k[0], k[1]))
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
