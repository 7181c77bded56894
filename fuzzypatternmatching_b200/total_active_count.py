"""Global per-superstep statistics from the per-rank count files of a result tree.

Replaces /root/reference/examples/scripts/total_active_count.py (Python 2; README.md:35-36 of the reference):

    python -m fuzzypatternmatching_b200.total_active_count <o>/0/all_ranks_active_vertices_count/ > /tmp/vertices_count

Every file of the directory (`active_vertices_<r>`, `active_edges_<r>` or `messages_<r>`) holds one row
"itr, LP|TP, index, count" per superstep / constraint; the rows of all ranks line up.  The output has the same
lines as the reference script's: a header, then one "itr,LP|TP,index,<sum over ranks>" line per row, then "Done.".
"""
import os
import sys


def total_counts(directory):
    """[(prefix fields as written, e.g. ("0", "LP", "1")), total count] per row, summed over the files of `directory`."""
    names = sorted(n for n in os.listdir(directory) if os.path.isfile(os.path.join(directory, n)))
    if not names:
        raise ValueError("no per-rank files in %s" % directory)
    prefixes, totals = None, None
    for name in names:
        rows = [[t.strip() for t in line.strip().split(",")] for line in open(os.path.join(directory, name)) if line.strip()]
        if prefixes is None:
            prefixes = [tuple(r[:-1]) for r in rows]
            totals = [0] * len(rows)
        if len(rows) != len(totals):
            raise ValueError("%s has %d rows, expected %d" % (name, len(rows), len(totals)))
        for i, r in enumerate(rows):
            totals[i] += int(r[-1])
    return list(zip(prefixes, totals)), len(names)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        sys.stderr.write("usage: python -m fuzzypatternmatching_b200.total_active_count <count directory>\n")
        return 2
    rows, n_files = total_counts(argv[0])
    print("%d files to process ... " % n_files)
    print("Total active vertices count")
    print("Counting total number of iterations ... ")
    print("Total number of iterations: %d" % len(rows))
    for prefix, total in rows:
        print("".join(p + "," for p in prefix) + str(total))
    print("Done.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
