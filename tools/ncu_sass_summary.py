"""Aggregate an ncu SASS source page by opcode: share of warp instructions and of stall samples."""
import collections
import csv
import subprocess
import sys


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    for b in out.split('"Kernel Name",')[1:]:
        lines = b.splitlines()
        rows = list(csv.reader(lines[1:]))
        h = rows[0]
        si, ie, te, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Source")
        body = [r for r in rows[1:] if len(r) > te and r[ie].isdigit()]
        tot_s, tot_i, tot_t = (sum(int(r[k]) for r in body) for k in (si, ie, te))
        print(lines[0][:90], "| samples", tot_s, "warp-inst", tot_i, "thread-inst", tot_t, "simt %.1f" % (tot_t / max(tot_i, 1)))
        ops, ops_s = collections.Counter(), collections.Counter()
        for r in body:
            toks = r[src].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            ops[op] += int(r[ie])
            ops_s[op] += int(r[si])
        print("  warp-inst %:", [(k, round(100 * v / tot_i, 1)) for k, v in ops.most_common(16)])
        print("  samples   %:", [(k, round(100 * v / tot_s, 1)) for k, v in ops_s.most_common(12)])


if __name__ == "__main__":
    main(sys.argv[1])
