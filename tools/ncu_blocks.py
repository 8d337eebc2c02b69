"""Group an ncu SASS page into runs of equal execution count (~basic blocks) and print the heaviest."""
import collections
import csv
import subprocess
import sys


def main(rep, kernel_substr, top=14):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    seen = set()
    for b in out.split('"Kernel Name",')[1:]:
        lines = b.splitlines()
        if kernel_substr not in lines[0] or lines[0] in seen:
            continue
        seen.add(lines[0])
        rows = list(csv.reader(lines[1:]))
        h = rows[0]
        si, ie, te, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Source")
        body = [(int(r[ie]), int(r[te]), int(r[si]), r[src]) for r in rows[1:] if len(r) > te and r[ie].isdigit()]
        tot = sum(x[0] for x in body)
        blocks, cur = [], []
        for x in body:
            if cur and x[0] != cur[-1][0]:
                blocks.append(cur)
                cur = []
            cur.append(x)
        if cur:
            blocks.append(cur)
        print(lines[0][:80], "total warp-inst", tot)
        for blk in sorted(blocks, key=lambda bl: -sum(x[0] for x in bl))[:top]:
            n = len(blk)
            execs = blk[0][0]
            ti = sum(x[1] for x in blk)
            wi = sum(x[0] for x in blk)
            ops = collections.Counter()
            for x in blk:
                toks = x[3].split()
                ops[(toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]] += 1
            print(f"  {100*wi/tot:5.1f}% | {n:4d} instr x {execs:9d} execs | simt {ti/max(wi,1):4.1f} | samples {sum(x[2] for x in blk):6d} | {dict(ops.most_common(8))}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
