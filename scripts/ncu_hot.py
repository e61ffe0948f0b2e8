"""Print the hottest SASS lines (by stall samples) of one kernel in an .ncu-rep source page.
usage: python scripts/ncu_hot.py REPORT [launch_skip] [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
rows = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
si, src, ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[si]) for r in rows)
print("total samples", tot, "rows", len(rows))
top = sorted(range(len(rows)), key=lambda i: -int(rows[i][si]))[:topn]
for i in sorted(top):
    r = rows[i]
    st = {h: int(r[hdr.index(h)]) for h in hdr if h.startswith("stall_") and "Not" not in h and int(r[hdr.index(h)]) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5d %6s %5.1f%% ex=%-8s %-62s %s" % (i, r[si], 100 * int(r[si]) / max(tot, 1), r[ex], r[src].strip()[:62], st))
