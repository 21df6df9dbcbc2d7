"""Developer tool: per-source-line stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] == "Line No":
        continue
    if r[2] != "-":      # SASS row
        continue
    try:
        n, ie = int(r[6] or 0), int(r[7] or 0)
    except ValueError:
        continue
    k = (cur, int(r[0]))
    a = agg.setdefault(k, [0, 0, r[1]])
    a[0] += n
    a[1] += ie
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %8d inst  %s:%d  %s" % (100.0 * a[0] / max(tot, 1), a[1], k[0], k[1], a[2].strip()[:120]))
