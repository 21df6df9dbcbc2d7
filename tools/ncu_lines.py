"""Developer tool: per-source-line digest of `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`:
the lines with the most stall samples and the most executed instructions, with lanes active per instruction."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] == "Line No" or r[2] != "-":      # keep the source rows, skip the SASS rows
        continue
    try:
        n, ie, te = int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur, int(r[0])), [0, 0, 0, r[1]])
    a[0] += n
    a[1] += ie
    a[2] += te
tot = max(sum(a[0] for a in agg.values()), 1)
ti = max(sum(a[1] for a in agg.values()), 1)
tt = sum(a[2] for a in agg.values())
print("samples %d, warp instructions %d, thread instructions %d, lanes per instruction %.2f" % (tot, ti, tt, tt / ti))
for title, key in (("by stall samples", 0), ("by executed instructions", 1)):
    print("\n-- top %d source lines %s --" % (top, title))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        print("%5.1f%% smp %5.1f%% inst %5.1f lanes  %s:%d  %s" % (100.0 * a[0] / tot, 100.0 * a[1] / ti, a[2] / max(a[1], 1), k[0], k[1], a[3].strip()[:110]))
