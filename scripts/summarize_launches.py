"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 2:]
ki, vi, gi, bi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size'), H.index('Block Size')
agg = collections.defaultdict(lambda: [0, 0.0])
per = collections.defaultdict(collections.Counter)
for r in data:
    if len(r) <= vi:
        continue
    n = re.sub(r'\(.*', '', r[ki])
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    agg[n][0] += 1; agg[n][1] += v
    per[n][(r[gi], round(v / 1000))] += 1
tot = sum(v[1] for v in agg.values())
print("total ms %.3f  launches %d" % (tot / 1e6, sum(v[0] for v in agg.values())))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6.2f%% %8.3f ms %5d  %s" % (100 * t / tot, t / 1e6, c, n[:120]))
    if len(sys.argv) > 3:
        for k, v in sorted(per[n].items(), key=lambda kv: -kv[0][1] * kv[1])[:int(sys.argv[3])]:
            print("            grid %-16s %6d us x %d" % (k[0], k[1], v))
