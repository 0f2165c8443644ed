"""Join a MELOGAN_TRACE=1 log of one cycle with the ncu launch list of the same cycle: time and TFLOP/s per tap-GEMM shape."""
import collections, csv, sys
trace, launches = sys.argv[1], sys.argv[2]
lines = [l.strip() for l in open(trace) if l.startswith('[tc_tap]')]
n = len(lines) // 3          # profile_cycle.py runs 2 warm-up cycles + 1 profiled
cyc = lines[-n:]
rows = list(csv.reader(open(launches)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 2:]
ki, vi, gi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size')
tc = [(float(r[vi].replace(',', '')) / 1000, r[gi]) for r in data if len(r) > vi and 'tc_tapgemm' in r[ki]]
assert len(cyc) == len(tc), (len(cyc), len(tc))
agg = collections.OrderedDict()
tot = 0
for l, (us, g) in zip(cyc, tc):
    m = dict(kv.split('=') for kv in l.split()[1:] if '=' in kv)
    fl = 2.0 * int(m['rows']) * int(m['N']) * int(m['K']) * int(m['taps'])
    a = agg.setdefault(l, [0, 0.0, fl, g]); a[0] += 1; a[1] += us; tot += us
print("tap-GEMM launches %d, %.2f ms per cycle" % (len(tc), tot / 1000))
for k, (c, us, fl, g) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("%7.1f us x%2d = %6.2f ms %5.0f TF/s  %-13s %s" % (us / c, c, us / 1000, fl / (us / c) / 1e6, g, k[9:]))
