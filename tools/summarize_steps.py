"""Summarise a GGML_B200_PROFILE_STEPS log (tools/profile_unet.py): time per step family and the heaviest shapes."""
import re, sys, collections
agg = collections.defaultdict(lambda: [0, 0.0]); byname = collections.defaultdict(float); tot = 0.0
for l in open(sys.argv[1]):
    m = re.match(r"step (\S+)\s+kind\s+(\d+)\s+([\d.]+) us\s+out\[([\d,]+)\] in0\[([\d,]+)\] M(\d+) N(\d+) K(\d+)", l)
    if not m: continue
    name, us = m.group(1), float(m.group(3))
    key = (name, m.group(4), m.group(5), m.group(6), m.group(7), m.group(8))
    agg[key][0] += 1; agg[key][1] += us; byname[name] += us; tot += us
print("total %.1f us" % tot)
for n, u in sorted(byname.items(), key=lambda x: -x[1]): print("%-16s %9.1f us %5.1f%%" % (n, u, 100 * u / tot))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for k, (c, u) in sorted(agg.items(), key=lambda x: -x[1][1])[:n]:
    print("%-12s out[%s] in[%s] M%s N%s K%s  x%d  %.1f us each, %.0f total" % (k[0], k[1], k[2], k[3], k[4], k[5], c, u / c, u))
