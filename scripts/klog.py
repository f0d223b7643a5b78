"""Aggregate a bench.py --kernel-log CSV by (class, description): launches, total ms, us per launch, TFLOP/s, GB/s."""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault((r["class"], r["desc"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += float(r["ms"]); a[2] += float(r["gflop"]); a[3] += float(r["mbytes"])
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms")
for (c, d), (n, ms, gf, mb) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{c} n={n:3d} ms={ms:7.3f} per={ms / n * 1e3:8.1f}us TF/s={gf / ms if ms else 0:7.1f} GB/s={mb / ms if ms else 0:7.0f} {d}")
