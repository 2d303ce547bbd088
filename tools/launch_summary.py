"""Summarise an ncu launch list CSV (--metrics gpu__time_duration.sum): per-kernel count / total / mean, and the sequence."""
import csv, sys, collections
path = sys.argv[1]
seq_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path, errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
seq = []
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: seq.append((r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", ""), float(r[mv].replace(",", "")) / 1e3))
    except ValueError: pass
tot = collections.OrderedDict()
for k, v in seq:
    c = tot.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
total = sum(v for _, v in seq)
print(f"{len(seq)} launches, {total:.1f} us")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[:70]:70s} n={n:4d} total={t:10.1f} us mean={t/n:9.1f} share={t/total*100:5.1f}%")
for k, v in seq[:seq_n]:
    print(f"    {k[:60]:60s} {v:9.1f}")
