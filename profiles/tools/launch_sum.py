"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python launch_sum.py list.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    k = r[ki][:60]
    tot[k] = tot.get(k, 0.0) + float(r[vi].replace(',', '')) / 1e6
    cnt[k] += 1
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'{v:9.3f} ms {cnt[k]:4d} launches {100 * v / s:5.1f}%  avg {v / cnt[k]:.4f} ms  {k}')
print(f'total {s:.3f} ms')
