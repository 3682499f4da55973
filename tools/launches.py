"""Summarise an ncu --metrics gpu__time_duration.sum launch list: the last `n` launches (one step) with durations."""
import csv, re, sys, collections
path, n = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
step = seq[-n:]
def short(s):
    s = re.sub(r'psvae::|__nv_bfloat16', lambda m: 'bf16' if m.group(0) != 'psvae::' else '', s)
    s = re.sub(r'\(.*', '', s)
    return s.replace('void ', '')[:110]
agg = collections.OrderedDict()
for name, v in step:
    print(f"{v/1000:8.1f} us  {short(name)}") if '-v' in sys.argv else None
    k = short(name); agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v / 1000
tot = sum(v for _, v in step) / 1000
print(f"-- {len(seq)} launches captured; last {n}: total {tot:.1f} us")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:8.1f} us {100*t/tot:5.1f}%  x{c:<3d} {k}")
