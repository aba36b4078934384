"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals and one training step's breakdown."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    try: t = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    u = row['Metric Unit']
    t = t / 1e3 if u == 'ns' else (t * 1e3 if u == 'ms' else t)
    seq.append((re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('wun::', ''), t, row['Grid Size']))
idx = [i for i, s in enumerate(seq) if s[0] == 'adam_kernel']
print("launches captured", len(seq), "adam at", idx)
if len(idx) >= 2: st = seq[idx[0] + 1: idx[1] + 1]
elif idx: st = seq[max(0, idx[0] - 340): idx[0] + 1]
else: st = seq
tot = sum(s[1] for s in st)
print("ONE STEP: %d launches, %.1f us (serialised, cold-cache device time)" % (len(st), tot))
a = collections.OrderedDict()
for s in st:
    a.setdefault(s[0], [0, 0.0]); a[s[0]][0] += 1; a[s[0]][1] += s[1]
for k, v in sorted(a.items(), key=lambda kv: -kv[1][1]):
    print("   %-34s n=%3d %9.1f us %5.1f%%" % (k[:34], v[0], v[1], 100 * v[1] / tot))
if len(sys.argv) > 2:
    for s in st: print("%-30s %8.1f %s" % (s[0][:30], s[1], s[2]))
