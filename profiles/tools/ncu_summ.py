import csv, sys, collections, re
fn = sys.argv[1]
rows = []
with open(fn) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = r['Kernel Name']; name = re.sub(r'\(.*', '', name); name = re.sub(r'<.*', '', name)
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
    tot[name][0] += 1; tot[name][1] += ms
all_ms = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print('%-70s %5d %10.3f ms %5.1f%%' % (k[:70], v[0], v[1], 100 * v[1] / all_ms))
print('total', all_ms)
