"""Hot spots of an `ncu --page source --csv --print-source sass` dump: stall mix and the top instructions by samples.
usage: python profiles/sass_hot.py dump.csv [n_top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
def f(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print('instructions', len(data), 'samples', int(tot), 'warp instr executed', int(sum(f(r, 'Instructions Executed') for r in data)))
stalls = [h for h in hdr if h.startswith('stall_')]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print({k[6:]: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]})
top = sorted(range(len(data)), key=lambda i: -f(data[i], '# Samples'))[:ntop]
for i in sorted(top):
    r = data[i]
    st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
    print(i, r[col['Source']][:64].ljust(64), int(f(r, '# Samples')), int(f(r, 'Instructions Executed')),
          [(s[6:], int(v)) for v, s in st], 'wf', r[col['L1 Wavefronts Shared']], 'ideal', r[col['L1 Wavefronts Shared Ideal']])
