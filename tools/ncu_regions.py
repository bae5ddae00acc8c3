"""Per-barrier-region breakdown of an ncu source-page CSV (development aid).

  ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_regions.py src.csv <events> <warps_per_cta>
"""
import bisect, collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
nev = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
nw = float(sys.argv[3]) if len(sys.argv) > 3 else 16.0
hi = next(i for i, r in enumerate(rows[:10]) if 'Source' in r)
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]


def f(r, h):
    try:
        return float(r[col[h]] or 0)
    except ValueError:
        return 0.0


stalls = ['stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_math', 'stall_short_sb', 'stall_not_selected',
          'stall_selected', 'stall_mio', 'stall_lg', 'stall_no_inst', 'stall_branch_resolving', 'stall_dispatch']
bars = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[col['Source']]]
reg = collections.defaultdict(collections.Counter)
ops = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for i, r in enumerate(data):
    g = bisect.bisect_left(bars, i)
    s = r[col['Source']].split()
    op = (s[1] if s[0].startswith('@') else s[0]).split('.')[0].rstrip(';')
    n = f(r, 'Instructions Executed') / nev / nw
    ops[g][op] += n
    reg[g]['inst'] += n
    reg[g]['samp'] += f(r, '# Samples')
    for h in stalls:
        reg[g][h] += f(r, h)
        tot[h] += f(r, h)
T = sum(v['samp'] for v in reg.values())
print('total samples', T, 'inst/thread/event', sum(v['inst'] for v in reg.values()))
print('stalls', ' '.join(f'{h[6:]}={100 * v / T:.1f}%' for h, v in tot.most_common()))
for g in sorted(reg):
    c = reg[g]
    if c['samp'] < 0.005 * T:
        continue
    print(f"R{g}: {100 * c['samp'] / T:5.1f}% inst {c['inst']:6.0f} | " +
          ' '.join(f"{h[6:]}={100 * c[h] / T:.1f}" for h in stalls if c[h] > 0.005 * T) + ' | ' +
          ' '.join(f'{o}={v:.0f}' for o, v in ops[g].most_common(14)))
