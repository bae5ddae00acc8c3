"""
Aggregate an ncu source-page CSV by source line of the kernel body (development aid).

  ncu -i rep.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all lib.so ; nvdisasm -gi x.cubin > dis.txt
  python tools/ncu_lines.py src.csv dis.txt <mangled-substring> <file> <lo> <hi>
"""
import csv, re, sys, collections

src_csv, dis, fnsub, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
# offset -> frames
frames = {}
cur = []
fn = None
pending = []
for line in open(dis):
    m = re.match(r'//-+ \.text\.(\S+)', line)
    if m:
        fn = m.group(1); pending = []; continue
    if fn is None or fnsub not in fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        pending.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', line)
    if m:
        if pending:
            cur = pending; pending = []
        frames[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    addr = int(r[col['Address']], 16)
    if base is None: base = addr
    off = addr - base
    fr = frames.get(off, [])
    key = None
    for f, l in fr:            # innermost first; pick first frame inside the body range
        if f == fname and lo <= l <= hi:
            key = l; break
    if key is None:
        key = fr[0] if fr else ('?', 0)
    a = agg[key]
    for name in ('# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_no_inst', 'stall_barrier', 'stall_wait',
                 'stall_short_sb', 'stall_mio', 'stall_not_selected', 'stall_selected', 'stall_math', 'stall_lg', 'stall_branch_resolving', 'stall_dispatch'):
        v = int(float(r[col[name]] or 0))
        a[name] += v; tot[name] += v
print('TOTAL', dict(tot))
print(f'{"line":>28} {"samp":>7} {"inst":>9} {"long_sb":>7} {"no_inst":>7} {"barrier":>7} {"wait":>6} {"short":>6} {"mio":>5} {"notsel":>6} {"sel":>6} {"math":>5} {"lg":>5}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['# Samples'])[:45]:
    print(f'{str(k):>28} {a["# Samples"]:7d} {a["Instructions Executed"]:9d} {a["stall_long_sb"]:7d} {a["stall_no_inst"]:7d} {a["stall_barrier"]:7d} {a["stall_wait"]:6d} {a["stall_short_sb"]:6d} {a["stall_mio"]:5d} {a["stall_not_selected"]:6d} {a["stall_selected"]:6d} {a["stall_math"]:5d} {a["stall_lg"]:5d}')
