"""Split an ncu report's samples / instructions into [lo, hi] SASS offset range vs the rest."""
import csv, subprocess, sys, collections
rep, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = None
agg = {True: [0, 0], False: [0, 0]}
st = {True: collections.Counter(), False: collections.Counter()}
sc = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
top = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ix['Address']], 16)
    if base is None: base = a
    off = a - base
    inr = lo <= off <= hi
    s = int(r[ix['# Samples']] or 0); n = int(r[ix['Instructions Executed']] or 0)
    agg[inr][0] += s; agg[inr][1] += n
    for c in sc:
        st[inr][c] += int(r[ix[c]] or 0)
    top.append((s, hex(off), r[ix['Source']].strip(), inr))
for k, name in ((True, 'in range'), (False, 'outside')):
    print(name, 'samples', agg[k][0], 'instr', agg[k][1], dict(st[k].most_common(6)))
top.sort(reverse=True)
for s, o, src, inr in top[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print(s, o, 'LOOP' if inr else '    ', src[:90])
