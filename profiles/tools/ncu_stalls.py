"""Per-source-line sample and instruction totals from an ncu report (source page, CUDA+SASS view)."""
import csv, sys, collections, subprocess
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]
iinst = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
agg = []; fname = None
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]
    if len(r) > iinst and r[0] not in ('', 'Line No', 'File Path', 'Function Name'):
        try: agg.append((int(r[isamp] or 0), int(r[iinst]), fname, r[0], r[1].strip()[:90]))
        except Exception: pass
ts = sum(a[0] for a in agg); ti = sum(a[1] for a in agg)
print('samples', ts, 'warp instr', ti)
agg.sort(reverse=True)
for s, n, f, l, src in agg[:top]:
    print(f'{100*s/ts:5.1f}% samp {100*n/ti:5.1f}% inst  {f}:{l}: {src}')
