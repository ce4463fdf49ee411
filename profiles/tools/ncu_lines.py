import csv, sys, collections
"""per-source-line executed instructions / stall samples of one kernel of an ncu report:
python ncu_lines.py rep.ncu-rep [warp-steps to normalise by] [lines to print] [kernel-name regex]"""
rep, steps = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 8*24*482560/32
import subprocess
kern = ['-k', 'regex:' + sys.argv[4]] if len(sys.argv) > 4 else []
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'] + kern,capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Line No'][0]
hdr=rows[hi]
iinst=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
agg=[]; fname=None
ops=collections.Counter()
for r in rows:
    if r and r[0]=='File Path': fname=r[1].split('/')[-1]
    if len(r)>iinst and r[0] not in ('','Line No','File Path','Function Name'):
        try: agg.append((int(r[iinst]), int(r[isamp] or 0), fname, r[0], r[1].strip()[:100]))
        except: pass
    elif len(r)>iinst and r[0]=='' and len(r)>3:
        try:
            n=int(r[iinst]); parts=r[3].split(); op=parts[1] if parts[0].startswith('@') else parts[0]
            ops[op.split('.')[0]]+=n
        except: pass
tot=sum(a[0] for a in agg)
print('total warp instr',tot,'per warp-step %.1f'%(tot/steps))
agg.sort(reverse=True)
for n,s,f,l,src in agg[:int(sys.argv[3]) if len(sys.argv)>3 else 45]:
    print(f'{100*n/tot:5.1f}% {n/steps:6.1f}/step samp {s:6d} {f}:{l}: {src}')
print(' '.join(f'{k}:{v/steps:.1f}' for k,v in ops.most_common(28)))
