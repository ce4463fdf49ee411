import csv, sys, collections, subprocess
rep=sys.argv[1]; steps=float(sys.argv[2]) if len(sys.argv)>2 else 8*24*482560/32
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
ops=collections.Counter(); tot=0; stall=collections.Counter()
sc=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    try: n=int(r[ix['Instructions Executed']])
    except: continue
    parts=r[ix['Source']].split(); op=parts[1] if parts[0].startswith('@') else parts[0]
    ops[op.split('.')[0]]+=n; tot+=n
    for s in sc:
        try: stall[s]+=int(r[ix[s]] or 0)
        except: pass
print('SASS lines',len(rows)-2,'total warp instr',tot,'per warp-step %.1f'%(tot/steps))
print(' '.join(f'{k}:{v/steps:.1f}' for k,v in ops.most_common(40)))
print({k:v for k,v in stall.most_common(8)})
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rr=list(csv.reader(raw.splitlines())); h=rr[0]; d=rr[2]
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','launch__registers_per_thread','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg']:
    if k in h: print(k, d[h.index(k)], rr[1][h.index(k)])
