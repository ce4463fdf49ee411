import re, subprocess, sys, collections
cubin, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True).stdout
for p in re.split(r'\n\s*Function : ', txt)[1:]:
    name = p.split('\n', 1)[0]
    if pat not in name: continue
    lines = [l for l in p.split('\n') if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', l)]
    flat = [re.sub(r'^\s+/\*([0-9a-f]{4,5})\*/\s+', r'\1 ', re.sub(r'\s*;\s*/\*.*$', '', l)) for l in lines]
    addr = {f.split()[0]: n for n, f in enumerate(flat)}
    # the inner loop: the backward branch right after the stores, nearest CREDUX before
    red = [n for n, f in enumerate(flat) if 'CREDUX' in f or 'REDUX' in f]
    best = None
    for n, f in enumerate(flat):
        m = re.search(r'BRA(?:\.U)? .*0x([0-9a-f]+)$', f) or re.search(r'BRA 0x([0-9a-f]+)$', f)
        if m and m.group(1) in addr and addr[m.group(1)] < n and red and addr[m.group(1)] > red[0] and n - addr[m.group(1)] > 80:
            best = (addr[m.group(1)], n); break
    if not best: print('no loop found'); continue
    a, b = best
    ops = collections.Counter()
    for f in flat[a:b+1]:
        t = f.split(); op = t[2] if t[1].startswith('@') else t[1]
        ops[op.split('.')[0]] += 1
    print(name[:60], 'loop', flat[a].split()[0], '-', flat[b].split()[0], 'len', b - a + 1)
    print(' '.join(f'{k}:{v}' for k, v in ops.most_common()))
    if len(sys.argv) > 3:
        open(sys.argv[3], 'w').write('\n'.join(flat[a:b+1]) + '\n')
