"""Dump one kernel's SASS from the built library as 'addr instr' lines (needs cuobjdump)."""
import re, subprocess, sys
so, pat, out = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
for p in re.split(r'\n\s*Function : ', txt)[1:]:
    name = p.split('\n', 1)[0]
    if pat in name:
        lines = [l for l in p.split('\n') if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', l)]
        flat = [re.sub(r'^\s+/\*([0-9a-f]{4,5})\*/\s+', r'\1 ', re.sub(r'\s*;\s*/\*.*$', '', l)) for l in lines]
        open(out, 'w').write('\n'.join(flat) + '\n')
        print(name, len(flat))
