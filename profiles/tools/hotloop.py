"""Instruction counts of the innermost store loop of a kernel in a built library.

    python profiles/tools/hotloop.py <lib.so> <mangled-name substring> [n_stores]

Finds the tightest backward branch whose body holds n_stores STG (default 6) and prints the body's
instruction mix -- a static count (branches inside the body are not weighed)."""
import collections
import re
import subprocess
import sys


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    n_st = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    for name, ins in kernels(lib).items():
        if pat not in name:
            continue
        best = None
        for k, (addr, text) in enumerate(ins):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt >= addr:
                continue
            body = [t for a, t in ins if tgt <= a <= addr]
            if sum("STG" in t for t in body) >= n_st and (best is None or len(body) < len(best[2])):
                best = (tgt, addr, body)
        if best is None:
            print(name, "no loop found")
            continue
        tgt, addr, body = best
        mix = collections.Counter()
        for t in body:
            op = re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0]
            mix[op] += 1
        print(f"{name[:60]}: total {len(ins)} instr; loop {tgt:#x}-{addr:#x}: {len(body)} instr")
        print("   " + "  ".join(f"{k}:{v}" for k, v in mix.most_common()))


if __name__ == "__main__":
    main()
