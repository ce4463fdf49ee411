"""Differently-compiled builds of the SAME sources for A/B timing on the GPU box.

    python profiles/tools/build_variants.py name=-DMSB_THERMO_MINB=6 other="-DMSB_THERMO_MINB=8 -DX=1" ...

A token ``SRC:disagg_interior.cu=profiles/tools/experiments/x.cu`` among a variant's flags compiles that
file in place of the product source.  Writes profiles/variants/lib_<name>.so (git-ignored like every .so; it travels with gpurun) plus
lib_base.so (no extra flags).  Only the sources that honour a knob are recompiled per variant; the
others are compiled once to objects.  Select one at run time with MSB_LIBRARY=<path>
(metsim_b200/_lib.py).  Prints registers / spills of the kernels the knobs touch.
"""
import os
import re
import shlex
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from metsim_b200 import build as B  # noqa: E402

OUTDIR = os.path.join(ROOT, "profiles", "variants")
KNOB_SOURCES = ("disagg.cu", "disagg_interior.cu", "disagg_streams.cu", "daily.cu", "solar.cu")
WATCH = ("thermo_knot_kernelIfLi", "sw_day_kernelIfLb0", "prec_day_kernelIfLb0", "daily_fused_kernel",
         "solar_rows_kernel", "solar_moments_kernel")


def run(cmd):
    subprocess.check_call(cmd)


def main():
    os.makedirs(OUTDIR, exist_ok=True)
    nvcc = B._nvcc()
    cflags = [f for f in B.FLAGS if f != "-shared"]
    objs = {}
    for s in B.SOURCES:
        if s in KNOB_SOURCES:
            continue
        o = os.path.join(OUTDIR, s.replace(".cu", ".o"))
        run([nvcc, *cflags, "-I", B.INCLUDE, "-c", os.path.join(B.CSRC, s), "-o", o])
        objs[s] = o
    variants = [("base", "")] + [a.split("=", 1) for a in sys.argv[1:]]
    for name, flags in variants:
        vobjs = []
        # a token SRC:<file.cu>=<path> compiles <path> in place of csrc/<file.cu> (experimental kernels
        # kept under profiles/tools/experiments/, outside the product sources and their hash)
        toks = shlex.split(flags)
        override = dict(t[4:].split("=", 1) for t in toks if t.startswith("SRC:"))
        toks = [t for t in toks if not t.startswith("SRC:")]
        for s in KNOB_SOURCES:
            o = os.path.join(OUTDIR, f"{name}_{s.replace('.cu', '.o')}")
            src = os.path.join(ROOT, override[s]) if s in override else os.path.join(B.CSRC, s)
            run([nvcc, *cflags, *toks, "-I", B.INCLUDE, "-I", B.CSRC, "-c", src, "-o", o])
            vobjs.append(o)
        out = os.path.join(OUTDIR, f"lib_{name}.so")
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs.values(), *vobjs])
        res = subprocess.run(["cuobjdump", "--dump-resource-usage", out], capture_output=True, text=True).stdout
        print(f"== {name}: {flags}")
        fn = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                fn = m.group(1)
            elif fn and any(w in fn for w in WATCH) and "REG" in line:
                short = re.sub(r"^_ZN3msb\d+", "", fn)[:40]
                print("   %-40s %s" % (short, " ".join(t for t in line.split() if t.split(":")[0] in ("REG", "STACK", "SHARED", "LOCAL"))))
                fn = None
        for o in vobjs:
            os.remove(o)
    for o in objs.values():
        os.remove(o)


if __name__ == "__main__":
    main()
