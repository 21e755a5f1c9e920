#!/usr/bin/env python
"""Static evidence for profiles/: per-kernel registers / spills / shared memory from `ptxas -v` and the
memory-instruction mix of the hot kernels from `cuobjdump -sass`, for the engine as the Makefile builds
it (sm_100a).  No GPU needed.
usage: python tools/sass_summary.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ["k2_query", "k2_cmat", "k3_mark", "k3_members", "k3_fix", "k3_bulk", "k_fill_part", "k_fill_apply",
       "k_phred", "k_polish_fill_warp", "k_probe_query"]
MEM = ("LDG", "STG", "ATOMG", "REDG", "RED.", "ATOMS", "LDL", "STL", "LDS", "STS", "BAR", "UTMA", "LDGSTS")


def main():
    with tempfile.TemporaryDirectory() as d:
        obj = os.path.join(d, "engine.o")
        cmd = ["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O3", "-lineinfo", "-gencode",
               "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-fopenmp", "-I" + ROOT + "/include",
               "-I" + ROOT + "/goldrush_b200/csrc", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v", "-c",
               ROOT + "/goldrush_b200/csrc/engine.cu", "-o", obj]
        log = subprocess.run(cmd, capture_output=True, text=True).stderr
        sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    print("# " + " ".join(cmd[:14]) + " ... engine.cu   (ptxas -v; cuobjdump -sass)")
    print("\n## registers / local memory / static shared memory per kernel (ptxas)\n")
    print(f"{'kernel':44s} {'regs':>5s} {'stack B':>8s} {'spill st B':>10s} {'spill ld B':>10s} {'smem B':>7s}")
    pat = (r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes"
           r" stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers"
           r"(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?")
    rows = []
    for m in re.finditer(pat, log):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        rows.append((name, int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(6) or 0)))
    for r in sorted(rows):
        print(f"{r[0]:44s} {r[1]:5d} {r[2]:8d} {r[3]:10d} {r[4]:10d} {r[5]:7d}")
    print("\n## memory / synchronisation instructions of the hot kernels (SASS, static counts)\n")
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        mangled = f.split("\n", 1)[0].strip()
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        if not any(name.startswith(h) for h in HOT):
            continue
        ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M))
        mem = {k: v for k, v in sorted(ops.items()) if k.startswith(MEM)}
        print(f"{name}: {sum(ops.values())} instructions")
        print("    " + ", ".join(f"{k} x{v}" for k, v in mem.items()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
