#!/usr/bin/env python
"""Inner-loop SASS of the hot kernels, for profiles/ (the evidence behind the instruction counts quoted in DESIGN.md).
For each kernel: resources (cuobjdump -res-usage), instruction histogram of the whole function, and the hottest loop -
the backward-branch-delimited region with the most FP64 (or FP32) math instructions - printed in full.
usage: tools/sass_excerpt.py pybnesian_b200/libpbn_cuda.so profiles/r2_sass"""
import collections
import re
import subprocess
import sys

lib, out_prefix = sys.argv[1], sys.argv[2]
KERNELS = [
    # (tag, mangled-name pattern, math opcodes, opcode the wanted loop must contain)
    # the f64 kernels hold three tile loops (dot-product form, difference form with / without the wide-range clamp); the
    # dot-product form - the one every benchmark runs - is the only one that adds the hoisted test-row norm to the
    # rounded exponent (VIADDMNMX)
    ("pair_f64_ckde_d4", r"pair_kernelIdLi4ELb1ELb0ELb0E", ("DFMA", "DADD", "DMUL"), "VIADDMNMX"),
    ("pair_f64_kde_d4", r"pair_kernelIdLi4ELb0ELb0ELb0E", ("DFMA", "DADD", "DMUL"), "VIADDMNMX"),
    ("pair_f32_ckde_d4", r"pair_kernelIfLi4ELb1ELb0ELb0E", ("FFMA2", "FADD2", "FFMA", "FADD", "MUFU"), "FFMA2"),
    ("ucv_f64_d4", r"ucv_kernelIdLi4E", ("DFMA", "DADD", "DMUL"), "VIADDMNMX"),
]
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
for tag, pat, math_ops, must in KERNELS:
    names = [n for n in funcs if re.search(pat, n)]
    if not names:
        print("no match for", pat)
        continue
    name = names[0]
    ins = funcs[name]
    addr = [int(re.search(r"/\*([0-9a-f]{4,})\*/", l).group(1), 16) for l in ins]
    op = lambda l: re.sub(r"^@!?U?P\d+\s+", "", re.sub(r"\s+/\*[0-9a-f]+\*/\s+", "", l).strip()).split()[0].split(".")[0]
    # loops: backward branches
    loops = []
    for i, l in enumerate(ins):
        m = re.search(r"BRA(?:\.\S+)?\s+(?:\S+,\s+)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) < addr[i] and int(m.group(1), 16) in addr:
            loops.append((addr.index(int(m.group(1), 16)), i))
    best = None
    for j, i in loops:  # innermost loops only: no other loop nested inside
        if any((j2, i2) != (j, i) and j <= j2 and i2 <= i for j2, i2 in loops):
            continue
        n_math = sum(1 for q in ins[j:i + 1] if op(q) in math_ops)
        if not any(must in q for q in ins[j:i + 1]):
            continue
        if best is None or n_math > best[0]:
            best = (n_math, j, i)
    with open("%s_%s.txt" % (out_prefix, tag), "w") as f:
        f.write("# %s\n# function %s\n" % (lib, name))
        r = re.search(re.escape(name) + r":\n\s+(.*)", res)
        f.write("# resources: %s\n" % (r.group(1) if r else "?"))
        hist = collections.Counter(op(l) for l in ins)
        f.write("# whole function, %d instructions: %s\n" % (len(ins), ", ".join("%s %d" % kv for kv in hist.most_common(14))))
        for key in ("UBLKCP", "SYNCS", "UTMALDG", "LDS", "FFMA2", "FADD2", "MUFU", "VIADDMNMX"):
            f.write("#   %-10s %d\n" % (key, sum(1 for l in ins if key in l)))
        if best:
            n_math, j, i = best
            loop = ins[j:i + 1]
            h = collections.Counter(op(l) for l in loop)
            f.write("# hottest loop: %d instructions, 0x%04x..0x%04x: %s\n" % (len(loop), addr[j], addr[i],
                                                                             ", ".join("%s %d" % kv for kv in h.most_common())))
            f.write("\n".join(loop) + "\n")
    print(tag, len(ins), best)
