#!/usr/bin/env python3
"""Expands the reference's templated OpenCL-C kernel source IN PLACE from the reference tree
(never copied into this repository's history): runs the reference's own `conv_template.py`
on `pybnesian/kde/opencl_kernels/KDE.cl.src` — exactly what its build does
(/root/reference/expand_sources.py:6-15) — and writes the result under oracle/_ref/
(git-ignored build output).

usage: expand_kernels.py <reference_root> <output.inc>
"""
import os
import sys


def main():
    ref_root, out = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ref_root)
    import conv_template  # the reference's own template expander
    src = os.path.join(ref_root, "pybnesian", "kde", "opencl_kernels", "KDE.cl.src")
    text = conv_template.process_file(src)
    # `#line N "file"` markers are C-compatible; keep them so compiler messages point at the reference file
    with open(out, "w") as f:
        f.write(text)


if __name__ == "__main__":
    main()
