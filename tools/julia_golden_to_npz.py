#!/usr/bin/env python
"""Converts the output directory of tools/julia_golden.jl (raw column-major arrays + manifest.txt) into one .npz with the
repository's array conventions: a Julia (n_a, n_e) matrix becomes a NumPy (n_e, n_a) C-order array, vectors stay vectors.
usage: python tools/julia_golden_to_npz.py <julia_golden_dir> tests/golden/julia_ks.npz"""
import os
import sys

import numpy as np

src, dst = sys.argv[1], sys.argv[2]
out = {}
for line in open(os.path.join(src, "manifest.txt")):
    name, dt, shape = line.split()
    dims = tuple(int(s) for s in shape.split("x"))
    a = np.fromfile(os.path.join(src, name + ".bin"), dtype="<" + dt)
    a = a.reshape(dims[::-1])                 # column-major (d1, d2) -> C-order (d2, d1): the repository's layout
    out[name] = a[0] if a.size == 1 and len(dims) == 1 else a
np.savez_compressed(dst, **out)
print("wrote", dst, "with", len(out), "arrays")
