#!/usr/bin/env python3
"""Derives sde-sim-rs_b200/data/joe_kuo_d6_21201.bin from scipy's copy of new-joe-kuo-6.21201.

The reference gets the same table from the `sobol` crate (JoeKuoD6::extended, src/rng/sobol.rs:16);
that crate is not in this image, scipy's npz is the only in-container source.
Layout (little endian): 8-byte magic "SDEJK601", u32 ndims, u32 stride(=18),
u32 poly[ndims], u32 minit[ndims][18].  Row 0 is dimension 1 (van der Corput).
"""
import os
import struct

import numpy as np
import scipy

src = os.path.join(os.path.dirname(scipy.__file__), "stats", "_sobol_direction_numbers.npz")
z = np.load(src)
poly = z["poly"].astype("<u4")
vinit = z["vinit"].astype("<u4")
assert poly.shape == (21201,) and vinit.shape == (21201, 18)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sde-sim-rs_b200", "data", "joe_kuo_d6_21201.bin")
with open(out, "wb") as f:
    f.write(b"SDEJK601")
    f.write(struct.pack("<II", poly.shape[0], vinit.shape[1]))
    f.write(poly.tobytes())
    f.write(np.ascontiguousarray(vinit).tobytes())
print("wrote", os.path.abspath(out), os.path.getsize(out), "bytes")
