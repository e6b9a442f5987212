#!/usr/bin/env python3
"""Derive bioem_b200/data/quat_<n>.npy from the reference's orientation lists.

Run in the build container (needs /root/reference); the .npy fixtures are
committed because /root/reference does not exist on the GPU box.  Values are
parsed exactly like the reference does (fixed 12-character columns,
param.cpp:1254-1264 — SURVEY quirk Q5), so they are the float32 numbers the
reference itself computes with.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bioem_b200.synth import parse_orientation_list  # noqa: E402

REF = os.environ.get("BIOEM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "bioem_b200", "data")

for n in (576, 4608, 36864):
    q = parse_orientation_list(os.path.join(REF, "Quaternions", f"QUATERNION_LIST_{n}_Orient"))
    assert q.shape == (n, 4)
    ws = np.loadtxt(os.path.join(REF, "Quaternions", f"QUATERNION_LIST_{n}_Orient"), skiprows=1,
                    dtype=np.float32)
    ndiff = int((ws != q).any(axis=1).sum())
    print(f"{n}: rows differing from a whitespace parse: {ndiff}")
    np.save(os.path.join(OUT, f"quat_{n}.npy"), q)
