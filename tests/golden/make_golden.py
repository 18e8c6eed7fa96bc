"""Generates tests/golden/biquad_ref.npz from the reference itself (oracle/_ref/libzg_ref_custom.so =
/root/reference/test/benchmark.cpp compiled where it lies).  Run in the build container:
    make -C oracle ref && python tests/golden/make_golden.py
Outputs of the reference's hand-written biquad loops (make_custom, benchmark.cpp:35-126) on the
benchmark's dirac (201 samples) and on 4096 samples of hash noise.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import flowz_oracle as fo  # noqa: E402

lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libzg_ref_custom.so"))
P = ctypes.POINTER(ctypes.c_float)
dirac = np.zeros(201, np.float32); dirac[0] = 1.0
noise = fo.noise(1, 4096, seed=3)[0].copy()
out = {"x_dirac": dirac, "x_noise": noise}
for form in (1, 2, 3, 4):
    for name, x in (("dirac", dirac), ("noise", noise)):
        y = np.zeros_like(x)
        assert lib.zg_ref_custom(form, x.ctypes.data_as(P), y.ctypes.data_as(P), ctypes.c_long(len(x))) == 0
        out[f"custom{form}_{name}"] = y
np.savez_compressed(os.path.join(HERE, "biquad_ref.npz"), **out)
print("wrote", os.path.join(HERE, "biquad_ref.npz"))
