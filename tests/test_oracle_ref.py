"""Pins the oracle's biquad numerics to the reference itself.

oracle/_ref/libzg_ref_custom.so is the reference's unmodified test/benchmark.cpp compiled where it lies
(oracle/Makefile; the half that holds only the reference's own loops and is linked without the product).  Its hand-written `make_custom` loops (test/benchmark.cpp:35-126) are the only
Boost-free executable statement of the benchmark filters; the oracle must agree with them bit for
bit wherever the reference's own flowz graph and custom loop share an association (DF1, DF2, DF1T;
the DF2T custom loop subtracts a1*y where the graph adds (-a1)*y' -- a different expression tree).
The same vectors are committed under tests/golden/ so the check also runs where /root/reference
and _ref are absent.

The second half, oracle/_ref/libzg_ref_flow.so, is the drop-in check: the SAME translation unit's four make_flow()
graphs (test/benchmark.cpp:29-32, 60-63, 83-86, 111-114), spelled exactly as the reference spells them, compile
against include/flowz/flowz.hpp and tick through libzignal_b200's host voice.  They must reproduce the reference's
hand-written loops: bit for bit for DF1 / DF2 / DF1T, to 1e-5 for DF2T (a different expression tree, see above).
"""
import ctypes
import os

import numpy as np
import pytest

import flowz_oracle as fo
import reference_vectors as rv

P = ctypes.POINTER(ctypes.c_float)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "biquad_ref.npz")


def _ref_run(lib, fn, form, x):
    y = np.zeros_like(x)
    assert getattr(lib, fn)(form, x.ctypes.data_as(P), y.ctypes.data_as(P), ctypes.c_long(len(x))) == 0
    return y


def _inputs():
    dirac = np.zeros(201, np.float32); dirac[0] = 1.0          # sum_dirac's input, benchmark.cpp:137-147
    return {"dirac": dirac, "noise": fo.noise(1, 4096, seed=3)[0].copy()}


@pytest.mark.parametrize("form", [1, 2, 3])
def test_oracle_equals_reference_custom_loops(ref_lib, form):
    for name, x in _inputs().items():
        want = _ref_run(ref_lib, "zg_ref_custom", form, x)
        got = fo.Oracle(rv.bench_graphs()[form]).process([x[None, :]])[0][0]
        assert np.array_equal(got, want), (form, name)


def test_df2t_graph_vs_custom_is_close_not_equal(ref_lib):
    x = _inputs()["noise"]
    want = _ref_run(ref_lib, "zg_ref_custom", 4, x)
    got = fo.Oracle(rv.bench_graphs()[4]).process([x[None, :]])[0][0]
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("form", [1, 2, 3, 4])
def test_sum_dirac(ref_lib, form):
    # the reference's whole benchmark workload: 1 tick of 1.f + 200 ticks of 0.f, summed in float
    x = _inputs()["dirac"]
    y = fo.Oracle(rv.bench_graphs()[form]).process([x[None, :]])[0][0]
    acc = np.float32(0)
    for v in y:
        acc = np.float32(acc + v)
    if form != 4:
        assert acc == np.float32(ref_lib.zg_ref_sum_dirac_custom(form))
    else:
        assert abs(acc - ref_lib.zg_ref_sum_dirac_custom(form)) < 1e-5


@pytest.mark.parametrize("form", [1, 2, 3, 4])
def test_oracle_equals_committed_reference_vectors(form):
    """Same check against the committed outputs of the reference (tests/golden/make_golden.py)."""
    g = np.load(GOLDEN)
    for name in ("dirac", "noise"):
        x = g[f"x_{name}"]
        got = fo.Oracle(rv.bench_graphs()[form]).process([x[None, :]])[0][0]
        if form != 4:
            assert np.array_equal(got, g[f"custom{form}_{name}"])
        else:
            assert np.abs(got - g[f"custom4_{name}"]).max() <= 1e-5 * np.abs(g[f"custom4_{name}"]).max()


# ---- the reference's own make_flow() graphs, unmodified, running on this repository's shim ----------------------

@pytest.mark.parametrize("form", [1, 2, 3, 4])
def test_reference_make_flow_on_the_shim_equals_its_custom_loop(ref_lib, ref_flow_lib, form):
    for name, x in _inputs().items():
        want = _ref_run(ref_lib, "zg_ref_custom", form, x)
        got = _ref_run(ref_flow_lib, "zg_ref_flow", form, x)
        if form != 4:
            assert np.array_equal(got, want), (form, name)
        else:
            assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
        # and the oracle agrees with the shim bit for bit on all four (same expression tree on both sides)
        assert np.array_equal(got, fo.Oracle(rv.bench_graphs()[form]).process([x[None, :]])[0][0]), (form, name)


@pytest.mark.parametrize("form", [1, 2, 3, 4])
def test_reference_sum_dirac_on_the_shim(ref_lib, ref_flow_lib, form):
    """sum_dirac (test/benchmark.cpp:137-147), the reference's whole benchmark workload, over make_flow() on the shim
    against the same loop over make_custom(); DF1 sums to 2.9444442."""
    flow, custom = ref_flow_lib.zg_ref_sum_dirac_flow(form), ref_lib.zg_ref_sum_dirac_custom(form)
    if form != 4:
        assert np.float32(flow) == np.float32(custom)
    else:
        assert abs(flow - custom) < 1e-5
    if form == 1:
        assert np.float32(flow) == np.float32(2.9444442)


def test_reference_only_half_does_not_link_the_product():
    import subprocess
    so = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libzg_ref_custom.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libzignal_b200" not in needed
