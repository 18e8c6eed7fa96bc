"""Benchmark workloads of BASELINE.json as flowz expression text + coefficients.

Pure text / coefficient builders: nothing here evaluates a graph.  bench.py's GPU arm and the tuning tools
take their graphs from this module; the oracle (oracle/flowz_oracle.py, test infrastructure) keeps its own
copies so that it stays independent of the product, and tests/test_kernel_class.py pins the two sets to be
identical.  Importable without the CUDA library (it does not import the package's C binding).
"""
from __future__ import annotations

import math
import struct
from typing import Sequence


def _f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", float(x)))[0]


def lit(x) -> str:
    """exact fp32 literal in flowz text"""
    return _f32(x).hex() + "f"


def rbj_lowpass(f: float, q: float = 0.707, sr: float = 44100.0):
    """RBJ low-pass section in flowz sign convention y = b0 x + b1 x1 + b2 x2 + a1 y1 + a2 y2
    (test/benchmark.cpp:25-26); formulae as in reactive_equations/reactive_filter_coeff.cpp:16-50."""
    w0 = 2.0 * math.pi * f / sr
    alpha = math.sin(w0) / (2.0 * q)
    a0 = 1.0 + alpha
    b1 = (1.0 - math.cos(w0)) / a0
    b0 = b1 / 2.0
    return (_f32(b0), _f32(b1), _f32(b0), _f32(2.0 * math.cos(w0) / a0), _f32(-(1.0 - alpha) / a0))


def biquad_df1(b0, b1, b2, a1, a2) -> str:
    """fwd |= bwd of test/benchmark.cpp:25-33"""
    return (f"({lit(b0)}*_1 + {lit(b1)}*_1[_1] + {lit(b2)}*_1[_2]"
            f" |= ~(_2 + {lit(a1)}*_1[_1] + {lit(a2)}*_1[_2]))")


def biquad_cascade(sections: int = 4) -> str:
    """`sections` stable RBJ low-pass DF1 sections in series, f = 440 * 2^k Hz; the octaves wrap after six
    sections so that every cutoff stays below Nyquist."""
    return " |= ".join(biquad_df1(*rbj_lowpass(440.0 * 2 ** (k % 6))) for k in range(sections))


def biquad_cascade_params(sections: int = 4) -> str:
    """Same cascade with every coefficient a run-time parameter $0..$(5*sections-1)."""
    parts = []
    for k in range(sections):
        p = 5 * k
        parts.append(f"(${p}*_1 + ${p+1}*_1[_1] + ${p+2}*_1[_2] |= ~(_2 + ${p+3}*_1[_1] + ${p+4}*_1[_2]))")
    return " |= ".join(parts)


def fir_taps(n: int = 256, cutoff: float = 0.25):
    """n-tap Hamming-windowed sinc low-pass, DC gain ~ 1, fp32 (list of floats)."""
    if n == 1:
        return [1.0]
    h = []
    for i in range(n):
        k = i - (n - 1) / 2.0
        x = k * cutoff
        sinc = 1.0 if x == 0 else math.sin(math.pi * x) / (math.pi * x)
        h.append(sinc * (0.54 - 0.46 * math.cos(2.0 * math.pi * i / (n - 1))))
    s = sum(h)
    return [_f32(v / s) for v in h]


def fir_expr(taps: Sequence[float]) -> str:
    """c0*_1 + c1*_1[_1] + ... (BASELINE configs[3]); C++ associates the sum to the left."""
    return " + ".join(f"{lit(c)}*_1" if k == 0 else f"{lit(c)}*_1[_{k}]" for k, c in enumerate(taps))


def osc_expr(f: float = 440.0, sr: float = 44100.0) -> str:
    """Recursive sine oscillator y = k*y1 - y2 + x, dirac-excited; k = 2 cos(2 pi f / sr)."""
    return f"~({lit(2.0 * math.cos(2.0 * math.pi * f / sr))}*_1[_1] - _1[_2] + _2)"


def osc_lp_expr(f: float = 440.0, a: float = 0.9) -> str:
    """BASELINE configs[2]: sine oscillator >> one-pole low-pass."""
    return f"{osc_expr(f)} |= ~(_2 + {lit(a)}*_1[_1])"


def poly_voice_expr(f: float = 440.0, g: float = 0.25) -> str:
    """BASELINE configs[4]: osc >> biquad >> (biquad inside a unit-delayed feedback loop of gain g)."""
    bq1 = biquad_df1(*rbj_lowpass(1760.0))
    bq2 = biquad_df1(*rbj_lowpass(3520.0))
    return f"{osc_expr(f)} |= {bq1} |= ~((_2 + {lit(g)}*_1[_1]) |= {bq2})"
