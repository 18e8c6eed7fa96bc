"""Known-answer vectors the reference's own tests pin for the flowz path.

Transcribed from /root/reference/test/tests.cpp and flowz/README.md as *data* (expression text in
flowz syntax + expected values); line numbers are cited per entry.  Used twice: against the oracle
(tests/test_oracle_golden.py) and against the product's front end / host voice
(tests/test_frontend.py), so a disagreement tells which side left the reference.
"""

# G12: un2bin(~(...)) == bin_fb(L, R)              test/tests.cpp:27-60
# (expression under ~, expected promise part L, expected future part R).  `|=` is right
# associative, written here exactly as in the reference.
CANONICAL_SPLITS = [
    ("~( _1[_1] )", "_1", "_1[_1]", 27),
    ("~( _1 |= _1[_1] )", "_1 |= _1", "_1[_1]", 28),
    ("~( (_1 |= _1) |= _1[_1] )", "_1 |= _1 |= _1", "_1[_1]", 30),
    ("~( _1 |= (_1 |= _1[_1]) )", "_1 |= _1 |= _1", "_1[_1]", 31),
    ("~( (_1 |= _1[_1]) |= (_1 |= _1[_1]) )", "_1 |= _1", "_1[_1] |= (_1 |= _1[_1])", 33),
    ("~( (_1 |= _1[_1] |= _1) |= (_1[_1]) )", "_1 |= _1", "(_1[_1] |= _1) |= _1[_1]", 34),
    ("~( (_1) |= (_1[_1] |= _1 |= _1[_1]) )", "_1 |= _1", "_1[_1] |= _1 |= _1[_1]", 35),
    ("~( _1 |= _1[_1] + _2 )", "_1 |= _1", "_1[_1] + _2", 37),
    ("~( _1+2 |= _1[_1] + _2 )", "_1 |= _1+2", "_1[_1] + _2", 38),
    ("~( _1+2 |= _1[_1] - 13 + _2 )", "_1 |= _1+2", "_1[_1] - 13 + _2", 39),
    ("~( _1 + _2 |= _1[_1] |= (_1,_1) )", "_1|_1 |= _1 + _2", "_1[_1] |= (_1,_1)", 41),
    ("~( _2 |= _1 )", "_1", "_2 |= _1", 43),
    ("~( _2 |= _1[_1] + _2[_1] )", "_1", "_2 |= _1[_1] + _2[_1]", 44),
    ("~( _2 |= _1 |= _1[_1] + _2[_1] )", "_1", "_2 |= _1 |= _1[_1] + _2[_1]", 45),
    ("~( _1 + _2 |= _1[_1] )", "_1 |= _1 + _2", "_1[_1]", 47),
    ("~( (_1,_1) |= _1[_1] + _2[_1] )", "_1 |= (_1,_1)", "_1[_1] + _2[_1]", 52),
    ("~( (_1,_1) |= _1 + _2 |= _1[_1] )", "_1 |= (_1,_1) |= _1 + _2", "_1[_1]", 53),
    ("~~( _1[_1] + _2[_1] )", "_1", "bfb( _1 , _1[_1] + _2[_1] )", 58),
    ("~(_1[_1] |= ~( _1[_1] + _2 ))", "_1", "_1[_1] |= bfb( _1 , _1[_1] + _2 )", 60),
]

# G10 / G11: arity and max delays of the canonical form       test/tests.cpp:67-77
CANONICAL_ARITY = [
    ("~( _1 + _2[_1] |= _1[_1] + _2 )", 2, 1, [1, 0], 67),
    ("~( _1 + _3[_1] |= _1[_1] + _2 )", 3, 1, [0, 1, 0], 73),
]

# G8 / G9: wires around boxes                                 test/tests.cpp:88-102
#   (expr, n_in, n_out, inputs, outputs).  Line 101 of the reference compares a 1-tuple with the
#   2-tuple result (ill-formed with current standard libraries); line 102 pins the size, the
#   fan-out rule (flowz.hpp:765-768) pins both values.
WIRES_AROUND = [
    ("_1 |= _2", 2, 1, (2, 1337), (1337,), 88),
    ("(_1,_1) |= _1", 1, 2, (1337,), (1337, 1337), 96),
]

# G1-G7: known-answer ticks, state persists across calls      test/tests.cpp:110-178
#   (expr, [(inputs, outputs), ...], line)
TICKS = [
    ("_1", [((1337,), (1337,)), ((42,), (42,))], 110),
    ("_1[_1]", [((1337,), (0,)), ((42,), (1337,)), ((17,), (42,))], 116),
    ("_1 - _1[_1]", [((1337,), (1337,)), ((42,), (42 - 1337,)), ((17,), (17 - 42,))], 123),
    ("~(_1[_1] + _2)", [((1337,), (1337,)), ((42,), (1337 + 42,)), ((17,), (1337 + 42 + 17,))], 130),
    ("_1 |= _1[_1]", [((1337,), (0,)), ((0,), (1337,)), ((0,), (0,))], 143),
    ("_1 |= (_1[_1],_2[_2])", [((1337, 42), (0, 0)), ((0, 0), (1337, 0)), ((0, 0), (0, 42))], 149),
    ("~(_1[_1] + _2 |= _1)", [((1337,), (1337,)), ((42,), (1337 + 42,)), ((17,), (1337 + 42 + 17,))], 168),
    ("~(_1 |= _1[_1] + _2)", [((1337,), (1337,)), ((42,), (1337 + 42,)), ((17,), (1337 + 42 + 17,))], 174),
    # flowz/README.md:9-22 (unit delay) and :25-38 (integrator)
    ("_1[_1]", [((1,), (0,)), ((2,), (1,)), ((3,), (2,))], 9),
    ("~(_1[_1] + _2)", [((1,), (1,)), ((2,), (3,)), ((3,), (6,)), ((4,), (10,))], 25),
]

# G15: ResultType                                             test/tests.cpp:182-232
#   (expression, expected types of the output wires, is a std::tuple, line); the input is tuple<float> throughout
#   (:192).  Types: "i" int, "f" float, "d" double, "c" std::complex<float>.  cplx{1,0} as in :188.  Lines 225-228
#   apply make_canonical first ("c(...)"): canonical=True.
RESULT_TYPES = [
    ("_1", "f", False, False, 198),
    ("_1 * 1.0", "d", False, False, 199),
    ("_1 |= _1", "f", True, False, 201),
    ("_1 * 1.0 |= _1", "d", True, False, 202),
    ("_1 |= 1.0 * _1", "d", True, False, 203),
    ("_1 |= cplx{1,0} * _1", "c", True, False, 205),
    ("2*_1 |= _1*cplx{1,0} |= _1", "c", True, False, 206),
    ("(_1,_1)", "ff", True, False, 208),
    ("(_1,_1*1.0)", "fd", True, False, 209),
    ("(_1*1.0,_1)", "df", True, False, 210),
    ("(1.0*_1,1.0*_1)", "dd", True, False, 211),
    ("(_1*1.0,_1) |= (_2,_1)", "fd", True, False, 212),
    ("(_1,1.0*_1) |= (_1|_1)", "fd", True, False, 213),
    ("_1[_1]", "f", False, False, 215),
    ("_1 |= _1[_1]", "f", True, False, 216),
    ("(_1[_1],1.0*_1) |= _2[_1]", "d", True, False, 217),
    ("~( _1[_1] + _2 )", "f", True, False, 219),
    ("~( 1.0*_1[_1] + _2 )", "d", True, False, 220),
    ("~( _1[_1] + 1.0*_2 )", "d", True, False, 221),
    ("~( _1[_1] + 1.0 )", "d", True, False, 222),
    ("~( _1[_1] + _2 )", "f", True, True, 226),
    ("~( 1.0*_1[_1] + _2 )", "d", True, True, 227),
    ("~( _1[_1] + 1.0*_2 )", "d", True, True, 228),
    ("~( _1[_1] + 1.0 )", "d", True, True, 229),
]
TYPE_CODE = {"i": 0, "f": 1, "d": 2, "c": 4}

# G14: the reference's benchmark graphs (test/benchmark.cpp:14-129), coefficients :18-23.
BENCH_COEF = dict(b0=0.2, b1=-0.3, b2=1.1, a1=-0.2, a2=0.8)


def bench_graphs():
    """flowz text of make_flow for DF1 / DF2 / DF1T / DF2T with the reference's float constants."""
    import numpy as np
    c = {k: float(np.float32(v)).hex() + "f" for k, v in BENCH_COEF.items()}
    n = {k: float(-np.float32(v)).hex() + "f" for k, v in BENCH_COEF.items()}     # -a2 is a float constant
    neg = lambda s: f"({s})" if s.startswith("-") else s
    c = {k: neg(v) for k, v in c.items()}
    n = {k: neg(v) for k, v in n.items()}
    fwd = f"({c['b0']}*_1 + {c['b1']}*_1[_1] + {c['b2']}*_1[_2])"
    bwd = f"~(_2 + {c['a1']}*_1[_1] + {c['a2']}*_1[_2])"
    da2 = "(_1[_1] + _2 |= _1[_1] + _2)"
    fwdt = f"(({c['b2']}*_1 , {c['b1']}*_1 , {c['b0']}*_1) |= {da2})"
    bwdt = f"(({n['a2']}*_1 , {n['a1']}*_1) |= {da2})"
    return {
        1: f"{fwd} |= {bwd}",            # :31
        2: f"{bwd} |= {fwd}",            # :62
        3: f"~{bwdt} |= {fwdt}",         # :85
        4: f"{fwdt} |= ~{bwdt}",         # :113
    }
