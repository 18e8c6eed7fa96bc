"""zignal_b200 -- B200-native block evaluator for flowz signal graphs (Python side of the C ABI).

This module is a thin ctypes binding of include/zignal_b200.h; it holds no numerics.  The compute
path is libzignal_b200.so (host C++ front end + sm_100a CUDA kernels).  Importing the package
without the built library raises; creating a Plan without a B200 raises ZgError -- there is no CPU
fallback for block evaluation.

Mirrors the reference's interface for the path (flowz/flowz.hpp, andre-bergner/zignal):
    compile(expr)            -> Graph           flowz.hpp:1233-1249
    Graph.voice()(x1..xN)    -> tuple           stateful_lambda::operator(), flowz.hpp:1225-1229
    Graph.plan(channels=...) -> Plan            new: C voices, block-wise, on one B200
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzignal_b200.so")

ZG_OK = 0
ZG_ERR_PARSE, ZG_ERR_GRAPH, ZG_ERR_ARG, ZG_ERR_UNSUPPORTED, ZG_ERR_CUDA, ZG_ERR_INTERNAL = -1, -2, -3, -4, -5, -6
I32, F32, F64 = 0, 1, 2
MODE_EXACT, MODE_FAST = 0, 1
PLANAR, INTERLEAVED = 0, 1
IN_BUFFER, IN_DIRAC, IN_ZERO = 0, 1, 2
TP_AUTO, TP_OFF, TP_WARMUP, TP_TWO_PASS = 0, 1, 2, 3     # zg_time_parallel
MAX_WIRES = 8
I32, F32, F64, BF16 = 0, 1, 2, 3            # zg_dtype (BF16: sample storage only)
NONLINEAR, AFFINE, LINEAR = 0, 1, 2          # zg_linearity
C64, C128, TYPE_OPEN = 4, 5, -1             # complex<float> / <double> and the open `absorber` type: result_types() only


class ZgError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[zg_status {status}] {message}")
        self.status = status


class GraphInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_in", "n_out", "n_params", "n_state", "n_lines", "n_nodes", "all_f32")]


class PlanOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("channels", C.c_int64), ("mode", C.c_int), ("layout", C.c_int),
                ("io_dtype", C.c_int), ("lanes_per_channel", C.c_int), ("input_kind", C.c_int * MAX_WIRES),
                ("force_jit", C.c_int), ("time_parallel", C.c_int), ("fir_tensor_cores", C.c_int), ("section_warps", C.c_int),
                ("reserved", C.c_int * 4)]


class PlanInfo(C.Structure):
    _fields_ = [("kernel", C.c_char * 96), ("jit", C.c_int), ("lanes_per_channel", C.c_int),
                ("host_chunks", C.c_int), ("regs_per_thread", C.c_int), ("smem_bytes", C.c_int),
                ("launches", C.c_int), ("threads_per_cta", C.c_int), ("stages", C.c_int),
                ("uniform_params", C.c_int), ("boxes", C.c_int), ("time_segments", C.c_int),
                ("segment_samples", C.c_int), ("warmup_samples", C.c_int), ("linearity", C.c_int)]


class SplitPlan(C.Structure):              # zg_split_plan
    _fields_ = [(n, C.c_int) for n in ("groups_per_cta", "warps_per_group", "sections_per_warp", "grid", "threads_per_cta",
                                      "stages", "boxes_per_tile", "boxes_per_handover", "smem_bytes",
                                      "segments", "segment_boxes", "warmup_boxes")]


def split_plan(sections: int, channels: int, samples: int, exact: bool = True, sm_count: int = 148,
               max_smem: int = 232448, segments: int = 1, warmup_samples: int = 0):
    """The geometry a launch of the section-split biquad kernel would get (host only), or None when the block stays on
    the other biquad kernels."""
    o = SplitPlan()
    if not lib.zg_split_plan_query(sections, int(exact), channels, samples, sm_count, max_smem, segments, warmup_samples, C.byref(o)):
        return None
    return o


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built; run `python zignal_b200/build.py` "
                          "(zignal_b200 has no pure-Python path)")
    lib = C.CDLL(LIB_PATH)
    vp, cp, ci, i64, sz = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_size_t
    P = C.POINTER
    sig = {
        "zg_last_error": (cp, []),
        "zg_version": (cp, []),
        "zg_expr_arity": (ci, [cp, P(ci), P(ci)]),
        "zg_expr_delays": (ci, [cp, ci, P(ci), ci, P(ci)]),
        "zg_expr_canonical": (ci, [cp, C.c_char_p, sz]),
        "zg_expr_result_types": (ci, [cp, P(ci), ci, P(ci), ci, P(ci), P(ci)]),
        "zg_graph_compile": (ci, [cp, P(vp)]),
        "zg_graph_destroy": (None, [vp]),
        "zg_graph_get_info": (ci, [vp, P(GraphInfo)]),
        "zg_graph_canonical": (cp, [vp]),
        "zg_graph_dump": (cp, [vp]),
        "zg_graph_kernel_class": (ci, [vp, C.c_char_p, sz]),
        "zg_graph_linearity": (ci, [vp, P(ci)]),
        "zg_graph_state_matrix": (ci, [vp, P(C.c_float), ci, P(C.c_double), sz]),
        "zg_graph_settling_time": (ci, [vp, P(C.c_float), ci, ci, ci, C.c_double, P(ci)]),
        "zg_voice_create": (ci, [vp, P(vp)]),
        "zg_voice_clone": (ci, [vp, P(vp)]),
        "zg_voice_destroy": (None, [vp]),
        "zg_voice_tick": (ci, [vp, P(C.c_double), P(ci), P(C.c_double), P(ci)]),
        "zg_voice_set_param": (ci, [vp, ci, C.c_float]),
        "zg_voice_state": (ci, [vp, P(P(C.c_float)), P(ci)]),
        "zg_plan_opts_default": (None, [P(PlanOpts)]),
        "zg_graph_kernel_compile": (ci, [vp, P(PlanOpts), ci, ci, C.c_char_p, sz, P(sz)]),
        "zg_plan_create": (ci, [vp, P(PlanOpts), P(vp)]),
        "zg_plan_destroy": (None, [vp]),
        "zg_plan_get_info": (ci, [vp, P(PlanInfo)]),
        "zg_process": (ci, [vp, P(vp), P(vp), i64, i64, i64, vp]),
        "zg_process_host": (ci, [vp, P(vp), P(vp), i64, i64, i64]),
        "zg_state_reset": (ci, [vp]),
        "zg_state_get": (ci, [vp, P(C.c_float), sz]),
        "zg_state_set": (ci, [vp, P(C.c_float), sz]),
        "zg_param_set": (ci, [vp, ci, P(C.c_float), i64]),
        "zg_param_set_device": (ci, [vp, ci, vp, i64]),
        "zg_shard_range": (ci, [i64, ci, ci, P(i64), P(i64)]),
        "zg_split_plan_query": (ci, [ci, ci, i64, i64, ci, ci, ci, ci, P(SplitPlan)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header and library disagree
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()
EXPORTED = ["zg_last_error", "zg_version", "zg_expr_arity", "zg_expr_delays", "zg_expr_canonical", "zg_expr_result_types",
            "zg_graph_compile", "zg_graph_destroy", "zg_graph_get_info", "zg_graph_canonical", "zg_graph_dump", "zg_graph_kernel_class", "zg_graph_linearity",
            "zg_graph_state_matrix", "zg_graph_settling_time",
            "zg_voice_create", "zg_voice_clone", "zg_voice_destroy", "zg_voice_tick", "zg_voice_set_param",
            "zg_voice_state", "zg_plan_opts_default", "zg_graph_kernel_compile", "zg_plan_create",
            "zg_plan_destroy", "zg_plan_get_info", "zg_process", "zg_process_host", "zg_state_reset",
            "zg_state_get", "zg_state_set", "zg_param_set", "zg_param_set_device", "zg_shard_range", "zg_split_plan_query"]


def _check(status: int) -> None:
    if status != ZG_OK:
        raise ZgError(status, lib.zg_last_error().decode(errors="replace"))


def version() -> str:
    return lib.zg_version().decode()


# ---- static analysis on bare expressions (flowz.hpp:162-246, 443-506, 794-805) ---------------------

def arity(expr: str):
    a, b = C.c_int(), C.c_int()
    _check(lib.zg_expr_arity(expr.encode(), C.byref(a), C.byref(b)))
    return a.value, b.value


def delays(expr: str, minimum: bool = False):
    cap = 512
    while True:                                   # *count is the number of wires, also when it exceeds the capacity
        buf = (C.c_int * cap)()
        n = C.c_int()
        _check(lib.zg_expr_delays(expr.encode(), 1 if minimum else 0, buf, cap, C.byref(n)))
        if n.value <= cap:
            return list(buf[:n.value])
        cap = n.value


def canonical(expr: str) -> str:
    buf = C.create_string_buffer(len(expr) * 16 + 4096)
    _check(lib.zg_expr_canonical(expr.encode(), buf, len(buf)))
    return buf.value.decode()


def result_types(expr: str, in_dtypes):
    """ResultType (flowz.hpp:515-644): (types of the output wires, is_tuple) for the given input wire types."""
    ins = (C.c_int * max(len(in_dtypes), 1))(*in_dtypes)
    out = (C.c_int * 64)()
    n, tup = C.c_int(), C.c_int()
    _check(lib.zg_expr_result_types(expr.encode(), ins, len(in_dtypes), out, 64, C.byref(n), C.byref(tup)))
    return list(out[:n.value]), bool(tup.value)


# ---- compile() ------------------------------------------------------------------------------------------

class Graph:
    """compile(expr): front panel + canonical form + tick program (flowz.hpp:1233-1249)."""

    def __init__(self, expr: str):
        h = C.c_void_p()
        _check(lib.zg_graph_compile(expr.encode(), C.byref(h)))
        self._h = h
        self.expr = expr
        info = GraphInfo()
        _check(lib.zg_graph_get_info(h, C.byref(info)))
        for f, _ in GraphInfo._fields_:
            setattr(self, f, getattr(info, f))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:                 # (module globals are gone at interpreter shutdown)
            lib.zg_graph_destroy(h)

    @property
    def canonical(self) -> str:
        return lib.zg_graph_canonical(self._h).decode()

    def dump(self) -> str:
        return lib.zg_graph_dump(self._h).decode()

    def kernel_class(self) -> str:
        """'biquad_df1:<S>' | 'fir:<N>' | 'generated' | 'host-only' (zg_graph_kernel_class)."""
        buf = C.create_string_buffer(64)
        _check(lib.zg_graph_kernel_class(self._h, buf, len(buf)))
        return buf.value.decode()

    def linearity(self) -> int:
        """NONLINEAR (0) | AFFINE (1) | LINEAR (2): is one tick a linear map of inputs and state (zg_graph_linearity)."""
        k = C.c_int()
        _check(lib.zg_graph_linearity(self._h, C.byref(k)))
        return k.value

    def state_space(self, params: Sequence[float] = ()):
        """(A, B, C, D) of a LINEAR graph, float64 [n_state, n_state], [n_state, n_in], [n_out, n_state], [n_out, n_in]:
        state' = A state + B x, y = C state + D x, with the state in zg_state_get order.  Read off the host tick by
        probing it with unit vectors (exact: every entry is one coefficient path of the tick program).  What a
        time-parallel evaluation of few, long channels would scan over."""
        import numpy as np
        if self.linearity() != LINEAR:
            raise ValueError("state_space() needs a LINEAR graph (zg_graph_linearity)")
        ns, ni, no = self.n_state, self.n_in, self.n_out
        A, B = np.zeros((ns, ns)), np.zeros((ns, ni))
        Cm, D = np.zeros((no, ns)), np.zeros((no, ni))
        v = Voice(self)
        for k, p in enumerate(params):
            v.set_param(k, float(p))
        p_state, n = C.POINTER(C.c_float)(), C.c_int()
        _check(lib.zg_voice_state(v._h, C.byref(p_state), C.byref(n)))
        for j in range(ns + ni):
            for i in range(ns):
                p_state[i] = 1.0 if i == j else 0.0
            y = v.tick(*[1.0 if ns + i == j else 0.0 for i in range(ni)], dtypes=[F32] * ni)
            col = [p_state[i] for i in range(ns)]
            if j < ns:
                A[:, j], Cm[:, j] = col, y
            else:
                B[:, j - ns], D[:, j - ns] = col, y
        return A, B, Cm, D

    def state_matrix(self, params: Sequence[float] = ()):
        """A of state' = A state + B x + c, float64 [n_state, n_state] (zg_graph_state_matrix): what the time-segmented
        launches of few, long channels are planned with.  Same matrix as state_space()[0]; affine ticks allowed."""
        import numpy as np
        n = self.n_state
        a = np.zeros((n, n), np.float64)
        prm = (C.c_float * max(len(params), 1))(*[float(p) for p in params])
        _check(lib.zg_graph_state_matrix(self._h, prm, len(params), a.ctypes.data_as(C.POINTER(C.c_double)), max(a.size, 1)))
        return a

    def settling_time(self, params: Sequence[float] = (), step: int = 128, k_max: int = 8192, tol: float = 2.0 ** -30) -> int:
        """Smallest K = m * step <= k_max with |A^K|_inf <= tol, 0 if the tick does not forget its state that fast
        (zg_graph_settling_time): the warm-up length of ZG_TP_WARMUP."""
        k = C.c_int()
        prm = (C.c_float * max(len(params), 1))(*[float(p) for p in params])
        _check(lib.zg_graph_settling_time(self._h, prm, len(params), step, k_max, tol, C.byref(k)))
        return k.value

    def voice(self) -> "Voice":
        return Voice(self)

    def plan(self, channels: int, **kw) -> "Plan":
        return Plan(self, channels, **kw)

    def kernel(self, cubin: bool = False, uniform_params: bool = True, **kw) -> bytes:
        """CUDA source / sm_100a cubin of the generated-tick kernel for this graph (no GPU needed)."""
        o = _opts(1, **kw)
        n = C.c_size_t()
        _check(lib.zg_graph_kernel_compile(self._h, C.byref(o), int(uniform_params), int(cubin), None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(lib.zg_graph_kernel_compile(self._h, C.byref(o), int(uniform_params), int(cubin), buf, n.value, C.byref(n)))
        return buf.raw[:n.value]


def compile(expr: str) -> Graph:   # noqa: A001  (the reference's name)
    return Graph(expr)


class Voice:
    """One voice on the host: stateful_lambda (flowz.hpp:1181-1230).  Call it with n_in numbers;
    Python ints are C++ ints, Python floats are C++ floats unless dtype=F64 is given."""

    def __init__(self, graph: Graph, _handle=None):
        self.graph = graph
        if _handle is None:
            _handle = C.c_void_p()
            _check(lib.zg_voice_create(graph._h, C.byref(_handle)))
        self._h = _handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.zg_voice_destroy(h)

    def clone(self) -> "Voice":
        h = C.c_void_p()
        _check(lib.zg_voice_clone(self._h, C.byref(h)))
        return Voice(self.graph, h)

    def set_param(self, index: int, value: float) -> None:
        _check(lib.zg_voice_set_param(self._h, index, value))

    def tick(self, *xs, dtypes: Optional[Sequence[int]] = None):
        n_in, n_out = self.graph.n_in, self.graph.n_out
        if len(xs) != n_in:
            raise TypeError(f"graph takes {n_in} inputs, got {len(xs)}")
        if dtypes is None:
            dtypes = [I32 if isinstance(x, int) and not isinstance(x, bool) else F32 for x in xs]
        ins = (C.c_double * max(n_in, 1))(*[float(x) for x in xs])
        dts = (C.c_int * max(n_in, 1))(*dtypes)
        outs = (C.c_double * n_out)()
        odt = (C.c_int * n_out)()
        _check(lib.zg_voice_tick(self._h, ins, dts, outs, odt))
        self.out_dtypes = tuple(odt)         # the C++ type of every returned value (zg_dtype), as the reference's tuple has it
        return tuple(int(v) if t == I32 else float(v) for v, t in zip(outs, odt))

    __call__ = tick

    @property
    def state(self):
        p = C.POINTER(C.c_float)()
        n = C.c_int()
        _check(lib.zg_voice_state(self._h, C.byref(p), C.byref(n)))
        return [p[i] for i in range(n.value)]


def _opts(channels: int, device: int = 0, mode: int = MODE_EXACT, layout: int = PLANAR,
          input_kind: Optional[Sequence[int]] = None, lanes_per_channel: int = 0, force_jit: bool = False,
          io_dtype: int = 1, time_parallel: int = 0, fir_tensor_cores: int = 0, section_warps: int = 0) -> PlanOpts:
    o = PlanOpts()
    lib.zg_plan_opts_default(C.byref(o))
    o.device, o.channels, o.mode, o.layout, o.io_dtype = device, channels, mode, layout, io_dtype
    o.lanes_per_channel, o.force_jit, o.time_parallel = lanes_per_channel, int(force_jit), time_parallel
    o.fir_tensor_cores = fir_tensor_cores
    o.section_warps = section_warps
    for i, k in enumerate(input_kind or []):
        o.input_kind[i] = k
    return o


def empty_block(rows: int, cols: int, device=None, dtype=None):
    """CUDA sample buffer [rows, cols] (float32, or bfloat16) whose row pitch is a multiple of 16 bytes,
    as the TMA tensor maps behind zg_process require; a view into a padded allocation otherwise."""
    import torch
    dtype = dtype or torch.float32
    m = 16 // torch.empty((), dtype=dtype).element_size()
    ld = (cols + m - 1) // m * m
    return torch.empty((rows, ld), dtype=dtype, device=device or "cuda")[:, :cols]


def to_block(array, device=None, dtype=None):
    """Host [rows, cols] float32 data -> device buffer with a legal pitch (see empty_block); with
    dtype=torch.bfloat16 the samples are rounded to nearest even on the way."""
    import numpy as np
    import torch
    a = torch.from_numpy(np.ascontiguousarray(array, np.float32))
    out = empty_block(a.shape[0], a.shape[1], device, dtype)
    out.copy_(a)
    return out


class Plan:
    """`channels` independent voices of one graph on one B200; process() evaluates one block.

    Buffers are torch CUDA tensors (float32): planar [channels, samples] or interleaved
    [samples, channels].  State persists across process() calls like consecutive ticks."""

    def __init__(self, graph: Graph, channels: int, **kw):
        self.graph = graph
        self.channels = channels
        self.opts = _opts(channels, **kw)
        h = C.c_void_p()
        _check(lib.zg_plan_create(graph._h, C.byref(self.opts), C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.zg_plan_destroy(h)

    @property
    def interleaved(self) -> bool:
        return self.opts.layout == INTERLEAVED

    def info(self) -> PlanInfo:
        i = PlanInfo()
        _check(lib.zg_plan_get_info(self._h, C.byref(i)))
        return i

    def _buffer_inputs(self):
        return [k for k in range(self.graph.n_in) if self.opts.input_kind[k] == IN_BUFFER]

    def process_ptrs(self, in_ptrs, out_ptrs, n_samples, ld_in, ld_out, stream=0) -> None:
        ins = (C.c_void_p * MAX_WIRES)(*in_ptrs)
        outs = (C.c_void_p * MAX_WIRES)(*out_ptrs)
        _check(lib.zg_process(self._h, ins, outs, n_samples, ld_in, ld_out, C.c_void_p(stream)))

    def process(self, inputs, outputs=None, n_samples: Optional[int] = None):
        """inputs: one CUDA tensor per graph input (None for synthesised ones).  Returns outputs."""
        import torch
        ins = list(inputs)
        if len(ins) != self.graph.n_in:
            raise TypeError(f"graph takes {self.graph.n_in} inputs")
        ref = next((t for t in ins if t is not None), None)
        if n_samples is None:
            if ref is None:
                raise TypeError("n_samples is required when every input is synthesised")
            n_samples = ref.shape[0] if self.interleaved else ref.shape[1]
        shape = (n_samples, self.channels) if self.interleaved else (self.channels, n_samples)
        dev = torch.device("cuda", self.opts.device)
        dt = torch.bfloat16 if self.opts.io_dtype == BF16 else torch.float32
        if outputs is None:
            outputs = [empty_block(shape[0], shape[1], dev, dt) for _ in range(self.graph.n_out)]
        for t in [t for t in ins if t is not None] + list(outputs):
            if t.dtype != dt or not t.is_cuda or t.stride(1) != 1 or tuple(t.shape) != shape:
                raise TypeError(f"buffers must be {dt} CUDA tensors of shape {shape} with unit inner stride")
        ld_in = ref.stride(0) if ref is not None else (shape[1] + 7) // 8 * 8
        # zg_process takes one pitch for all inputs and one for all outputs
        if any(t is not None and t.stride(0) != ld_in for t in ins) or any(t.stride(0) != outputs[0].stride(0) for t in outputs):
            raise TypeError("all input buffers must share one row pitch, and all output buffers one row pitch")
        in_ptrs = [t.data_ptr() if t is not None else None for t in ins]
        out_ptrs = [t.data_ptr() for t in outputs]
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.process_ptrs(in_ptrs, out_ptrs, n_samples, ld_in, outputs[0].stride(0), stream)
        return outputs

    def process_host(self, inputs, outputs=None, n_samples: Optional[int] = None):
        """Same with host arrays (numpy float32 or CPU torch tensors): H2D, kernel, D2H, synchronise."""
        import numpy as np
        bf16 = self.opts.io_dtype == BF16                       # bf16 blocks: CPU torch.bfloat16 tensors
        ins = [None if x is None else (x if hasattr(x, "data_ptr") else np.ascontiguousarray(x, np.float32)) for x in inputs]
        ref = next((t for t in ins if t is not None), None)
        if n_samples is None:
            n_samples = ref.shape[0] if self.interleaved else ref.shape[1]
        shape = (n_samples, self.channels) if self.interleaved else (self.channels, n_samples)
        if bf16:
            import torch
            if any(t is not None and not (hasattr(t, "data_ptr") and t.dtype == torch.bfloat16) for t in ins):
                raise TypeError("a bf16 plan takes torch.bfloat16 host tensors")
            if outputs is None:
                outputs = [torch.empty(shape, dtype=torch.bfloat16) for _ in range(self.graph.n_out)]
        elif outputs is None:
            outputs = [np.empty(shape, np.float32) for _ in range(self.graph.n_out)]

        for t in [t for t in ins if t is not None] + list(outputs):
            is_torch = hasattr(t, "data_ptr")
            ok = tuple(t.shape) == shape and (t.is_contiguous() if is_torch else t.flags["C_CONTIGUOUS"])
            if not bf16:
                ok = ok and str(t.dtype).endswith("float32")
            if not ok:
                raise TypeError(f"host buffers must be contiguous {'bfloat16' if bf16 else 'float32'} arrays of shape {shape}")

        def ptr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

        in_ptrs = (C.c_void_p * MAX_WIRES)(*[None if t is None else ptr(t) for t in ins])
        out_ptrs = (C.c_void_p * MAX_WIRES)(*[ptr(t) for t in outputs])
        _check(lib.zg_process_host(self._h, in_ptrs, out_ptrs, n_samples, shape[1], shape[1]))
        return outputs

    def reset(self) -> None:
        _check(lib.zg_state_reset(self._h))

    def get_state(self):
        import numpy as np
        a = np.empty((self.graph.n_state, self.channels), np.float32)
        _check(lib.zg_state_get(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), a.size))
        return a

    def set_state(self, a) -> None:
        import numpy as np
        a = np.ascontiguousarray(a, np.float32)
        _check(lib.zg_state_set(self._h, a.ctypes.data_as(C.POINTER(C.c_float)), a.size))

    def set_param(self, index: int, values) -> None:
        import numpy as np
        a = np.ascontiguousarray(np.atleast_1d(values), np.float32)
        _check(lib.zg_param_set(self._h, index, a.ctypes.data_as(C.POINTER(C.c_float)), a.size))

    def set_param_device(self, index: int, values) -> None:
        """values: float32 CUDA tensor [channels] (e.g. coefficients computed with torch on the device)."""
        import torch
        if not (values.is_cuda and values.dtype == torch.float32 and values.is_contiguous()):
            raise TypeError("set_param_device needs a contiguous float32 CUDA tensor (one value per channel)")
        _check(lib.zg_param_set_device(self._h, index, C.c_void_p(values.data_ptr()), values.numel()))
