"""The C ABI boundary: the library loads, exports every symbol include/zignal_b200.h declares, and
fails loudly (never silently on a CPU) when there is no B200.  CPU only; no compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

import flowz_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "zignal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zg_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(zg):
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(zg.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/zignal_b200.h but not exported"
    assert sorted(zg.EXPORTED) == names           # and the Python binding covers the whole header


def test_library_has_sm100a_kernels_and_tma():
    so = os.path.join(ROOT, "zignal_b200", "libzignal_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "UTMASTG" in sass          # TMA loads and stores
    assert "FFMA" in sass and "LDS.128" in sass
    # the tensor-core FIR: tcgen05.mma (UTCHMMA), tcgen05.ld / st (LDTM / STTM), tcgen05.commit (UTCBAR)
    assert "UTCHMMA" in sass and "LDTM" in sass and "STTM" in sass and "UTCBAR" in sass


def test_no_gpu_is_a_loud_error(zg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = zg.compile(fo.biquad_cascade(4))
    with pytest.raises(zg.ZgError) as e:
        g.plan(channels=64)
    assert e.value.status == zg.ZG_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("layout", ["planar", "interleaved"])
def test_generated_kernels_compile_for_sm100a(zg, mode, layout):
    """NVRTC needs no device: every BASELINE graph's specialised kernel must build offline."""
    kw = dict(mode=zg.MODE_EXACT if mode == "exact" else zg.MODE_FAST,
              layout=zg.PLANAR if layout == "planar" else zg.INTERLEAVED)
    graphs = ["~(_2 + 0.9f*_1[_1])",
              "~(0x1.fcp0f*_1[_1] - _1[_2] + _2) |= ~(_2 + 0.9f*_1[_1])",
              fo.biquad_cascade(4), fo.biquad_cascade_params(2),
              "(_1 | _1[_1]) |= ~(_2 + _3 + 0.25f*_1[_2])"]
    for expr in graphs:
        g = zg.compile(expr)
        for uniform in (True, False):
            cubin = g.kernel(cubin=True, uniform_params=uniform, **kw)
            assert cubin[:4] == b"\x7fELF" and len(cubin) > 4096


def test_exact_mode_source_has_no_contractable_arithmetic(zg):
    src = zg.compile(fo.biquad_cascade(1)).kernel(mode=zg.MODE_EXACT).decode()
    tick = src[src.index("struct ZgTick"):]
    assert "__fmul_rn" in tick and "__fadd_rn" in tick
    assert not re.search(r"v\d+ [*+] v\d+", tick)


def test_bf16_storage_kernels_compile_for_sm100a(zg):
    """BASELINE configs[4]: the polyphonic voice chain with bf16 sample storage and a synthesised dirac
    input; the kernel converts with cvt.rn.bf16x2.f32 (F2FP in SASS) and keeps fp32 arithmetic."""
    g = zg.compile(fo.poly_voice_expr())
    for layout in (zg.PLANAR, zg.INTERLEAVED):
        cubin = g.kernel(cubin=True, io_dtype=zg.BF16, layout=layout, input_kind=[zg.IN_DIRAC])
        assert cubin[:4] == b"\x7fELF"
    src = g.kernel(io_dtype=zg.BF16).decode()
    assert "stream_block<ZgTick, false, true, 2, false>" in src
    with pytest.raises(zg.ZgError) as e:
        g.kernel(io_dtype=zg.I32)
    assert e.value.status == zg.ZG_ERR_ARG


@pytest.mark.parametrize("expr,n_ring_in,n_ring_out", [
    ("~(_2 + 0.5f*_1[_100])", 1, 1),                                        # feedback comb
    ("_1 + 0.5f*_1[_37] - 0.25f*_1[_1000]", 2, 1),                          # two far reads of one long line
    ("~(_2 + 0.5f*_1[_3] + 0.25f*_1[_500])", 1, 1),                         # the near read stays in a register window
    ("~(_2 + 0.4f*_1[_17]) |= ~(_2 - 0.3f*_1[_64]) |= (_1 , _1[_33])", 3, 2),
    ("~(_2 + 0.5f*_1[_16])", 0, 0),                                         # 16 floats: still register-resident
])
def test_long_delay_lines_become_ring_wires_of_the_generated_kernel(zg, expr, n_ring_in, n_ring_out):
    """Host-side split of long delay lines (zg_ir.cpp: split_long_lines): far reads become extra kernel inputs,
    pushes extra outputs; the kernel builds for sm_100a in every layout / storage (NVRTC, no device needed)."""
    g = zg.compile(expr)
    src = g.kernel().decode()
    tick = src[src.index("struct ZgTick"):]
    if n_ring_out:
        assert f"N_RING_IN = {n_ring_in}, N_RING_OUT = {n_ring_out}" in tick
        assert f"N_IN = {g.n_in + n_ring_in}, N_OUT = {g.n_out + n_ring_out}" in tick
    else:
        assert "N_RING_IN" not in tick
    for kw in (dict(), dict(layout=zg.INTERLEAVED), dict(io_dtype=zg.BF16), dict(mode=zg.MODE_FAST)):
        assert g.kernel(cubin=True, **kw)[:4] == b"\x7fELF"


def test_too_many_far_reads_is_unsupported_not_wrong(zg):
    many = " + ".join(f"0.1f*_1[_{100 + 10 * k}]" for k in range(12))
    with pytest.raises(zg.ZgError) as e:
        zg.compile(many).kernel()
    assert e.value.status == zg.ZG_ERR_UNSUPPORTED


def test_kernel_cache_on_disk(tmp_path):
    """ZG_KERNEL_CACHE_DIR keeps NVRTC cubins across processes; damaged or foreign files are ignored and replaced."""
    import struct
    import subprocess
    import sys
    prog = ("import sys, hashlib; sys.path.insert(0, %r); import zignal_b200 as zg; "
            "c = zg.compile('~(_2 + 0.5f*_1[_1]) |= _1 - 0.25f*_1[_3]').kernel(cubin=True); "
            "print(len(c), hashlib.sha1(c).hexdigest())" % ROOT)
    env = dict(os.environ, ZG_KERNEL_CACHE_DIR=str(tmp_path))

    def run():
        out = subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        return out.stdout.split()

    first = run()
    files = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(files) == 1 and not [f for f in os.listdir(tmp_path) if ".tmp" in f]
    path = os.path.join(tmp_path, files[0])
    blob = open(path, "rb").read()
    magic, key_len, check, n, body_hash = struct.unpack("<5Q", blob[:40])
    assert n == int(first[0]) == len(blob) - 40 and blob[40:44] == b"\x7fELF"

    def fnv1a(b):
        h = 0xcbf29ce484222325
        for c in b:
            h = ((h ^ c) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
        return h
    assert body_hash == fnv1a(blob[40:])
    # a well-formed entry is what the next process returns: swap the payload for a marker of the same size (with its hash)
    marker = b"M" * n
    open(path, "wb").write(struct.pack("<5Q", magic, key_len, check, n, fnv1a(marker)) + marker)
    import hashlib
    assert run() == [str(n), hashlib.sha1(marker).hexdigest()]
    # a body of the right length that does not match its hash (bit rot), a truncated file and one whose key check differs
    # are ignored, recompiled and rewritten
    open(path, "wb").write(blob[:40] + marker)
    assert run() == first and open(path, "rb").read() == blob
    open(path, "wb").write(blob[:100])
    assert run() == first and open(path, "rb").read() == blob
    open(path, "wb").write(struct.pack("<5Q", magic, key_len, check ^ 1, n, fnv1a(marker)) + marker)
    assert run() == first and open(path, "rb").read() == blob
